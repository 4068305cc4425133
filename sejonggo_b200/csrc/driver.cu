// driver.cu — the native per-step driver: one MCTS search step of EVERY game in one ABI call
// (select -> gather the leaves awaiting evaluation -> tower forward with the symmetry gathers
// fused -> expand -> backup), and the per-ply record packer.  The Python host only loops.
//
//   mode A step = self_play.simulate for all games            (self_play.py:28-120)
//   mode B step = nomodel_self_play.async_simulate2 (one wave, incl. the "No best leaf" resumes,
//                 nomodel_self_play.py:59-82) for all games
#include "engine.h"

static inline cudaStream_t S_(void *s) { return (cudaStream_t)s; }

// leaf slots awaiting evaluation whose game uses network slot `want` (model_of_game NULL = all -> slot 0);
// also gathers the symmetry id of each: per game (mode A: one draw per simulate batch, symmetry.py:128 via
// self_play.py:70) or per leaf slot (mode B: one draw per put_predict_request, predicting_queue_worker.py:88-92)
__global__ void k_leaf_gather(const LeafRef *leaf_refs, int total, int L, const int32_t *model_of_game, int want,
                              const int32_t *sym_in, int sym_per_leaf, int32_t *index, int32_t *sym_out, int32_t *count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    bool v = i < total && leaf_refs[i].state == 1;
    if (v && model_of_game) v = model_of_game[i / L] == want;
    unsigned bal = __ballot_sync(SGO_FULL, v);
    if (!bal) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(bal));
    base = __shfl_sync(SGO_FULL, base, 0);
    if (v) {
        int o = base + __popc(bal & ((1u << lane) - 1u));
        index[o] = i;
        if (sym_out) sym_out[o] = sym_in ? sym_in[sym_per_leaf ? i : i / L] : 0;
    }
}

static int ensure_step_buffers(sgo_engine *e)
{
    if (e->step_policy) return 0;
    size_t GL = (size_t)e->G * e->L;
    SGO_CUDA_OK(e, cudaMalloc(&e->step_policy, sizeof(float) * GL * e->A));
    SGO_CUDA_OK(e, cudaMalloc(&e->step_value, sizeof(float) * GL));
    SGO_CUDA_OK(e, cudaMalloc(&e->step_index, sizeof(int32_t) * GL));
    SGO_CUDA_OK(e, cudaMalloc(&e->step_sym, sizeof(int32_t) * GL));
    SGO_CUDA_OK(e, cudaMemset(e->step_policy, 0, sizeof(float) * GL * e->A));
    SGO_CUDA_OK(e, cudaMemset(e->step_value, 0, sizeof(float) * GL));
    return 0;
}

extern "C" int sgo_tower_max_positions(sgo_engine *e, int32_t slot);

// evaluate every leaf awaiting evaluation with the network slot(s) and leave policy/value in the
// per-slot step buffers; returns the number evaluated
static int eval_pending_leaves(sgo_engine *e, const int32_t *d_model_of_game, const int32_t *d_sym_game, int sym_per_leaf, int *n_out, void *stream)
{
    int total = e->G * e->L, done = 0;
    for (int slot = 0; slot < (d_model_of_game ? 2 : 1); slot++) {
        SGO_CUDA_OK(e, cudaMemsetAsync(e->counters + 2, 0, sizeof(int32_t), S_(stream)));
        k_leaf_gather<<<(total + 255) / 256, 256, 0, S_(stream)>>>(e->leaf_refs, total, e->L, d_model_of_game, slot, d_sym_game, sym_per_leaf,
                                                                   e->step_index, d_sym_game ? e->step_sym : nullptr, e->counters + 2);
        SGO_LAUNCHED(e);
        SGO_CUDA_OK(e, cudaMemcpyAsync(e->h_pinned + 7, e->counters + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, S_(stream)));
        SGO_CUDA_OK(e, cudaStreamSynchronize(S_(stream)));
        int n = e->h_pinned[7];
        int mp = sgo_tower_max_positions(e, slot);
        if (n > 0 && mp <= 0) return sgo_fail(e, "selfplay_step: network slot has no weights");
        for (int s = 0; s < n; s += mp) {
            int c = n - s < mp ? n - s : mp;
            int rc = sgo_tower_forward(e, slot, 1, e->step_index + s, c, d_sym_game ? e->step_sym + s : nullptr, 1,
                                       e->step_policy, e->step_value, stream);
            if (rc) return rc;
        }
        done += n;
    }
    *n_out = done;
    return 0;
}

extern "C" int sgo_selfplay_step(sgo_engine *e, int32_t mode, const int32_t *d_tree_sel, const int32_t *d_model_of_game,
                                 int32_t leaves, int32_t total_energy, const int32_t *d_sym_game, int32_t sym_per_leaf,
                                 int32_t *h_leaves_done, void *stream)
{
    int rc = ensure_step_buffers(e);
    if (rc) return rc;
    int done = 0, n = 0;
    if (mode == 0) {
        rc = sgo_tree_select_a(e, d_tree_sel, leaves, stream);
        if (rc) return rc;
        rc = eval_pending_leaves(e, d_model_of_game, d_sym_game, sym_per_leaf, &n, stream);
        if (rc) return rc;
        done = n;
        rc = sgo_tree_expand(e, d_tree_sel, e->step_policy, e->step_value, stream);
        if (rc) return rc;
        rc = sgo_tree_backup_a(e, d_tree_sel, stream);
        if (rc) return rc;
    } else {
        int restart = 1;
        for (;;) {
            int32_t counts[2];
            rc = sgo_tree_select_b_sync(e, d_tree_sel, leaves, restart, counts, stream);
            if (rc) return rc;
            restart = 0;
            if (counts[0] == 0) break;
            rc = eval_pending_leaves(e, d_model_of_game, d_sym_game, sym_per_leaf, &n, stream);
            if (rc) return rc;
            done += n;
            rc = sgo_tree_expand(e, d_tree_sel, e->step_policy, e->step_value, stream);
            if (rc) return rc;
            if (counts[1] == 0) break;
        }
        rc = sgo_tree_backup_b(e, d_tree_sel, total_energy, stream);
        if (rc) return rc;
    }
    if (h_leaves_done) *h_leaves_done = done;
    // a game whose allocation failed ran no (or fewer) simulations in THIS step: say so now, not at the end of the run
    SGO_CUDA_OK(e, cudaMemcpyAsync(e->h_pinned + 12, e->err_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, S_(stream)));
    SGO_CUDA_OK(e, cudaStreamSynchronize(S_(stream)));
    if (e->h_pinned[12] & SGO_ERR_ARENA)
        return sgo_fail(e, "MCTS node pool exhausted during this search step: raise arena_blocks (sgo_pool_stats_sync shows the low-water mark)", -5);
    return 0;
}

// ---------------------------------------------------------------- records
// One record per game per ply, the content of the reference's move_data (self_play.py:207-214)
// in packed form: [0..PW) packed board (16 planes x W words + to_move), [PW] move index,
// [PW+1] value bits (f32), [PW+2] tree valid flag, [PW+3 .. PW+3+A) policy target = root priors as f32.
__global__ void k_records_pack(const Board *boards, int S, int G, int T, Pool pool, const TreeMeta *meta,
                               const double *root_p64, const int32_t *tree_sel, const int32_t *moves, const float *values,
                               uint32_t *out, int rec_words)
{
    __shared__ uint32_t scratch[4][SGO_AWORDS];
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= G) return;
    uint32_t *dst = out + (size_t)g * rec_words;
    uint32_t *sc = scratch[(threadIdx.x >> 5) & 3];
    const Board *b = boards + g;
    int W = (S * S + 31) / 32, PW = 16 * W + 1, A = S * S + 1;
    int tm = b->to_move, head = b->head;
    uint32_t rm = row_mask(S, lane);
    for (int k = 0; k < SGO_HIST; k++) {
        int slot = (head + SGO_HIST - k) & (SGO_HIST - 1);
        uint32_t bl = lane < SGO_ROWW ? b->st[slot][0][lane] : 0u, wh = lane < SGO_ROWW ? b->st[slot][1][lane] : 0u;
        uint32_t own = tm == 1 ? bl : wh, opp = tm == 1 ? wh : bl;
        illegal_rows_to_words(own & rm, S, lane, sc);
        if (lane < W) dst[(2 * k) * W + lane] = sc[lane];
        __syncwarp();
        illegal_rows_to_words(opp & rm, S, lane, sc);
        if (lane < W) dst[(2 * k + 1) * W + lane] = sc[lane];
        __syncwarp();
    }
    int tsel = tree_sel ? tree_sel[g] : 0;
    int tree = g * T + (tsel < 0 ? 0 : tsel);
    TreeMeta m = meta[tree];
    bool tv = m.valid && tsel >= 0;
    if (lane == 0) {
        dst[16 * W] = (uint32_t)tm;
        dst[PW] = (uint32_t)(moves ? moves[g] : -1);
        dst[PW + 1] = values ? __float_as_uint(values[g]) : 0u;
        dst[PW + 2] = tv ? 1u : 0u;
    }
    const NodeBlock *nb = pool.blk + (tv ? m.root : 0);
    for (int it = 0; it < SGO_AWORDS; it++) {
        int slot = it * 32 + lane;
        if (slot >= A) continue;
        bool ex = tv && ((nb->exist[it] >> lane) & 1u);
        float p = ex ? (m.root_f64 ? (float)root_p64[(size_t)tree * SGO_APAD + slot] : nb->prior[slot]) : 0.f;
        dst[PW + 3 + slot] = __float_as_uint(p);
    }
}

extern "C" int sgo_record_words(sgo_engine *e)
{
    int W = (e->S * e->S + 31) / 32;
    return 16 * W + 1 + 3 + e->A;
}

extern "C" int sgo_records_pack(sgo_engine *e, const int32_t *d_tree_sel, const int32_t *d_moves, const float *d_values,
                                uint32_t *d_out, void *stream)
{
    k_records_pack<<<(e->G + 3) / 4, 128, 0, S_(stream)>>>(e->boards, e->S, e->G, e->T, sgo_pool(e), e->meta, e->root_p64,
                                                          d_tree_sel, d_moves, d_values, d_out, sgo_record_words(e));
    SGO_LAUNCHED(e);
    return 0;
}

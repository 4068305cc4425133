// tree.cu — flat-tree PUCT MCTS kernels for both of the reference's search drivers.
//
//   mode A  self_play.py:28-152   (batched top-N leaves, no virtual loss)
//   mode B  tree_util.py:4-24 + nomodel_self_play.py:40-140 (busy-flag waves)
//   play.py:308-352 selectors, play.py:376-421 expansion.
//
// One warp per game.  A node's children live in one NodeBlock (SoA); lane l owns
// action slots l, l+32, ... so every statistic is read with coalesced 128 B loads
// and the argmax over <=362 children is a register scan + 5 shuffle steps.
//
// Memory: every tree draws its blocks from ONE pool (engine.h Pool): a block is addressed by its pool id, a
// tree is its root id plus the child / parent links.  Re-rooting (self_play.py:223-238) keeps the chosen
// child's subtree where it is and returns every other block to the pool, so a game whose search concentrates
// on one line borrows the room that flatter trees do not need; nothing is copied and nothing is sized per tree.
//
// Bit-exactness (SURVEY "hard parts", Q18/Q19): PUCT is evaluated with explicitly
// rounded float32 ops (__fmul_rn/__fdiv_rn/__fadd_rn, never contracted to FMA),
// or float64 at a Dirichlet-noised root; ties go to the lowest action index;
// value sums are applied in the reference's order by a single lane.
#include "engine.h"

#define TREE_WARPS 4
static inline dim3 warp_grid(int n) { return dim3((n + TREE_WARPS - 1) / TREE_WARPS); }
static inline cudaStream_t S_(void *s) { return (cudaStream_t)s; }

#define NEG_INF (-__longlong_as_double(0x7ff0000000000000LL))

// ------------------------------------------------------------ device helpers
// PUCT scores of all children of `nb` (play.py:318-319,331-332,344-345).
// sc[it] = score of slot it*32+lane; bit `it` of the returned mask = selectable.
__device__ __forceinline__ uint32_t node_scores(const NodeBlock *nb, const double *p64, bool skip_busy, int lane, double sc[SGO_AWORDS])
{
    uint32_t ok = 0;
    int cnt[SGO_AWORDS];
    int local = 0;
#pragma unroll
    for (int it = 0; it < SGO_AWORDS; it++) {
        uint32_t ex = nb->exist[it];
        bool e = (ex >> lane) & 1u;
        cnt[it] = e ? nb->n[it * 32 + lane] : 0;
        local += cnt[it];
        if (e) {
            bool busy = skip_busy && ((nb->busy[it] >> lane) & 1u);
            if (!busy) ok |= 1u << it;
        }
    }
    int sum_n = warp_sum(local);                       // sum over ALL children, busy ones included
    double tn = sqrt((double)sum_n);
    if (tn == 0.0) tn = 1.0;
    float tnf = (float)tn;
#pragma unroll
    for (int it = 0; it < SGO_AWORDS; it++) {
        sc[it] = NEG_INF;
        if ((ok >> it) & 1u) {
            int slot = it * 32 + lane;
            int n = cnt[it];
            float mean = n > 0 ? __fdiv_rn(nb->w[slot], (float)n) : 0.0f;
            if (p64) {
                double u = __ddiv_rn(__dmul_rn(p64[slot], tn), __dadd_rn(1.0, (double)n));
                sc[it] = __dadd_rn((double)mean, u);
            } else {
                float u = __fdiv_rn(__fmul_rn(nb->prior[slot], tnf), (float)(1 + n));
                sc[it] = (double)__fadd_rn(mean, u);
            }
        }
    }
    return ok;
}

// argmax with strict '>' from `sentinel`, ties -> lowest slot; returns slot or -1
__device__ __forceinline__ int warp_argmax(const double sc[SGO_AWORDS], uint32_t ok, double sentinel, int lane)
{
    double best = sentinel;
    int bslot = 0x7fffffff;
#pragma unroll
    for (int it = 0; it < SGO_AWORDS; it++)
        if (((ok >> it) & 1u) && sc[it] > best) { best = sc[it]; bslot = it * 32 + lane; }
    bool has = bslot != 0x7fffffff;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        double ob = __shfl_xor_sync(SGO_FULL, best, o);
        int os = __shfl_xor_sync(SGO_FULL, bslot, o);
        bool oh = os != 0x7fffffff;
        if (oh && (!has || ob > best || (ob == best && os < bslot))) { best = ob; bslot = os; has = true; }
    }
    return has ? bslot : -1;
}

__device__ __forceinline__ void set_busy(NodeBlock *nb, int slot, bool on, int lane)
{
    if (lane == 0) {
        if (on) nb->busy[slot >> 5] |= 1u << (slot & 31);
        else nb->busy[slot >> 5] &= ~(1u << (slot & 31));
    }
    __syncwarp();
}

// ------------------------------------------------------------------ kernels
__global__ void k_tree_reset(TreeMeta *meta, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TreeMeta m = meta[i];
    m.root = -1; m.n_blocks = 0; m.valid = 0; m.root_f64 = 0; m.root_count = 0; m.root_value = 0.f; m.overflow = 0;
    meta[i] = m;
}

// every block back on the free stack (ids handed out in ascending order)
__global__ void k_pool_reset(Pool pool, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pool.free_list[i] = (int32_t)(n - 1 - i);
    if (i == 0) { pool.ctl[0] = (int32_t)n; pool.ctl[1] = (int32_t)n; pool.ctl[2] = 0; }
}

// Returns block `start` and all its descendants to the pool, except the subtree hanging off block `keep`
// (-1: none).  The blocks still to visit form a list threaded through their (now meaningless) parent_block
// fields, so no stack or queue is needed however large the discarded part is.  Whole warp; returns #freed.
__device__ int free_subtree(const Pool &pool, int start, int keep, int lane)
{
    NodeBlock *pb = pool.blk;
    if (lane == 0) pb[start].parent_block = -1;                      // list terminator
    __syncwarp();
    int head = start, freed = 0;
    int mine = -1;                                                   // lane (freed % 32) holds the id freed at that count: 32 ids per push
    while (head >= 0) {
        const int b = head;
        int next = pb[b].parent_block;
        int chs[SGO_AWORDS];
#pragma unroll
        for (int it = 0; it < SGO_AWORDS; it++) chs[it] = pb[b].child[it * 32 + lane];      // 12 independent loads in flight
#pragma unroll
        for (int it = 0; it < SGO_AWORDS; it++) {
            int ch = chs[it];
            if (ch == keep) ch = -1;
            const unsigned bal = __ballot_sync(SGO_FULL, ch >= 0);
            if (!bal) continue;
            const unsigned higher = lane == 31 ? 0u : (bal & ~((2u << lane) - 1u));
            const int succ = __shfl_sync(SGO_FULL, ch, higher ? __ffs(higher) - 1 : lane);
            if (ch >= 0) pb[ch].parent_block = higher ? succ : next;  // chain this block's children in front of the list
            next = __shfl_sync(SGO_FULL, ch, __ffs(bal) - 1);
        }
        if (lane == (freed & 31)) mine = b;
        freed++;
        if ((freed & 31) == 0) {                                      // one atomic per 32 blocks instead of one per block
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&pool.ctl[0], 32);
            pos = __shfl_sync(SGO_FULL, pos, 0);
            pool.free_list[pos + lane] = mine;
        }
        __syncwarp();
        head = next;
    }
    const int rest = freed & 31;
    if (rest) {
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&pool.ctl[0], rest);
        pos = __shfl_sync(SGO_FULL, pos, 0);
        if (lane < rest) pool.free_list[pos + lane] = mine;
    }
    __syncwarp();
    return freed;
}

// drop the trees of the games flagged in `mask` (NULL = tree_sel-selected trees that are valid): selfplay_worker.py:81-124
// starts the next game in the same worker; here the slot's blocks go back to the pool first
__global__ void k_tree_free(int G, int T, Pool pool, TreeMeta *meta, const int32_t *game_mask, const int32_t *tree_sel)
{
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (g >= G) return;
    if (game_mask && !game_mask[g]) return;
    for (int t = 0; t < T; t++) {
        if (tree_sel && tree_sel[g] != t) continue;
        TreeMeta m = meta[g * T + t];
        if (m.valid && m.root >= 0) free_subtree(pool, m.root, -1, lane);
        __syncwarp();
        if (lane == 0) {
            m.root = -1; m.n_blocks = 0; m.valid = 0; m.root_f64 = 0; m.root_count = 0; m.root_value = 0.f; m.overflow = 0;
            meta[g * T + t] = m;
        }
    }
}

// play.py:376-421 new_tree / new_subtree at the root
__global__ void k_tree_new(const Board *boards, int S, int G, int T, Pool pool, TreeMeta *meta, double *root_p64,
                           const int32_t *tree_sel, const float *policy, const double *noise, double keep, double eps, int32_t *err)
{
    __shared__ uint32_t scratch[TREE_WARPS][SGO_AWORDS];
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (g >= G) return;
    int tsel = tree_sel ? tree_sel[g] : 0;
    int tree = g * T + (tsel < 0 ? 0 : tsel);
    if (tsel < 0) return;
    TreeMeta m = meta[tree];
    if (m.valid) return;                                   // (a forced rebuild frees the old tree first: sgo_tree_new)
    int root = -1;
    if (lane == 0) {
        int base = pool_pop(pool, 1);
        if (base >= 0) root = pool.free_list[base - 1];
        else { atomicOr(err, SGO_ERR_ARENA); meta[tree].overflow = 1; }
    }
    root = __shfl_sync(SGO_FULL, root, 0);
    if (root < 0) return;
    uint32_t *sc = scratch[threadIdx.x >> 5];
    uint32_t ill = board_illegal(boards + g, S, lane);
    illegal_rows_to_words(ill, S, lane, sc);
    NodeBlock *nb = pool.blk + root;
    int A = S * S + 1;
    double *p64 = root_p64 + (size_t)tree * SGO_APAD;
    for (int it = 0; it < SGO_AWORDS; it++) {
        int slot = it * 32 + lane;
        bool legal = slot < A && !((sc[it] >> lane) & 1u);
        float p = legal ? policy[(size_t)g * A + slot] : 0.f;
        nb->prior[slot] = p;
        nb->n[slot] = 0;
        nb->w[slot] = 0.f;
        nb->child[slot] = -1;
        if (noise) {
            // numpy.ma promotes to float64: 0.75*float64(p) + 0.25*noise (play.py:403, [probe])
            double v = legal ? __dadd_rn(__dmul_rn(keep, (double)p), __dmul_rn(eps, noise[(size_t)g * A + slot])) : 0.0;
            p64[slot] = v;
        }
        uint32_t ex = __ballot_sync(SGO_FULL, legal);
        if (lane == 0) { nb->exist[it] = ex; nb->busy[it] = 0; }
    }
    if (lane == 0) {
        nb->parent_block = -1; nb->parent_slot = -1;
        m.root = root; m.n_blocks = 1; m.valid = 1; m.root_f64 = noise ? 1 : 0; m.root_count = 0; m.root_value = 0.f; m.overflow = 0;
        meta[tree] = m;
    }
}

// self_play.py:28-66 — mode A selection
__global__ void k_select_a(const Board *boards, int S, int G, int T, int L, Pool pool, TreeMeta *meta,
                           const double *root_p64, const int32_t *tree_sel, int batch, Board *leaf_boards,
                           LeafRef *leaf_refs, int32_t *leaf_count, int32_t *err)
{
    __shared__ Board sboard[TREE_WARPS];
    __shared__ int16_t ssel[TREE_WARPS][128];
    int wib = threadIdx.x >> 5;
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (g >= G) return;
    int tsel = tree_sel ? tree_sel[g] : 0;
    int tree = g * T + (tsel < 0 ? 0 : tsel);
    TreeMeta m = meta[tree];
    if (lane == 0) leaf_count[g] = 0;
    for (int i = lane; i < L; i += 32) leaf_refs[(size_t)g * L + i].state = 0;
    if (!m.valid || tsel < 0) return;
    NodeBlock *ar = pool.blk;
    Board *sb = &sboard[wib];
    board_copy(sb, boards + g, lane);
    double sc[SGO_AWORDS];
    uint32_t ok;
    int blk = m.root;
    for (int depth = 0;; depth++) {                       // self_play.py:117-120
        const double *p64 = (blk == m.root && m.root_f64) ? root_p64 + (size_t)tree * SGO_APAD : nullptr;
        ok = node_scores(ar + blk, p64, false, lane, sc);
        int best = warp_argmax(sc, ok, NEG_INF, lane);
        if (best < 0) { if (lane == 0) atomicOr(err, SGO_ERR_NOACTION); return; }
        int c = ar[blk].child[best];
        if (c < 0) break;
        board_play(sb, S, best, 0, lane);
        blk = c;
        if (depth > SGO_MAXDEPTH) { if (lane == 0) atomicOr(err, SGO_ERR_DEPTH); return; }
    }
    // top-`batch` of the frontier node in (score desc, index asc) order (play.py:337-352)
    int n = 0;
    for (; n < batch && n < 128; n++) {
        int s = warp_argmax(sc, ok, NEG_INF, lane);
        if (s < 0) break;
        if ((s & 31) == lane) ok &= ~(1u << (s >> 5));
        if (lane == 0) ssel[wib][n] = (int16_t)s;
    }
    __syncwarp();
    int base = 0;
    if (lane == 0) base = pool_pop(pool, n);              // one block per leaf, taken before anything is written
    base = __shfl_sync(SGO_FULL, base, 0);
    if (base < 0) {                                       // pool exhausted: this game selects nothing this step; the step reports it
        if (lane == 0) { atomicOr(err, SGO_ERR_ARENA); meta[tree].overflow = 1; }
        return;
    }
    for (int i = 0; i < n; i++) {                         // self_play.py:41-66
        int s = ssel[wib][i];
        size_t li = (size_t)g * L + i;
        Board *lb = leaf_boards + li;
        board_copy(lb, sb, lane);
        board_play(lb, S, s, 0, lane);
        int cb = blk, cs = s, c = ar[cb].child[cs];
        while (c >= 0) {                                  // greedy top_one_action descent
            double sc2[SGO_AWORDS];
            uint32_t ok2 = node_scores(ar + c, nullptr, false, lane, sc2);
            int s2 = warp_argmax(sc2, ok2, -1.0, lane);   // sentinel -1 (play.py:329)
            if (s2 < 0) { if (lane == 0) atomicOr(err, SGO_ERR_NOACTION); break; }
            board_play(lb, S, s2, 0, lane);
            cb = c; cs = s2; c = ar[cb].child[cs];
        }
        if (lane == 0) {
            LeafRef r;
            r.block = cb; r.slot = cs; r.to_move = lb->to_move; r.state = 1; r.new_block = pool.free_list[base - 1 - i]; r.sv = 0.f;
            r.pad[0] = r.pad[1] = 0;
            leaf_refs[li] = r;
        }
    }
    if (lane == 0) { leaf_count[g] = n; meta[tree].n_blocks = m.n_blocks + n; }
}

// play.py:391-421 new_subtree for one selected leaf (one warp per leaf slot)
__global__ void k_expand(const Board *boards, int S, int G, int T, int L, Pool pool, const TreeMeta *meta,
                         const int32_t *tree_sel, const Board *leaf_boards, LeafRef *leaf_refs,
                         const float *policy, const float *value)
{
    __shared__ uint32_t scratch[TREE_WARPS][SGO_AWORDS];
    size_t li = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = lane_id();
    if (li >= (size_t)G * L) return;
    LeafRef r = leaf_refs[li];
    if (r.state != 1) return;
    int g = (int)(li / L);
    int tsel = tree_sel ? tree_sel[g] : 0;
    if (tsel < 0) return;
    NodeBlock *ar = pool.blk;
    uint32_t *sc = scratch[threadIdx.x >> 5];
    uint32_t ill = board_illegal(leaf_boards + li, S, lane);
    illegal_rows_to_words(ill, S, lane, sc);
    NodeBlock *nb = ar + r.new_block;
    int A = S * S + 1;
    for (int it = 0; it < SGO_AWORDS; it++) {
        int slot = it * 32 + lane;
        bool legal = slot < A && !((sc[it] >> lane) & 1u);
        nb->prior[slot] = legal ? policy[li * A + slot] : 0.f;
        nb->n[slot] = 0;
        nb->w[slot] = 0.f;
        nb->child[slot] = -1;
        uint32_t ex = __ballot_sync(SGO_FULL, legal);
        if (lane == 0) { nb->exist[it] = ex; nb->busy[it] = 0; }
    }
    if (lane == 0) {
        nb->parent_block = r.block; nb->parent_slot = r.slot;
        ar[r.block].child[r.slot] = r.new_block;
        float v = value[li];
        // self_play.py:100-102 / simulation_workers.py:50
        leaf_refs[li].sv = (r.to_move == boards[g].to_move) ? v : -v;
        leaf_refs[li].state = 2;
    }
}

// self_play.py:108-116 — mode A backup, leaves in rank order, one lane walks
__global__ void k_backup_a(int G, int T, int L, Pool pool, TreeMeta *meta, const int32_t *tree_sel,
                           LeafRef *leaf_refs, const int32_t *leaf_count)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int tsel = tree_sel ? tree_sel[g] : 0;
    int tree = g * T + (tsel < 0 ? 0 : tsel);
    TreeMeta m = meta[tree];
    if (!m.valid || tsel < 0) return;
    NodeBlock *ar = pool.blk;
    int n = leaf_count[g];
    for (int i = 0; i < n; i++) {
        LeafRef r = leaf_refs[(size_t)g * L + i];
        if (r.state != 2) continue;
        int b = r.block, s = r.slot;
        while (b >= 0) {
            ar[b].n[s] += 1;
            ar[b].w[s] = __fadd_rn(ar[b].w[s], r.sv);
            s = ar[b].parent_slot;
            b = ar[b].parent_block;
        }
        m.root_count += 1;
        m.root_value = __fadd_rn(m.root_value, r.sv);
        leaf_refs[(size_t)g * L + i].state = 0;
    }
    meta[tree].root_count = m.root_count;
    meta[tree].root_value = m.root_value;
}

// nomodel_self_play.py:40-56 back_propagation of pending entry `li` (single lane)
__device__ __forceinline__ void backprop_b(NodeBlock *ar, TreeMeta *m, LeafRef *ref)
{
    int b = ref->block, s = ref->slot;
    ar[b].n[s] += 1;                                          // the worker's copy: count/value/mean
    float lw = __fadd_rn(ar[b].w[s], ref->sv);
    ar[b].w[s] = lw;
    ar[b].busy[s >> 5] &= ~(1u << (s & 31));
    for (;;) {                                                // closest_parent upwards
        int pb = ar[b].parent_block, ps = ar[b].parent_slot;
        if (pb < 0) {
            m->root_count += 1;
            m->root_value = __fadd_rn(m->root_value, lw);
            break;
        }
        ar[pb].n[ps] += 1;
        ar[pb].w[ps] = __fadd_rn(ar[pb].w[ps], lw);
        ar[pb].busy[ps >> 5] &= ~(1u << (ps & 31));
        b = pb;
    }
    ref->state = 0;
}

// wave state words
#define WV_ENERGY 0
#define WV_PREBP 1
#define WV_HEAD 2
#define WV_TAIL 3
#define WV_STALL 4

// tree_util.py:4-24 + nomodel_self_play.py:59-75 — mode B selection (resumable)
__global__ void k_select_b(const Board *boards, int S, int G, int T, int L, Pool pool, TreeMeta *meta,
                           const double *root_p64, const int32_t *tree_sel, int energy, int restart, Board *leaf_boards,
                           LeafRef *leaf_refs, int32_t *leaf_count, int32_t *wave, int32_t *counters, int32_t *err)
{
    __shared__ int16_t spath[TREE_WARPS][SGO_MAXDEPTH];
    int wib = threadIdx.x >> 5;
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (g >= G) return;
    int tsel = tree_sel ? tree_sel[g] : 0;
    int tree = g * T + (tsel < 0 ? 0 : tsel);
    TreeMeta m = meta[tree];
    int32_t *wv = wave + (size_t)g * 8;
    int energy_left, pre_bp, head, tail;
    if (restart) {
        energy_left = energy; pre_bp = 0; head = 0; tail = 0;
        for (int i = lane; i < L; i += 32) leaf_refs[(size_t)g * L + i].state = 0;
        __syncwarp();
    } else {
        energy_left = wv[WV_ENERGY]; pre_bp = wv[WV_PREBP]; head = wv[WV_HEAD]; tail = wv[WV_TAIL];
    }
    int stalled = 0, newly = 0;
    if (!m.valid || tsel < 0) energy_left = 0;                  // node['subtree'] == {} -> return (:60-61)
    NodeBlock *ar = pool.blk;
    const int rootb = m.root;
    while (energy_left > 0) {
        // ---- find_best_leaf_virtual_loss
        int blk = rootb, lslot = -1;
        for (int guard = 0; guard < 8 * SGO_MAXDEPTH; guard++) {
            double sc[SGO_AWORDS];
            const double *p64 = (blk == rootb && m.root_f64) ? root_p64 + (size_t)tree * SGO_APAD : nullptr;
            uint32_t ok = node_scores(ar + blk, p64, true, lane, sc);
            int s = warp_argmax(sc, ok, -100.0, lane);          // sentinel -100 (play.py:313)
            if (s < 0) {
                if (blk == rootb) break;                        // (None, None)
                int pb = ar[blk].parent_block, ps = ar[blk].parent_slot;
                set_busy(ar + pb, ps, true, lane);              // node['virtual_loss'] = 2; node = parent
                blk = pb;
                continue;
            }
            int c = ar[blk].child[s];
            if (c < 0) { lslot = s; break; }
            blk = c;
        }
        if (lslot >= 0) {
            set_busy(ar + blk, lslot, true, lane);
            if (ar[blk].n[lslot] > 0) { energy_left--; pre_bp++; continue; }     // :66-69
            int nblk = -1;
            if (lane == 0 && tail < L) {
                int base = pool_pop(pool, 1);
                if (base >= 0) nblk = pool.free_list[base - 1];
            }
            nblk = __shfl_sync(SGO_FULL, nblk, 0);
            if (nblk < 0) {                                     // no leaf slot / pool exhausted: the wave ends here and the step reports it
                if (lane == 0) { atomicOr(err, SGO_ERR_ARENA); meta[tree].overflow = 1; }
                set_busy(ar + blk, lslot, false, lane);
                energy_left = 0;
                break;
            }
            // basic_tasks2 (simulation_workers.py:43-45): replay the move list from the root board
            int depth = 0;
            if (lane == 0) {
                spath[wib][0] = (int16_t)lslot;
                int b = blk, d = 1;
                while (ar[b].parent_block >= 0 && d < SGO_MAXDEPTH) {
                    spath[wib][d++] = (int16_t)ar[b].parent_slot;
                    b = ar[b].parent_block;
                }
                if (ar[b].parent_block >= 0) atomicOr(err, SGO_ERR_DEPTH);
                depth = d;
            }
            depth = __shfl_sync(SGO_FULL, depth, 0);
            __syncwarp();
            size_t li = (size_t)g * L + tail;
            Board *lb = leaf_boards + li;
            board_copy(lb, boards + g, lane);
            for (int d = depth - 1; d >= 0; d--) board_play(lb, S, spath[wib][d], 0, lane);
            if (lane == 0) {
                LeafRef r;
                r.block = blk; r.slot = lslot; r.to_move = lb->to_move; r.state = 1; r.new_block = nblk; r.sv = 0.f;
                r.pad[0] = r.pad[1] = 0;
                leaf_refs[li] = r;
            }
            m.n_blocks += 1;
            tail++; energy_left--; newly++;
        } else {
            // "No best leaf": consume one finished result, FIFO (:70-75)
            int st = head < tail ? leaf_refs[(size_t)g * L + head].state : -1;
            if (st == 2) {
                if (lane == 0) backprop_b(ar, &m, &leaf_refs[(size_t)g * L + head]);
                m.root_count = __shfl_sync(SGO_FULL, m.root_count, 0);
                m.root_value = __shfl_sync(SGO_FULL, m.root_value, 0);
                __syncwarp();
                head++; pre_bp++;
                continue;
            }
            stalled = 1;                                        // pending results not evaluated yet
            break;
        }
    }
    if (lane == 0) {
        wv[WV_ENERGY] = energy_left; wv[WV_PREBP] = pre_bp; wv[WV_HEAD] = head; wv[WV_TAIL] = tail; wv[WV_STALL] = stalled;
        leaf_count[g] = tail;
        TreeMeta *mm = meta + tree;
        mm->n_blocks = m.n_blocks; mm->root_count = m.root_count; mm->root_value = m.root_value;
        if (newly) atomicAdd(&counters[0], newly);
        if (stalled) atomicAdd(&counters[1], 1);
    }
}

// nomodel_self_play.py:80-82 — the remaining back-props of a wave, FIFO
__global__ void k_backup_b(int G, int T, int L, Pool pool, TreeMeta *meta, const int32_t *tree_sel,
                           LeafRef *leaf_refs, int32_t *wave, int total_energy)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int tsel = tree_sel ? tree_sel[g] : 0;
    int tree = g * T + (tsel < 0 ? 0 : tsel);
    TreeMeta m = meta[tree];
    if (!m.valid || tsel < 0) return;
    NodeBlock *ar = pool.blk;
    int32_t *wv = wave + (size_t)g * 8;
    int head = wv[WV_HEAD], tail = wv[WV_TAIL], todo = total_energy - wv[WV_PREBP];
    for (int i = 0; i < todo && head < tail; i++, head++) {
        LeafRef *r = &leaf_refs[(size_t)g * L + head];
        if (r->state == 2) backprop_b(ar, &m, r);
    }
    wv[WV_HEAD] = head;
    meta[tree].root_count = m.root_count;
    meta[tree].root_value = m.root_value;
}

// self_play.py:138-152 — move pick
__global__ void k_pick(int S, int G, int T, Pool pool, const TreeMeta *meta, const int32_t *tree_sel,
                       const int32_t *temperature, const double *u01, const int32_t *forced, int32_t *move_out)
{
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (g >= G) return;
    int tsel = tree_sel ? tree_sel[g] : 0;
    int tree = g * T + (tsel < 0 ? 0 : tsel);
    TreeMeta m = meta[tree];
    if (!m.valid || tsel < 0) { if (lane == 0) move_out[g] = -1; return; }
    if (forced && forced[g] >= 0) { if (lane == 0) move_out[g] = forced[g]; return; }
    const NodeBlock *nb = pool.blk + m.root;
    int temp = temperature ? temperature[g] : 0;
    if (temp == 0) {
        // max over (count, mean_value, action) tuples -> ties to the HIGHER index (Q18)
        int bn = -1, bs = -1; float bm = 0.f;
        for (int it = 0; it < SGO_AWORDS; it++) {
            int slot = it * 32 + lane;
            if (!((nb->exist[it] >> lane) & 1u)) continue;
            int n = nb->n[slot];
            float mean = n > 0 ? __fdiv_rn(nb->w[slot], (float)n) : 0.f;
            if (n > bn || (n == bn && (mean > bm || (mean == bm && slot > bs)))) { bn = n; bm = mean; bs = slot; }
        }
        for (int o = 16; o; o >>= 1) {
            int on = __shfl_xor_sync(SGO_FULL, bn, o), os = __shfl_xor_sync(SGO_FULL, bs, o);
            float om = __shfl_xor_sync(SGO_FULL, bm, o);
            if (on > bn || (on == bn && (om > bm || (om == bm && os > bs)))) { bn = on; bm = om; bs = os; }
        }
        if (lane == 0) move_out[g] = bs;
    } else {
        // np.random.choice(moves, size=1, p=N/total) over the visited children in ascending index (self_play.py:140-149),
        // reproduced from ONE uniform draw exactly as numpy computes it: p_i = N_i / float(total) in fp64, cdf = sequential
        // cumulative sum, cdf /= cdf[-1], first index with cdf > u (searchsorted side='right').  The sums are order
        // dependent in fp64, so one lane walks the children (once per ply per game; tests/test_cpu_choice.py pins the
        // algorithm against numpy).
        int local = 0;
        for (int it = 0; it < SGO_AWORDS; it++)
            if ((nb->exist[it] >> lane) & 1u) local += nb->n[it * 32 + lane];
        const int total = warp_sum(local);
        if (lane == 0) {
            const int A = S * S + 1;
            const double u = u01 ? u01[g] : 0.5, tot = (double)total;
            double fin = 0.0;
            for (int slot = 0; slot < A; slot++) {
                if (!((nb->exist[slot >> 5] >> (slot & 31)) & 1u)) continue;
                const int n = nb->n[slot];
                if (n > 0) fin = __dadd_rn(fin, __ddiv_rn((double)n, tot));
            }
            double run = 0.0;
            int pick = -1, last = -1;
            for (int slot = 0; slot < A && pick < 0; slot++) {
                if (!((nb->exist[slot >> 5] >> (slot & 31)) & 1u)) continue;
                const int n = nb->n[slot];
                if (n <= 0) continue;
                run = __dadd_rn(run, __ddiv_rn((double)n, tot));
                last = slot;
                if (__ddiv_rn(run, fin) > u) pick = slot;
            }
            move_out[g] = pick >= 0 ? pick : last;
        }
    }
}

// self_play.py:223-238 — cut each tree of the game to the child `move`: the child's block becomes the root
// where it lies ("mcts_tree['parent'] = None") and every block outside its subtree goes back to the pool.
__global__ void k_reroot(int G, int T, Pool pool, TreeMeta *meta, const int32_t *moves)
{
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (g >= G) return;
    int mv = moves[g];
    if (mv < 0) return;
    for (int t = 0; t < T; t++) {
        int tree = g * T + t;
        TreeMeta m = meta[tree];
        if (!m.valid) continue;
        NodeBlock *root = pool.blk + m.root;
        if (!((root->exist[mv >> 5] >> (mv & 31)) & 1u)) continue;        // `index in subtree` false
        int c = root->child[mv];
        int rc = root->n[mv];
        float rv = root->w[mv];
        __syncwarp();
        int freed = free_subtree(pool, m.root, c, lane);                  // c < 0: the whole tree
        if (lane == 0) {
            if (c < 0) {                                                  // child has subtree {} -> new tree later
                m.root = -1; m.valid = 0; m.n_blocks = 0;
            } else {
                pool.blk[c].parent_block = -1; pool.blk[c].parent_slot = -1;
                m.root = c; m.n_blocks -= freed;
            }
            m.root_f64 = 0; m.root_count = rc; m.root_value = rv;
            meta[tree] = m;
        }
        __syncwarp();
    }
}

// Cheney copy of a tree into a compact buffer (block 0 = root, children in breadth-first slot order, links rewritten
// to buffer indices): the canonical form sgo_tree_download_sync hands to the host.  One warp.
__global__ void k_tree_export(Pool pool, const TreeMeta *meta, int tree, NodeBlock *dst, int max_blocks, int32_t *n_out)
{
    int lane = lane_id();
    TreeMeta m = meta[tree];
    if (!m.valid || max_blocks < 1) { if (lane == 0) *n_out = 0; return; }
    const NodeBlock *src = pool.blk;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + m.root);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (int i = lane; i < (int)(sizeof(NodeBlock) / 16); i += 32) d4[i] = s4[i];
    __syncwarp();
    if (lane == 0) { dst[0].parent_block = -1; dst[0].parent_slot = -1; }
    int f = 1;
    bool full = false;
    for (int s = 0; s < f && !full; s++) {
        for (int it = 0; it < SGO_AWORDS && !full; it++) {
            int slot = it * 32 + lane;
            int ch = dst[s].child[slot];
            unsigned bal = __ballot_sync(SGO_FULL, ch >= 0);
            if (!bal) continue;
            if (f + __popc(bal) > max_blocks) { full = true; break; }
            int mine = f + __popc(bal & ((1u << lane) - 1u));
            if (ch >= 0) dst[s].child[slot] = mine;
            while (bal) {
                int l = __ffs(bal) - 1;
                bal &= bal - 1;
                int old = __shfl_sync(SGO_FULL, ch, l);
                const uint4 *a = reinterpret_cast<const uint4 *>(src + old);
                uint4 *b = reinterpret_cast<uint4 *>(dst + f);
                for (int i = lane; i < (int)(sizeof(NodeBlock) / 16); i += 32) b[i] = a[i];
                __syncwarp();
                if (lane == 0) { dst[f].parent_block = s; dst[f].parent_slot = it * 32 + l; }
                f++;
            }
            __syncwarp();
        }
    }
    if (lane == 0) *n_out = full ? -f : f;
}

// the inverse: n blocks in the compact form (block 0 = root, links = buffer indices) into freshly popped pool blocks
__global__ void k_tree_import(Pool pool, TreeMeta *meta, int tree, const NodeBlock *src, int n, TreeMeta hm, int32_t *err)
{
    __shared__ int s_base;
    int lane = lane_id();
    if (lane == 0) s_base = n > 0 ? pool_pop(pool, n) : 0;
    __syncwarp();
    int base = s_base;
    if (base < 0) { if (lane == 0) { atomicOr(err, SGO_ERR_ARENA); meta[tree].overflow = 1; } return; }
    for (int b = 0; b < n; b++) {
        NodeBlock *d = pool.blk + pool.free_list[base - 1 - b];
        const uint4 *a = reinterpret_cast<const uint4 *>(src + b);
        uint4 *o = reinterpret_cast<uint4 *>(d);
        for (int i = lane; i < (int)(sizeof(NodeBlock) / 16); i += 32) o[i] = a[i];
        __syncwarp();
        for (int it = 0; it < SGO_AWORDS; it++) {
            int c = d->child[it * 32 + lane];
            if (c >= 0) d->child[it * 32 + lane] = pool.free_list[base - 1 - c];
        }
        if (lane == 0 && d->parent_block >= 0) d->parent_block = pool.free_list[base - 1 - d->parent_block];
        __syncwarp();
    }
    if (lane == 0) {
        hm.root = n > 0 ? pool.free_list[base - 1] : -1;
        hm.n_blocks = n;
        if (n == 0) hm.valid = 0;
        meta[tree] = hm;
    }
}

__global__ void k_child_stats(int S, int G, int T, Pool pool, const TreeMeta *meta, const double *root_p64,
                              const int32_t *tree_sel, double *prior, int32_t *count, float *value)
{
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (g >= G) return;
    int tsel = tree_sel ? tree_sel[g] : 0;
    int tree = g * T + (tsel < 0 ? 0 : tsel);
    TreeMeta m = meta[tree];
    const bool tv = m.valid && tsel >= 0;
    const NodeBlock *nb = pool.blk + (tv ? m.root : 0);
    int A = S * S + 1;
    for (int it = 0; it < SGO_AWORDS; it++) {
        int slot = it * 32 + lane;
        if (slot >= A) continue;
        bool e = tv && ((nb->exist[it] >> lane) & 1u);
        size_t o = (size_t)g * A + slot;
        if (prior) prior[o] = e ? (m.root_f64 ? root_p64[(size_t)tree * SGO_APAD + slot] : (double)nb->prior[slot]) : 0.0;
        if (count) count[o] = e ? nb->n[slot] : 0;
        if (value) value[o] = e ? nb->w[slot] : 0.f;
    }
}

__global__ void k_tree_valid(int G, int T, const TreeMeta *meta, const int32_t *tree_sel, int32_t *valid)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int ts = tree_sel ? tree_sel[g] : 0;
    valid[g] = ts < 0 ? 0 : meta[g * T + ts].valid;
}


// indices of leaf slots awaiting evaluation (state == 1), warp-aggregated append
__global__ void k_leaf_compact(const LeafRef *leaf_refs, int total, int32_t *index, int32_t *count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x, lane = lane_id();
    bool v = i < total && leaf_refs[i].state == 1;
    unsigned bal = __ballot_sync(SGO_FULL, v);
    if (!bal) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(bal));
    base = __shfl_sync(SGO_FULL, base, 0);
    if (v) index[base + __popc(bal & ((1u << lane) - 1u))] = i;
}

// ------------------------------------------------------------------ C ABI
#define LAUNCH_OK(e) SGO_LAUNCHED(e)

extern "C" int sgo_tree_reset(sgo_engine *e, void *stream)
{
    int n = e->G * e->T;
    k_tree_reset<<<(n + 127) / 128, 128, 0, S_(stream)>>>(e->meta, n);
    LAUNCH_OK(e);
    k_pool_reset<<<(unsigned)((e->pool_blocks + 255) / 256), 256, 0, S_(stream)>>>(sgo_pool(e), e->pool_blocks);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_free(sgo_engine *e, const int32_t *d_game_mask, void *stream)
{
    k_tree_free<<<warp_grid(e->G), TREE_WARPS * 32, 0, S_(stream)>>>(e->G, e->T, sgo_pool(e), e->meta, d_game_mask, nullptr);
    LAUNCH_OK(e);
    return 0;
}

__global__ void k_tree_sizes(int n, const TreeMeta *meta, int32_t *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = meta[i].valid ? meta[i].n_blocks : 0;
}

extern "C" int sgo_tree_sizes(sgo_engine *e, int32_t *d_out, void *stream)
{
    int n = e->G * e->T;
    k_tree_sizes<<<(n + 127) / 128, 128, 0, S_(stream)>>>(n, e->meta, d_out);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_pool_stats_sync(sgo_engine *e, int64_t *h_out, void *stream)
{
    SGO_CUDA_OK(e, cudaMemcpyAsync(e->h_pinned + 8, e->pool_ctl, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, S_(stream)));
    SGO_CUDA_OK(e, cudaStreamSynchronize(S_(stream)));
    h_out[0] = e->pool_blocks; h_out[1] = e->h_pinned[8]; h_out[2] = e->h_pinned[9]; h_out[3] = e->h_pinned[10];
    return 0;
}

extern "C" int sgo_tree_new(sgo_engine *e, const int32_t *d_tree_sel, const float *d_policy, const double *d_noise,
                            double eps, int32_t force, void *stream)
{
    if (force) {              // the old trees' blocks go back to the pool in a launch of their own (no pop may race a push)
        k_tree_free<<<warp_grid(e->G), TREE_WARPS * 32, 0, S_(stream)>>>(e->G, e->T, sgo_pool(e), e->meta, nullptr,
                                                                        d_tree_sel ? d_tree_sel : e->zero_sel);
        LAUNCH_OK(e);
    }
    k_tree_new<<<warp_grid(e->G), TREE_WARPS * 32, 0, S_(stream)>>>(e->boards, e->S, e->G, e->T, sgo_pool(e), e->meta, e->root_p64,
                                                                   d_tree_sel, d_policy, d_noise, 1.0 - eps, eps, e->err_flags);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_select_a(sgo_engine *e, const int32_t *d_tree_sel, int32_t batch, void *stream)
{
    if (batch < 1 || batch > e->L || batch > 128) return sgo_fail(e, "batch must be 1..min(max_leaves,128)");
    k_select_a<<<warp_grid(e->G), TREE_WARPS * 32, 0, S_(stream)>>>(e->boards, e->S, e->G, e->T, e->L, sgo_pool(e), e->meta,
                                                                   e->root_p64, d_tree_sel, batch, e->leaf_boards, e->leaf_refs,
                                                                   e->leaf_count, e->err_flags);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_select_b_sync(sgo_engine *e, const int32_t *d_tree_sel, int32_t energy, int32_t restart,
                                      int32_t *h_counts, void *stream)
{
    if (energy < 1 || energy > e->L) return sgo_fail(e, "energy must be 1..max_leaves");
    SGO_CUDA_OK(e, cudaMemsetAsync(e->counters, 0, sizeof(int32_t) * 2, S_(stream)));
    k_select_b<<<warp_grid(e->G), TREE_WARPS * 32, 0, S_(stream)>>>(e->boards, e->S, e->G, e->T, e->L, sgo_pool(e), e->meta,
                                                                   e->root_p64, d_tree_sel, energy, restart, e->leaf_boards,
                                                                   e->leaf_refs, e->leaf_count, e->wave, e->counters, e->err_flags);
    LAUNCH_OK(e);
    SGO_CUDA_OK(e, cudaMemcpyAsync(e->h_pinned + 4, e->counters, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, S_(stream)));
    SGO_CUDA_OK(e, cudaStreamSynchronize(S_(stream)));
    if (h_counts) { h_counts[0] = e->h_pinned[4]; h_counts[1] = e->h_pinned[5]; }
    return 0;
}

extern "C" int sgo_tree_expand(sgo_engine *e, const int32_t *d_tree_sel, const float *d_policy, const float *d_value, void *stream)
{
    size_t n = (size_t)e->G * e->L;
    k_expand<<<(unsigned)((n + TREE_WARPS - 1) / TREE_WARPS), TREE_WARPS * 32, 0, S_(stream)>>>(
        e->boards, e->S, e->G, e->T, e->L, sgo_pool(e), e->meta, d_tree_sel, e->leaf_boards, e->leaf_refs, d_policy, d_value);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_backup_a(sgo_engine *e, const int32_t *d_tree_sel, void *stream)
{
    k_backup_a<<<(e->G + 63) / 64, 64, 0, S_(stream)>>>(e->G, e->T, e->L, sgo_pool(e), e->meta, d_tree_sel, e->leaf_refs, e->leaf_count);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_backup_b(sgo_engine *e, const int32_t *d_tree_sel, int32_t total_energy, void *stream)
{
    k_backup_b<<<(e->G + 63) / 64, 64, 0, S_(stream)>>>(e->G, e->T, e->L, sgo_pool(e), e->meta, d_tree_sel, e->leaf_refs, e->wave, total_energy);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_pick(sgo_engine *e, const int32_t *d_tree_sel, const int32_t *d_temperature, const double *d_u01,
                             const int32_t *d_forced, int32_t *d_move_out, void *stream)
{
    k_pick<<<warp_grid(e->G), TREE_WARPS * 32, 0, S_(stream)>>>(e->S, e->G, e->T, sgo_pool(e), e->meta, d_tree_sel, d_temperature,
                                                               d_u01, d_forced, d_move_out);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_reroot(sgo_engine *e, const int32_t *d_moves, void *stream)
{
    k_reroot<<<warp_grid(e->G), TREE_WARPS * 32, 0, S_(stream)>>>(e->G, e->T, sgo_pool(e), e->meta, d_moves);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_child_stats(sgo_engine *e, const int32_t *d_tree_sel, double *d_prior, int32_t *d_count, float *d_value, void *stream)
{
    k_child_stats<<<warp_grid(e->G), TREE_WARPS * 32, 0, S_(stream)>>>(e->S, e->G, e->T, sgo_pool(e), e->meta, e->root_p64, d_tree_sel,
                                                                      d_prior, d_count, d_value);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_tree_valid(sgo_engine *e, const int32_t *d_tree_sel, int32_t *d_valid, void *stream)
{
    k_tree_valid<<<(e->G + 127) / 128, 128, 0, S_(stream)>>>(e->G, e->T, e->meta, d_tree_sel, d_valid);
    LAUNCH_OK(e);
    return 0;
}

extern "C" int sgo_leaf_counts(sgo_engine *e, int32_t *d_counts, void *stream)
{
    SGO_CUDA_OK(e, cudaMemcpyAsync(d_counts, e->leaf_count, sizeof(int32_t) * e->G, cudaMemcpyDeviceToDevice, S_(stream)));
    return 0;
}


extern "C" int sgo_leaf_compact_sync(sgo_engine *e, int32_t *d_index, int32_t *h_count, void *stream)
{
    int total = e->G * e->L;
    SGO_CUDA_OK(e, cudaMemsetAsync(e->counters + 2, 0, sizeof(int32_t), S_(stream)));
    k_leaf_compact<<<(total + 255) / 256, 256, 0, S_(stream)>>>(e->leaf_refs, total, d_index, e->counters + 2);
    LAUNCH_OK(e);
    SGO_CUDA_OK(e, cudaMemcpyAsync(e->h_pinned + 7, e->counters + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, S_(stream)));
    SGO_CUDA_OK(e, cudaStreamSynchronize(S_(stream)));
    if (h_count) *h_count = e->h_pinned[7];
    return 0;
}

// staging buffer for tree download / upload (tests, checkpointing), grown on demand
static int ensure_stage(sgo_engine *e, int blocks)
{
    if (e->stage_blocks >= blocks) return 0;
    if (e->stage) cudaFree(e->stage);
    e->stage = nullptr; e->stage_blocks = 0;
    SGO_CUDA_OK(e, cudaMalloc(&e->stage, sizeof(NodeBlock) * (size_t)blocks));
    e->stage_blocks = blocks;
    return 0;
}

extern "C" int sgo_tree_download_sync(sgo_engine *e, int32_t tree, void *h_blocks, int32_t max_blocks, void *h_meta, double *h_root_p64)
{
    if (tree < 0 || tree >= e->G * e->T) return sgo_fail(e, "tree index out of range");
    TreeMeta m;
    SGO_CUDA_OK(e, cudaDeviceSynchronize());
    SGO_CUDA_OK(e, cudaMemcpy(&m, e->meta + tree, sizeof(TreeMeta), cudaMemcpyDeviceToHost));
    int n = 0;
    if (h_blocks && m.valid && max_blocks > 0) {
        int want = m.n_blocks < max_blocks ? m.n_blocks : max_blocks;
        int rc = ensure_stage(e, want);
        if (rc) return rc;
        k_tree_export<<<1, 32>>>(sgo_pool(e), e->meta, tree, e->stage, want, e->counters + 3);
        SGO_LAUNCHED(e);
        SGO_CUDA_OK(e, cudaMemcpy(&n, e->counters + 3, sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (n < 0) n = -n;                                    // truncated at max_blocks
        if (n > 0) SGO_CUDA_OK(e, cudaMemcpy(h_blocks, e->stage, sizeof(NodeBlock) * (size_t)n, cudaMemcpyDeviceToHost));
    }
    if (h_meta) { m.root = 0; memcpy(h_meta, &m, sizeof(TreeMeta)); }    // in the exported form the root is block 0
    if (h_root_p64)
        SGO_CUDA_OK(e, cudaMemcpy(h_root_p64, e->root_p64 + (size_t)tree * SGO_APAD, sizeof(double) * SGO_APAD, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sgo_tree_upload_sync(sgo_engine *e, int32_t tree, const void *h_blocks, int32_t n_blocks, const void *h_meta, const double *h_root_p64)
{
    if (tree < 0 || tree >= e->G * e->T) return sgo_fail(e, "tree index out of range");
    if (n_blocks < 0 || n_blocks > e->pool_blocks) return sgo_fail(e, "tree larger than the node pool");
    TreeMeta m;
    memcpy(&m, h_meta, sizeof(TreeMeta));
    SGO_CUDA_OK(e, cudaDeviceSynchronize());
    // whatever the slot held goes back to the pool first
    int g = tree / e->T, t = tree % e->T;
    int32_t *sel = nullptr, *mask = nullptr;
    SGO_CUDA_OK(e, cudaMalloc(&sel, sizeof(int32_t) * e->G * 2));
    mask = sel + e->G;
    SGO_CUDA_OK(e, cudaMemset(sel, 0, sizeof(int32_t) * e->G * 2));
    int32_t one = 1;
    SGO_CUDA_OK(e, cudaMemcpy(sel + g, &t, sizeof(int32_t), cudaMemcpyHostToDevice));
    SGO_CUDA_OK(e, cudaMemcpy(mask + g, &one, sizeof(int32_t), cudaMemcpyHostToDevice));
    k_tree_free<<<warp_grid(e->G), TREE_WARPS * 32>>>(e->G, e->T, sgo_pool(e), e->meta, mask, sel);
    SGO_LAUNCHED(e);
    SGO_CUDA_OK(e, cudaDeviceSynchronize());
    cudaFree(sel);
    int rc = ensure_stage(e, n_blocks > 0 ? n_blocks : 1);
    if (rc) return rc;
    if (n_blocks > 0) SGO_CUDA_OK(e, cudaMemcpy(e->stage, h_blocks, sizeof(NodeBlock) * (size_t)n_blocks, cudaMemcpyHostToDevice));
    k_tree_import<<<1, 32>>>(sgo_pool(e), e->meta, tree, e->stage, n_blocks, m, e->err_flags);
    SGO_LAUNCHED(e);
    if (h_root_p64)
        SGO_CUDA_OK(e, cudaMemcpy(e->root_p64 + (size_t)tree * SGO_APAD, h_root_p64, sizeof(double) * SGO_APAD, cudaMemcpyHostToDevice));
    SGO_CUDA_OK(e, cudaDeviceSynchronize());
    return 0;
}

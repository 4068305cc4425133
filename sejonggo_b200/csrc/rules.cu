// rules.cu — engine lifetime + the rules engine kernels (play.py) behind the C ABI.
// One warp per game; lane r holds row r of every bitboard (common.cuh).
#include "engine.h"
#include <cstdio>
#include <cstring>
#include <new>

#define WARPS_PER_BLOCK 4

static inline dim3 warp_grid(int n) { return dim3((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }
static inline cudaStream_t S_(void *s) { return (cudaStream_t)s; }

// ---------------------------------------------------------------- kernels
__global__ void k_games_reset(Board *boards, int first, int n)
{
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= n) return;
    uint32_t *p = reinterpret_cast<uint32_t *>(boards + first + w);
    for (int i = lane; i < (int)(sizeof(Board) / 4); i += 32) p[i] = 0;
    __syncwarp();
    if (lane == 0) { boards[first + w].head = 0; boards[first + w].to_move = 1; }
}

__global__ void k_apply_moves(Board *boards, int S, int first, int n, const int32_t *moves, const int32_t *colors, int32_t *err)
{
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= n) return;
    int mv = moves[w];
    if (mv < 0) return;
    int color = colors ? colors[w] : 0;
    if (!board_play(boards + first + w, S, mv, color, lane) && lane == 0) atomicOr(err, SGO_ERR_OCCUPIED);
}

__global__ void k_legal_masks(const Board *boards, int S, int first, int n, uint8_t *out)
{
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= n) return;
    uint32_t ill = board_illegal(boards + first + w, S, lane);
    int A = S * S + 1;
    uint8_t *o = out + (size_t)w * A;
    if (lane < S)
        for (int x = 0; x < S; x++) o[lane * S + x] = (ill >> x) & 1u;
    if (lane == 0) o[S * S] = 0;
}

// play.py:244-292: area score.  reach_c = empties connected to colour c.
__device__ __forceinline__ void bb_score(uint32_t black, uint32_t white, uint32_t rm, int lane, int &bp, int &wp)
{
    uint32_t empty = ~(black | white) & rm;
    uint32_t rb = bb_flood(bb_nbr(black, rm, lane) & empty, empty, lane);
    uint32_t rw = bb_flood(bb_nbr(white, rm, lane) & empty, empty, lane);
    bp = warp_sum(__popc(black) + __popc(rb & ~rw));
    wp = warp_sum(__popc(white) + __popc(rw & ~rb));
}

__global__ void k_score(const Board *boards, int S, float komi, int first, int n, int32_t *out)
{
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= n) return;
    int head, tm, bp, wp;
    uint32_t bl, wh;
    board_load_cur(boards + first + w, lane, bl, wh, head, tm);
    uint32_t rm = row_mask(S, lane);
    bb_score(bl & rm, wh & rm, rm, lane, bp, wp);
    if (lane == 0) {
        float white = (float)wp + komi;          // komi 5.5 is exact in fp32/fp64 alike
        out[w * 3 + 0] = (float)bp > white ? 1 : ((float)bp == white ? 0 : -1);
        out[w * 3 + 1] = bp;
        out[w * 3 + 2] = wp;
    }
}

// reference tensor [n][S][S][17] int32 <-> Board.  Plane 2k = side-to-move's stones
// k plies ago, 2k+1 = the opponent's, plane 16 = +-1 (play.py:295-299).
__global__ void k_import_boards(Board *boards, int S, int first, int n, const int32_t *in)
{
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= n) return;
    Board *b = boards + first + w;
    const int32_t *src = in + (size_t)w * S * S * 17;
    int tm = src[16] >= 0 ? 1 : -1;
    for (int k = 0; k < SGO_HIST; k++) {
        uint32_t own = 0, opp = 0;
        if (lane < S)
            for (int x = 0; x < S; x++) {
                const int32_t *c = src + ((size_t)lane * S + x) * 17;
                own |= (c[2 * k] != 0 ? 1u : 0u) << x;
                opp |= (c[2 * k + 1] != 0 ? 1u : 0u) << x;
            }
        int slot = (SGO_HIST - k) & (SGO_HIST - 1);        // head = 0, k plies ago = slot -k
        if (lane < SGO_ROWW) {
            b->st[slot][0][lane] = tm == 1 ? own : opp;
            b->st[slot][1][lane] = tm == 1 ? opp : own;
        }
    }
    if (lane == 0) { b->head = 0; b->to_move = tm; b->pad[0] = b->pad[1] = 0; }
}

__device__ __forceinline__ void board_plane_rows(const Board *b, int k, int lane, uint32_t &own, uint32_t &opp)
{
    int slot = (b->head + SGO_HIST - k) & (SGO_HIST - 1);
    int tm = b->to_move;
    uint32_t bl = lane < SGO_ROWW ? b->st[slot][0][lane] : 0u;
    uint32_t wh = lane < SGO_ROWW ? b->st[slot][1][lane] : 0u;
    own = tm == 1 ? bl : wh;
    opp = tm == 1 ? wh : bl;
}

__global__ void k_export_boards(const Board *boards, int S, int first, int n, int32_t *out)
{
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= n) return;
    const Board *b = boards + first + w;
    int32_t *dst = out + (size_t)w * S * S * 17;
    int tm = b->to_move;
    for (int k = 0; k < SGO_HIST; k++) {
        uint32_t own, opp;
        board_plane_rows(b, k, lane, own, opp);
        if (lane < S)
            for (int x = 0; x < S; x++) {
                int32_t *c = dst + ((size_t)lane * S + x) * 17;
                c[2 * k] = (own >> x) & 1u;
                c[2 * k + 1] = (opp >> x) & 1u;
                if (k == 0) c[16] = tm;
            }
    }
}

// comparison format of oracle/go_oracle.c orc_pack_board: 16 planes x W words + to_move
__global__ void k_export_packed(const Board *boards, int S, int first, int n, uint32_t *out)
{
    __shared__ uint32_t scratch[WARPS_PER_BLOCK][SGO_AWORDS];
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= n) return;
    const Board *b = boards + first + w;
    int W = (S * S + 31) / 32;
    uint32_t *dst = out + (size_t)w * (16 * W + 1);
    uint32_t *sc = scratch[(threadIdx.x >> 5)];
    for (int k = 0; k < SGO_HIST; k++) {
        uint32_t own, opp;
        board_plane_rows(b, k, lane, own, opp);
        uint32_t rm = row_mask(S, lane);
        illegal_rows_to_words(own & rm, S, lane, sc);
        if (lane < W) dst[(2 * k) * W + lane] = sc[lane];
        __syncwarp();
        illegal_rows_to_words(opp & rm, S, lane, sc);
        if (lane < W) dst[(2 * k + 1) * W + lane] = sc[lane];
        __syncwarp();
    }
    if (lane == 0) dst[16 * W] = (uint32_t)b->to_move;
}

// symmetry.py:45-114 as probed: out[y,x] = in[g(y,x)]
__device__ __forceinline__ void sym_src(int S, int sym, int y, int x, int &sy, int &sx)
{
    int m = S - 1;
    switch (sym) {
    case 0: sy = y; sx = x; break;
    case 1: sy = x; sx = y; break;
    case 2: sy = y; sx = m - x; break;
    case 3: sy = m - y; sx = x; break;
    case 4: sy = x; sx = m - y; break;
    case 5: sy = m - y; sx = m - x; break;
    case 6: sy = m - x; sx = y; break;
    default: sy = m - x; sx = m - y; break;
    }
}

// float32 planes [n][S][S][17] with the symmetry gather fused: one block per position,
// the 16 bitboards staged in shared memory, output written fully coalesced.
__global__ void k_export_planes(const Board *boards, int S, int first, const int32_t *index, int n, int sym, const int32_t *syms, float *out)
{
    __shared__ uint32_t rows[16][SGO_ROWW];
    __shared__ int tm_s;
    int g = blockIdx.x;
    if (g >= n) return;
    const Board *b = boards + (index ? index[g] : first + g);
    int tm = b->to_move, head = b->head;
    if (syms) sym = syms[g] & 7;
    for (int i = threadIdx.x; i < 16 * SGO_ROWW; i += blockDim.x) {
        int p = i / SGO_ROWW, r = i - p * SGO_ROWW;
        int k = p >> 1, slot = (head + SGO_HIST - k) & (SGO_HIST - 1);
        int own_c = tm == 1 ? 0 : 1;
        int c = (p & 1) ? (1 - own_c) : own_c;
        rows[p][r] = b->st[slot][c][r];
    }
    if (threadIdx.x == 0) tm_s = tm;
    __syncthreads();
    int total = S * S * 17;
    float *dst = out + (size_t)g * total;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int cell = i / 17, p = i - cell * 17;
        int y = cell / S, x = cell - y * S, sy, sx;
        sym_src(S, sym, y, x, sy, sx);
        float v = p == 16 ? (float)tm_s : (float)((rows[p][sy] >> sx) & 1u);
        dst[i] = v;
    }
}

__global__ void k_policy_unsym(int S, int n, int sym, const int32_t *syms, const float *in, float *out)
{
    int A = S * S + 1;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * A) return;
    int b = i / A, a = i - b * A;
    int src = a;
    if (syms) sym = syms[b] & 7;
    if (a < S * S) {
        int y = a / S, x = a - y * S, sy, sx;
        sym_src(S, sym, y, x, sy, sx);
        src = sy * S + sx;
    }
    out[i] = in[(size_t)b * A + src];
}

// device-resident random legal playouts, one warp per game.
__global__ void k_random_playouts(Board *boards, int S, int first, int n, uint64_t seed, int max_plies,
                                  int16_t *moves, int32_t *nplies)
{
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= n) return;
    Board *b = boards + first + w;
    int16_t *mv = moves + (size_t)w * max_plies;
    int passes = 0, t = 0;
    for (; t < max_plies && passes < 2; t++) {
        uint32_t ill = board_illegal(b, S, lane);
        uint32_t legal = ~ill & row_mask(S, lane);
        int cnt = __popc(legal), incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(SGO_FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int total = __shfl_sync(SGO_FULL, incl, 31);
        int move = S * S;
        if (total > 0) {
            uint64_t r = sgo_mix64(seed ^ sgo_mix64(((uint64_t)(first + w) << 32) | (uint32_t)t));
            int k = (int)(r % (uint64_t)total);
            int excl = incl - cnt;
            bool mine = k >= excl && k < incl;
            int x = 0;
            if (mine) {
                uint32_t v = legal;
                for (int j = k - excl; j > 0; j--) v &= v - 1;
                x = __ffs(v) - 1;
            }
            unsigned bal = __ballot_sync(SGO_FULL, mine);
            int src = __ffs(bal) - 1;
            x = __shfl_sync(SGO_FULL, x, src);
            move = src * S + x;
        }
        board_play(b, S, move, 0, lane);
        if (lane == 0) mv[t] = (int16_t)move;
        passes = move == S * S ? passes + 1 : 0;
    }
    for (int i = t + lane; i < max_plies; i += 32) mv[i] = -1;
    if (lane == 0) nplies[w] = t;
}

// ---------------------------------------------------------------- C ABI
extern "C" int sgo_abi_version(void) { return 1; }

extern "C" int64_t sgo_launch_count(sgo_engine *e) { return e ? (int64_t)e->launches : 0; }

extern "C" const char *sgo_last_error(sgo_engine *e) { return e ? e->last_error.c_str() : "null engine"; }


extern "C" int sgo_create(const sgo_config *cfg, sgo_engine **out)
{
    if (!cfg || !out) return -1;
    if (cfg->size < 2 || cfg->size > SGO_MAXS || cfg->n_games < 1 || cfg->trees_per_game < 1 ||
        cfg->trees_per_game > 2 || cfg->max_leaves < 1 || cfg->arena_blocks < 2)
        return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || cfg->device >= ndev) return -3;   // no CPU fallback
    sgo_engine *e = new (std::nothrow) sgo_engine();
    if (!e) return -4;
    e->cfg = *cfg;
    e->S = cfg->size; e->A = cfg->size * cfg->size + 1; e->G = cfg->n_games; e->T = cfg->trees_per_game;
    e->L = cfg->max_leaves; e->NB = cfg->arena_blocks;
    e->tower[0] = e->tower[1] = nullptr;
    e->step_policy = e->step_value = nullptr; e->step_index = e->step_sym = nullptr;
    e->launches = 0;
    e->stage = nullptr; e->stage_blocks = 0;
    e->pool_blocks = (long long)cfg->n_games * cfg->trees_per_game * cfg->arena_blocks;
    if (e->pool_blocks > 0x7fffffffLL) { delete e; return -1; }
    *out = e;
    SGO_CUDA_OK(e, cudaSetDevice(cfg->device));
    size_t GL = (size_t)e->G * e->L, GT = (size_t)e->G * e->T;
    SGO_CUDA_OK(e, cudaMalloc(&e->boards, sizeof(Board) * e->G));
    SGO_CUDA_OK(e, cudaMalloc(&e->leaf_boards, sizeof(Board) * GL));
    SGO_CUDA_OK(e, cudaMalloc(&e->leaf_refs, sizeof(LeafRef) * GL));
    SGO_CUDA_OK(e, cudaMalloc(&e->leaf_count, sizeof(int32_t) * e->G));
    SGO_CUDA_OK(e, cudaMalloc(&e->arena, sizeof(NodeBlock) * (size_t)e->pool_blocks));
    SGO_CUDA_OK(e, cudaMalloc(&e->free_list, sizeof(int32_t) * (size_t)e->pool_blocks));
    SGO_CUDA_OK(e, cudaMalloc(&e->pool_ctl, sizeof(int32_t) * 4));
    SGO_CUDA_OK(e, cudaMalloc(&e->zero_sel, sizeof(int32_t) * e->G));
    SGO_CUDA_OK(e, cudaMemset(e->zero_sel, 0, sizeof(int32_t) * e->G));
    SGO_CUDA_OK(e, cudaMalloc(&e->meta, sizeof(TreeMeta) * GT));
    SGO_CUDA_OK(e, cudaMalloc(&e->root_p64, sizeof(double) * SGO_APAD * GT));
    SGO_CUDA_OK(e, cudaMalloc(&e->wave, sizeof(int32_t) * 8 * e->G));
    SGO_CUDA_OK(e, cudaMalloc(&e->err_flags, sizeof(int32_t)));
    SGO_CUDA_OK(e, cudaMalloc(&e->counters, sizeof(int32_t) * 4));
    SGO_CUDA_OK(e, cudaMallocHost(&e->h_pinned, sizeof(int32_t) * 16));
    SGO_CUDA_OK(e, cudaMemset(e->err_flags, 0, sizeof(int32_t)));
    SGO_CUDA_OK(e, cudaMemset(e->counters, 0, sizeof(int32_t) * 4));
    SGO_CUDA_OK(e, cudaMemset(e->leaf_refs, 0, sizeof(LeafRef) * GL));
    SGO_CUDA_OK(e, cudaMemset(e->leaf_count, 0, sizeof(int32_t) * e->G));
    SGO_CUDA_OK(e, cudaMemset(e->leaf_boards, 0, sizeof(Board) * GL));
    SGO_CUDA_OK(e, cudaMemset(e->meta, 0, sizeof(TreeMeta) * GT));
    SGO_CUDA_OK(e, cudaMemset(e->root_p64, 0, sizeof(double) * SGO_APAD * GT));
    SGO_CUDA_OK(e, cudaMemset(e->wave, 0, sizeof(int32_t) * 8 * e->G));
    k_games_reset<<<warp_grid(e->G), WARPS_PER_BLOCK * 32>>>(e->boards, 0, e->G);
    SGO_CUDA_OK(e, cudaGetLastError());
    int rc = sgo_tree_reset(e, nullptr);                 // metas invalid, every pool block on the free stack
    if (rc) return rc;
    SGO_CUDA_OK(e, cudaDeviceSynchronize());
    return 0;
}

extern "C" int sgo_tower_free(sgo_engine *e, int32_t slot);

extern "C" int sgo_destroy(sgo_engine *e)
{
    if (!e) return 0;
    cudaSetDevice(e->cfg.device);
    sgo_tower_free(e, 0);
    sgo_tower_free(e, 1);
    cudaFree(e->boards); cudaFree(e->leaf_boards); cudaFree(e->leaf_refs); cudaFree(e->leaf_count);
    cudaFree(e->arena); cudaFree(e->free_list); cudaFree(e->pool_ctl); cudaFree(e->zero_sel); cudaFree(e->stage);
    cudaFree(e->meta); cudaFree(e->root_p64);
    cudaFree(e->step_policy); cudaFree(e->step_value); cudaFree(e->step_index); cudaFree(e->step_sym);
    cudaFree(e->wave); cudaFree(e->err_flags); cudaFree(e->counters); cudaFreeHost(e->h_pinned);
    delete e;
    return 0;
}

extern "C" int sgo_check_errors_sync(sgo_engine *e, void *stream, int32_t *h_flags)
{
    SGO_CUDA_OK(e, cudaMemcpyAsync(e->h_pinned, e->err_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, S_(stream)));
    SGO_CUDA_OK(e, cudaMemsetAsync(e->err_flags, 0, sizeof(int32_t), S_(stream)));
    SGO_CUDA_OK(e, cudaStreamSynchronize(S_(stream)));
    if (h_flags) *h_flags = e->h_pinned[0];
    return 0;
}

#define RANGE_OK(e, first, n, limit) \
    if ((first) < 0 || (n) < 0 || (first) + (n) > (limit)) return sgo_fail(e, "range out of bounds")

extern "C" int sgo_games_reset(sgo_engine *e, int32_t first, int32_t n, void *stream)
{
    RANGE_OK(e, first, n, e->G);
    if (n == 0) return 0;
    k_games_reset<<<warp_grid(n), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(e->boards, first, n);
    SGO_LAUNCHED(e);
    return 0;
}

__global__ void k_games_restart(Board *boards, int G, const int32_t *mask)
{
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (w >= G || !mask[w]) return;
    uint32_t *p = reinterpret_cast<uint32_t *>(boards + w);
    for (int i = lane; i < (int)(sizeof(Board) / 4); i += 32) p[i] = 0;
    __syncwarp();
    if (lane == 0) { boards[w].head = 0; boards[w].to_move = 1; }
}

extern "C" int sgo_games_restart(sgo_engine *e, const int32_t *d_game_mask, void *stream)
{
    if (!d_game_mask) return sgo_fail(e, "games_restart needs a mask");
    k_games_restart<<<warp_grid(e->G), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(e->boards, e->G, d_game_mask);
    SGO_LAUNCHED(e);
    return sgo_tree_free(e, d_game_mask, stream);
}

extern "C" int sgo_apply_moves(sgo_engine *e, int32_t first, int32_t n, const int32_t *d_moves, const int32_t *d_colors, void *stream)
{
    RANGE_OK(e, first, n, e->G);
    if (n == 0) return 0;
    k_apply_moves<<<warp_grid(n), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(e->boards, e->S, first, n, d_moves, d_colors, e->err_flags);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_legal_masks(sgo_engine *e, int32_t first, int32_t n, uint8_t *d_mask, void *stream)
{
    RANGE_OK(e, first, n, e->G);
    if (n == 0) return 0;
    k_legal_masks<<<warp_grid(n), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(e->boards, e->S, first, n, d_mask);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_score(sgo_engine *e, int32_t first, int32_t n, int32_t *d_out, void *stream)
{
    RANGE_OK(e, first, n, e->G);
    if (n == 0) return 0;
    k_score<<<warp_grid(n), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(e->boards, e->S, e->cfg.komi, first, n, d_out);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_import_boards(sgo_engine *e, int32_t first, int32_t n, const int32_t *d_boards, void *stream)
{
    RANGE_OK(e, first, n, e->G);
    if (n == 0) return 0;
    k_import_boards<<<warp_grid(n), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(e->boards, e->S, first, n, d_boards);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_export_boards(sgo_engine *e, int32_t first, int32_t n, int32_t *d_boards, void *stream)
{
    RANGE_OK(e, first, n, e->G);
    if (n == 0) return 0;
    k_export_boards<<<warp_grid(n), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(e->boards, e->S, first, n, d_boards);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_export_packed(sgo_engine *e, int32_t which, int32_t first, int32_t n, uint32_t *d_out, void *stream)
{
    RANGE_OK(e, first, n, which ? e->G * e->L : e->G);
    if (n == 0) return 0;
    k_export_packed<<<warp_grid(n), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(which ? e->leaf_boards : e->boards, e->S, first, n, d_out);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_export_planes(sgo_engine *e, int32_t which, int32_t first, int32_t n, int32_t sym, const int32_t *d_sym, float *d_out, void *stream)
{
    RANGE_OK(e, first, n, which ? e->G * e->L : e->G);
    if (sym < 0 || sym > 7) return sgo_fail(e, "symmetry id out of range");
    if (n == 0) return 0;
    k_export_planes<<<n, 256, 0, S_(stream)>>>(which ? e->leaf_boards : e->boards, e->S, first, nullptr, n, sym, d_sym, d_out);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_export_planes_indexed(sgo_engine *e, int32_t which, const int32_t *d_index, int32_t n, const int32_t *d_sym,
                                         float *d_out, void *stream)
{
    if (n < 0 || n > (which ? e->G * e->L : e->G)) return sgo_fail(e, "export_planes_indexed: n out of range");
    if (n == 0) return 0;
    k_export_planes<<<n, 256, 0, S_(stream)>>>(which ? e->leaf_boards : e->boards, e->S, 0, d_index, n, 0, d_sym, d_out);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_policy_unsym(sgo_engine *e, int32_t n, int32_t sym, const int32_t *d_sym, const float *d_in, float *d_out, void *stream)
{
    if (sym < 0 || sym > 7) return sgo_fail(e, "symmetry id out of range");
    if (d_in == d_out) return sgo_fail(e, "policy_unsym must not run in place");
    if (n == 0) return 0;
    int total = n * e->A;
    k_policy_unsym<<<(total + 255) / 256, 256, 0, S_(stream)>>>(e->S, n, sym, d_sym, d_in, d_out);
    SGO_LAUNCHED(e);
    return 0;
}

extern "C" int sgo_random_playouts(sgo_engine *e, int32_t first, int32_t n, uint64_t seed, int32_t max_plies,
                                   int16_t *d_moves, int32_t *d_nplies, void *stream)
{
    RANGE_OK(e, first, n, e->G);
    if (n == 0) return 0;
    k_random_playouts<<<warp_grid(n), WARPS_PER_BLOCK * 32, 0, S_(stream)>>>(e->boards, e->S, first, n, seed, max_plies, d_moves, d_nplies);
    SGO_LAUNCHED(e);
    return 0;
}

// common.cuh — shared device types and warp-level bitboard primitives.
//
// Data layout in HBM:
//   Board      one game position incl. the 8-ply history the reference keeps in
//              planes 0..15 (play.py:295-299), stored as ABSOLUTE-colour row
//              bitboards in a ring, so the reference's "shift 14 planes + swap
//              pairs" (play.py:219-242) becomes "write one ring slot".
//   NodeBlock  the children of one expanded MCTS node (play.py:376-421 dict
//              node), SoA over the S*S+1 action slots so a warp reads priors /
//              counts / value sums with coalesced 128 B transactions.
//
// One WARP owns one game: lane r holds row r (bit x = column x) of a bitboard.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SGO_MAXS 19
#define SGO_ROWW 20                 // words per bitboard (19 rows + 1 pad -> 80 B, 16 B aligned)
#define SGO_HIST 8
#define SGO_APAD 384                // action slots per node block (12 x 32 >= 19*19+1)
#define SGO_AWORDS 12
#define SGO_FULL 0xffffffffu

struct __align__(16) Board {
    uint32_t st[SGO_HIST][2][SGO_ROWW];   // [ring slot][0 black,1 white][row]
    int32_t head;                         // ring slot of the current position
    int32_t to_move;                      // +1 black, -1 white  (plane 16)
    int32_t pad[2];
};                                        // 1296 B
static_assert(sizeof(Board) == 1296, "Board layout");

struct __align__(16) NodeBlock {
    float prior[SGO_APAD];                // p, float32 (play.py:415)
    int32_t n[SGO_APAD];                  // count
    float w[SGO_APAD];                    // value (float32 running sum)
    int32_t child[SGO_APAD];              // block index of the expanded child, -1 = subtree {}
    uint32_t exist[SGO_AWORDS];           // slot has a child entry (legal at expansion time)
    uint32_t busy[SGO_AWORDS];            // virtual_loss > 0 (mode B)
    int32_t parent_block, parent_slot;    // entry of this node in its parent, -1 at the root
    int32_t pad[6];
};                                        // 6272 B
static_assert(sizeof(NodeBlock) == 6272, "NodeBlock layout");

struct TreeMeta {
    int32_t root;          // pool id of the root block (meaningful while valid)
    int32_t n_blocks;      // blocks this tree owns in the pool
    int32_t valid;         // 0: tree is None / has an empty subtree (self_play.py:195)
    int32_t root_f64;      // root priors are float64 (Dirichlet-noised, play.py:401-403)
    int32_t root_count;    // the root node's own count / value (self_play.py:110-112)
    float root_value;
    int32_t overflow;      // an allocation for this tree failed: the node pool was exhausted (sticky until the tree is freed)
    int32_t pad;
};

struct LeafRef {           // one selected leaf = entry `slot` of block `block`
    int32_t block, slot;   // pool id of the parent block, action slot in it
    int32_t to_move;       // side to move at the leaf position
    int32_t state;         // 0 empty, 1 selected (awaiting evaluation), 2 evaluated/expanded
    int32_t new_block;     // pool id of the block allocated for its children
    float sv;              // signed value to back up
    int32_t pad[2];
};

// error bits (device-side sticky flags, sgo_check_errors)
#define SGO_ERR_OCCUPIED 1      // make_play on an occupied point (reference assert, play.py:233)
#define SGO_ERR_ARENA 2         // node arena exhausted
#define SGO_ERR_NOACTION 4      // selector found no child (reference would raise)
#define SGO_ERR_DEPTH 8         // path deeper than SGO_MAXDEPTH
#define SGO_MAXDEPTH 1024

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t row_mask(int S, int lane) { return lane < S ? ((1u << S) - 1u) : 0u; }

// 4-neighbourhood of a bitboard (without the bits themselves)
__device__ __forceinline__ uint32_t bb_nbr(uint32_t v, uint32_t rm, int lane)
{
    uint32_t up = __shfl_up_sync(SGO_FULL, v, 1);
    uint32_t dn = __shfl_down_sync(SGO_FULL, v, 1);
    if (lane == 0) up = 0;
    if (lane == 31) dn = 0;
    return ((v << 1) | (v >> 1) | up | dn) & rm;
}

// connected closure of `seed` inside `within` (4-connectivity): the warp-level
// flood fill that replaces play.py:160-180's recursive capture_group.
__device__ __forceinline__ uint32_t bb_flood(uint32_t seed, uint32_t within, int lane)
{
    uint32_t g = seed & within;
    for (;;) {
        uint32_t prev;
        do {                                   // close along the row first
            prev = g;
            g |= ((g << 1) | (g >> 1)) & within;
        } while (g != prev);
        __syncwarp();
        uint32_t up = __shfl_up_sync(SGO_FULL, g, 1);
        uint32_t dn = __shfl_down_sync(SGO_FULL, g, 1);
        if (lane == 0) up = 0;
        if (lane == 31) dn = 0;
        uint32_t add = (up | dn) & within & ~g;
        if (!__any_sync(SGO_FULL, add != 0)) break;
        g |= add;
    }
    return g;
}

__device__ __forceinline__ int warp_sum(int v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(SGO_FULL, v, o);
    return v;
}

// play.py:182-217 take_stones on row bitboards: `own` = mover (stone already
// placed), `opp` = other colour.  Captures first, then self-capture (Q5).
__device__ __forceinline__ void bb_resolve(uint32_t &own, uint32_t &opp, uint32_t placed, uint32_t rm, int lane)
{
    uint32_t empty = ~(own | opp) & rm;
    uint32_t seeds = bb_nbr(placed, rm, lane) & opp;
    uint32_t captured = 0;
    while (__any_sync(SGO_FULL, seeds != 0)) {
        unsigned b = __ballot_sync(SGO_FULL, seeds != 0);
        int sl = __ffs(b) - 1;
        uint32_t sv = __shfl_sync(SGO_FULL, seeds, sl);
        uint32_t one = sv & (0u - sv);
        uint32_t g = bb_flood(lane == sl ? one : 0u, opp, lane);
        if (!__any_sync(SGO_FULL, (bb_nbr(g, rm, lane) & empty) != 0)) captured |= g;
        seeds &= ~g;
    }
    opp &= ~captured;
    empty |= captured;
    uint32_t g = bb_flood(placed, own, lane);
    if (!__any_sync(SGO_FULL, (bb_nbr(g, rm, lane) & empty) != 0)) own &= ~g;
}

// play.py:71-104 legal_moves.  Returns this lane's ILLEGAL bits (1 = illegal)
// for the side to move: own/opp = current stones, prev_own = side-to-move's
// stones one ply ago (plane 2).
__device__ __forceinline__ uint32_t bb_illegal(uint32_t own, uint32_t opp, uint32_t prev_own, uint32_t rm, int lane)
{
    uint32_t occ = own | opp;
    uint32_t empty = ~occ & rm;
    uint32_t ko = prev_own & ~own & rm;                       // plane2 - plane0 == 1
    int nko = warp_sum(__popc(ko));
    if (nko != 1) ko = 0;
    uint32_t has_empty_nbr = bb_nbr(empty, rm, lane);
    uint32_t cand = empty & ~has_empty_nbr;                   // all neighbours are stones/edge
    uint32_t l1 = 0;                                          // sole liberties of opponent groups in atari
    if (__any_sync(SGO_FULL, cand != 0)) {
        uint32_t seeds = bb_nbr(cand, rm, lane) & opp;
        while (__any_sync(SGO_FULL, seeds != 0)) {
            unsigned b = __ballot_sync(SGO_FULL, seeds != 0);
            int sl = __ffs(b) - 1;
            uint32_t sv = __shfl_sync(SGO_FULL, seeds, sl);
            uint32_t one = sv & (0u - sv);
            uint32_t g = bb_flood(lane == sl ? one : 0u, opp, lane);
            uint32_t libs = bb_nbr(g, rm, lane) & empty;
            if (warp_sum(__popc(libs)) == 1) l1 |= libs;
            seeds &= ~g;
        }
    }
    uint32_t legal = empty & ~ko & (has_empty_nbr | l1);
    return ~legal & rm;
}

// ---- Board accessors (warp-collective: lane = row) -------------------------
__device__ __forceinline__ void board_load_cur(const Board *b, int lane, uint32_t &black, uint32_t &white, int &head, int &to_move)
{
    head = b->head;
    to_move = b->to_move;
    black = lane < SGO_ROWW ? b->st[head][0][lane] : 0u;
    white = lane < SGO_ROWW ? b->st[head][1][lane] : 0u;
}

// make_play (play.py:226-242) on a Board in global/shared memory.  mv = y*S+x or
// S*S for pass; color 0 = side to move.  Returns false if the point is occupied.
__device__ __forceinline__ bool board_play(Board *b, int S, int mv, int color, int lane)
{
    int head, tm;
    uint32_t bl, wh;
    board_load_cur(b, lane, bl, wh, head, tm);
    if (color != 0) tm = color;                         // swap_player when colour differs
    uint32_t rm = row_mask(S, lane);
    bl &= rm; wh &= rm;
    uint32_t own = tm == 1 ? bl : wh, opp = tm == 1 ? wh : bl;
    bool ok = true;
    if (mv != S * S) {
        int y = mv / S, x = mv - y * S;
        uint32_t placed = (lane == y) ? (1u << x) : 0u;
        if (__any_sync(SGO_FULL, (placed & (own | opp)) != 0)) ok = false;
        else {
            own |= placed;
            bb_resolve(own, opp, placed, rm, lane);
        }
    }
    if (!ok) return false;
    int nh = (head + 1) & (SGO_HIST - 1);
    __syncwarp();
    if (lane < SGO_ROWW) {
        b->st[nh][0][lane] = tm == 1 ? own : opp;
        b->st[nh][1][lane] = tm == 1 ? opp : own;
    }
    if (lane == 0) { b->head = nh; b->to_move = -tm; }
    __syncwarp();
    return true;
}

// copy a Board with the whole warp (1296 B = 81 x 16 B)
__device__ __forceinline__ void board_copy(Board *dst, const Board *src, int lane)
{
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = lane; i < (int)(sizeof(Board) / 16); i += 32) d[i] = s[i];
    __syncwarp();
}

// legality of the position in *b for its side to move; returns lane's illegal bits
__device__ __forceinline__ uint32_t board_illegal(const Board *b, int S, int lane)
{
    int head, tm;
    uint32_t bl, wh;
    board_load_cur(b, lane, bl, wh, head, tm);
    uint32_t rm = row_mask(S, lane);
    int ph = (head + SGO_HIST - 1) & (SGO_HIST - 1);
    uint32_t prev_own = lane < SGO_ROWW ? b->st[ph][tm == 1 ? 0 : 1][lane] : 0u;
    uint32_t own = (tm == 1 ? bl : wh) & rm, opp = (tm == 1 ? wh : bl) & rm;
    return bb_illegal(own, opp, prev_own & rm, rm, lane);
}

// Row-bitboard illegal mask -> 12 action words (bit a = y*S+x; pass bit clear).
// Uses a small shared scratch (>= 12 words per warp).
__device__ __forceinline__ void illegal_rows_to_words(uint32_t ill, int S, int lane, uint32_t *scratch /*[12]*/)
{
    if (lane < SGO_AWORDS) scratch[lane] = 0;
    __syncwarp();
    if (lane < S) {
        int base = lane * S;                       // bits base .. base+S-1
        int w0 = base >> 5, sh = base & 31;
        uint64_t v = (uint64_t)ill << sh;
        atomicOr(&scratch[w0], (uint32_t)v);
        if ((v >> 32) != 0) atomicOr(&scratch[w0 + 1], (uint32_t)(v >> 32));
    }
    __syncwarp();
}

// splitmix64 — counter-based RNG for throughput runs (parity runs inject draws)
__device__ __host__ __forceinline__ uint64_t sgo_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// engine.h — host-side engine object behind the C ABI (include/sejonggo_b200.h).
#pragma once
#include "common.cuh"
#include "../../include/sejonggo_b200.h"
#include <string>

struct sgo_tower;

struct sgo_engine {
    sgo_config cfg;
    int S, A, G, T, L, NB;        // size, actions, games, trees/game, leaf slots/game, node blocks per tree ON AVERAGE (pool = G*T*NB)
    Board *boards;                // [G]
    Board *leaf_boards;           // [G*L]
    LeafRef *leaf_refs;           // [G*L]
    int32_t *leaf_count;          // [G] leaves selected by the last select
    NodeBlock *arena;             // [pool_blocks] ONE pool of node blocks shared by every tree (a tree is a linked set of block ids)
    int32_t *free_list;           // [pool_blocks] stack of free block ids
    int32_t *pool_ctl;            // [4] free_top, low-water mark of free_top, failed allocations, pad
    long long pool_blocks;
    NodeBlock *stage;             // staging buffer of tree download / upload (grown on demand)
    int stage_blocks;
    int32_t *zero_sel;            // [G] zeros: "tree 0 of every game"
    TreeMeta *meta;               // [G*T]
    double *root_p64;             // [G*T][APAD]
    int32_t *wave;                // [G][8] mode-B wave state: energy_left, pre_bp, head, tail, stalled
    int32_t *err_flags;           // [1] sticky device error bits
    int32_t *counters;            // [4] device scratch counters
    int32_t *h_pinned;            // [16] pinned host mirror
    float *step_policy, *step_value;   // [G*L][A], [G*L] evaluator outputs of the current step (driver.cu, lazily allocated)
    int32_t *step_index, *step_sym;    // [G*L] compacted leaf slots awaiting evaluation and their symmetry ids
    struct sgo_tower *tower[2];   // network weight slots (tower.cu): 0 = model1/best, 1 = model2/tested
    unsigned long long launches;  // kernels launched through the ABI (bench.py gpu_launches)
    std::string last_error;
};

#define SGO_CUDA_OK(e, call)                                                              \
    do {                                                                                  \
        cudaError_t _err = (call);                                                        \
        if (_err != cudaSuccess) {                                                        \
            (e)->last_error = std::string(#call) + ": " + cudaGetErrorString(_err);       \
            return -2;                                                                    \
        }                                                                                 \
    } while (0)

#define SGO_LAUNCHED(e)                              \
    do {                                             \
        (e)->launches++;                             \
        SGO_CUDA_OK(e, cudaGetLastError());          \
    } while (0)

static inline int sgo_fail(sgo_engine *e, const char *msg, int code = -1)
{
    e->last_error = msg;
    return code;
}

// The node pool as the kernels see it.  Allocation (pool_pop) happens only in the select / new-tree kernels and
// release (pool_push) only in the re-root / free kernels, never in the same launch: a pop and a push racing on the
// same stack slot would hand out an id before it is written.
struct Pool {
    NodeBlock *blk;
    int32_t *free_list;
    int32_t *ctl;                 // [0] free_top, [1] low-water mark, [2] failed allocations
};
static inline Pool sgo_pool(sgo_engine *e) { Pool p; p.blk = e->arena; p.free_list = e->free_list; p.ctl = e->pool_ctl; return p; }

#ifdef __CUDACC__
// takes n ids off the stack (one thread): ids are free_list[base - 1 - i], i = 0..n-1; returns base or -1 when exhausted
__device__ __forceinline__ int pool_pop(const Pool &p, int n)
{
    int old = atomicSub(&p.ctl[0], n);
    if (old < n) { atomicAdd(&p.ctl[0], n); atomicAdd(&p.ctl[2], 1); return -1; }
    atomicMin(&p.ctl[1], old - n);
    return old;
}
__device__ __forceinline__ void pool_push(const Pool &p, int b)
{
    int pos = atomicAdd(&p.ctl[0], 1);
    p.free_list[pos] = b;
}
#endif

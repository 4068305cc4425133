// engine.h — host-side engine object behind the C ABI (include/sejonggo_b200.h).
#pragma once
#include "common.cuh"
#include "../../include/sejonggo_b200.h"
#include <string>

struct sgo_tower;

struct sgo_engine {
    sgo_config cfg;
    int S, A, G, T, L, NB;        // size, actions, games, trees/game, leaf slots/game, blocks/arena half
    Board *boards;                // [G]
    Board *leaf_boards;           // [G*L]
    LeafRef *leaf_refs;           // [G*L]
    int32_t *leaf_count;          // [G] leaves selected by the last select
    NodeBlock *arena;             // [G*T][2][NB]
    TreeMeta *meta;               // [G*T]
    double *root_p64;             // [G*T][APAD]
    int32_t *wave;                // [G][8] mode-B wave state: energy_left, pre_bp, head, tail, stalled
    int32_t *err_flags;           // [1] sticky device error bits
    int32_t *counters;            // [4] device scratch counters
    int32_t *h_pinned;            // [8] pinned host mirror
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

__host__ __device__ static inline NodeBlock *tree_arena(NodeBlock *arena, int NB, int tree, int side)
{
    return arena + ((size_t)tree * 2 + side) * NB;
}

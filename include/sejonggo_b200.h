/*
 * sejonggo_b200.h — C ABI of the B200-native batched self-play engine.
 *
 * The reference (drsagitn/sejonggo) is pure Python and has no FFI: its "plugin
 * interface" for this path is Python duck typing.  This header
 * is therefore the NEW boundary a maintainer binds with ctypes;
 * each entry cites the reference function(s) it replaces.
 *
 * Conventions
 *   - every entry returns 0 on success, <0 on error; sgo_last_error() gives text.
 *   - one sgo_engine per GPU; entries are not re-entrant per engine.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); work is
 *     enqueued asynchronously unless the entry name ends in _sync or copies to
 *     host memory.
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers.
 *   - moves are action indices a = y*S + x, pass = S*S (play.py:31-37).
 *   - no torch types, no C++ types, no hidden global state.
 */
#ifndef SEJONGGO_B200_H
#define SEJONGGO_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgo_engine sgo_engine;

typedef struct sgo_config {
    int32_t device;          /* CUDA device ordinal */
    int32_t size;            /* board size S (conf['SIZE']), 2..19 */
    int32_t n_games;         /* concurrent games G held in HBM */
    int32_t trees_per_game;  /* 1 self-play (shared tree, Q16), 2 match play */
    int32_t max_leaves;      /* leaves per game per step: MCTS_BATCH_SIZE (mode A) or ENERGY (mode B) */
    int32_t arena_blocks;    /* node blocks per tree ON AVERAGE: all trees share one pool of n_games*trees_per_game*arena_blocks */
    float komi;              /* conf['KOMI'] */
    int32_t reserved[9];
} sgo_config;

/* sizes of the device structs, for hosts that read raw arenas (tests) */
#define SGO_BOARD_BYTES 1296
#define SGO_NODEBLOCK_BYTES 6272
#define SGO_APAD 384

int sgo_create(const sgo_config *cfg, sgo_engine **out);
int sgo_destroy(sgo_engine *e);
const char *sgo_last_error(sgo_engine *e);
/* reads and clears the sticky device error flags (synchronises `stream`) */
int sgo_check_errors_sync(sgo_engine *e, void *stream, int32_t *h_flags);
int sgo_abi_version(void);
/* number of kernels this engine has launched so far (bench.py's gpu_launches) */
int64_t sgo_launch_count(sgo_engine *e);

/* ---- rules engine: play.py ------------------------------------------------ */
/* play.py:295-299 game_init for games [first, first+n) */
int sgo_games_reset(sgo_engine *e, int32_t first, int32_t n, void *stream);
/* play.py:226-242 make_play: d_moves[i] applied to game first+i; move <0 = skip;
 * d_colors NULL or per-game colour (0 = side to move) */
int sgo_apply_moves(sgo_engine *e, int32_t first, int32_t n, const int32_t *d_moves, const int32_t *d_colors, void *stream);
/* play.py:71-104 legal_moves: d_mask uint8 [n][S*S+1], 1 = illegal */
int sgo_legal_masks(sgo_engine *e, int32_t first, int32_t n, uint8_t *d_mask, void *stream);
/* play.py:274-292 get_winner: d_out int32 [n][3] = winner(+1/0/-1), black points, white stones+territory (komi not added) */
int sgo_score(sgo_engine *e, int32_t first, int32_t n, int32_t *d_out, void *stream);
/* reference tensor <-> packed state: int32 [n][S][S][17] (play.py:295 layout) */
int sgo_import_boards(sgo_engine *e, int32_t first, int32_t n, const int32_t *d_boards, void *stream);
int sgo_export_boards(sgo_engine *e, int32_t first, int32_t n, int32_t *d_boards, void *stream);
/* comparison format shared with the oracle: uint32 [n][16*ceil(S*S/32)+1] */
int sgo_export_packed(sgo_engine *e, int32_t which /*0 games,1 leaves*/, int32_t first, int32_t n, uint32_t *d_out, void *stream);
/* whole random-legal playouts on device.  d_moves int16
 * [n][max_plies] (-1 padded), d_nplies int32 [n].  Uniform legal non-pass move
 * from splitmix64(seed, game, ply); pass iff none; stop at 2 passes / max_plies. */
int sgo_random_playouts(sgo_engine *e, int32_t first, int32_t n, uint64_t seed, int32_t max_plies,
                        int16_t *d_moves, int32_t *d_nplies, void *stream);
/* self_play.py input planes for the evaluator protocol: float32 [n][S][S][17] of
 * leaf slots (which=1) or games (which=0), with symmetry.py's board transform
 * `sym` (0..7; or per-position ids d_sym[n] when non-NULL — the reference draws one
 * symmetry per predict batch = per game per step) fused into the gather (symmetry.py:45-114) */
int sgo_export_planes(sgo_engine *e, int32_t which, int32_t first, int32_t n, int32_t sym, const int32_t *d_sym, float *d_out, void *stream);
/* the same for an arbitrary list of positions: d_index int32 [n] (games or leaf slots), d_sym int32 [n] or NULL;
 * the evaluator-protocol path (model.predict_on_batch on host planes) copies only what it asks for */
int sgo_export_planes_indexed(sgo_engine *e, int32_t which, const int32_t *d_index, int32_t n, const int32_t *d_sym,
                              float *d_out, void *stream);
/* symmetry.py "reverse" policy gather (same map, quirk Q8): float32 [n][S*S+1], out of place */
int sgo_policy_unsym(sgo_engine *e, int32_t n, int32_t sym, const int32_t *d_sym, const float *d_in, float *d_out, void *stream);

/* ---- MCTS: play.py:308-421, self_play.py:28-152, tree_util.py, nomodel_self_play.py:40-140 ---- */
/* d_tree_sel int32 [G]: tree index (0..trees_per_game-1) each game searches with; NULL = 0 */
/* play.py:376-421 new_tree for games whose selected tree is invalid (or all if force):
 * d_policy f32 [G][S*S+1]; d_noise f64 [G][S*S+1] or NULL (add_noise); eps = DIRICHLET_EPSILON */
int sgo_tree_new(sgo_engine *e, const int32_t *d_tree_sel, const float *d_policy, const double *d_noise,
                 double eps, int32_t force, void *stream);
int sgo_tree_reset(sgo_engine *e, void *stream);   /* all trees -> None, every node block back in the pool */
/* drop every tree of the games with d_game_mask[g] != 0 (int32 [G]) and return their blocks to the pool */
int sgo_tree_free(sgo_engine *e, const int32_t *d_game_mask, void *stream);
/* selfplay_worker.py:81-124 — a worker starts its next game as soon as one ends: for the games with
 * d_game_mask[g] != 0, game_init (play.py:295-299) + drop their trees, leaving every other game untouched */
int sgo_games_restart(sgo_engine *e, const int32_t *d_game_mask, void *stream);
/* node blocks each tree owns: d_out int32 [G*trees_per_game] */
int sgo_tree_sizes(sgo_engine *e, int32_t *d_out, void *stream);
/* node pool occupancy: h_out[4] = capacity in blocks, free now, fewest free ever seen, failed allocations (synchronises) */
int sgo_pool_stats_sync(sgo_engine *e, int64_t *h_out, void *stream);
/* self_play.py:28-66 simulate (mode A) select: descend by top_n[0], take top-`batch`
 * at the frontier, greedy top_one_action below expanded ones; fills the leaf slots */
int sgo_tree_select_a(sgo_engine *e, const int32_t *d_tree_sel, int32_t batch, void *stream);
/* tree_util.py:4-24 + nomodel_self_play.py:59-75 (mode B): `energy` sequential
 * busy-flag selections per game; restart=1 begins a wave, 0 resumes after a stall.
 * h_counts[0] = leaves newly selected, h_counts[1] = games stalled (synchronises). */
int sgo_tree_select_b_sync(sgo_engine *e, const int32_t *d_tree_sel, int32_t energy, int32_t restart,
                           int32_t *h_counts, void *stream);
/* play.py:391-421 new_subtree for every selected leaf: d_policy f32 [G*max_leaves][S*S+1],
 * d_value f32 [G*max_leaves] (evaluator outputs, leaf-slot order) */
int sgo_tree_expand(sgo_engine *e, const int32_t *d_tree_sel, const float *d_policy, const float *d_value, void *stream);
/* self_play.py:95-116 backup in rank order (mode A) */
int sgo_tree_backup_a(sgo_engine *e, const int32_t *d_tree_sel, void *stream);
/* nomodel_self_play.py:40-56,80-82 back_propagation FIFO (mode B, end of wave) */
int sgo_tree_backup_b(sgo_engine *e, const int32_t *d_tree_sel, int32_t total_energy, void *stream);
/* self_play.py:138-152: temperature 0 -> max (count, mean, index); temperature 1 ->
 * np.random.choice(moves, p=count/total) from ONE uniform d_u01[g] in [0,1) per game, computed as numpy does (fp64 p,
 * sequential cumulative sum, cdf /= cdf[-1], first cdf > u); d_forced int32 [G] >= 0 overrides (injected choice).
 * d_move_out int32 [G] */
int sgo_tree_pick(sgo_engine *e, const int32_t *d_tree_sel, const int32_t *d_temperature, const double *d_u01,
                  const int32_t *d_forced, int32_t *d_move_out, void *stream);
/* self_play.py:223-238: cut every tree of each game to the child `d_moves[g]` (<0 = skip) */
int sgo_tree_reroot(sgo_engine *e, const int32_t *d_moves, void *stream);
/* root children: d_prior f64 [G][S*S+1] (0 where no child = policy_target, self_play.py:203-205),
 * d_count int32, d_value f32; any may be NULL */
int sgo_tree_child_stats(sgo_engine *e, const int32_t *d_tree_sel, double *d_prior, int32_t *d_count, float *d_value, void *stream);
/* raw access for tests / checkpointing: a tree to/from host memory in canonical form — block 0 is the root, the other
 * blocks follow in breadth-first slot order, child / parent links are indices into that array; h_meta[0] = 0 */
int sgo_tree_download_sync(sgo_engine *e, int32_t tree, void *h_blocks, int32_t max_blocks, void *h_meta /*8 x int32*/, double *h_root_p64);
int sgo_tree_upload_sync(sgo_engine *e, int32_t tree, const void *h_blocks, int32_t n_blocks, const void *h_meta, const double *h_root_p64);
/* leaf bookkeeping: int32 [G] leaves selected in the last select call */
int sgo_leaf_counts(sgo_engine *e, int32_t *d_counts, void *stream);
/* device list of leaf slots awaiting evaluation (unordered); h_count = how many (synchronises) */
int sgo_leaf_compact_sync(sgo_engine *e, int32_t *d_index, int32_t *h_count, void *stream);
int sgo_tree_valid(sgo_engine *e, const int32_t *d_tree_sel, int32_t *d_valid, void *stream);

/* ---- network: model.py:37-96 (Keras/TF1.7 residual tower; the arithmetic lives in that
 * third-party dependency, so parity is against an fp32 restatement — DESIGN.md) ---- */
typedef struct sgo_tower_weights {
    int32_t n_blocks;          /* conf['N_RESIDUAL_BLOCKS'] */
    int32_t channels;          /* 256 */
    int32_t size;              /* board size S; the tower runs on (S-2)x(S-2) after the valid stem conv (Q11) */
    int32_t reserved;
    /* all DEVICE pointers; BatchNorm already folded into weights/biases by the host */
    const float *stem_w;       /* [3][3][17][C]            model.py:57-59 kernel layout (kh,kw,in,out) */
    const float *stem_b;       /* [C] */
    const uint16_t *conv_w;    /* bf16 [2*n_blocks][3][3][C out][C in] */
    const float *conv_b;       /* [2*n_blocks][C] */
    const float *pol_conv_w;   /* [C][2]                   model.py:73 */
    const float *pol_conv_b;   /* [2] */
    const float *pol_fc_w;     /* [2*(S-2)^2][S*S+1]       model.py:80, rows in HWC flatten order */
    const float *pol_fc_b;     /* [S*S+1] */
    const float *val_conv_w;   /* [C][2]                   model.py:83 */
    const float *val_conv_b;   /* [2] */
    const float *val_fc1_w;    /* [2*(S-2)^2][256]         model.py:90 */
    const float *val_fc1_b;    /* [256] */
    const float *val_fc2_w;    /* [256]                    model.py:91 */
    const float *val_fc2_b;    /* [1] */
} sgo_tower_weights;
/* copies the weights into engine-owned HBM and sizes activations for max_positions per forward */
int sgo_tower_load_weights(sgo_engine *e, int32_t slot, const sgo_tower_weights *w, int32_t max_positions, void *stream);
int sgo_tower_free(sgo_engine *e, int32_t slot);
/* model.predict_on_batch (self_play.py:70,187) + symmetry.random_symmetry_predict (symmetry.py:127-132):
 * n positions taken from games (which=0) or leaf slots (which=1) by d_index (NULL = 0..n-1), per-position
 * symmetry ids d_sym (NULL = identity); d_policy f32 [..][S*S+1] softmax, d_value f32 [..] tanh; rows are
 * compact (i) or scattered to d_index[i] when scatter != 0 */
int sgo_tower_forward(sgo_engine *e, int32_t slot, int32_t which, const int32_t *d_index, int32_t n,
                      const int32_t *d_sym, int32_t scatter, float *d_policy, float *d_value, void *stream);
/* positions one forward of this slot can take (0 = no weights loaded) */
int sgo_tower_max_positions(sgo_engine *e, int32_t slot);
/* sticky tower error flags (16 = an mbarrier wait timed out); synchronises */
int sgo_tower_check_sync(sgo_engine *e, int32_t slot, int32_t *h_flags, void *stream);
/* live CUDA-event timing of the tower kernels on the launching stream (bench.py roofline):
 * h_out[6] = ms in stem, ms in conv layers, ms in heads, conv launches, positions, forward calls */
int sgo_tower_profile(sgo_engine *e, int32_t slot, int32_t enable);
int sgo_tower_profile_read_sync(sgo_engine *e, int32_t slot, double *h_out);
/* test / profiling hooks: one tensor-core conv layer in isolation; raw activation buffers
 * bf16 [n*(S-1)+1][S-1][C] (row 0 and every (S-1)-th row, and the last pixel of every row, are zero padding) */
int sgo_tower_debug_conv(sgo_engine *e, int32_t slot, int32_t n, int32_t layer, int32_t in, int32_t out, int32_t skip, void *stream);
int sgo_tower_act_copy(sgo_engine *e, int32_t slot, int32_t buf, int32_t n, void *d_data, int32_t to_tower, void *stream);

/* ---- the per-step driver: one MCTS search step of EVERY game in one call ------------------
 * mode 0 = self_play.simulate (self_play.py:28-120): select top-`leaves` + greedy descents,
 *          one evaluation of all leaves, expand + backup in rank order;
 * mode 1 = one nomodel_self_play.async_simulate2 wave (nomodel_self_play.py:59-82): `leaves`
 *          (= ENERGY) busy-flag selections, evaluation, expand, resumed while any game reports
 *          "No best leaf", then the FIFO back-propagation of `total_energy` results.
 * Leaves are evaluated by the network slot d_model_of_game[g] (NULL = slot 0 for every game)
 * with symmetry ids (symmetry.random_symmetry_predict; NULL = identity): d_sym[g], one draw per game
 * per batch as in mode A (self_play.py:70), or with sym_per_leaf != 0 d_sym[g*max_leaves + l], one draw
 * per predict request as in mode B (predicting_queue_worker.py:88-92).
 * *h_leaves_done = simulations performed (synchronises). */
int sgo_selfplay_step(sgo_engine *e, int32_t mode, const int32_t *d_tree_sel, const int32_t *d_model_of_game,
                      int32_t leaves, int32_t total_energy, const int32_t *d_sym, int32_t sym_per_leaf,
                      int32_t *h_leaves_done, void *stream);
/* self_play.py:203-214 move_data for every game, packed: uint32 [G][sgo_record_words()] =
 * board (16 x ceil(S*S/32) plane words + to_move), move index, value (f32 bits), tree-valid
 * flag, policy_target = root priors as f32 [S*S+1] (Q14).  d_moves / d_values may be NULL. */
int sgo_record_words(sgo_engine *e);
int sgo_records_pack(sgo_engine *e, const int32_t *d_tree_sel, const int32_t *d_moves, const float *d_values,
                     uint32_t *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif

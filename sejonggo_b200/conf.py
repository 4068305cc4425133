"""The conf keys the hot path reads (reference conf.py:3-106), same names and
defaults.  Unlike the reference nothing is frozen at import: every entry point
takes explicit arguments and falls back to this dict at call time.  Hosts,
credentials and directory names of the reference's conf.py are deliberately absent."""

conf = {
    'GPUs': [0, 1, 2, 3, 4, 5, 6, 7],
    'N_RESIDUAL_BLOCKS': 20,
    'N_GAMES': 5000,
    'MCTS_SIMULATIONS': 1600,
    'N_GAME_PROCESS': 32,
    'ENERGY': 8,
    'SIZE': 19,
    'KOMI': 5.5,
    'STOP_EXPLORATION': 30,
    'MCTS_BATCH_SIZE': 100,
    'DIRICHLET_ALPHA': .03,
    'DIRICHLET_EPSILON': .25,
    'RESIGNATION_PERCENT': .10,
    'RESIGNATION_ALLOWED_ERROR': .05,
    'EVALUATE_N_GAMES': 100,
    'EVALUATE_MARGIN': .55,
    'PREDICTING_BATCH_SIZE': 32,
    'SELF_PLAY_DIR': 'sp_self_play_data',
    'MODEL_DIR': 'sp_models',
    'EVAL_DIR': 'sp_eval_games',
    'GAMES_DIR': 'sp_eval_games',
    'SGF_ENABLED': False,
    'BEST_MODEL': 'best_model.npz',      # the reference's best_model.h5 (Keras .h5 I/O is out of scope)
    # engine-only knobs
    'CONCURRENT_GAMES': 1024,
    'ARENA_BLOCKS': None,
}

"""Mirror of the reference's self_play.py (mode A) on the batched GPU engine.
Same function names, positional order, defaults and game_data keys
(self_play.py:164, 293, 343); `play_games` is the batched form they wrap."""
from random import random
import numpy as np

from .conf import conf
from .batched import BatchedGames, HostRng


def play_games(model1, model2, n_games, mcts_simulations, stop_exploration, self_play=False, num_moves=None,
               resign_model1=None, resign_model2=None, size=None, mcts_batch_size=None, rng=None, rngs=None,
               record_boards='full', use_symmetry=True, arena_blocks=None, device=0):
    if mcts_simulations is None:
        mcts_simulations = conf['MCTS_SIMULATIONS']
    bg = BatchedGames((model1, model2), n_games, size=size or conf['SIZE'], mode='a',
                      mcts_batch_size=mcts_batch_size or conf['MCTS_BATCH_SIZE'], mcts_simulations=mcts_simulations,
                      stop_exploration=stop_exploration, self_play=self_play, num_moves=num_moves,
                      resign=(resign_model1, resign_model2), komi=conf['KOMI'], dirichlet_eps=conf['DIRICHLET_EPSILON'],
                      use_symmetry=use_symmetry, rng=rng, rngs=rngs, arena_blocks=arena_blocks or conf['ARENA_BLOCKS'],
                      device=device, record_boards=record_boards)
    return bg.run()


def play_game(model1, model2, mcts_simulations, stop_exploration, self_play=False, num_moves=None,
              resign_model1=None, resign_model2=None, **kw):
    """self_play.py:164 — one game (a batch of 1)."""
    return play_games(model1, model2, 1, mcts_simulations, stop_exploration, self_play, num_moves,
                      resign_model1, resign_model2, **kw)[0]


def _calibrated_self_play(model, n_games, mcts_simulations, concurrent, **kw):
    """self_play.py:343-378 / 293-340: n games with the resignation calibration —
    RESIGNATION_PERCENT of the games play without resignation and their winners'
    minimum values set the threshold.  Games run `concurrent` at a time; the threshold
    is updated between batches (the reference updates it between single games)."""
    games_data, min_values, current_resign = [], [], None
    done = 0
    while done < n_games:
        n = min(concurrent, n_games - done)
        resign = np.array([np.nan if (random() <= conf['RESIGNATION_PERCENT'] or current_resign is None)
                           else current_resign for _ in range(n)])
        batch = play_games(model, model, n, mcts_simulations, conf['STOP_EXPLORATION'], self_play=True,
                           resign_model1=resign, resign_model2=resign, **kw)
        for g, gd in enumerate(batch):
            if np.isnan(resign[g]) and gd['moves']:
                mv = gd['moves'][::2] if gd['winner'] == 1 else gd['moves'][1::2]
                if mv:
                    min_values.append(min(float(m['value']) for m in mv))
                idx = int(conf['RESIGNATION_ALLOWED_ERROR'] * len(min_values))
                if idx > 0:
                    current_resign = min_values[idx]
        games_data.extend(batch)
        done += n
    return games_data


def self_play(model, n_games, mcts_simulations, concurrent=None, **kw):
    return _calibrated_self_play(model, n_games, mcts_simulations, concurrent or conf['CONCURRENT_GAMES'], **kw)


def model_self_play(model, one_game_only=-1, concurrent=None, **kw):
    n = 1 if one_game_only >= 0 else conf['N_GAMES']
    return _calibrated_self_play(model, n, conf['MCTS_SIMULATIONS'], concurrent or conf['CONCURRENT_GAMES'], **kw)

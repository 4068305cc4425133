"""Mirror of the reference's self_play.py (mode A) on the batched GPU engine.
Same function names, positional order, defaults and game_data keys
(self_play.py:164, 293, 343); `play_games` is the batched form they wrap."""
import os
from random import random

from .conf import conf
from .batched import BatchedGames
from .sgfsave import save_game_data, save_self_play_data


def play_games(model1, model2, n_games, mcts_simulations, stop_exploration, self_play=False, num_moves=None,
               resign_model1=None, resign_model2=None, size=None, mcts_batch_size=None, rng=None, rngs=None,
               record_boards='full', use_symmetry=True, arena_blocks=None, device=0, concurrent=None,
               on_game_start=None, on_game_end=None, rng_for_game=None):
    """n_games games, `concurrent` of them (default: all) in HBM at a time; a slot starts its next game as soon
    as one ends.  Returns the game_data list in game order."""
    if mcts_simulations is None:
        mcts_simulations = conf['MCTS_SIMULATIONS']
    slots = min(n_games, concurrent or n_games)
    bg = BatchedGames((model1, model2), slots, size=size or conf['SIZE'], mode='a',
                      mcts_batch_size=mcts_batch_size or conf['MCTS_BATCH_SIZE'], mcts_simulations=mcts_simulations,
                      stop_exploration=stop_exploration, self_play=self_play, num_moves=num_moves,
                      resign=(resign_model1, resign_model2), komi=conf['KOMI'], dirichlet_eps=conf['DIRICHLET_EPSILON'],
                      use_symmetry=use_symmetry, rng=rng, rngs=rngs, arena_blocks=arena_blocks or conf['ARENA_BLOCKS'],
                      device=device, record_boards=record_boards, n_total=n_games, on_game_start=on_game_start,
                      on_game_end=on_game_end, rng_for_game=rng_for_game)
    return bg.run()


def play_game(model1, model2, mcts_simulations, stop_exploration, self_play=False, num_moves=None,
              resign_model1=None, resign_model2=None, **kw):
    """self_play.py:164 — one game (a batch of 1)."""
    return play_games(model1, model2, 1, mcts_simulations, stop_exploration, self_play, num_moves,
                      resign_model1, resign_model2, **kw)[0]


class ResignationCalibrator(object):
    """The resignation bookkeeping the reference repeats in self_play (self_play.py:343-378), model_self_play
    (:293-340) and NoModelSelfPlayWorker.run (selfplay_worker.py:76-112): at the START of a game a lottery decides
    whether it may resign (`random() > RESIGNATION_PERCENT` -> the current threshold, which is None until enough
    games have been seen); at its END a game that played without a threshold contributes the minimum root value of
    the winner's plies (black = even plies when winner == 1, else the odd plies), and the threshold becomes
    min_values[int(RESIGNATION_ALLOWED_ERROR * len(min_values))] — an index into the UNSORTED list, as the reference
    does it.  With one game in flight this is the reference's sequence exactly; with many, a game sees the
    threshold as of the moment it starts."""

    def __init__(self, percent=None, allowed_error=None, rand=random):
        self.percent = conf['RESIGNATION_PERCENT'] if percent is None else percent
        self.allowed_error = conf['RESIGNATION_ALLOWED_ERROR'] if allowed_error is None else allowed_error
        self.rand = rand
        self.current_resign = None
        self.min_values = []
        self.resign_of = {}

    def start(self, game_id):
        resign = self.current_resign if self.rand() > self.percent else None
        self.resign_of[game_id] = resign
        return resign, resign

    def end(self, game_id, game_data):
        if self.resign_of.get(game_id) is not None:
            return
        moves = game_data['moves'][::2] if game_data['winner'] == 1 else game_data['moves'][1::2]
        if not moves:
            return                       # (the reference's min() of an empty list would raise)
        self.min_values.append(min(m['value'] for m in moves))
        idx = int(self.allowed_error * len(self.min_values))
        if idx > 0:
            self.current_resign = self.min_values[idx]


def self_play(model, n_games, mcts_simulations, concurrent=None, record_boards='packed', save=True, rand=random,
              calibrator=None, **kw):
    """self_play.py:343-378: n self-play games with the resignation calibration, each saved through
    save_game_data(model.name, game, game_data) as it finishes."""
    cal = calibrator or ResignationCalibrator(rand=rand)
    size = kw.get('size') or conf['SIZE']

    def on_end(game, gd):
        cal.end(game, gd)
        if save:
            save_game_data(model.name, game, gd, size=size)

    return play_games(model, model, n_games, mcts_simulations, conf['STOP_EXPLORATION'], self_play=True,
                      concurrent=concurrent or conf['CONCURRENT_GAMES'], record_boards=record_boards,
                      on_game_start=cal.start, on_game_end=on_end, **kw)


def model_self_play(model, one_game_only=-1, concurrent=None, record_boards='packed', rand=random, calibrator=None, **kw):
    """self_play.py:293-340: conf['N_GAMES'] games of conf['MCTS_SIMULATIONS'], skipping games whose directory
    SELF_PLAY_DIR/<model>/game_%05d already exists (resume; a directory is claimed by creating it), each saved
    through save_self_play_data as it finishes."""
    root = os.path.join(conf['SELF_PLAY_DIR'], model.name)
    size = kw.get('size') or conf['SIZE']
    todo = [g for g in range(conf['N_GAMES']) if (one_game_only < 0 or g == one_game_only)
            and not os.path.isdir(os.path.join(root, "game_%05d" % g))]
    cal = calibrator or ResignationCalibrator(rand=rand)

    def on_start(i):
        try:
            os.makedirs(os.path.join(root, "game_%05d" % todo[i]))
        except OSError:
            return False                 # someone else took it meanwhile
        return cal.start(i)

    def on_end(i, gd):
        cal.end(i, gd)
        gd['game'] = todo[i]
        os.rmdir(os.path.join(root, "game_%05d" % todo[i]))       # save re-creates it with the move directories
        save_self_play_data(model.name, todo[i], gd, size=size)

    if not todo:
        return []
    return play_games(model, model, len(todo), conf['MCTS_SIMULATIONS'], conf['STOP_EXPLORATION'], self_play=True,
                      concurrent=concurrent or conf['CONCURRENT_GAMES'], record_boards=record_boards,
                      on_game_start=on_start, on_game_end=on_end, **kw)

"""Batched stand-in for the reference's NoModelSelfPlayWorker (selfplay_worker.py:61-130) and the
main_selfplay.py launcher (main_selfplay.py:9-29).

ONE worker per GPU keeps `concurrent` games in HBM; a slot starts the next game the moment one ends, which is the
reference's per-process `for game in games` loop (selfplay_worker.py:81-124) run for all slots at once.  Kept from
the reference: resume by skipping game directories that already exist (:83-89) and claiming one by creating it (so
several workers / GPUs / ranks can share SELF_PLAY_DIR), the resignation calibration (:91-112), dropping empty
games (:115-118) and the per-game save through sgfsave.save_self_play_data.  scp sync is out of scope.
"""
import os
from random import random


from .conf import conf
from . import predicting_queue_worker as pq
from .nomodel_self_play import play_games_async
from .self_play import ResignationCalibrator
from .sgfsave import save_self_play_data


def run_selfplay(n_games=None, concurrent=None, energy=None, device=0, size=None, record_boards='packed', games=None,
                 rand=random, calibrator=None, **kw):
    """Plays the still-unclaimed games among `games` (default range(N_GAMES)); returns the game numbers saved."""
    n_games = conf['N_GAMES'] if n_games is None else n_games
    concurrent = concurrent or conf['CONCURRENT_GAMES']
    energy = energy or conf['ENERGY']
    model_name = pq.put_name_request("BEST_NAME")
    root = os.path.join(conf['SELF_PLAY_DIR'], model_name)
    todo = [g for g in (range(n_games) if games is None else games) if not os.path.isdir(os.path.join(root, "game_%05d" % g))]
    if not todo:
        return []
    cal = calibrator or ResignationCalibrator(rand=rand)
    saved = []

    def on_start(i):
        try:
            os.makedirs(os.path.join(root, "game_%05d" % todo[i]))
        except OSError:
            return False                                          # another worker took it
        return cal.start(i)

    def on_end(i, gd):
        cal.end(i, gd)
        os.rmdir(os.path.join(root, "game_%05d" % todo[i]))       # save re-creates it with the move dirs; empty games leave nothing
        if not gd['moves']:
            return
        save_self_play_data(model_name, todo[i], gd, size=size or conf['SIZE'])
        saved.append(todo[i])

    play_games_async("BEST_SYM", "BEST_SYM", len(todo), energy, conf['STOP_EXPLORATION'], self_play=True, size=size,
                     device=device, record_boards=record_boards, exact_rng_order=False, concurrent=min(concurrent, len(todo)),
                     on_game_start=on_start, on_game_end=on_end, **kw)
    return saved

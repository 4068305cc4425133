"""Batched stand-in for the reference's NoModelSelfPlayWorker (selfplay_worker.py:61-130):
ONE worker per GPU plays `concurrent` games at a time instead of one game per OS process.
Kept from the reference: resume by skipping game directories that already exist (:83-89),
the resignation calibration (:91-112; RESIGNATION_PERCENT of the games play without
resignation, their winners' minimum values feed the threshold), dropping empty games (:115-118)
and the per-game save through sgfsave.save_self_play_data.  scp sync is out of scope."""
import os
from random import random

import numpy as np

from .conf import conf
from . import predicting_queue_worker as pq
from .nomodel_self_play import play_games_async
from .sgfsave import save_self_play_data


def run_selfplay(n_games=None, concurrent=None, energy=None, device=0, size=None, record_boards='packed', **kw):
    n_games = conf['N_GAMES'] if n_games is None else n_games
    concurrent = concurrent or conf['CONCURRENT_GAMES']
    energy = energy or conf['ENERGY']
    model_name = pq.put_name_request("BEST_NAME")
    root = conf['SELF_PLAY_DIR']
    todo = []
    for game in range(n_games):                                   # dir-skip resume
        directory = os.path.join(root, model_name, "game_%05d" % game)
        if os.path.isdir(directory):
            continue
        todo.append(game)
    current_resign, min_values, saved = None, [], []
    for s in range(0, len(todo), concurrent):
        ids = todo[s:s + concurrent]
        claimed = []
        for g in ids:
            try:
                os.makedirs(os.path.join(root, model_name, "game_%05d" % g))
                claimed.append(g)
            except OSError:
                continue                                          # another worker took it
        if not claimed:
            continue
        resign = np.array([current_resign if (random() > conf['RESIGNATION_PERCENT'] and current_resign is not None) else np.nan
                           for _ in claimed])
        games = play_games_async("BEST_SYM", "BEST_SYM", len(claimed), energy, conf['STOP_EXPLORATION'], self_play=True,
                                 resign_model1=resign, resign_model2=resign, size=size, device=device,
                                 record_boards=record_boards, exact_rng_order=False, **kw)
        for g, r, gd in zip(claimed, resign, games):
            if np.isnan(r) and gd['moves']:
                mv = gd['moves'][::2] if gd['winner'] == 1 else gd['moves'][1::2]
                if mv:
                    min_values.append(min(float(m['value']) for m in mv))
                idx = int(conf['RESIGNATION_ALLOWED_ERROR'] * len(min_values))
                if idx > 0:
                    current_resign = min_values[idx]
            if not gd['moves']:
                os.rmdir(os.path.join(root, model_name, "game_%05d" % g))
                continue
            os.rmdir(os.path.join(root, model_name, "game_%05d" % g))   # save re-creates it with the move dirs
            save_self_play_data(model_name, g, gd, size=size or conf['SIZE'])
            saved.append(g)
    return saved

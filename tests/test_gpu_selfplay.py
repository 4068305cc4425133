"""GPU: the n-game drivers with the resignation calibration — self_play.self_play / model_self_play
(self_play.py:293-378) and the batched NoModelSelfPlayWorker (selfplay_worker.py:61-130) — against a run of the
unmodified reference with recorded draws (tests/golden/selfplay_calib_s5.npz) and through their bookkeeping."""
import os
import numpy as np
import pytest

from oracle import game_loop as gl
from oracle.fake_eval import FakeModel
from tests.conftest import GOLDEN
from tests.test_oracle_golden import _calib_checks

pytestmark = pytest.mark.gpu


def _conf(tmp_path, **kw):
    from sejonggo_b200.conf import conf
    old = dict(conf)
    conf.update(SELF_PLAY_DIR=str(tmp_path / "sp"), GAMES_DIR=str(tmp_path / "games"), MODEL_DIR=str(tmp_path / "models"))
    conf.update(kw)
    return conf, old


def test_self_play_equals_reference_run_with_calibration(tmp_path):
    """One game in flight (concurrent=1): the lottery, the thresholds and every game equal the reference's
    self_play(model, 16, 8) — thresholds appear from the 4th game on and some games resign."""
    from sejonggo_b200 import self_play as sp
    z = np.load(os.path.join(GOLDEN, "selfplay_calib_s5.npz"))
    S, batch, sims, seed, n = int(z["size"]), int(z["batch"]), int(z["sims"]), int(z["seed"]), int(z["n_games"])
    conf, old = _conf(tmp_path, SIZE=S, MCTS_BATCH_SIZE=batch, KOMI=float(z["komi"]), STOP_EXPLORATION=int(z["stop_exploration"]),
                      RESIGNATION_PERCENT=float(z["percent"]), RESIGNATION_ALLOWED_ERROR=float(z["allowed_error"]))
    try:
        lot = list(z["lottery"])
        rng = gl.ReplayRng(coin=z["coin"], noise=z["noise"], choice=z["choice"])
        games = sp.self_play(FakeModel("model_1", salt=seed, sharp=True), n, sims, concurrent=1, rand=lambda: lot.pop(0),
                             rng=rng, use_symmetry=False)
        _calib_checks(games, z)
        assert not lot
        # every game was saved as it finished (save_game_data: GAMES_DIR/<model>/game_%03d/move_%03d)
        for g, gd in enumerate(games):
            d = tmp_path / "games" / "model_1" / ("game_%03d" % g)
            assert len(os.listdir(d)) == len(gd['moves']) if gd['moves'] else not d.exists()
    finally:
        conf.clear()
        conf.update(old)


def test_self_play_concurrent_games_calibrate_as_they_finish(tmp_path):
    """Four games in flight: a game takes the threshold as of the moment it starts, so the first four play without
    one and later ones pick up values that finished no-resign games contributed."""
    from sejonggo_b200 import self_play as sp
    conf, old = _conf(tmp_path, SIZE=5, MCTS_BATCH_SIZE=4, KOMI=0.5, STOP_EXPLORATION=3, RESIGNATION_PERCENT=0.3,
                      RESIGNATION_ALLOWED_ERROR=0.34)
    try:
        cal = sp.ResignationCalibrator(rand=np.random.RandomState(5).random_sample)
        games = sp.self_play(FakeModel("model_1", salt=31, sharp=True), 20, 8, concurrent=4, calibrator=cal,
                             rng=gl.SeededRng(9), use_symmetry=False, save=False)
        assert len(games) == 20 and [g['game_id'] for g in games] == list(range(20))
        assert all(cal.resign_of[g] is None for g in range(4))
        used = [cal.resign_of[g] for g in range(20) if cal.resign_of[g] is not None]
        assert used and all(any(np.float32(u) == np.float32(m) for m in cal.min_values) for u in used)
        for g, gd in enumerate(games):
            assert gd['resign_model1'] == (None if cal.resign_of[g] is None else float(cal.resign_of[g]))
            if gd['end_reason'] == 'resign':
                assert gd['resign_model1'] is not None and gd['result'].endswith("+R")
        n_free = sum(1 for g in range(20) if cal.resign_of[g] is None and games[g]['moves'])
        assert len(cal.min_values) in (n_free, n_free - 1, n_free - 2)          # (a one-ply game won by white has no winner's ply)
    finally:
        conf.clear()
        conf.update(old)


def test_model_self_play_resumes_saves_and_one_game_only(tmp_path):
    from sejonggo_b200 import self_play as sp
    conf, old = _conf(tmp_path, SIZE=5, MCTS_BATCH_SIZE=4, MCTS_SIMULATIONS=8, KOMI=0.5, STOP_EXPLORATION=2, N_GAMES=6,
                      RESIGNATION_PERCENT=0.5, RESIGNATION_ALLOWED_ERROR=0.34)
    try:
        m = FakeModel("model_4", salt=3, sharp=True)
        root = tmp_path / "sp" / "model_4"
        os.makedirs(root / "game_00001")                                         # played elsewhere already
        one = sp.model_self_play(m, one_game_only=3, rng=gl.SeededRng(1), use_symmetry=False, num_moves=6)
        assert len(one) == 1 and one[0]['game'] == 3 and sorted(os.listdir(root)) == ["game_00001", "game_00003"]
        games = sp.model_self_play(m, concurrent=2, rng=gl.SeededRng(2), use_symmetry=False, num_moves=6)
        assert sorted(g['game'] for g in games) == [0, 2, 4, 5]
        assert sorted(os.listdir(root)) == ["game_%05d" % g for g in range(6)]
        assert os.listdir(root / "game_00001") == []
        for gd in games:
            d = root / ("game_%05d" % gd['game'])
            assert sorted(os.listdir(d)) == ["move_%03d" % k for k in range(len(gd['moves']))]
            zf = np.load(d / "move_000" / "sample.npz")
            assert zf["board"].shape == (1, 5, 5, 17) and zf["policy_target"].shape == (26,)
            last = np.load(d / ("move_%03d" % (len(gd['moves']) - 1)) / "sample.npz")
            player = gd['moves'][-1]['player']
            assert float(last["value_target"]) == (1.0 if gd['winner'] == player else -1.0)       # Q15
        assert sp.model_self_play(m, rng=gl.SeededRng(3), use_symmetry=False, num_moves=6) == []   # nothing left
    finally:
        conf.clear()
        conf.update(old)


def test_run_selfplay_worker_calibrates_and_resigns(tmp_path):
    """selfplay_worker.py:76-124 with RESIGNATION_PERCENT < 1: thresholds get set from finished games and later
    games resign on them; every saved game has its move directories."""
    from sejonggo_b200 import selfplay_worker as sw, predicting_queue_worker as pq, self_play as sp
    conf, old = _conf(tmp_path, SIZE=5, ENERGY=4, MCTS_SIMULATIONS=8, KOMI=0.5, STOP_EXPLORATION=2, N_GAMES=24,
                      RESIGNATION_PERCENT=0.2, RESIGNATION_ALLOWED_ERROR=0.34)
    try:
        m = FakeModel("model_9", salt=31, sharp=True)
        pq.register_models(best=m, latest=m)
        cal = sp.ResignationCalibrator(rand=np.random.RandomState(7).random_sample)
        saved = sw.run_selfplay(concurrent=3, size=5, calibrator=cal, rng=gl.SeededRng(4))
        assert sorted(saved) == [g for g in range(24) if os.path.isdir(tmp_path / "sp" / "model_9" / ("game_%05d" % g))]
        assert len(cal.min_values) >= 3 and cal.current_resign is not None
        assert any(v is not None for v in cal.resign_of.values())
        for g in saved:
            d = tmp_path / "sp" / "model_9" / ("game_%05d" % g)
            assert sorted(os.listdir(d))[0] == "move_000"
    finally:
        conf.clear()
        conf.update(old)

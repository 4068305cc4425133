"""GPU parity: whole games through the reference-named Python API
(sejonggo_b200.self_play.play_game / nomodel_self_play.play_game_async / evaluator)
against game fixtures recorded from the unmodified reference."""
import os
import glob
import numpy as np
import pytest

from oracle import game_loop as gl
from oracle.fake_eval import FakeModel
from tests.conftest import GOLDEN
from tests.test_oracle_golden import check_game, game_kwargs

pytestmark = pytest.mark.gpu

GAMES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "game_*.npz")))


@pytest.mark.parametrize("name", GAMES)
def test_game_fixture(name):
    from sejonggo_b200 import self_play as sp, nomodel_self_play as nsp, predicting_queue_worker as pq
    from sejonggo_b200.conf import conf
    z = np.load(os.path.join(GOLDEN, name))
    S, mode, batch, sims, seed = int(z["size"]), str(z["mode"]), int(z["batch"]), int(z["sims"]), int(z["seed"])
    kw = game_kwargs(z)
    live = "live_sym" in z.files and int(z["live_sym"])
    rng = gl.ReplayRng(coin=z["coin"], noise=z["noise"], choice=z["choice"], symmetry=z["symmetry"] if live else None)
    kind = str(z["evalkind"]) if "evalkind" in z.files else "fake"
    m1 = FakeModel("model_1", salt=seed, sharp=True, kind=kind)
    m2 = m1 if kw['self_play'] else FakeModel("model_2", salt=seed + 1, sharp=True, kind=kind)
    old = dict(conf)
    try:
        conf.update(SIZE=S, MCTS_BATCH_SIZE=batch, ENERGY=batch, MCTS_SIMULATIONS=sims, KOMI=5.5)
        if mode == 'a':
            gd = sp.play_game(m1, m2, sims, kw['stop_exploration'], kw['self_play'], kw['num_moves'],
                              kw['resign_model1'], kw['resign_model2'], rng=rng, use_symmetry=bool(live))
            calls = m1.calls + [-1] + (m2.calls if m2 is not m1 else [])
            assert calls == list(z["calls"])
        else:
            pq.register_models(best=m1, latest=m2)
            i1, i2 = ("BEST", "BEST") if kw['self_play'] else ("BEST", "LATEST")
            if live:            # the *_SYM tags: one symmetry per request, LATEST_SYM served by the best network (Q21)
                i1, i2 = ("BEST_SYM", "BEST_SYM") if kw['self_play'] else ("BEST_SYM", "LATEST_SYM")
            gd = nsp.play_game_async(i1, i2, batch, kw['stop_exploration'], 0, kw['self_play'], kw['num_moves'],
                                     kw['resign_model1'], kw['resign_model2'], rng=rng)
    finally:
        conf.clear()
        conf.update(old)
    check_game(gd, z)
    if live:
        assert len(rng._sym) == 0


def test_batched_games_match_single_games():
    """8 concurrent self-play games with per-game seeded RNGs == the same 8 games from the oracle."""
    from sejonggo_b200 import self_play as sp
    from sejonggo_b200.conf import conf
    S, batch, sims, G = 9, 8, 24, 8
    model = FakeModel("m", salt=5, sharp=True)
    old = dict(conf)
    try:
        conf.update(SIZE=S, MCTS_BATCH_SIZE=batch, KOMI=5.5)
        games = sp.play_games(model, model, G, sims, 4, self_play=True, num_moves=10,
                              rngs=[gl.SeededRng(50 + g) for g in range(G)])
    finally:
        conf.clear()
        conf.update(old)
    for g in range(G):
        ref = gl.play_game(FakeModel("m", salt=5, sharp=True), None, sims, 4, self_play=True, num_moves=10, size=S,
                           mcts_batch_size=batch, rng=gl.SeededRng(50 + g)) if False else None
    # oracle run needs model2 is model1 for self-play
    for g in range(G):
        m = FakeModel("m", salt=5, sharp=True)
        ref = gl.play_game(m, m, sims, 4, self_play=True, num_moves=10, size=S, mcts_batch_size=batch,
                           rng=gl.SeededRng(50 + g))
        got = games[g]
        assert len(got['moves']) == len(ref['moves'])
        for a, b in zip(got['moves'], ref['moves']):
            assert a['move'] == b['move'] and a['player'] == b['player']
            assert np.array_equal(a['board'], b['board'])
            assert np.array_equal(np.asarray(a['policy']).view(np.uint64), np.asarray(b['policy']).view(np.uint64))
        assert got['result'] == ref['result'] and got['winner'] == ref['winner']


def test_evaluate_runs_two_trees():
    from sejonggo_b200 import evaluator
    from sejonggo_b200.conf import conf
    old = dict(conf)
    try:
        conf.update(SIZE=5, MCTS_BATCH_SIZE=4, MCTS_SIMULATIONS=8, EVALUATE_N_GAMES=6, KOMI=5.5)
        best, tested = FakeModel("best", salt=1, sharp=True), FakeModel("tested", salt=2, sharp=True)
        games = evaluator.evaluate_games(best, tested, num_moves=12, rng=gl.SeededRng(3))
        assert len(games) == 6 and all(len(g['moves']) > 0 for g in games)
        assert isinstance(evaluator.evaluate(best, tested, num_moves=6, rng=gl.SeededRng(4)), bool)
    finally:
        conf.clear()
        conf.update(old)


def test_selfplay_worker_shim_resume_and_records(tmp_path):
    """batched NoModelSelfPlayWorker stand-in — dir-skip resume, per-ply samples."""
    from sejonggo_b200 import selfplay_worker as sw, predicting_queue_worker as pq
    from sejonggo_b200.conf import conf
    old = dict(conf)
    try:
        conf.update(SIZE=5, ENERGY=4, MCTS_SIMULATIONS=8, KOMI=5.5, SELF_PLAY_DIR=str(tmp_path / "sp"), STOP_EXPLORATION=2,
                    RESIGNATION_PERCENT=1.0)
        m = FakeModel("model_9", salt=3, sharp=True)
        pq.register_models(best=m, latest=m)
        os.makedirs(tmp_path / "sp" / "model_9" / "game_00001")           # already played elsewhere -> skipped
        saved = sw.run_selfplay(n_games=4, concurrent=2, size=5, num_moves=6)
        assert sorted(saved) == [0, 2, 3]
        for g in saved:
            d = tmp_path / "sp" / "model_9" / ("game_%05d" % g)
            moves = sorted(os.listdir(d))
            assert moves and moves[0] == "move_000"
            z = np.load(d / "move_000" / "sample.npz")
            assert z["board"].shape == (1, 5, 5, 17) and z["policy_target"].shape == (26,)
            assert float(z["board"][0, 0, 0, 16]) == 1.0 and float(z["board"][..., :16].sum()) == 0.0
    finally:
        conf.clear()
        conf.update(old)


def _same_game(got, ref):
    assert len(got['moves']) == len(ref['moves'])
    for a, b in zip(got['moves'], ref['moves']):
        assert a['move'] == b['move'] and a['player'] == b['player'] and a['move_n'] == b['move_n']
        assert np.array_equal(a['board'], b['board'])
        assert np.float32(a['value']).view(np.uint32) == np.float32(b['value']).view(np.uint32)
        assert np.array_equal(np.asarray(a['policy']).view(np.uint64), np.asarray(b['policy']).view(np.uint64))
    assert got['result'] == ref['result'] and got['winner'] == ref['winner'] and got['winner_model'] == ref['winner_model']
    assert got['modelB_name'] == ref['modelB_name'] and got['modelW_name'] == ref['modelW_name']


def test_match_play_with_random_symmetries_vs_oracle():
    """BASELINE config 4 semantics (evaluator.py:23-36): two models, two trees, stop_exploration=0, one of the
    7 symmetries drawn per predict batch (incl. the Q8 rot90/270 quirk) — engine vs oracle, 6 concurrent games."""
    from sejonggo_b200 import self_play as sp
    from sejonggo_b200.conf import conf
    S, batch, sims, G = 9, 8, 24, 6
    old = dict(conf)
    try:
        conf.update(SIZE=S, MCTS_BATCH_SIZE=batch, KOMI=5.5)
        best, tested = FakeModel("best", salt=11, sharp=True), FakeModel("tested", salt=12, sharp=True)
        games = sp.play_games(best, tested, G, sims, 0, self_play=False, num_moves=9,
                              rngs=[gl.SeededRng(70 + g) for g in range(G)], use_symmetry=True)
    finally:
        conf.clear()
        conf.update(old)
    for g in range(G):
        b, t = FakeModel("best", salt=11, sharp=True), FakeModel("tested", salt=12, sharp=True)
        ref = gl.play_game(b, t, sims, 0, self_play=False, num_moves=9, size=S, mcts_batch_size=batch, rng=gl.SeededRng(70 + g))
        _same_game(games[g], ref)


def test_mode_b_selfplay_with_symmetries_vs_oracle():
    """main_selfplay.py's path: play_game_async("BEST_SYM","BEST_SYM") — every request (root and leaves) draws a
    symmetry; engine (exact RNG order) vs oracle, 4 concurrent games on 7x7."""
    from sejonggo_b200 import nomodel_self_play as nsp, predicting_queue_worker as pq
    from sejonggo_b200.conf import conf
    from oracle.fake_eval import evaluate
    S, energy, sims, G = 7, 8, 24, 4
    old = dict(conf)
    try:
        conf.update(SIZE=S, ENERGY=energy, MCTS_SIMULATIONS=sims, KOMI=5.5)
        m = FakeModel("model_1", salt=21, sharp=True)
        pq.register_models(best=m, latest=m)
        games = nsp.play_games_async("BEST_SYM", "BEST_SYM", G, energy, 3, self_play=True, num_moves=8,
                                     rngs=[gl.SeededRng(90 + g) for g in range(G)])
    finally:
        conf.clear()
        conf.update(old)
    for g in range(G):
        mm = FakeModel("model_1", salt=21, sharp=True)
        ref = gl.play_game_async("BEST_SYM", "BEST_SYM", energy, 3, 0, self_play=True, num_moves=8, size=S, conf_sims=sims,
                                 conf_energy=energy, rng=gl.SeededRng(90 + g), names={"BEST_SYM": "model_1"},
                                 predict=lambda tag, b, sym: (lambda p, v: (p[0], v[0]))(*gl.sym_predict(mm, b, sym)))
        _same_game(games[g], ref)


def test_run_evaluation_bookkeeping(tmp_path):
    """evaluate_worker.run_evaluation: claims EVAL_DIR/<latest>/game_%03d, plays the claimed games concurrently,
    touches the winner file; eval_statistic reads it back; a second worker finds nothing left (modes A and B)."""
    from sejonggo_b200 import evaluate_worker as ew, evaluator
    from sejonggo_b200.conf import conf
    old = dict(conf)
    try:
        for mode, latest_name in (('a', "model_2"), ('b', "model_3")):
            conf.update(SIZE=9, MCTS_BATCH_SIZE=8, ENERGY=8, MCTS_SIMULATIONS=16, KOMI=5.5, EVALUATE_N_GAMES=6,
                        EVAL_DIR=str(tmp_path / "eval"), GAMES_DIR=str(tmp_path / "games"), MODEL_DIR=str(tmp_path / "models"))
            best, latest = FakeModel("model_1", salt=1, sharp=True), FakeModel(latest_name, salt=2, sharp=True)
            os.makedirs(os.path.join(conf['EVAL_DIR'], latest_name, "game_004"))          # taken by another worker
            wins, total = ew.run_evaluation(best, latest, concurrent=4, mode=mode, num_moves=12)
            assert total == 5 and 0 <= wins <= 5
            stat = evaluator.eval_statistic()
            assert abs(stat[latest_name] - wins / 6.0) < 1e-12                             # game_004 counts as played, not won
            for g in (0, 1, 2, 3, 5):
                files = os.listdir(os.path.join(conf['EVAL_DIR'], latest_name, "game_%03d" % g))
                assert len(files) == 1 and files[0] in ("model_1", latest_name, "None")
            assert ew.run_evaluation(best, latest, concurrent=4, mode=mode, num_moves=12) == (0, 0)
            if mode == 'b':                                                                # eval games double as training data
                assert os.path.isdir(os.path.join(conf['GAMES_DIR'], latest_name, "eval_game_000", "move_000"))
                assert len(latest.calls) > 0            # the candidate network really was evaluated (LATEST_SYM -> latest model)
        # ... unless the reference's Q21 is asked for: LATEST_SYM is then served by the BEST network and `latest` is never called
        conf.update(EVAL_DIR=str(tmp_path / "eval_q21"))
        best, latest = FakeModel("model_1", salt=1, sharp=True), FakeModel("model_4", salt=2, sharp=True)
        wins, total = ew.run_evaluation(best, latest, concurrent=3, n_games=3, mode='b', num_moves=6, reference_q21=True)
        assert total == 3 and len(latest.calls) == 0 and len(best.calls) > 0
        assert ew.run_evaluation(best, best, mode='a') == (0, 0)                           # "No new trained model"
    finally:
        conf.clear()
        conf.update(old)


def test_gtp_engine_matches_oracle_search():
    """sejonggo.GTPEngine (batch of one): genmove = root eval + new tree + sims/B simulate steps + T=0 pick, with the
    tree re-rooted by play/genmove — every generated move equals the oracle's mode-A search on the same position."""
    from sejonggo_b200 import sejonggo as gtp
    from sejonggo_b200.conf import conf
    from oracle import oracle as o
    S, B, SIMS = 9, 8, 32
    old = dict(conf)
    try:
        conf.update(SIZE=S, MCTS_BATCH_SIZE=B, MCTS_SIMULATIONS=SIMS, KOMI=5.5)
        model = FakeModel("model_7", salt=11, sharp=True)
        g = gtp.GTPEngine(model=model, mcts_simulations=SIMS, size=S, mcts_batch_size=B, use_symmetry=False)
        assert g.parse_command("protocol_version") == "= 2\n\n"
        assert g.parse_command("name") == "= SejongGo - model_7 - 32 simulations\n\n"
        assert g.parse_command("boardsize 9") == "=\n\n"
        with pytest.raises(Exception):
            g.parse_command("boardsize 19")
        assert g.parse_move("A9") == (0, 0) and g.parse_move("J1") == (8, 8) and g.parse_move("pass") == (0, S)
        assert g.print_move(0, 0) == "A9" and g.print_move(8, 8) == "J1"
        # oracle side: the same game, searching each genmove from scratch or from the kept subtree
        om = FakeModel("model_7", salt=11, sharp=True)
        board, _ = o.game_init(S)
        tree = None

        def reroot(t, index):                      # sejonggo.py:38-43
            c = t.child(index) if t is not None else None
            return c.detach() if c is not None else None

        script = [("play", "B", "E5"), ("genmove", "W"), ("play", "B", "C3"), ("genmove", "W"), ("genmove", "B"), ("play", "W", "pass"),
                  ("genmove", "B")]
        for cmd in script:
            if cmd[0] == "play":
                x, y = g.parse_move(cmd[2])
                index = S * S if y == S else y * S + x
                assert g.parse_command("play %s %s" % (cmd[1], cmd[2])) == "=\n\n"
                tree = reroot(tree, index)
                o.make_play(x, y, board, gtp.COLOR_TO_PLAYER[cmd[1]])
            else:
                p, v = om.predict_on_batch(board)
                if tree is None or tree.nchild == 0:
                    tree = o.new_tree(p[0], board)
                for _ in range(SIMS // B):
                    o.simulate(tree, np.copy(board), lambda b: om.predict_on_batch(b), B, int(board[0, 0, 0, 16]))
                index = o.pick_t0(tree)
                x, y = index % S, index // S
                out = g.parse_command("genmove %s" % cmd[1])
                assert out == "= %s\n\n" % g.print_move(x, y), (cmd, out)
                tree = reroot(tree, index)
                o.make_play(x, y, board, gtp.COLOR_TO_PLAYER[cmd[1]])
            assert np.array_equal(g.board, board), cmd
        assert g.parse_command("clear_board") == "=\n\n"
        assert np.array_equal(g.board, o.game_init(S)[0])
    finally:
        conf.clear()
        conf.update(old)


def test_main_selfplay_launcher_single_rank(tmp_path, capsys):
    """main_selfplay.main (main_selfplay.py:9-29) in one process: loads / creates the best model, plays the unplayed
    games with slot refill, keeps the record rows on the device, "gathers" them (world 1) and writes the reference's
    directory tree; a second round finds no new best model and stops; a re-run skips the games already on disk."""
    from sejonggo_b200 import main_selfplay as ms, model
    from sejonggo_b200.conf import conf
    old = dict(conf)
    try:
        conf.update(SIZE=9, N_RESIDUAL_BLOCKS=1, MCTS_SIMULATIONS=16, ENERGY=8, MCTS_BATCH_SIZE=8, STOP_EXPLORATION=2,
                    SELF_PLAY_DIR=str(tmp_path / "sp"), MODEL_DIR=str(tmp_path / "models"), N_GAMES=7, CONCURRENT_GAMES=3,
                    RESIGNATION_PERCENT=1.0)
        os.makedirs(tmp_path / "sp" / "model_1" / "game_00002")               # already played elsewhere
        ms.main(["--games", "7", "--concurrent", "3", "--sims", "16", "--mode", "b", "--num-moves", "5", "--gather-every", "2"])
        out = capsys.readouterr().out
        assert "SELF-PLAYING BEST MODEL  model_1" in out and "No new best model for self-playing. Stopping.." in out
        assert "6 games saved" in out
        root = tmp_path / "sp" / "model_1"
        for g in (0, 1, 3, 4, 5, 6):
            moves = sorted(os.listdir(root / ("game_%05d" % g)))
            assert moves and moves[0] == "move_000" and len(moves) <= 5
            z = np.load(root / ("game_%05d" % g) / "move_000" / "sample.npz")
            assert z["board"].shape == (1, 9, 9, 17) and float(z["board"][..., :16].sum()) == 0.0 and float(z["board"][0, 0, 0, 16]) == 1.0
            assert z["policy_target"].shape == (82,) and abs(float(z["policy_target"].sum()) - 1.0) < 0.05
            assert float(z["value_target"]) in (1.0, -1.0)
        assert os.listdir(root / "game_00002") == []                           # the claimed one was left alone
        ms.main(["--games", "7", "--concurrent", "3", "--sims", "16", "--mode", "a", "--num-moves", "5", "--max-rounds", "1"])
        assert "0 games saved" in capsys.readouterr().out                      # resume: nothing left to play
    finally:
        conf.clear()
        conf.update(old)

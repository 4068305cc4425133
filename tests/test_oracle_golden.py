"""CPU suite: pins the C oracle (oracle/go_oracle.c) against fixtures produced by
the unmodified reference (oracle/gen_golden.py) — SURVEY.md §8c."""
import os
import glob
import numpy as np
import pytest

from oracle import oracle as o
from oracle import game_loop as gl
from oracle.fake_eval import evaluate, evaluate_kind, FakeModel
from tests.treeio import oracle_rows, rows_equal
from tests.conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.mark.parametrize("size", [5, 9, 19])
def test_rules_fixture(size):
    z = _load("rules_s%d.npz" % size)
    offs, moves = z["move_offsets"], z["moves"]
    srow = 0
    for g in range(len(offs) - 1):
        mv = moves[offs[g]:offs[g + 1]]
        w, b, wh, states, masks = o.replay(size, mv, komi=float(z["komi"]))
        n = len(mv) + 1
        assert np.array_equal(states, z["states"][srow:srow + n])
        assert np.array_equal(masks, z["masks"][srow:srow + n])
        assert (w, b, wh) == tuple(z["results"][g])
        srow += n


@pytest.mark.parametrize("size", [9, 19])
def test_symmetry_fixture(size):
    z = _load("symmetry_s%d.npz" % size)
    A = size * size + 1
    assert list(z["symmetries"]) == ["_id", "left_diagonal", "vertical_axis", "horizontal_axis",
                                     "rotation_90", "rotation_180", "rotation_270"]
    idx = np.arange(size * size, dtype=np.int32).reshape(1, size, size, 1).repeat(17, axis=3)
    pol = np.arange(A, dtype=np.float32).reshape(1, A)
    for k in range(8):
        assert np.array_equal(o.sym_board(k, idx)[0, :, :, 0].reshape(-1), z["board_src"][k])
        assert np.array_equal(o.sym_policy(k, pol, size)[0].astype(np.int32), z["policy_src"][k])


MCTS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "mcts_*.npz")))


@pytest.mark.parametrize("name", MCTS)
def test_mcts_fixture(name):
    z = _load(name)
    S, mode, batch, steps, seed = int(z["size"]), str(z["mode"]), int(z["batch"]), int(z["steps"]), int(z["seed"])
    ev = lambda b: evaluate(b, seed, True)
    board, _ = o.game_init(S)
    pol, _ = ev(board)
    tree = o.new_tree(pol[0], board, noise=z["noise"])
    offs = z["tree_offsets"]
    for ply in range(len(z["picks"])):
        op = int(board[0, 0, 0, 16])
        for _ in range(steps):
            if mode == 'a':
                o.simulate(tree, np.copy(board), ev, batch, op)
            else:
                o.async_simulate2(tree, np.copy(board), ev, batch, op)
        ok, why = rows_equal(oracle_rows(tree), z["trees"][offs[ply]:offs[ply + 1]])
        assert ok, "ply %d: %s" % (ply, why)
        assert o.pick_t0(tree) == z["t0picks"][ply]
        sel = int(z["picks"][ply])
        tree = tree.child(sel).detach()
        o.make_play(sel % S, sel // S, board)
        if tree.nchild == 0:
            assert z["newtree"][ply] == 1
            pol, _ = ev(board)
            tree = o.new_tree(pol[0], board)
        else:
            assert z["newtree"][ply] == 0


GAMES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "game_*.npz")))


def check_game(gd, z):
    S = int(z["size"])
    assert len(gd['moves']) == len(z["move"])
    for i, m in enumerate(gd['moves']):
        assert np.array_equal(o.pack_board(np.ascontiguousarray(m['board'], dtype=np.int32)), z["boards"][i]), i
        assert tuple(m['move']) == tuple(z["move"][i]), i
        assert m['move_n'] == z["move_n"][i] and m['player'] == z["player"][i], i
        assert np.float32(m['value']).view(np.uint32) == z["value"][i].view(np.uint32), i
        assert np.array_equal(np.asarray(m['policy'], dtype=np.float64).view(np.uint64), z["policy"][i].view(np.uint64)), i
    assert gd['modelB_name'] == str(z["modelB_name"]) and gd['modelW_name'] == str(z["modelW_name"])
    assert (-1 if gd['winner'] is None else gd['winner']) == int(z["winner"])
    assert str(gd['winner_model']) == str(z["winner_model"])
    assert gd['result'] == str(z["result"])


def game_kwargs(z):
    resign = None if np.isnan(float(z["resign"])) else float(z["resign"])
    nm = int(z["num_moves"])
    return dict(stop_exploration=int(z["stop_exploration"]), self_play=bool(int(z["self_play"])),
                num_moves=None if nm < 0 else nm, resign_model1=resign, resign_model2=resign)


@pytest.mark.parametrize("name", GAMES)
def test_game_fixture(name):
    z = _load(name)
    S, mode, batch, sims, seed = int(z["size"]), str(z["mode"]), int(z["batch"]), int(z["sims"]), int(z["seed"])
    live = "live_sym" in z.files and int(z["live_sym"])
    rng = gl.ReplayRng(coin=z["coin"], noise=z["noise"], choice=z["choice"], symmetry=z["symmetry"] if live else None)
    kw = game_kwargs(z)
    kind = str(z["evalkind"]) if "evalkind" in z.files else "fake"
    if mode == 'a':
        m1 = FakeModel("model_1", salt=seed, sharp=True, kind=kind)
        m2 = m1 if kw['self_play'] else FakeModel("model_2", salt=seed + 1, sharp=True, kind=kind)
        gd = gl.play_game(m1, m2, sims, size=S, mcts_batch_size=batch, rng=rng, **kw)
        calls = m1.calls + [-1] + (m2.calls if m2 is not m1 else [])
        assert calls == list(z["calls"])
    else:
        salts = {"BEST_SYM": seed, "LATEST_SYM": seed + 1}
        names = {"BEST_SYM": "model_1", "LATEST_SYM": "model_2"}

        def predict(tag, board, sym):
            if live:        # all 7 symmetries in play; LATEST_SYM is served by the BEST network (Q21)
                p, v = gl.sym_predict(FakeModel("model_1", salt=seed, sharp=True, kind=kind), board, sym)
                return p[0], v[0]
            p, v = evaluate_kind(board, salts[tag], True, kind)
            return p[0], v[0]

        i1, i2 = ("BEST_SYM", "BEST_SYM") if kw['self_play'] else ("BEST_SYM", "LATEST_SYM")
        gd = gl.play_game_async(i1, i2, batch, process_id=0, size=S, conf_sims=sims, conf_energy=batch,
                                rng=rng, predict=predict, names=names, **kw)
    check_game(gd, z)
    if live:
        assert len(rng._sym) == 0 and len(set(z["symmetry"].tolist())) >= 5      # every recorded draw consumed; most of the 7 maps occurred


def _calib_checks(games, z):
    offs = z["move_offsets"]
    assert len(games) == int(z["n_games"])
    S = int(z["size"])
    for g, gd in enumerate(games):
        want = z["resign"][g]
        got = gd['resign_model1']
        assert (got is None and np.isnan(want)) or (got is not None and np.float32(got) == np.float32(want)), (g, got, want)
        mv = z["move"][offs[g]:offs[g + 1]]
        assert [tuple(m['move']) for m in gd['moves']] == [tuple(x) for x in mv.tolist()], g
        assert np.array_equal(np.array([np.float32(m['value']) for m in gd['moves']], np.float32).view(np.uint32),
                              z["value"][offs[g]:offs[g + 1]].view(np.uint32)), g
        assert gd['result'] == str(z["result"][g]) and (-1 if gd['winner'] is None else gd['winner']) == int(z["winner"][g]), g


def test_self_play_resignation_calibration_fixture():
    """self_play.self_play (self_play.py:343-378) run by the unmodified reference with recorded draws: 16 games, the
    lottery, thresholds appearing from the 4th game on (index into the UNSORTED min_values list) and games that
    resign — the oracle's restatement reproduces every game and every threshold."""
    z = _load("selfplay_calib_s5.npz")
    S, batch, sims, seed = int(z["size"]), int(z["batch"]), int(z["sims"]), int(z["seed"])
    rng = gl.ReplayRng(coin=z["coin"], noise=z["noise"], choice=z["choice"])
    lot = list(z["lottery"])
    games = gl.self_play(FakeModel("model_1", salt=seed, sharp=True), int(z["n_games"]), sims, lambda: lot.pop(0),
                         float(z["percent"]), float(z["allowed_error"]), int(z["stop_exploration"]), size=S,
                         mcts_batch_size=batch, rng=rng, komi=float(z["komi"]))
    _calib_checks(games, z)
    assert not lot and np.isfinite(z["resign"]).sum() >= 4 and any(str(r).endswith("+R") for r in z["result"])


def test_rules_64_long_19x19_games_hashed():
    """64 games played by the unmodified reference on 19x19 (several run into the 2*S*S = 722-ply cap; suicides
    and ko retakes included): the oracle's replay reproduces the packed state and the legality mask at every one of
    the ~25,000 plies (64-bit row hashes) and every final result."""
    from oracle.fake_eval import hash_rows
    z = _load("ruleshash_s19_64.npz")
    S = int(z["size"])
    offs = z["move_offsets"]
    lens = np.diff(offs)
    assert len(lens) == 64 and (lens == 2 * S * S).sum() >= 5 and lens.sum() > 20000
    pos = 0
    for g in range(len(lens)):
        mv = z["moves"][offs[g]:offs[g + 1]].astype(np.int32)
        w, b, wh, states, masks = o.replay(S, mv, float(z["komi"]))
        n = len(mv) + 1
        assert np.array_equal(hash_rows(states), z["state_hash"][pos:pos + n]), g
        assert np.array_equal(hash_rows(masks), z["mask_hash"][pos:pos + n]), g
        assert (w, b, wh) == tuple(z["results"][g]), g
        pos += n
    assert pos == len(z["state_hash"])

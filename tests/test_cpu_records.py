"""sgfsave-compatible record writer: layout, dtypes, Q14/Q15 semantics."""
import os
import numpy as np
import pytest

from oracle import oracle as o
from sejonggo_b200 import sgfsave
from sejonggo_b200.conf import conf


def test_sample_layout_and_value_target(tmp_path):
    S = 9
    board, _ = o.game_init(S)
    o.make_play(2, 3, board)
    packed = o.pack_board(board)
    assert np.array_equal(sgfsave.unpack_board(packed, S), board.astype(np.float32))
    pol = np.linspace(0, 1, S * S + 1)
    gd = dict(winner=1, moves=[dict(board=board.copy(), policy=pol, value=np.float32(0.1), move=(0, 0), move_n=0, player=1),
                               dict(board=packed, policy=pol, value=np.float32(0.2), move=(1, 0), move_n=1, player=1),
                               dict(board=packed, policy=pol, value=np.float32(0.3), move=(2, 0), move_n=2, player=-1)])
    old = dict(conf)
    try:
        conf.update(SELF_PLAY_DIR=str(tmp_path / "sp"), SIZE=S)
        sgfsave.save_self_play_data("model_7", 3, gd)
        used = sgfsave.save_self_play_data("model_7", 3, gd)          # clash -> the reference bumps the game number
        assert used == 4
    finally:
        conf.clear()
        conf.update(old)
    for mv, vt in ((0, 1.0), (1, 1.0), (2, -1.0)):
        d = tmp_path / "sp" / "model_7" / "game_00003" / ("move_%03d" % mv)
        z = np.load(os.path.join(d, "sample.npz"))
        assert z["board"].shape == (1, S, S, 17) and z["board"].dtype == np.float32
        assert z["policy_target"].shape == (S * S + 1,) and z["policy_target"].dtype == np.float32
        assert z["value_target"].shape == () and float(z["value_target"]) == vt
        assert np.array_equal(z["board"], board.astype(np.float32))
    # Q15: with winner == 0 (white won) a white ply (player -1) still gets -1
    assert float(sgfsave.sample_arrays(dict(board=board, policy=pol, player=-1), 0)["value_target"]) == -1.0
    assert os.path.isdir(tmp_path / "sp" / "model_7" / "game_00004" / "move_002")


def test_resignation_calibrator_follows_the_reference_sequence():
    """self_play.ResignationCalibrator against a literal transcription of the reference's bookkeeping
    (self_play.py:343-378): same lottery draws, same games -> same thresholds, including the index into the UNSORTED list."""
    from sejonggo_b200.self_play import ResignationCalibrator
    rs = np.random.RandomState(3)
    n = 60
    lottery = rs.random_sample(n)
    games = []
    for g in range(n):
        k = int(rs.randint(1, 12))
        games.append(dict(winner=[1, 0, None][int(rs.randint(3))],
                          moves=[dict(value=np.float32(rs.uniform(-1, 1))) for _ in range(k)]))
    # the reference, line by line
    want, current_resign, min_values = [], None, []
    for g in range(n):
        resign = current_resign if lottery[g] > 0.25 else None
        want.append(resign)
        if resign is None:
            gd = games[g]
            mv = gd['moves'][::2] if gd['winner'] == 1 else gd['moves'][1::2]
            if mv:                                       # (min() of an empty list would raise in the reference)
                min_values.append(min(m['value'] for m in mv))
                idx = int(0.2 * len(min_values))
                if idx > 0:
                    current_resign = min_values[idx]
    it = iter(lottery)
    cal = ResignationCalibrator(percent=0.25, allowed_error=0.2, rand=lambda: next(it))
    got = []
    for g in range(n):
        r1, r2 = cal.start(g)
        assert r1 is r2
        got.append(r1)
        cal.end(g, games[g])
    assert got == want and any(w is not None for w in want)
    assert cal.min_values == min_values and min_values != sorted(min_values)


def test_record_rows_round_trip_and_partial_games():
    """records.rows_from_game_data -> games_from_rows: every field of game_data that sgfsave / the calibration read
    survives; rows may arrive in any order and mixed over games; a game without its footer is not reported."""
    from sejonggo_b200 import records
    from oracle import oracle as o, game_loop as gl
    from oracle.fake_eval import FakeModel
    S = 5
    rows, ref = [], {}
    for gid in (3, 7):
        m = FakeModel("model_2", salt=gid, sharp=True)
        gd = gl.play_game(m, m, 8, 1, self_play=True, num_moves=6, size=S, mcts_batch_size=4, rng=gl.SeededRng(gid),
                          resign_model1=-0.9 if gid == 7 else None, resign_model2=-0.9 if gid == 7 else None)
        gd['model1_isblack'] = gid == 3
        ref[gid] = gd
        rows.append(records.rows_from_game_data(S, gid, gd, o.pack_board))
    allrows = np.concatenate(rows)
    rs = np.random.RandomState(0)
    shuffled = allrows[rs.permutation(len(allrows))]
    partial = shuffled[~((shuffled[:, 0] == 7) & (shuffled[:, 1] == records.FOOTER))]        # game 7 loses its footer
    got = records.games_from_rows(partial, S, names=("model_2", "model_2"))
    assert sorted(got) == [3]
    got = records.games_from_rows(shuffled, S, names=("model_2", "model_2"))
    assert sorted(got) == [3, 7]
    for gid, gd in got.items():
        r = ref[gid]
        assert gd['result'] == r['result'] and gd['winner'] == r['winner'] and gd['end_reason'] == r['end_reason']
        assert gd['resign_model1'] == (None if r['resign_model1'] is None else float(np.float32(r['resign_model1'])))
        assert gd['model1_isblack'] == r['model1_isblack'] and gd['game_id'] == gid
        assert len(gd['moves']) == len(r['moves'])
        for a, b in zip(gd['moves'], r['moves']):
            assert a['move'] == b['move'] and a['move_n'] == b['move_n'] and a['player'] == b['player']
            assert np.float32(a['value']) == np.float32(b['value'])
            assert np.array_equal(a['board'], o.pack_board(np.ascontiguousarray(b['board'], dtype=np.int32)))
            assert np.array_equal(a['policy'], np.asarray(b['policy'], np.float32))
    with pytest.raises(ValueError):
        records.games_from_rows(allrows[1:], S)          # a ply row of a finished game is missing


def test_sgf_game_record(tmp_path):
    """sgfsave.save_game_sgf (sgfsave.py:130-167, written without sgfmill): properties, one node per ply, colours from the
    reference's lagging `player` field, SGF letter coordinates (column x, row y from the top, empty = pass), file
    number bumped on a clash; switched on by conf['SGF_ENABLED'] from save_self_play_data like the reference."""
    import re
    from oracle import game_loop as gl
    from oracle.fake_eval import FakeModel
    old = dict(conf)
    try:
        conf.update(SIZE=5, KOMI=5.5, GAMES_DIR=str(tmp_path / "games"), SELF_PLAY_DIR=str(tmp_path / "sp"), SGF_ENABLED=True)
        m = FakeModel("model_8", salt=4, sharp=True)
        gd = gl.play_game(m, m, 8, 2, self_play=True, num_moves=30, size=5, mcts_batch_size=4, rng=gl.SeededRng(6))
        sgfsave.save_self_play_data("model_8", 0, gd, size=5)
        sgfsave.save_self_play_data("model_8", 0, gd, size=5)            # the same game number again: sample dirs and sgf both bump
        files = sorted(os.listdir(tmp_path / "games" / "model_8"))
        assert files == ["game_000.sgf", "game_001.sgf"]
        text = open(tmp_path / "games" / "model_8" / "game_000.sgf").read()
        assert text.startswith("(;FF[4]") and text.rstrip().endswith(")")
        props = dict(re.findall(r"(SZ|PB|PW|KM|RE)\[([^\]]*)\]", text))
        assert props == {"SZ": "5", "PB": "model_8", "PW": "model_8", "KM": "5.5", "RE": gd['result']}
        nodes = re.findall(r";([BW])\[([a-z]{0,2})\]C\[", text)
        assert len(nodes) == len(gd['moves'])
        for (color, coord), md in zip(nodes, gd['moves']):
            x, y = md['move']
            assert color == ('B' if md['player'] == 1 else 'W')
            assert coord == ("" if y == 5 else "abcde"[x] + "abcde"[y])
        assert "Value " in text
    finally:
        conf.clear()
        conf.update(old)

"""sgfsave-compatible record writer: layout, dtypes, Q14/Q15 semantics."""
import os
import numpy as np

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

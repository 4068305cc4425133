"""GPU parity: rules kernels (csrc/rules.cu) vs the reference fixtures and the C oracle,
called through the C ABI."""
import os
import numpy as np
import pytest
import torch

from oracle import oracle as o
from tests import ref_cases as rc
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from sejonggo_b200.engine import Engine
    return Engine(**kw)


class EngineApi(object):
    """sejonggo_b200.play — the reference-named rules API over the CUDA kernels."""
    @staticmethod
    def game_init():
        from sejonggo_b200 import play
        return play.game_init(rc.S)

    @staticmethod
    def make_play(x, y, board, color=None):
        from sejonggo_b200 import play
        return play.make_play(x, y, board, color)

    @staticmethod
    def legal_moves(board):
        from sejonggo_b200 import play
        return play.legal_moves(board)

    @staticmethod
    def get_winner(board):
        from sejonggo_b200 import play
        return play.get_winner(board)


@pytest.mark.parametrize("case", rc.RULE_CASES, ids=lambda c: c.__name__)
def test_reference_rule_cases(case):
    case(EngineApi)


def test_make_play_occupied_asserts():
    from sejonggo_b200 import play
    b, _ = play.game_init(9)
    play.make_play(3, 3, b)
    with pytest.raises(AssertionError):
        play.make_play(3, 3, b)


def test_rules_64_long_19x19_games_lockstep():
    """64 reference-played 19x19 games (several reach the 722-ply cap; suicides and ko retakes executed) advanced in
    lock-step on the device: packed state and legality mask of every game at every ply against the reference's row
    hashes, and the final area scores."""
    from oracle.fake_eval import hash_rows
    z = np.load(os.path.join(GOLDEN, "ruleshash_s19_64.npz"))
    S = int(z["size"])
    offs, moves = z["move_offsets"], z["moves"].astype(np.int32)
    G = len(offs) - 1
    lens = np.diff(offs)
    srow = np.concatenate([[0], np.cumsum(lens + 1)])[:-1]
    e = _engine(size=S, n_games=G, max_leaves=1, arena_blocks=2)
    e.reset()
    checked = 0
    for t in range(int(lens.max()) + 1):
        sh = hash_rows(e.export_packed(0).cpu().numpy().view(np.uint32))
        mh = hash_rows(e.legal_masks().cpu().numpy())
        live = np.nonzero(t <= lens)[0]
        assert np.array_equal(sh[live], z["state_hash"][srow[live] + t]), t
        assert np.array_equal(mh[live], z["mask_hash"][srow[live] + t]), t
        checked += len(live)
        mv = np.where(t < lens, moves[np.minimum(offs[:-1] + t, len(moves) - 1)], -1).astype(np.int32)
        e.apply_moves(mv)
    assert checked == len(z["state_hash"]) > 20000
    sc = e.score().cpu().numpy()
    for g in range(G):
        w, b, wh = z["results"][g]
        assert (sc[g, 0], sc[g, 1], sc[g, 2] + 5.5) == (w, b, wh), g
    e.check_errors()


@pytest.mark.parametrize("size", [5, 9, 19])
def test_rules_fixture_lockstep(size):
    z = np.load(os.path.join(GOLDEN, "rules_s%d.npz" % size))
    offs, moves = z["move_offsets"], z["moves"]
    G = len(offs) - 1
    lens = np.diff(offs)
    srow = np.concatenate([[0], np.cumsum(lens + 1)])[:-1]
    e = _engine(size=size, n_games=G, max_leaves=1, arena_blocks=2)
    e.reset()
    for t in range(int(lens.max()) + 1):
        packed = e.export_packed(0).cpu().numpy().view(np.uint32)
        masks = e.legal_masks().cpu().numpy()
        mv = np.full(G, -1, np.int32)
        for g in range(G):
            if t <= lens[g]:
                assert np.array_equal(packed[g], z["states"][srow[g] + t]), (g, t)
                assert np.array_equal(masks[g], z["masks"][srow[g] + t]), (g, t)
            if t < lens[g]:
                mv[g] = moves[offs[g] + t]
        e.apply_moves(mv)
    sc = e.score().cpu().numpy()
    for g in range(G):
        w, b, wh = z["results"][g]
        assert (sc[g, 0], sc[g, 1], sc[g, 2] + 5.5) == (w, b, wh)
    e.check_errors()


def test_import_export_roundtrip_and_planes():
    z = np.load(os.path.join(GOLDEN, "rules_s9.npz"))
    states = z["states"][5:200:7]
    n = len(states)
    boards = np.concatenate([o.unpack_board(s, 9) for s in states])
    e = _engine(size=9, n_games=n, max_leaves=1, arena_blocks=2)
    e.import_boards(boards)
    assert np.array_equal(e.export_boards().cpu().numpy(), boards)
    assert np.array_equal(e.export_packed(0).cpu().numpy().view(np.uint32), states)
    zs = np.load(os.path.join(GOLDEN, "symmetry_s9.npz"))
    for k in range(8):
        planes = e.export_planes(0, 0, n, sym=k).cpu().numpy()
        assert np.array_equal(planes, o.sym_board(k, boards).astype(np.float32)), k
        assert np.array_equal(planes.reshape(n, 81, 17), boards.reshape(n, 81, 17)[:, zs["board_src"][k], :].astype(np.float32))
        pol = np.random.RandomState(k).rand(n, 82).astype(np.float32)
        assert np.array_equal(e.policy_unsym(pol, sym=k).cpu().numpy(), pol[:, zs["policy_src"][k]])
    syms = np.arange(n, dtype=np.int32) % 7
    planes = e.export_planes(0, 0, n, syms=syms).cpu().numpy()
    for i in range(n):
        assert np.array_equal(planes[i], o.sym_board(int(syms[i]), boards[i:i + 1])[0].astype(np.float32))


def test_config2_random_playouts_4096_bitexact():
    """BASELINE config 2: 4096 concurrent 19x19 random playouts; the GPU's move lists are
    replayed through the oracle: every ply for 256 games, final state + score for all."""
    S, G, P = 19, 4096, 722
    e = _engine(size=S, n_games=G, max_leaves=1, arena_blocks=2)
    e.reset()
    moves, nplies = e.random_playouts(seed=20260, max_plies=P)
    final = e.export_packed(0).cpu().numpy().view(np.uint32)
    sc = e.score().cpu().numpy()
    mv, npl = moves.cpu().numpy().astype(np.int32), nplies.cpu().numpy()
    assert npl.min() >= 2 and npl.max() <= P
    for g in range(G):
        w, b, wh, states, _ = o.replay(S, mv[g, :npl[g]], want_masks=False)
        assert np.array_equal(states[-1], final[g]), g
        assert (w, b, wh) == (sc[g, 0], sc[g, 1], sc[g, 2] + 5.5), g
    # per-ply comparison for the first 256 games, replayed in lock-step on the device
    K = 256
    e2 = _engine(size=S, n_games=K, max_leaves=1, arena_blocks=2)
    e2.reset()
    ref = [o.replay(S, mv[g, :npl[g]]) for g in range(K)]
    for t in range(int(npl[:K].max()) + 1):
        packed = e2.export_packed(0).cpu().numpy().view(np.uint32)
        masks = e2.legal_masks().cpu().numpy()
        step = np.full(K, -1, np.int32)
        for g in range(K):
            if t <= npl[g]:
                assert np.array_equal(packed[g], ref[g][3][t]), (g, t)
                assert np.array_equal(masks[g], ref[g][4][t]), (g, t)
                if t < npl[g]:
                    assert masks[g][mv[g, t]] == 0          # the device only ever picked legal moves
                    step[g] = mv[g, t]
        e2.apply_moves(step)
    e.check_errors()
    e2.check_errors()

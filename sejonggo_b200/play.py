"""Rules API mirror of the reference's play.py (the real rules engine; go_game.py
is a dead wrapper — SURVEY §0).  Same names, argument meaning and error behaviour,
on the reference's own board tensor int32 [n,S,S,17]; the work happens in the CUDA
rules kernels (csrc/rules.cu) through the C ABI.  Batched variants take n boards."""
import numpy as np

from .conf import conf
from .engine import Engine, EngineError

_engines = {}


def _engine(size, n):
    key = (size, n)
    if key not in _engines:
        _engines[key] = Engine(size=size, n_games=n, trees_per_game=1, max_leaves=1, arena_blocks=2, komi=conf['KOMI'])
    e = _engines[key]
    e.komi = conf['KOMI']
    return e


def index2coord(index, size=None):          # play.py:31-34
    size = size or conf['SIZE']
    y = index // size
    return index - size * y, y


def coord2index(x, y, size=None):           # play.py:36-37
    return y * (size or conf['SIZE']) + x


def game_init(size=None, n=1):              # play.py:295-299
    size = size or conf['SIZE']
    board = np.zeros((n, size, size, 17), dtype=np.int32)
    board[:, :, :, -1] = 1
    return board, 1


def _check(board):
    b = np.asarray(board)
    assert b.ndim == 4 and b.shape[1] == b.shape[2] and b.shape[3] == 17
    return b


def make_plays(moves, boards, colors=None):
    """Batched make_play: moves[i] (action index, S*S = pass) on boards[i]; in place."""
    b = _check(boards)
    n, S = b.shape[0], b.shape[1]
    e = _engine(S, n)
    e.import_boards(b.astype(np.int32))
    mover = e.export_packed(0)[:, -1].cpu().numpy().astype(np.int32)
    if colors is not None:
        mover = np.where(np.asarray(colors) != 0, np.asarray(colors), mover).astype(np.int32)
    e.apply_moves(np.asarray(moves, np.int32), None if colors is None else np.asarray(colors, np.int32))
    try:
        e.check_errors()
    except EngineError as ex:
        raise AssertionError(str(ex))      # the reference asserts the target is empty (play.py:233-234)
    boards[...] = e.export_boards().cpu().numpy().astype(boards.dtype)
    return boards, mover


def make_play(x, y, board, color=None):     # play.py:226-242
    S = np.asarray(board).shape[1]
    mv = S * S if y == S else y * S + x
    _, mover = make_plays([mv], board, None if color is None else [color])
    return board, int(mover[0])


def legal_moves_batch(boards):
    b = _check(boards)
    e = _engine(b.shape[1], b.shape[0])
    e.import_boards(b.astype(np.int32))
    return e.legal_masks().cpu().numpy().astype(np.int64)


def legal_moves(board):                      # play.py:71-104 -> int64[A], 1 = illegal
    return legal_moves_batch(board)[0]


def get_winners(boards):
    b = _check(boards)
    e = _engine(b.shape[1], b.shape[0])
    e.import_boards(b.astype(np.int32))
    sc = e.score().cpu().numpy()
    out = []
    for w, bp, wp in sc:
        black, white = int(bp), int(wp) + conf['KOMI']
        out.append((1 if black > white else (0 if black == white else -1), black, white))
    return out


def get_winner(board):                       # play.py:274-284
    return get_winners(board)[0]

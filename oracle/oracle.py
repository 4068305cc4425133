"""TEST INFRASTRUCTURE ONLY — ctypes front-end to oracle/go_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package (sejonggo_b200/)
never does: it fails loudly when its CUDA library is missing.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgo_oracle.so")
_lib = None

EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float))


def build(force=False):
    """Compiles go_oracle.c when the library is missing or older than the source — under a file lock, so that several
    processes starting together (torchrun ranks, pytest-xdist workers) do not run `make` into the same file at once."""
    import fcntl
    src = os.path.join(_HERE, "go_oracle.c")

    def stale():
        return force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src)

    if stale():
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        with open(os.path.join(os.path.dirname(_SO), ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if stale():
                    subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.orc_node_new.restype = vp
        L.orc_node_new.argtypes = [C.c_int, C.c_double, C.c_int]
        L.orc_node_free.argtypes = [vp]
        L.orc_node_add_child.argtypes = [vp, vp]
        L.orc_node_child.restype = vp
        L.orc_node_child.argtypes = [vp, C.c_int]
        L.orc_node_child_at.restype = vp
        L.orc_node_child_at.argtypes = [vp, C.c_int]
        L.orc_node_parent.restype = vp
        L.orc_node_parent.argtypes = [vp]
        L.orc_node_detach.argtypes = [vp]
        L.orc_node_set.argtypes = [vp, C.c_int64, C.c_float, C.c_float, C.c_int]
        L.orc_node_count.restype = C.c_int64
        L.orc_node_count.argtypes = [vp]
        L.orc_node_value.restype = C.c_float
        L.orc_node_value.argtypes = [vp]
        L.orc_node_mean.restype = C.c_float
        L.orc_node_mean.argtypes = [vp]
        L.orc_node_p.restype = C.c_double
        L.orc_node_p.argtypes = [vp]
        for f in ("orc_node_busy", "orc_node_nchild", "orc_node_move", "orc_tree_depth", "orc_pick_t0"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [vp]
        L.orc_node_children.argtypes = [vp] + [vp] * 7
        L.orc_new_tree.restype = vp
        L.orc_new_tree.argtypes = [C.c_int, vp, vp, vp, C.c_double]
        L.orc_new_subtree.argtypes = [C.c_int, vp, vp, vp, vp, C.c_double]
        L.orc_top_one_action.restype = vp
        L.orc_top_one_action.argtypes = [vp]
        L.orc_top_one_vl.restype = vp
        L.orc_top_one_vl.argtypes = [vp]
        L.orc_top_n_actions.argtypes = [vp, C.c_int, vp, vp]
        L.orc_simulate_a.argtypes = [C.c_int, vp, vp, C.c_int, C.c_int, EVAL_FN, vp]
        L.orc_find_best_leaf_vl.restype = vp
        L.orc_find_best_leaf_vl.argtypes = [vp, vp, vp]
        L.orc_node_by_moves.restype = vp
        L.orc_node_by_moves.argtypes = [vp, vp, C.c_int]
        L.orc_back_propagation.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_float]
        L.orc_wave_b.argtypes = [C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, EVAL_FN, vp]
        L.orc_make_play.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_legal_moves.argtypes = [C.c_int, vp, vp]
        L.orc_get_winner.argtypes = [C.c_int, vp, C.c_double, vp, vp]
        L.orc_get_points.argtypes = [C.c_int, vp, vp]
        L.orc_color_board.argtypes = [C.c_int, vp, C.c_int, vp]
        L.orc_capture_group.argtypes = [C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.orc_game_init.argtypes = [C.c_int, vp]
        L.orc_sym_board.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_sym_policy.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_pack_board.argtypes = [C.c_int, vp, vp]
        L.orc_unpack_board.argtypes = [C.c_int, vp, vp]
        L.orc_replay.argtypes = [C.c_int, C.c_int, vp, C.c_double, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- rules API
# Same names / argument meaning as play.py; S is inferred from the board.

def game_init(size):
    board = np.zeros((1, size, size, 17), dtype=np.int32)
    lib().orc_game_init(size, _p(board))
    return board, 1


def _check(board):
    assert board.dtype == np.int32 and board.flags.c_contiguous and board.shape[-1] == 17
    return board.shape[-2]


def make_play(x, y, board, color=None):
    S = _check(board)
    r = lib().orc_make_play(S, int(x), int(y), _p(board), 0 if color is None else int(color))
    if r == -99:
        raise AssertionError("make_play on an occupied point (play.py:233-234)")
    return board, r


def legal_moves(board):
    S = _check(board)
    mask = np.zeros(S * S + 1, dtype=np.int64)
    lib().orc_legal_moves(S, _p(board), _p(mask))
    return mask


def get_winner(board, komi=5.5):
    S = _check(board)
    b, w = C.c_double(), C.c_double()
    r = lib().orc_get_winner(S, _p(board), float(komi), C.byref(b), C.byref(w))
    return r, b.value, w.value


def get_points(real_board):
    rb = np.ascontiguousarray(real_board, dtype=np.int32)
    assert rb.shape[0] == rb.shape[1]
    pts = np.zeros(5, dtype=np.int32)
    lib().orc_get_points(rb.shape[0], _p(rb), _p(pts))
    return {k - 2: int(v) for k, v in enumerate(pts) if v}


def color_board(real_board, color):
    rb = np.ascontiguousarray(real_board, dtype=np.int32)
    assert rb.shape[0] == rb.shape[1], "oracle colours square boards only"
    out = np.zeros_like(rb)
    lib().orc_color_board(rb.shape[0], _p(rb), int(color), _p(out))
    return out


def capture_group(x, y, real_board):
    rb = np.ascontiguousarray(real_board, dtype=np.int32)
    S = rb.shape[0]
    gx = np.zeros(S * S, dtype=np.int32)
    gy = np.zeros(S * S, dtype=np.int32)
    n = lib().orc_capture_group(S, _p(rb), int(x), int(y), _p(gx), _p(gy))
    if n < 0:
        return None
    return [(int(gx[i]), int(gy[i])) for i in range(n)]


def sym_board(sym, boards):
    boards = np.ascontiguousarray(boards, dtype=np.int32)
    out = np.empty_like(boards)
    lib().orc_sym_board(boards.shape[-2], sym, boards.shape[0], _p(boards), _p(out))
    return out


def sym_policy(sym, policy, size):
    policy = np.ascontiguousarray(policy, dtype=np.float32)
    out = np.empty_like(policy)
    lib().orc_sym_policy(size, sym, policy.shape[0], _p(policy), _p(out))
    return out


def packed_words(size):
    return 16 * ((size * size + 31) // 32) + 1


def pack_board(board):
    S = _check(board)
    out = np.zeros(packed_words(S), dtype=np.uint32)
    lib().orc_pack_board(S, _p(board), _p(out))
    return out


def unpack_board(packed, size):
    packed = np.ascontiguousarray(packed, dtype=np.uint32)
    board = np.zeros((1, size, size, 17), dtype=np.int32)
    lib().orc_unpack_board(size, _p(packed), _p(board))
    return board


def replay(size, moves, komi=5.5, want_states=True, want_masks=True):
    """Replay a move list; returns (winner, black, white, states[T+1,PW], masks[T+1,A])."""
    moves = np.ascontiguousarray(moves, dtype=np.int32)
    T = len(moves)
    states = np.zeros((T + 1, packed_words(size)), dtype=np.uint32) if want_states else None
    masks = np.zeros((T + 1, size * size + 1), dtype=np.uint8) if want_masks else None
    score = np.zeros(2, dtype=np.float64)
    w = lib().orc_replay(size, T, _p(moves), float(komi),
                         _p(states) if want_states else None,
                         _p(masks) if want_masks else None, _p(score))
    if w == -98:
        raise AssertionError("replay: move onto an occupied point")
    return w, score[0], score[1], states, masks


# ----------------------------------------------------------------- tree API

class Node(object):
    """Handle on an oracle tree node (play.py:376-421 dict node)."""

    def __init__(self, ptr, owner=True):
        self.ptr = ptr
        self.owner = owner

    @staticmethod
    def new(move=-1, p=1.0, p64=True):
        return Node(lib().orc_node_new(move, float(p), 1 if p64 else 0))

    @staticmethod
    def from_dict(d, move=-1):
        """Build from a reference-style dict tree (hand-built test trees)."""
        n = Node.new(move, d.get('p', 1), True)
        lib().orc_node_set(n.ptr, int(d.get('count', 0)), float(d.get('value', 0)),
                           float(d.get('mean_value', 0)), int(d.get('virtual_loss', 0)))
        for m, c in d.get('subtree', {}).items():
            ch = Node.from_dict(c, m)
            ch.owner = False
            lib().orc_node_add_child(n.ptr, ch.ptr)
        return n

    def free(self):
        if self.owner and self.ptr:
            lib().orc_node_free(self.ptr)
            self.ptr = None

    def child(self, move):
        p = lib().orc_node_child(self.ptr, int(move))
        return Node(p, owner=False) if p else None

    count = property(lambda s: lib().orc_node_count(s.ptr))
    value = property(lambda s: np.float32(lib().orc_node_value(s.ptr)))
    mean_value = property(lambda s: np.float32(lib().orc_node_mean(s.ptr)))
    p = property(lambda s: lib().orc_node_p(s.ptr))
    virtual_loss = property(lambda s: lib().orc_node_busy(s.ptr))
    nchild = property(lambda s: lib().orc_node_nchild(s.ptr))
    move = property(lambda s: lib().orc_node_move(s.ptr))
    depth = property(lambda s: lib().orc_tree_depth(s.ptr))

    def detach(self):
        """Cut from the parent and become an owning root (self_play.py:228-233).
        The former parent keeps its (now dangling) entry: free only via this root."""
        lib().orc_node_detach(self.ptr)
        self.owner = True
        return self

    def children(self):
        n = self.nchild
        moves = np.zeros(n, np.int32); counts = np.zeros(n, np.int64)
        values = np.zeros(n, np.float32); means = np.zeros(n, np.float32)
        ps = np.zeros(n, np.float64); busy = np.zeros(n, np.int32); exp = np.zeros(n, np.int32)
        lib().orc_node_children(self.ptr, _p(moves), _p(counts), _p(values), _p(means), _p(ps), _p(busy), _p(exp))
        return dict(moves=moves, counts=counts, values=values, means=means, ps=ps, busy=busy, expanded=exp)

    def to_dict(self):
        """Recursive dump {move: (count, value, mean, p, busy, subtree)} for comparisons."""
        out = {}
        for i in range(self.nchild):
            c = Node(lib().orc_node_child_at(self.ptr, i), owner=False)
            out[c.move] = dict(count=c.count, value=float(c.value), mean_value=float(c.mean_value),
                               p=c.p, virtual_loss=c.virtual_loss, subtree=c.to_dict())
        return out


def new_tree(policy, board, noise=None, eps=0.25):
    S = _check(board)
    policy = np.ascontiguousarray(policy, dtype=np.float32)
    nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
    return Node(lib().orc_new_tree(S, _p(policy), _p(board), None if nz is None else _p(nz), float(eps)))


def new_subtree(node, policy, board):
    S = _check(board)
    policy = np.ascontiguousarray(policy, dtype=np.float32)
    lib().orc_new_subtree(S, node.ptr, _p(policy), _p(board), None, 0.0)


def _wrap_eval(evaluator, size):
    """evaluator(boards int32[n,S,S,17]) -> (policy f32[n,A], value f32[n])"""
    A = size * size + 1

    def cb(ctx, n, boards, policy, value):
        b = np.ctypeslib.as_array(boards, shape=(n, size, size, 17)).copy()
        p, v = evaluator(b)
        np.ctypeslib.as_array(policy, shape=(n, A))[:] = np.asarray(p, dtype=np.float32).reshape(n, A)
        np.ctypeslib.as_array(value, shape=(n,))[:] = np.asarray(v, dtype=np.float32).reshape(n)

    return EVAL_FN(cb)


def simulate(node, board, evaluator, mcts_batch_size, original_player):
    """self_play.py:28 simulate (mode A); mutates `board` like the reference."""
    S = _check(board)
    cb = _wrap_eval(evaluator, S)
    return lib().orc_simulate_a(S, node.ptr, _p(board), int(mcts_batch_size), int(original_player), cb, None)


def find_best_leaf_virtual_loss(node):
    moves = np.zeros(4096, np.int32)
    n = C.c_int(0)
    p = lib().orc_find_best_leaf_vl(node.ptr, _p(moves), C.byref(n))
    if not p:
        return None, None
    return Node(p, owner=False), [int(m) for m in moves[:n.value]]


def get_node_by_moves(node, moves):
    """tree_util.py:27-32; raises like the reference on an invalid path."""
    mv = np.ascontiguousarray(moves, dtype=np.int32)
    p = lib().orc_node_by_moves(node.ptr, _p(mv), len(mv))
    if not p:
        raise Exception("ERROR: Unable to get node: Invalid moves array")
    return Node(p, owner=False)


def back_propagation(result, node):
    """nomodel_self_play.py:40-56: result = (leaf dict with 'count'/'value', moves)."""
    leaf, moves = result
    mv = np.ascontiguousarray(moves, dtype=np.int32)
    rc = lib().orc_back_propagation(node.ptr, _p(mv), len(mv), int(leaf.get('count', 0)), float(leaf.get('value', 0)))
    if rc != 0:
        raise Exception("ERROR: Unable to get node: Invalid moves array")


def async_simulate2(node, board, evaluator, energy, original_player, total_energy=None):
    """nomodel_self_play.py:59 (mode B wave) under the synchronous pool."""
    S = _check(board)
    cb = _wrap_eval(evaluator, S)
    return lib().orc_wave_b(S, node.ptr, _p(board), int(energy), int(total_energy or energy), int(original_player), cb, None)


def pick_t0(node):
    return lib().orc_pick_t0(node.ptr)

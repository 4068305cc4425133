"""Reference unit tests (test/tests.py, test/tree_util_tests.py) against the CPU oracle."""
import numpy as np
import pytest

from oracle import oracle as o
from tests import ref_cases as rc


class OracleApi(object):
    @staticmethod
    def game_init():
        return o.game_init(rc.S)

    make_play = staticmethod(o.make_play)
    legal_moves = staticmethod(o.legal_moves)
    get_winner = staticmethod(o.get_winner)


@pytest.mark.parametrize("case", rc.RULE_CASES, ids=lambda c: c.__name__)
def test_rules_case(case):
    case(OracleApi)


def test_capture_group_order():                   # tests.py:137-213 (embedded at the 9x9 corner)
    def emb(a):
        b = np.zeros((9, 9), np.int32)
        a = np.asarray(a)
        b[:a.shape[0], :a.shape[1]] = a
        # seal the embedding so the outside is not a liberty
        return b
    assert o.capture_group(1, 1, emb([[0, 1, 0], [1, -1, 1], [0, 1, 0]])) == [(1, 1)]
    b = emb([[0, 1, 0], [1, -1, 1], [1, -1, 1], [0, 1, 0]])
    assert o.capture_group(1, 1, b) == [(1, 1), (1, 2)]
    assert o.capture_group(1, 2, b) == [(1, 2), (1, 1)]
    assert o.capture_group(0, 0, emb([[-1, 1, 0], [1, 0, 0], [0, 0, 0]])) == [(0, 0)]
    b = emb([[-1, -1, 1], [1, 1, 0], [0, 0, 0]])
    assert o.capture_group(0, 0, b) == [(0, 0), (1, 0)]
    assert o.capture_group(1, 0, b) == [(1, 0), (0, 0)]
    circ = emb([[0, 1, 1, 1, 0], [1, -1, -1, -1, 1], [1, -1, 1, -1, 1], [1, -1, -1, -1, 1], [0, 1, 1, 1, 0]])
    tgt = [(1, 1), (2, 1), (3, 1), (3, 2), (3, 3), (2, 3), (1, 3), (1, 2)]
    for x, y in tgt:
        assert sorted(o.capture_group(x, y, circ)) == sorted(tgt)


def test_coloring():                              # tests.py:55-117
    big = np.array(rc.BIG)
    tgt = np.array([[1, 1, 1, 2, 0, -2, -1, -1, -1], [1, 1, 1, 2, 0, -2, -1, -1, -1], [1, 1, 1, 2, 0, -2, -1, -1, -1],
                    [1, 1, 1, 2, -2, -1, -1, -2, -1], [2, 2, 2, -2, -1, -2, -2, -1, -1], [0, 0, 0, 2, -2, 0, 0, -2, -2],
                    [0, 0, 0, 2, 0, -2, 0, 0, 0], [0, 0, 0, 2, 0, -2, 0, 2, 0], [0, 0, 0, 0, 0, -2, 0, 0, 0]])
    assert np.array_equal(o.color_board(big, 1) + o.color_board(big, -1), tgt)
    assert o.get_points(big) == {0: 29, 1: 12, 2: 11, -1: 15, -2: 14}


def test_find_best_leaf_order():                  # tree_util_tests.py:69-86
    tree = {'count': 0, 'mean_value': 0, 'virtual_loss': 0, 'value': 0, 'subtree': {
        0: {'count': 0, 'p': 1, 'value': 1, 'mean_value': 0, 'virtual_loss': 0, 'subtree': {
            3: {'count': 0, 'p': 1, 'value': 1, 'mean_value': 0, 'virtual_loss': 0, 'subtree': {}},
            4: {'count': 0, 'p': 0, 'value': 0, 'mean_value': 0, 'virtual_loss': 0, 'subtree': {}}}},
        1: {'count': 0, 'p': 0, 'value': 0, 'mean_value': 0, 'virtual_loss': 0, 'subtree': {}}}}
    t = o.Node.from_dict(tree)
    n1, m1 = o.find_best_leaf_virtual_loss(t)
    assert m1 == [0, 3] and n1.virtual_loss > 0
    n2, m2 = o.find_best_leaf_virtual_loss(t)
    assert m2 == [0, 4]
    n3, m3 = o.find_best_leaf_virtual_loss(t)
    assert m3 == [1] and t.child(0).virtual_loss > 0


def test_all_leaf_busy():                         # tree_util_tests.py:88-122
    tree = {'subtree': {0: {'p': 1, 'value': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}}
    t = o.Node.from_dict(tree)
    o.find_best_leaf_virtual_loss(t)
    o.find_best_leaf_virtual_loss(t)
    assert o.find_best_leaf_virtual_loss(t) == (None, None)


def _dummy_eval(b):                               # tests.py:34-49 DummyModel: policy ~ [82..1], value +1
    n = b.shape[0]
    p = np.tile(np.arange(82, 0, -1, dtype=np.float32), (n, 1))
    p /= p.sum(axis=1, keepdims=True)
    return p, np.ones(n, np.float32)


def test_mcts_leaf():                             # tests.py:729-744
    tree = {'subtree': {0: {'p': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}}
    t = o.Node.from_dict(tree)
    board, _ = o.game_init(9)
    o.simulate(t, board, _dummy_eval, 2, 1)
    assert t.child(0).count == 1 and t.child(1).count == 1
    assert t.child(0).value == -1 and t.child(1).value == -1
    assert t.count == 2 and t.value == -2 and t.mean_value == -1


def test_mcts_nested_other_leaves():              # tests.py:1000-1068
    tree = {'subtree': {0: {'p': .75, 'subtree': {}},
                        1: {'p': .25, 'subtree': {0: {'p': 1, 'subtree': {}}, 2: {'p': 0, 'subtree': {}}}},
                        2: {'p': 0, 'subtree': {}}}}
    t = o.Node.from_dict(tree)
    board, _ = o.game_init(9)
    o.simulate(t, board, _dummy_eval, 2, 1)
    assert t.child(0).count == 1 and t.child(0).value == -1
    assert t.child(1).value == 1 and t.child(1).count == 1
    assert t.child(1).child(0).count == 1 and t.child(1).child(0).value == 1 and t.child(1).child(2).count == 0
    assert t.count == 2 and t.mean_value == 0 and t.child(2).count == 0 and t.child(2).nchild == 0
    assert t.depth == 4


def test_mcts_nested_selected():                  # tests.py:945-998
    tree = {'subtree': {0: {'p': 1, 'subtree': {1: {'p': 0, 'subtree': {}}, 2: {'p': 1, 'subtree': {}}}},
                        1: {'p': 0, 'subtree': {}}}}
    t = o.Node.from_dict(tree)
    board, _ = o.game_init(9)
    o.simulate(t, board, _dummy_eval, 2, 1)
    assert t.child(0).count == 2 and t.child(0).child(1).count == 1 and t.child(0).child(2).count == 1
    assert t.child(1).count == 0 and t.child(0).value == 2 and t.child(0).mean_value == 1 and t.child(1).value == 0


def test_tree_depth():                            # tests.py:723-726, tree_util_tests.py:65-67
    assert o.Node.from_dict({'subtree': {0: {'p': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}}).depth == 2
    assert o.Node.from_dict({'subtree': {0: {'p': 1, 'subtree': {3: {'p': 1, 'subtree': {}}, 4: {'p': 0, 'subtree': {}}}},
                                         1: {'p': 0, 'subtree': {}}}}).depth == 3


def _boards_after(*seqs):
    out = []
    for seq in seqs:
        b, _ = o.game_init(9)
        for x, y in seq:
            o.make_play(x, y, b)
        out.append(b)
    return out


def _checking_eval(expected):
    """tests.py:756-771 DummyModel: asserts the boards presented, policy[:, 0] = 1, value = 1."""
    seen = []

    def ev(X):
        assert X.shape[0] == len(expected)
        for i, e in enumerate(expected):
            assert np.array_equal(X[i:i + 1], e), i
        seen.append(X.shape[0])
        p = np.zeros((X.shape[0], 82), np.float32)
        p[:, 0] = 1
        return p, np.ones(X.shape[0], np.float32)

    return ev, seen


def test_model_evaluation_boards():               # tests.py:747-774: leaves (0,0) and (1,0) are evaluated, in that order
    t = o.Node.from_dict({'subtree': {0: {'p': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}})
    board, _ = o.game_init(9)
    ev, seen = _checking_eval(_boards_after([(0, 0)], [(1, 0)]))
    o.simulate(t, board, ev, 2, 1)
    assert seen == [2]


def test_model_evaluation_nested_boards():        # tests.py:776-850
    t = o.Node.from_dict({'subtree': {0: {'p': 1, 'subtree': {1: {'p': 1, 'subtree': {}}, 2: {'p': 0, 'subtree': {}}}},
                                      1: {'p': 0, 'subtree': {}}}})
    assert t.depth == 3
    board, _ = o.game_init(9)
    ev, seen = _checking_eval(_boards_after([(0, 0), (1, 0)], [(0, 0), (2, 0)]))
    o.simulate(t, board, ev, 2, 1)
    assert seen == [2]


def test_model_evaluation_other_nested_boards():  # tests.py:852-924
    t = o.Node.from_dict({'subtree': {0: {'p': 1, 'subtree': {}},
                                      1: {'p': 0, 'subtree': {0: {'p': 0, 'subtree': {}}, 2: {'p': 1, 'subtree': {}}}}}})
    assert t.depth == 3
    board, _ = o.game_init(9)
    ev, seen = _checking_eval(_boards_after([(0, 0)], [(1, 0), (2, 0)]))
    o.simulate(t, board, ev, 2, 1)
    assert seen == [2]


def test_small_batch_size():                      # tests.py:926-938
    t = o.Node.from_dict({'subtree': {0: {'p': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}})
    board, _ = o.game_init(9)
    o.simulate(t, board, _dummy_eval, 1, 1)
    assert t.child(0).count == 1 and t.child(0).value == -1 and t.child(0).nchild > 0
    assert t.child(1).count == 0 and t.child(1).value == 0 and t.child(1).nchild == 0


def test_get_leaf_by_moves():                     # tree_util_tests.py:126-137
    t = o.Node.from_dict({'subtree': {0: {'p': 1, 'value': 1, 'subtree': {3: {'p': 1, 'value': 1, 'subtree': {}}, 4: {'p': 0, 'subtree': {}}}},
                                      1: {'p': 0, 'subtree': {}}}})
    assert o.get_node_by_moves(t, [0]).move == 0
    n = o.get_node_by_moves(t, [0, 4])
    assert n.move == 4 and n.p == 0
    assert o.get_node_by_moves(t, [1]).move == 1
    with pytest.raises(Exception):
        o.get_node_by_moves(t, [0, 8])


def test_back_prop():                             # tree_util_tests.py:139-193
    t = o.Node.from_dict({'subtree': {0: {'p': 1, 'value': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}})
    o.back_propagation((dict(count=0, p=0.5, value=1, mean_value=0, virtual_loss=0, subtree={}), [0]), t)
    assert t.count == 1 and t.value == 1 and t.mean_value == 1
    assert t.child(0).value == 1 and t.child(0).virtual_loss == 0


class _Dummy(object):
    """tests.py:34-49 DummyModel as a model object (mode A asks model.predict_on_batch)."""
    name = "dummy"

    def predict_on_batch(self, b):
        return _dummy_eval(np.asarray(b))


class _IdentityRng(object):
    """The PlayTestCase setup (tests.py / async_sim_tests.py:1046-1052): SYMMETRIES cut to the identity, a seeded RNG."""

    def __init__(self, seed):
        from oracle import game_loop as gl
        self.inner = gl.SeededRng(seed)

    def __getattr__(self, k):
        return getattr(self.inner, k)

    def symmetry(self):
        return 0


@pytest.mark.parametrize("self_play,sims,stop,num_moves,expect", [
    (True, 8, 30, 5, 1),       # async_sim_tests.py:1059-1071 test_new_tree_called_once_self_play: one shared tree, re-rooted every ply
    (False, 32, 0, 2, 2),      # async_sim_tests.py:1074-1089 test_new_tree_called_twice_evaluation: one tree per model
])
def test_new_tree_call_count(monkeypatch, self_play, sims, stop, num_moves, expect):
    from oracle import game_loop as gl
    calls = []
    real = o.new_tree

    def counting(*a, **k):
        calls.append(1)
        return real(*a, **k)

    monkeypatch.setattr(o, "new_tree", counting)
    m = _Dummy()
    gd = gl.play_game(m, m, sims, stop, self_play=self_play, num_moves=num_moves, size=9, mcts_batch_size=8, rng=_IdentityRng(0))
    assert len(gd['moves']) == num_moves
    assert len(calls) == expect

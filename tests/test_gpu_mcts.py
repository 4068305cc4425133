"""GPU parity: MCTS kernels (csrc/tree.cu) vs fixtures from the reference and vs the
C oracle on fresh seeds, with IDENTICAL injected evaluator outputs (oracle/fake_eval.py)."""
import os
import glob
import numpy as np
import pytest
import torch

from oracle import oracle as o
from oracle.fake_eval import evaluate
from tests.treeio import oracle_rows, engine_rows, rows_equal
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from sejonggo_b200.engine import Engine
    return Engine(**kw)


def eval_slots(e, which, slots, ev, syms=None):
    """Host evaluator on selected positions; returns dense policy/value buffers."""
    lim = e.G * e.L if which else e.G
    pol = np.zeros((lim, e.A), np.float32)
    val = np.zeros(lim, np.float32)
    if len(slots):
        full = None
        if syms is not None:
            full = np.zeros(lim, np.int32)
            full[slots] = syms
        planes = e.export_planes(which, 0, lim, syms=full).cpu().numpy()[slots]
        p, v = ev(planes)
        if syms is not None:
            p = e.policy_unsym(p, syms=syms).cpu().numpy()
        pol[slots] = p
        val[slots] = v
    return pol, val


def step_a(e, ev, batch, syms_per_game=None):
    e.select_a(batch)
    counts = e.leaf_counts().cpu().numpy()
    slots = np.concatenate([g * e.L + np.arange(counts[g]) for g in range(e.G)]).astype(np.int64)
    sy = None if syms_per_game is None else syms_per_game[slots // e.L]
    pol, val = eval_slots(e, 1, slots, ev, sy)
    e.expand(pol, val)
    e.backup_a()
    return counts


def wave_b(e, ev, energy, syms_per_slot=None):
    restart, prev = True, np.zeros(e.G, np.int64)
    phases = 0
    while True:
        newly, stalled = e.select_b(energy, restart)
        restart = False
        if newly == 0:
            break
        counts = e.leaf_counts().cpu().numpy().astype(np.int64)
        slots = np.concatenate([g * e.L + np.arange(prev[g], counts[g]) for g in range(e.G)]).astype(np.int64)
        prev = counts
        pol, val = eval_slots(e, 1, slots, ev, None)
        e.expand(pol, val)
        phases += 1
        if stalled == 0:
            break
    e.backup_b(energy)
    return phases


MCTS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "mcts_*.npz")))


@pytest.mark.parametrize("name", MCTS)
def test_mcts_fixture(name):
    z = np.load(os.path.join(GOLDEN, name))
    S, mode, batch, steps, seed = int(z["size"]), str(z["mode"]), int(z["batch"]), int(z["steps"]), int(z["seed"])
    A = S * S + 1
    ev = lambda b: evaluate(b, seed, True)
    G = 3                                            # the same game three times: batch independence
    e = _engine(size=S, n_games=G, trees_per_game=1, max_leaves=batch, arena_blocks=max(2048, 8 * steps * batch))
    e.reset()
    e.tree_reset()
    pol, _ = eval_slots(e, 0, np.arange(G), ev)
    e.tree_new(pol, noise=np.tile(z["noise"], (G, 1)))
    offs = z["tree_offsets"]
    saw_stall = False
    for ply in range(len(z["picks"])):
        for _ in range(steps):
            if mode == 'a':
                step_a(e, ev, batch)
            else:
                saw_stall |= wave_b(e, ev, batch) > 1
        for g in (0, G - 1):
            blocks, meta, p64 = e.download_tree(g)
            ok, why = rows_equal(engine_rows(blocks, meta, p64, A), z["trees"][offs[ply]:offs[ply + 1]])
            assert ok, "ply %d game %d: %s" % (ply, g, why)
        e.check_errors()
        t0 = e.pick(np.zeros(G, np.int32)).cpu().numpy()
        assert (t0 == z["t0picks"][ply]).all()
        sel = int(z["picks"][ply])
        forced = e.pick(np.ones(G, np.int32), None, np.full(G, sel, np.int32)).cpu().numpy()
        assert (forced == sel).all()
        e.reroot(np.full(G, sel, np.int32))
        e.apply_moves(np.full(G, sel, np.int32))
        valid = e.tree_valid().cpu().numpy()
        assert (valid == (0 if z["newtree"][ply] else 1)).all()
        if z["newtree"][ply]:
            pol, _ = eval_slots(e, 0, np.arange(G), ev)
            e.tree_new(pol)
    if "stall" in name:
        assert saw_stall
    e.check_errors()


@pytest.mark.parametrize("mode,S,batch,steps,plies", [('a', 9, 8, 6, 8), ('b', 9, 8, 6, 8), ('a', 19, 100, 2, 3),
                                                      ('b', 19, 8, 10, 3), ('a', 5, 16, 4, 12), ('b', 4, 8, 4, 14)])
def test_mcts_vs_oracle_with_symmetries(mode, S, batch, steps, plies):
    """Fresh seeds, different evaluator salt per game, random symmetry ids per batch:
    engine vs C oracle tree-for-tree."""
    A, G = S * S + 1, 4
    rs = np.random.RandomState(100 + S + batch)
    e = _engine(size=S, n_games=G, trees_per_game=1, max_leaves=batch, arena_blocks=max(2048, 8 * steps * batch))
    e.reset()
    e.tree_reset()
    salts = [7 * g + 1 for g in range(G)]
    obs, ots = [], []
    noise = rs.dirichlet([0.03] * A, size=G)
    pol0 = np.zeros((G, A), np.float32)
    for g in range(G):
        b, _ = o.game_init(S)
        pol0[g] = evaluate(b, salts[g], True)[0][0]
        obs.append(b)
        ots.append(o.new_tree(pol0[g], b, noise=noise[g]))
    e.tree_new(pol0, noise=noise)

    def ev_all(planes_by_game):
        raise NotImplementedError

    for ply in range(plies):
        for _ in range(steps):
            syms = rs.randint(7, size=G).astype(np.int32)
            # oracle side
            for g in range(G):
                sg = int(syms[g])

                def oev(bb, g=g, sg=sg):
                    sb = o.sym_board(sg, bb)
                    p, v = evaluate(sb, salts[g], True)
                    return o.sym_policy(sg, p, S), v
                op = int(obs[g][0, 0, 0, 16])
                if mode == 'a':
                    o.simulate(ots[g], np.copy(obs[g]), oev, batch, op)
                else:
                    o.async_simulate2(ots[g], np.copy(obs[g]), oev, batch, op)
            # engine side: per-game salt -> evaluate game by game on the exported (symmetric) planes
            def run_eval(slots):
                lim = e.G * e.L
                full = np.zeros(lim, np.int32)
                full[slots] = syms[slots // e.L]
                planes = e.export_planes(1, 0, lim, syms=full).cpu().numpy()
                pol = np.zeros((lim, A), np.float32)
                val = np.zeros(lim, np.float32)
                for s in slots:
                    p, v = evaluate(planes[s:s + 1], salts[s // e.L], True)
                    pol[s], val[s] = p[0], v[0]
                if len(slots):
                    pol[slots] = e.policy_unsym(pol[slots], syms=full[slots]).cpu().numpy()
                return pol, val
            if mode == 'a':
                e.select_a(batch)
                counts = e.leaf_counts().cpu().numpy()
                slots = np.concatenate([g * e.L + np.arange(counts[g]) for g in range(G)]).astype(np.int64)
                pol, val = run_eval(slots)
                e.expand(pol, val)
                e.backup_a()
            else:
                restart, prev = True, np.zeros(G, np.int64)
                while True:
                    newly, stalled = e.select_b(batch, restart)
                    restart = False
                    if newly == 0:
                        break
                    counts = e.leaf_counts().cpu().numpy().astype(np.int64)
                    slots = np.concatenate([g * e.L + np.arange(prev[g], counts[g]) for g in range(G)]).astype(np.int64)
                    prev = counts
                    pol, val = run_eval(slots)
                    e.expand(pol, val)
                    if stalled == 0:
                        break
                e.backup_b(batch)
        picks = np.zeros(G, np.int32)
        e.check_errors()
        for g in range(G):
            blocks, meta, p64 = e.download_tree(g)
            ok, why = rows_equal(engine_rows(blocks, meta, p64, A), oracle_rows(ots[g]))
            assert ok, "ply %d game %d: %s" % (ply, g, why)
            ch = ots[g].children()
            vis = ch['moves'][ch['counts'] > 0]
            picks[g] = o.pick_t0(ots[g]) if ply % 2 == 0 else int(vis[(ply * 5 + g) % len(vis)])
        t0 = e.pick(np.zeros(G, np.int32)).cpu().numpy()
        assert all(t0[g] == o.pick_t0(ots[g]) for g in range(G))
        e.reroot(picks)
        e.apply_moves(picks)
        valid = e.tree_valid().cpu().numpy()
        newpol = np.zeros((G, A), np.float32)
        for g in range(G):
            ots[g] = ots[g].child(int(picks[g])).detach()
            o.make_play(int(picks[g]) % S, int(picks[g]) // S, obs[g])
            assert valid[g] == (1 if ots[g].nchild else 0)
            if not ots[g].nchild:
                newpol[g] = evaluate(obs[g], salts[g], True)[0][0]
                ots[g] = o.new_tree(newpol[g], obs[g])
        e.tree_new(newpol)
        assert np.array_equal(e.export_boards().cpu().numpy(), np.concatenate(obs))
    e.check_errors()


def _upload_dict_tree(e, tree_dict, A):
    """Hand-built reference-style dict tree -> engine arena (tests.py:684-1068 trees)."""
    from sejonggo_b200.engine import NODEBLOCK_DTYPE
    blocks = []

    def build(d, pb, ps):
        idx = len(blocks)
        blk = np.zeros((), dtype=NODEBLOCK_DTYPE)
        blk['child'][:] = -1
        blk['parent_block'], blk['parent_slot'] = pb, ps
        blocks.append(blk)
        for m, c in d['subtree'].items():
            blk['exist'][m >> 5] |= np.uint32(1 << (m & 31))
            blk['prior'][m] = c.get('p', 0)
            blk['n'][m] = c.get('count', 0)
            blk['w'][m] = c.get('value', 0)
            if c.get('subtree'):
                blk['child'][m] = build(c, idx, m)
        return idx

    build(tree_dict, -1, -1)
    e.upload_tree(0, np.array(blocks, dtype=NODEBLOCK_DTYPE), dict(valid=1, root_f64=0))


def _dummy_eval(b):                               # tests.py:34-49 DummyModel
    n = b.shape[0]
    p = np.tile(np.arange(82, 0, -1, dtype=np.float32), (n, 1))
    p /= p.sum(axis=1, keepdims=True)
    return p, np.ones(n, np.float32)


def test_reference_mcts_cases_on_engine():
    """tests.py:729-744 test_leaf, :945-998 nested_selected, :1000-1068 nested_other_leaves and the
    exact leaf boards of :746-943, on hand-built trees uploaded into the arena."""
    S, A = 9, 82
    e = _engine(size=S, n_games=1, trees_per_game=1, max_leaves=2, arena_blocks=32)

    def run(tree):
        e.reset()
        _upload_dict_tree(e, tree, A)
        e.select_a(2)
        n = int(e.leaf_counts().cpu().numpy()[0])
        leaf_boards = e.export_boards  # noqa
        packed = e.export_packed(1, 0, n).cpu().numpy().view(np.uint32)
        pol, val = eval_slots(e, 1, np.arange(n), _dummy_eval)
        e.expand(pol, val)
        e.backup_a()
        blocks, meta, _ = e.download_tree(0)
        return blocks, meta, packed

    def ref_board(moves):
        b, _ = o.game_init(S)
        for x, y in moves:
            o.make_play(x, y, b)
        return o.pack_board(b)

    leaf = {'subtree': {}}
    blocks, meta, packed = run({'subtree': {0: {'p': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}})
    assert list(blocks[0]['n'][:2]) == [1, 1] and list(blocks[0]['w'][:2]) == [-1, -1]
    assert meta['root_count'] == 2 and meta['root_value'] == -2
    assert np.array_equal(packed[0], ref_board([(0, 0)])) and np.array_equal(packed[1], ref_board([(1, 0)]))
    # nested: boards presented are (0,0)->(1,0) and (0,0)->(2,0)  (tests.py:773-833)
    blocks, meta, packed = run({'subtree': {0: {'p': 1, 'subtree': {1: {'p': 1, 'subtree': {}}, 2: {'p': 0, 'subtree': {}}}},
                                            1: {'p': 0, 'subtree': {}}}})
    assert np.array_equal(packed[0], ref_board([(0, 0), (1, 0)])) and np.array_equal(packed[1], ref_board([(0, 0), (2, 0)]))
    # other nested: (0,0) and (1,0)->(2,0)  (tests.py:852-920)
    blocks, meta, packed = run({'subtree': {0: {'p': 1, 'subtree': {}},
                                            1: {'p': 0, 'subtree': {0: {'p': 0, 'subtree': {}}, 2: {'p': 1, 'subtree': {}}}}}})
    assert np.array_equal(packed[0], ref_board([(0, 0)])) and np.array_equal(packed[1], ref_board([(1, 0), (2, 0)]))
    # nested_selected counts/values (tests.py:945-998)
    blocks, meta, _ = run({'subtree': {0: {'p': 1, 'subtree': {1: {'p': 0, 'subtree': {}}, 2: {'p': 1, 'subtree': {}}}},
                                       1: {'p': 0, 'subtree': {}}}})
    c = int(blocks[0]['child'][0])
    assert blocks[0]['n'][0] == 2 and blocks[c]['n'][1] == 1 and blocks[c]['n'][2] == 1 and blocks[0]['n'][1] == 0
    assert blocks[0]['w'][0] == 2 and blocks[0]['w'][1] == 0
    # nested_other_leaves (tests.py:1000-1068)
    blocks, meta, _ = run({'subtree': {0: {'p': .75, 'subtree': {}},
                                       1: {'p': .25, 'subtree': {0: {'p': 1, 'subtree': {}}, 2: {'p': 0, 'subtree': {}}}},
                                       2: {'p': 0, 'subtree': {}}}})
    c = int(blocks[0]['child'][1])
    assert blocks[0]['n'][0] == 1 and blocks[0]['w'][0] == -1 and blocks[0]['w'][1] == 1 and blocks[0]['n'][1] == 1
    assert blocks[c]['n'][0] == 1 and blocks[c]['w'][0] == 1 and blocks[c]['n'][2] == 0
    assert meta['root_count'] == 2 and meta['root_value'] == 0 and blocks[0]['n'][2] == 0 and blocks[0]['child'][2] == -1
    # small batch (tests.py:922-943)
    e.reset()
    _upload_dict_tree(e, {'subtree': {0: {'p': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}}, A)
    e.select_a(1)
    pol, val = eval_slots(e, 1, np.arange(1), _dummy_eval)
    e.expand(pol, val)
    e.backup_a()
    blocks, meta, _ = e.download_tree(0)
    assert blocks[0]['n'][0] == 1 and blocks[0]['w'][0] == -1 and blocks[0]['child'][0] >= 0
    assert blocks[0]['n'][1] == 0 and blocks[0]['child'][1] == -1


def test_reference_find_best_leaf_cases_on_engine():
    """tree_util_tests.py:69-122 (visit order [0,3] -> [0,4] -> [1]; all busy -> None)."""
    S, A = 9, 82
    e = _engine(size=S, n_games=1, trees_per_game=1, max_leaves=4, arena_blocks=32)
    e.reset()
    tree = {'subtree': {0: {'p': 1, 'subtree': {3: {'p': 1, 'subtree': {}}, 4: {'p': 0, 'subtree': {}}}},
                        1: {'p': 0, 'subtree': {}}}}
    _upload_dict_tree(e, tree, A)
    newly, stalled = e.select_b(3, True)
    assert (newly, stalled) == (3, 0)
    packed = e.export_packed(1, 0, 3).cpu().numpy().view(np.uint32)

    def ref_board(moves):
        b, _ = o.game_init(S)
        for m in moves:
            o.make_play(m % S, m // S, b)
        return o.pack_board(b)
    assert np.array_equal(packed[0], ref_board([0, 3])) and np.array_equal(packed[1], ref_board([0, 4]))
    assert np.array_equal(packed[2], ref_board([1]))
    blocks, _, _ = e.download_tree(0)
    assert blocks[0]['busy'][0] & 1                      # node 0 itself went busy when its children ran out
    e.reset()
    _upload_dict_tree(e, {'subtree': {0: {'p': 1, 'subtree': {}}, 1: {'p': 0, 'subtree': {}}}}, A)
    newly, stalled = e.select_b(3, True)
    assert (newly, stalled) == (2, 1)                    # third selection finds (None, None)


def test_pick_temperature1_is_numpy_choice():
    """k_pick at temperature 1 = np.random.choice(moves, p=N/total) (self_play.py:140-149) from the same uniform draw:
    fp64 p, sequential cumulative sum, cdf /= cdf[-1], searchsorted 'right' (tests/test_cpu_choice.py pins that
    algorithm against numpy).  Also temperature 0 = max (count, mean, index)."""
    from tests.test_cpu_choice import choice_from_u
    S, G, B = 9, 24, 8
    e = _engine(size=S, n_games=G, max_leaves=B, arena_blocks=256)
    e.reset()
    ev = lambda planes: evaluate(planes, 3, True)
    mv, _ = e.random_playouts(seed=5, max_plies=6)                      # 24 different roots
    pol, _ = eval_slots(e, 0, np.arange(G), ev)
    e.tree_new(pol, force=True)
    for _ in range(5):
        step_a(e, ev, B)
    _, count, value = e.child_stats(want=("count", "value"))
    count, value = count.cpu().numpy(), value.cpu().numpy()
    rs = np.random.RandomState(7)
    for trial in range(8):
        u = rs.random_sample(G)
        if trial == 0:
            u[:4] = [0.0, 1.0 - 2.0 ** -53, 0.5, 1e-300]               # the ends of [0, 1)
        got = e.pick(np.ones(G, np.int32), u, None).cpu().numpy()
        for g in range(G):
            moves = np.nonzero(count[g])[0]
            assert got[g] == choice_from_u(moves, count[g][moves], u[g]), (trial, g)
    got0 = e.pick(np.zeros(G, np.int32), None, None).cpu().numpy()
    legal = e.legal_masks().cpu().numpy()
    for g in range(G):
        best = max((int(count[g][a]), np.float32(value[g][a]) / np.float32(count[g][a]) if count[g][a] else np.float32(0), a)
                   for a in range(S * S + 1) if legal[g][a] == 0)
        assert got0[g] == best[2], g
    e.check_errors()

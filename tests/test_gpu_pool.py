"""GPU: the shared node pool (csrc/tree.cu: pool_pop / free_subtree / k_reroot) and continuous game turnover
(BatchedGames n_total > n_games; selfplay_worker.py:81-124)."""
import numpy as np
import pytest
import torch

from oracle import game_loop as gl
from oracle.fake_eval import FakeModel, board_key, _mix

pytestmark = pytest.mark.gpu


class PeakedModel(object):
    """Evaluator whose policy puts ~all its mass on one hash-chosen point and whose value is strongly signed:
    the search piles onto one line, so re-rooting keeps almost the whole tree ply after ply."""
    name = "peaked"

    def predict_on_batch(self, X):
        X = np.asarray(X)
        n, S = X.shape[0], X.shape[1]
        A = S * S + 1
        key = board_key(X)
        empty = (X[..., 0] == 0) & (X[..., 1] == 0)
        pol = np.full((n, A), 1e-6, np.float32)
        for i in range(n):
            cells = np.nonzero(empty[i].reshape(-1))[0]
            if len(cells):
                pol[i, cells[int(key[i] % np.uint64(len(cells)))]] = 0.99
            else:
                pol[i, A - 1] = 0.99
        val = ((_mix(key + np.uint64(5)) % np.uint64(2001)).astype(np.float32) - 1000) / 1000
        return pol, val.reshape(-1, 1).astype(np.float32)


def _pool_invariant(bg):
    e = bg.eng
    st = e.pool_stats()
    owned = 0
    for t in range(e.G * e.T):
        blocks, meta, _ = e.download_tree(t)
        if meta['valid']:
            assert meta['n_blocks'] == len(blocks), (t, meta['n_blocks'], len(blocks))    # n_blocks == what a traversal reaches
            owned += meta['n_blocks']
        else:
            assert meta['n_blocks'] == 0
    assert owned + st['free'] == st['capacity'], (owned, st)                               # nothing leaked, nothing double-freed
    return st, owned


def test_peaked_evaluator_long_games_default_sizing_equals_oracle():
    """A sharply peaked evaluator at temperature 0 until the games end (60-160 plies, trees re-used and re-rooted
    every ply) with the DEFAULT pool sizing: no allocation fails, no tree is dropped, every block is accounted for,
    and every game equals the oracle's move for move."""
    from sejonggo_b200.batched import BatchedGames
    S, G, sims, batch = 9, 6, 96, 8
    m = PeakedModel()
    bg = BatchedGames((m, m), G, size=S, mode='a', mcts_batch_size=batch, mcts_simulations=sims, stop_exploration=0,
                      self_play=True, rngs=[gl.SeededRng(40 + g) for g in range(G)], use_symmetry=True, num_moves=2 * S * S)
    # (the oracle's play_game draws one symmetry per simulate batch from the same rng: the engine has to draw them too)
    bg.start()
    plies, sizes = 0, []
    bg.after_search = lambda b, ts: sizes.append(int(b.eng.tree_sizes().max().item()))      # after the search, before the re-root
    while bg.step_ply(record=True):
        plies += 1
        if plies % 20 == 0:
            st, owned = _pool_invariant(bg)
            assert st['failed_allocs'] == 0
    st, owned = _pool_invariant(bg)
    assert st['failed_allocs'] == 0 and bg.trees_dropped == 0 and plies >= 40
    assert max(sizes) > 2 * sims                        # deep re-use: trees several plies' worth of nodes large
    bg.eng.check_errors()
    games = bg.finish()
    assert len(games) == G
    for g, got in enumerate(games):
        mm = PeakedModel()
        ref = gl.play_game(mm, mm, sims, 0, self_play=True, num_moves=2 * S * S, size=S, mcts_batch_size=batch, rng=gl.SeededRng(40 + g))
        assert [x['move'] for x in got['moves']] == [x['move'] for x in ref['moves']], g
        assert got['result'] == ref['result']


def test_tight_pool_drops_largest_trees_instead_of_starving_a_game():
    """The same games in a pool far too small for them: before a ply that could run short the games with the
    largest trees give theirs up (and search from a new tree, as the reference does for an empty one); no allocation
    ever fails, nothing leaks, all games finish."""
    from sejonggo_b200.batched import BatchedGames
    S, G, sims, batch = 9, 6, 96, 8
    m = PeakedModel()
    bg = BatchedGames((m, m), G, size=S, mode='a', mcts_batch_size=batch, mcts_simulations=sims, stop_exploration=0,
                      self_play=True, rng=gl.SeededRng(1), use_symmetry=False, arena_blocks=200, num_moves=2 * S * S)
    bg.start()
    plies = 0
    while bg.step_ply(record=False):
        plies += 1
        if plies % 20 == 0:
            _pool_invariant(bg)
    st, owned = _pool_invariant(bg)
    assert st['failed_allocs'] == 0 and bg.trees_dropped > 0 and plies >= 40
    bg.eng.check_errors()
    assert len(bg.finish()) == G


def test_tree_outgrows_its_share_without_overflow():
    """Two busy games next to idle slots: their trees take more than the per-tree average from the pool."""
    from sejonggo_b200.batched import BatchedGames
    S, G, sims, batch = 9, 8, 128, 16
    m = FakeModel("f", salt=3, sharp=True)
    bg = BatchedGames((m, m), G, size=S, mode='a', mcts_batch_size=batch, mcts_simulations=sims, stop_exploration=0,
                      self_play=True, rng=gl.SeededRng(2), use_symmetry=False, arena_blocks=140, num_moves=30,
                      n_total=2)                     # only two of the eight slots ever play
    sizes = []
    bg.after_search = lambda b, ts: sizes.append(int(b.eng.tree_sizes().max().item()))      # after the search, before the re-root
    bg.start()
    while bg.step_ply(record=False):
        pass
    peak = max(sizes)
    st = bg.eng.pool_stats()
    assert st['failed_allocs'] == 0 and bg.trees_dropped == 0
    assert peak > 140, peak                          # more than the average share: a fixed per-tree arena of that size would have overflowed
    assert len(bg.finish()) == 2


def test_pool_exhaustion_is_reported_at_once():
    from sejonggo_b200 import model
    from sejonggo_b200.batched import BatchedGames, HostRng
    from sejonggo_b200.engine import Engine, EngineError
    S, G = 9, 4
    m = model.TowerModel("m", size=S, n_blocks=1, seed=1, max_positions=64)
    # a pool that cannot hold one ply is refused before the search starts
    bg = BatchedGames((m, m), G, size=S, mode='a', mcts_batch_size=8, mcts_simulations=64, stop_exploration=0,
                      self_play=True, rng=HostRng(3), arena_blocks=24)
    bg.start()
    with pytest.raises(EngineError, match="pool too small"):
        bg.step_ply(record=False)
    # at the ABI level (no host policy in front): the step in which an allocation fails returns the error itself
    e = Engine(size=S, n_games=G, trees_per_game=1, max_leaves=8, arena_blocks=8)       # 32 blocks in all
    m.attach(e, 0)
    e.tree_new(np.full((G, S * S + 1), 1.0 / (S * S + 1), np.float32))
    n1 = e.selfplay_step('a', 6)                     # 4 roots + 24 leaves = 28 blocks
    assert n1 == 24
    with pytest.raises(EngineError, match="pool exhausted"):
        e.selfplay_step('a', 6)
    st = e.pool_stats()
    assert st['failed_allocs'] > 0 and st['capacity'] == 32
    over = [e.download_tree(t)[1]['overflow'] for t in range(G)]
    assert any(over)
    with pytest.raises(EngineError):
        e.check_errors()
    # freeing the trees gives every block back
    e.tree_free(np.ones(G, np.int32))
    assert e.pool_stats()['free'] == 32


@pytest.mark.parametrize("mode", ['a', 'b'])
def test_slot_refill_games_equal_single_games(mode):
    """10 games through 3 slots: a slot starts its next game the ply after one ends (boards reset, trees back in the
    pool).  Every game — including those played in re-used slots — equals the same game from the oracle."""
    from sejonggo_b200 import self_play as sp, nomodel_self_play as nsp, predicting_queue_worker as pq
    from sejonggo_b200.conf import conf
    from oracle.fake_eval import evaluate
    S, batch, sims, N, G = 7, 8, 24, 10, 3
    old = dict(conf)
    started, ended = [], []
    try:
        conf.update(SIZE=S, MCTS_BATCH_SIZE=batch, ENERGY=batch, MCTS_SIMULATIONS=sims, KOMI=5.5)
        model = FakeModel("model_1", salt=23, sharp=True)
        pq.register_models(best=model, latest=model)
        from sejonggo_b200.batched import BatchedGames
        bg = BatchedGames((model, model), G, size=S, mode=mode, mcts_batch_size=batch, energy=batch, mcts_simulations=sims,
                          stop_exploration=3, self_play=True, use_symmetry=(mode == 'a'), n_total=N, resign=(-0.55, -0.55), num_moves=12,
                          rng_for_game=lambda gid: gl.SeededRng(900 + gid),
                          on_game_start=lambda gid: started.append(gid), on_game_end=lambda gid, gd: ended.append(gid))
        games = bg.run()                                # the games resign at different plies, so slots free up at different times
    finally:
        conf.clear()
        conf.update(old)
    assert started == list(range(N)) and sorted(ended) == list(range(N)) and len(games) == N
    assert [g['game_id'] for g in games] == list(range(N))
    st = bg.eng.pool_stats()
    assert st['failed_allocs'] == 0
    for gid, got in enumerate(games):
        m = FakeModel("model_1", salt=23, sharp=True)
        if mode == 'a':
            ref = gl.play_game(m, m, sims, 3, self_play=True, num_moves=12, size=S, mcts_batch_size=batch, rng=gl.SeededRng(900 + gid),
                               resign_model1=-0.55, resign_model2=-0.55)
        else:
            ref = gl.play_game_async("BEST", "BEST", batch, 3, 0, self_play=True, num_moves=12, size=S, conf_sims=sims,
                                     conf_energy=batch, rng=gl.SeededRng(900 + gid), names={"BEST": "model_1"},
                                     resign_model1=-0.55, resign_model2=-0.55,
                                     predict=lambda tag, b, sym: (lambda p, v: (p[0], v[0]))(*evaluate(b, 23, True)))
        assert len(got['moves']) == len(ref['moves']), gid
        for a, b in zip(got['moves'], ref['moves']):
            assert a['move'] == b['move'] and a['player'] == b['player'] and a['move_n'] == b['move_n'], gid
            assert np.array_equal(a['board'], b['board'])
            assert np.array_equal(np.asarray(a['policy']).view(np.uint64), np.asarray(b['policy']).view(np.uint64))
        assert got['result'] == ref['result'] and got['winner'] == ref['winner'] and got['end_reason'] == ref['end_reason']
    assert len(set(len(g['moves']) for g in games)) > 2      # the games really did differ in length

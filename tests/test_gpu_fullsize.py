"""Full-size checks at BASELINE.json configs[2]: 1024 concurrent 19x19 games, 800 sims/ply, through
size-independent invariants (the oracle is far too slow for 820k evaluations per ply)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_config3_one_ply_invariants():
    from sejonggo_b200 import model
    from sejonggo_b200.batched import BatchedGames, HostRng
    from tests.treeio import engine_rows
    G, S, SIMS, B = 1024, 19, 800, 100
    m = model.TowerModel("m", size=S, n_blocks=1, seed=0, max_positions=16384)      # small tower: the search is what is checked
    bg = BatchedGames((m, m), G, size=S, mode='a', mcts_batch_size=B, mcts_simulations=SIMS, stop_exploration=30, self_play=True,
                      rng=HostRng(5), arena_blocks=4 * (SIMS + B), record_boards='packed')
    e = bg.eng
    bg.start()
    for ply in range(2):
        legal_before = e.legal_masks().cpu().numpy()
        sims0 = bg.sim_count
        # run the search of this ply by hand to look at the tree BEFORE the re-root
        assert bg.step_ply(record=True)
        assert bg.sim_count - sims0 == G * SIMS                        # every game expanded exactly 800 leaves
    e.check_errors()
    m.check(e, 0)
    # records: ply-0 policy target = noised root priors over the 362 legal moves, sums to ~1; boards replay
    for g in (0, 511, 1023):
        mv = bg.moves_rec[g]
        assert len(mv) == 2 and mv[0]['player'] == 1 and mv[1]['player'] == 1      # the reference's lagging player field
        assert abs(mv[0]['policy'].sum() - 1.0) < 1e-3 and (mv[0]['policy'] > 0).all()
    # tree invariants after two plies (re-rooted twice) on a sample of games
    prior, count, value = e.child_stats()
    count = count.cpu().numpy()
    legal = e.legal_masks().cpu().numpy()
    valid = e.tree_valid().cpu().numpy()
    assert (valid == 1).all()
    assert ((count > 0) <= (legal == 0)).all()                          # visited children are legal moves
    for g in (0, 300, 1023):
        blocks, meta, p64 = e.download_tree(g)
        assert meta['valid'] and meta['n_blocks'] == len(blocks) and not meta['overflow']
        rows = engine_rows(blocks, meta, p64, S * S + 1)
        # every expanded non-root node: count == 1 + sum(children counts)  (mode A: one visit expands, the rest pass through)
        nb = len(blocks)
        for b in range(1, nb):
            pb, ps = int(blocks[b]['parent_block']), int(blocks[b]['parent_slot'])
            own = int(blocks[pb]['n'][ps])
            ex = np.unpackbits(blocks[b]['exist'].view(np.uint8), bitorder='little')[:S * S + 1].astype(bool)
            assert own == 1 + int(blocks[b]['n'][:S * S + 1][ex].sum()), (g, b)
            assert int(blocks[pb]['child'][ps]) == b
            assert (blocks[b]['busy'] == 0).all()
        # |W| <= N because |value| <= 1
        for b in range(nb):
            assert (np.abs(blocks[b]['w']) <= blocks[b]['n'] + 1e-3).all()
        assert meta['root_count'] == int(count[g].sum()) + 1 or meta['root_count'] == int(count[g].sum())


def test_tower_full_batch_is_position_independent():
    """16,384 positions per forward (the bench's chunk size): every output row depends only on its own position —
    64 distinct positions tiled over the whole batch give bit-identical rows wherever they sit (tiles straddle
    positions, 281 tile waves per pair), and equal the rows of a 64-position forward that the small-batch tests pin to
    the fp32 reference."""
    from sejonggo_b200 import model
    from sejonggo_b200.engine import Engine
    S, N, K = 19, 16384, 64
    e = Engine(size=S, n_games=N, max_leaves=1, arena_blocks=2)
    e.reset(0, K)
    e.random_playouts(seed=11, max_plies=150, first=0, n=K)                   # 64 distinct mid-game positions
    boards = e.export_boards(0, K).cpu().numpy()
    e.import_boards(np.tile(boards, (N // K, 1, 1, 1)))
    m = model.TowerModel("t", size=S, n_blocks=2, seed=3, max_positions=N)
    idx = torch.arange(N, dtype=torch.int32, device=e.device)
    pol, val = m.evaluate(e, 0, idx, None, slot=0)
    m.check(e, 0)
    pol, val = pol.view(N // K, K, -1), val.view(N // K, K)
    assert torch.equal(pol, pol[0:1].expand_as(pol)) and torch.equal(val, val[0:1].expand_as(val))
    m2 = model.TowerModel("t2", params=m.params, max_positions=K)
    p2, v2 = m2.evaluate(e, 0, idx[:K], None, slot=1)
    assert torch.equal(p2, pol[0]) and torch.equal(v2, val[0])
    assert float((pol[0].sum(dim=1) - 1).abs().max()) < 1e-4
    e.close()

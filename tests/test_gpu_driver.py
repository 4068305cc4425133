"""GPU: the native per-step driver (sgo_selfplay_step) and the record packer (sgo_records_pack)
against the same work composed call by call through the finer-grained ABI entries, which the
other GPU tests pin to the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _play(native, mode, self_play, use_symmetry, plies=4, G=12, S=9, sims=32, batch=8, seed=3):
    from sejonggo_b200 import model
    from sejonggo_b200.batched import BatchedGames, HostRng
    m1 = model.TowerModel("m1", size=S, n_blocks=2, seed=1, max_positions=64)       # 64 < G*batch: exercises the chunked forward
    m2 = m1 if self_play else model.TowerModel("m2", size=S, n_blocks=2, seed=2, max_positions=64)
    bg = BatchedGames((m1, m2), G, size=S, mode=mode, mcts_batch_size=batch, energy=batch, mcts_simulations=sims,
                      stop_exploration=2, self_play=self_play, rng=HostRng(seed), use_symmetry=use_symmetry,
                      record_boards='packed', native_step=native)
    assert bg.native_step == native and bg.fast
    bg.start()
    for _ in range(plies):
        bg.step_ply(record=True)
    bg.eng.check_errors()
    trees = [bg.eng.download_tree(t) for t in range(G * bg.eng.T)]
    return bg, trees


@pytest.mark.parametrize("mode,self_play,sym", [('a', True, True), ('a', False, True), ('b', True, False), ('b', False, False),
                                                ('b', True, True), ('b', False, True)])
def test_native_step_equals_composed_calls(mode, self_play, sym):
    a, ta = _play(True, mode, self_play, sym)
    b, tb = _play(False, mode, self_play, sym)
    assert a.sim_count == b.sim_count and a.sim_count > 0
    for g in range(a.G):
        assert [m['move'] for m in a.moves_rec[g]] == [m['move'] for m in b.moves_rec[g]]
        for x, y in zip(a.moves_rec[g], b.moves_rec[g]):
            assert np.array_equal(x['board'], y['board'])
            assert np.array_equal(x['policy'].view(np.uint32), y['policy'].view(np.uint32))
            assert np.float32(x['value']).view(np.uint32) == np.float32(y['value']).view(np.uint32)
    for (ba, ma, pa), (bb, mb, pb) in zip(ta, tb):
        assert ma == mb
        assert ba.tobytes() == bb.tobytes()
        if ma['root_f64']:                                   # the fp64 side array only means something at a noised root
            assert np.array_equal(pa.view(np.uint64), pb.view(np.uint64))


def test_records_pack_matches_fine_grained_exports():
    from sejonggo_b200 import model
    from sejonggo_b200.batched import BatchedGames, HostRng
    G, S = 10, 9
    m = model.TowerModel("m", size=S, n_blocks=1, seed=4, max_positions=128)
    bg = BatchedGames((m, m), G, size=S, mode='a', mcts_batch_size=8, mcts_simulations=16, stop_exploration=30,
                      self_play=True, rng=HostRng(9), record_boards='packed')
    e = bg.eng
    bg.start()
    for _ in range(3):
        bg.step_ply(record=True)
    # a fresh search on the current position, then compare the packed rows with the separate exports
    ts = torch.zeros(G, dtype=torch.int32, device=e.device)
    ts[3] = -1                                                      # one game without a tree this ply
    moves = torch.arange(G, dtype=torch.int32, device=e.device) * 3
    values = torch.linspace(-1, 1, G, device=e.device)
    rec = e.records_pack(ts, moves, values).cpu().numpy().view(np.uint32)
    PW, A = e.packed_words, e.A
    assert rec.shape == (G, e.record_words()) and e.record_words() == PW + 3 + A
    packed = e.export_packed(0).cpu().numpy().view(np.uint32)
    prior, _, _ = e.child_stats(ts.clamp(min=0), want=("prior",))
    prior = prior.cpu().numpy().astype(np.float32)
    valid = e.tree_valid(ts.clamp(min=0)).cpu().numpy()
    assert np.array_equal(rec[:, :PW], packed)
    assert np.array_equal(rec[:, PW].view(np.int32), moves.cpu().numpy())
    assert np.array_equal(rec[:, PW + 1], values.cpu().numpy().view(np.uint32))
    for g in range(G):
        want_valid = int(valid[g] == 1 and g != 3)
        assert rec[g, PW + 2] == want_valid
        want = prior[g] if want_valid else np.zeros(A, np.float32)
        assert np.array_equal(rec[g, PW + 3:].view(np.float32).view(np.uint32), want.view(np.uint32)), g


def test_selfplay_step_rejects_missing_weights():
    from sejonggo_b200.engine import Engine, EngineError
    e = Engine(size=9, n_games=2, trees_per_game=1, max_leaves=4, arena_blocks=64)
    e.tree_new(np.full((2, 82), 1 / 82, np.float32))
    with pytest.raises(EngineError):
        e.selfplay_step('a', 4)

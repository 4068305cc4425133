"""GPU: the device record store (records.RecordStore, record_boards='device') and slot refill on the NATIVE search path
(TowerModel evaluators, sgo_selfplay_step)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rng_factory(base):
    from sejonggo_b200.batched import HostRng

    def make(gid):
        return HostRng(base + gid)

    make.device_pick = True                      # per-game rngs that allow the device-side pick keep the native path
    return make


def test_device_rows_equal_host_records():
    """The same games recorded twice: rows kept in HBM and rebuilt with games_from_rows vs the per-ply host copies
    ('packed'): boards, moves, values, policy targets, results — bit for bit, with games ending at different plies."""
    from sejonggo_b200 import model, records
    from sejonggo_b200.batched import BatchedGames
    S, G, N = 9, 4, 9
    m = model.TowerModel("model_5", size=S, n_blocks=1, seed=2, max_positions=128)
    lens = [3, 5, 2, 6, 4, 1, 5, 3, 2]

    def play(rec):
        bg = BatchedGames((m, m), G, size=S, mode='a', mcts_batch_size=8, mcts_simulations=16, stop_exploration=0, self_play=True,
                          use_symmetry=False, rng_for_game=_rng_factory(300), record_boards=rec, n_total=N,
                          on_game_start=lambda gid: dict(num_moves=lens[gid]))
        assert bg.native_step
        games = bg.run()
        return bg, games

    bg_h, host = play('packed')
    bg_d, dev = play('device')
    rows = bg_d.store.take().cpu().numpy().view(np.uint32)
    assert rows.shape == (sum(len(g['moves']) for g in host) + N, records.row_words(S))
    rebuilt = records.games_from_rows(rows, S, names=("model_5", "model_5"))
    assert sorted(rebuilt) == list(range(N))
    for gid in range(N):
        a, b = host[gid], rebuilt[gid]
        assert len(a['moves']) == len(b['moves']) == lens[gid] or a['end_reason'] == 'BOTH_PASSED'
        assert a['result'] == b['result'] and a['winner'] == b['winner'] and a['end_reason'] == b['end_reason']
        assert a['model1_isblack'] == b['model1_isblack'] and a['game_id'] == b['game_id'] == gid
        for x, y in zip(a['moves'], b['moves']):
            assert x['move'] == y['move'] and x['move_n'] == y['move_n'] and x['player'] == y['player']
            assert np.float32(x['value']).view(np.uint32) == np.float32(y['value']).view(np.uint32)
            assert np.array_equal(np.asarray(x['board'], np.uint32), y['board'])
            assert np.array_equal(np.asarray(x['policy'], np.float32).view(np.uint32), y['policy'].view(np.uint32))
    # the light host records of the device mode still carry what the game flow and the calibration need
    for gid in range(N):
        assert [mv['move'] for mv in dev[gid]['moves']] == [mv['move'] for mv in host[gid]['moves']]
        assert all(mv['board'] is None and mv['policy'] is None for mv in dev[gid]['moves'])


def test_native_refill_games_equal_games_played_alone():
    """Seven games through three slots on the native path (device evaluators, sgo_selfplay_step, slots restarted as
    their games end at different plies) equal the same seven games each played alone in a one-slot engine."""
    from sejonggo_b200 import model
    from sejonggo_b200.batched import BatchedGames
    S, N = 9, 7
    m = model.TowerModel("model_6", size=S, n_blocks=2, seed=4, max_positions=256)
    lens = [4, 2, 6, 3, 5, 1, 4]

    def play(G, ids):
        sub = [lens[i] for i in ids]
        base = _rng_factory(500)
        f = lambda k: base(ids[k])
        f.device_pick = True
        bg = BatchedGames((m, m), G, size=S, mode='b', energy=8, mcts_simulations=32, stop_exploration=0, self_play=True,
                          use_symmetry=False, rng_for_game=f, record_boards='packed', n_total=len(ids),
                          on_game_start=lambda k: dict(num_moves=sub[k]))
        assert bg.native_step
        out = bg.run()
        assert bg.eng.pool_stats()['failed_allocs'] == 0
        return out

    together = play(3, list(range(N)))
    for gid in range(N):
        alone = play(1, [gid])[0]
        got = together[gid]
        assert len(got['moves']) == len(alone['moves'])
        for x, y in zip(got['moves'], alone['moves']):
            assert x['move'] == y['move'] and x['player'] == y['player']
            assert np.float32(x['value']).view(np.uint32) == np.float32(y['value']).view(np.uint32)
            assert np.array_equal(x['board'], y['board'])
            assert np.array_equal(np.asarray(x['policy']).view(np.uint32), np.asarray(y['policy']).view(np.uint32))
        assert got['result'] == alone['result']

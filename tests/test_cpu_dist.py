"""world_size-2 gloo tests of the multi-GPU host logic: game sharding, the weight-blob broadcast, and the
gather of full record rows to rank 0, which rebuilds another rank's game and writes the same sample files that
rank wrote locally."""
import os
import sys
import subprocess
import textwrap
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np, torch
    from sejonggo_b200 import dist as sd, model, records, sgfsave
    from sejonggo_b200.conf import conf
    from oracle import game_loop as gl, oracle as o
    from oracle.fake_eval import FakeModel
    out = %(out)r
    rank, world, local = sd.init(backend="gloo")
    assert world == 2
    ids = sd.shard_games(7, rank, world)
    assert list(ids) == list(range(rank, 7, 2))

    # ---- weight blob: rank 0's network reaches rank 1 bit for bit, without rank 1 knowing its shape in advance
    S = 9
    mine = model.TowerModel("model_%%d" %% (7 + rank), params=model.init_params(S, 2, seed=rank, randomize_bn=True, random_bias=True))
    got = sd.broadcast_model(mine if rank == 0 else None, src=0, max_positions=64)
    ref = model.folded_arrays(model.init_params(S, 2, seed=0, randomize_bn=True, random_bias=True))
    assert got.name == "model_7" and got.size == S and got.n_blocks == 2
    f = got.folded()
    assert sorted(f) == sorted(ref)
    for k in ref:
        a, b = f[k].contiguous(), ref[k].contiguous()
        assert a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape), k
        assert torch.equal(a.view(torch.uint8), b.view(torch.uint8)), k
    assert sd.blob_bytes(got) == model.blob_layout(S, 2)[1] > 2 * 2 * 9 * 256 * 256 * 2

    # ---- records: each rank plays its own games (oracle, CPU), keeps the rows in a store, rank 0 gathers them
    conf['SELF_PLAY_DIR'] = os.path.join(out, "gathered")
    S = 5
    store = records.RecordStore(S, 8, "cpu")
    local_games = {}
    for local_id in range(2 + rank):
        m = FakeModel("model_7", salt=10 * rank + local_id, sharp=True)
        gd = gl.play_game(m, m, 16, 2, self_play=True, num_moves=5 + local_id, size=S, mcts_batch_size=4, rng=gl.SeededRng(rank * 100 + local_id))
        gd['model1_isblack'] = True
        local_games[local_id] = gd
        store.append_host_rows(records.rows_from_game_data(S, local_id, gd, o.pack_board))
    # a game still in flight: plies but no footer yet -> must not be reported as finished
    partial = records.rows_from_game_data(S, 99, local_games[0], o.pack_board)[:2]
    store.append_host_rows(partial)
    rows = store.take()
    assert store.n == 0 and rows.shape[1] == records.row_words(S)
    got = sd.gather_rows(rows, dst=0)
    if rank == 1:
        assert got is None
        conf['SELF_PLAY_DIR'] = os.path.join(out, "local_rank1")
        for local_id, gd in local_games.items():
            sgfsave.save_self_play_data("model_7", local_id, gd, size=S)
    else:
        assert len(got) == 2 and [int(t.shape[0]) for t in got] == [sum(len(g['moves']) + 1 for g in local_games.values()) + 2, got[1].shape[0]]
        games1 = records.games_from_rows(got[1].numpy().view(np.uint32), S, names=("model_7", "model_7"))
        assert sorted(games1) == [0, 1, 2]                     # rank 1 played three; its partial game 99 is not there
        for local_id, gd in games1.items():
            sgfsave.save_self_play_data("model_7", 10 + local_id, gd, size=S)
        games0 = records.games_from_rows(got[0].numpy().view(np.uint32), S, names=("model_7", "model_7"))
        for local_id, gd in games0.items():
            ref_gd = local_games[local_id]
            assert gd['result'] == ref_gd['result'] and gd['winner'] == ref_gd['winner'] and gd['end_reason'] == ref_gd['end_reason']
            assert [m['move'] for m in gd['moves']] == [m['move'] for m in ref_gd['moves']]
            assert [m['player'] for m in gd['moves']] == [m['player'] for m in ref_gd['moves']]
    torch.distributed.barrier()
    if rank == 0:
        # what rank 0 wrote for rank 1's games == what rank 1 wrote itself
        n = 0
        for local_id in range(3):
            a = os.path.join(out, "gathered", "model_7", "game_%%05d" %% (10 + local_id))
            b = os.path.join(out, "local_rank1", "model_7", "game_%%05d" %% local_id)
            assert sorted(os.listdir(a)) == sorted(os.listdir(b)) and len(os.listdir(a)) >= 1
            for mv in os.listdir(a):
                za, zb = np.load(os.path.join(a, mv, "sample.npz")), np.load(os.path.join(b, mv, "sample.npz"))
                for k in ("board", "policy_target", "value_target"):
                    assert za[k].dtype == zb[k].dtype and np.array_equal(za[k], zb[k]), (local_id, mv, k)
                n += 1
        assert n >= 5
    d = sd.gather_per_rank([float(rank), 2.5])
    assert d.shape == (2, 2) and d[1, 0] == 1.0
    sys.stdout.write("rank %%d ok\\n" %% rank)
    sys.stdout.flush()
''')


def _launch(script):
    import socket
    with socket.socket() as sk:                      # a free port: a fixed one can still be in TIME_WAIT from the last run
        sk.bind(("127.0.0.1", 0))
        port = str(sk.getsockname()[1])
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                          env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)


def test_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % dict(root=ROOT, out=str(tmp_path)))
    r = _launch(script)
    if r.returncode != 0 and "Address already in use" in r.stdout:       # the probed port was taken in between: once more
        r = _launch(script)
    if r.returncode != 0 or "rank 0 ok" not in r.stdout:
        with open("/tmp/sgo_dist_fail.log", "w") as f:
            f.write(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout, r.stdout[-3000:]

"""world_size-2 gloo tests of the multi-GPU host logic (sharding, weight broadcast, record gather)."""
import os
import sys
import subprocess
import textwrap
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch
    from sejonggo_b200 import dist as sd, model
    rank, world, local = sd.init(backend="gloo")
    assert world == 2
    ids = sd.shard_games(7, rank, world)
    assert list(ids) == list(range(rank, 7, 2))
    p = model.init_params(9, 1, seed=rank)            # different weights per rank before the broadcast
    ref = model.init_params(9, 1, seed=0)
    sd.broadcast_params(p, src=0)
    for (k, kk), t in sd._flat_items(p):
        r = ref[k] if kk is None else ref[k][kk]
        assert torch.equal(t, r), (k, kk)
    games = [dict(moves=[dict(move=(g, rank + 1)), dict(move=(0, 9))], winner=[1, 0, None][(g + rank) %% 3]) for g in range(2 + rank)]
    got = sd.gather_records(sd.pack_records(games), dst=0)
    if rank == 0:
        assert len(got) == 2
        all_games = [sd.unpack_records(b) for b in got]
        assert [len(x) for x in all_games] == [2, 3]
        assert all_games[1][0]['moves'] == [(0, 2), (0, 9)] and all_games[0][1]['winner'] == 0
    else:
        assert got is None
    print("rank", rank, "ok")
''')


def test_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    import socket
    with socket.socket() as sk:                      # a free port: a fixed one can still be in TIME_WAIT from the last run
        sk.bind(("127.0.0.1", 0))
        port = str(sk.getsockname()[1])
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout

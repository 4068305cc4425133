"""Board/tree kernels at the BASELINE configs[2] scale (1024 games, 19x19, trees grown by 800-sim
plies), for `ncu` (profiles/): two warm plies, then ONE profiled ply per mode between
cudaProfilerStart/Stop (run ncu with --profile-from-start off).  The tower is cut to one residual
block so the run is short; the conv kernel is profiled separately (tools/bench_tower.py).

    python tools/prof_kernels.py [games] > gpurun_out/prof_units.json
"""
import json
import sys
import time
import torch

sys.path.insert(0, ".")
from sejonggo_b200 import model
from sejonggo_b200.batched import BatchedGames, HostRng
from sejonggo_b200.engine import Engine


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = 19
    units = dict(games=G, size=S)
    m = model.TowerModel("m", size=S, n_blocks=1, seed=0, max_positions=16384)
    for mode, warm_sims, prof_sims, batch in (('a', 800, 100, 100), ('b', 800, 16, 8)):
        bg = BatchedGames((m, m), G, size=S, mode=mode, mcts_batch_size=100, energy=8, mcts_simulations=warm_sims,
                          stop_exploration=30, self_play=True, rng=HostRng(7), arena_blocks=4 * 900, record_boards='packed')
        bg.start()
        for _ in range(2):
            bg.step_ply(record=False)
        bg.sims = prof_sims
        torch.cuda.synchronize()
        s0 = bg.sim_count
        t0 = time.time()
        torch.cuda.profiler.start()
        bg.step_ply(record=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        units["mode_%s" % mode] = dict(leaves=bg.sim_count - s0, steps=prof_sims // batch, ply_s=time.time() - t0)
        bg.eng.check_errors()
        bg.eng.close()
        del bg
    # rules kernels on 4096 mid-game positions (configs[1])
    R = 4096
    e = Engine(size=S, n_games=R, max_leaves=1, arena_blocks=2)
    e.reset()
    e.random_playouts(seed=1, max_plies=200)
    mv = torch.full((R,), S * S, dtype=torch.int32, device=e.device)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e.apply_moves(mv)
    e.legal_masks()
    e.score()
    e.export_planes(0, 0, R, sym=4)
    e.export_boards()
    e.export_packed(0)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    units["rules_games"] = R
    print(json.dumps(units))


if __name__ == "__main__":
    main()

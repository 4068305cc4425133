#!/usr/bin/env python
"""bench.py — self-play MCTS throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode a|b]

A "step" is one ply of EVERY concurrent game: root evaluation, new trees where needed,
MCTS_SIMULATIONS simulations (mode A: sims/100 batched top-N steps, self_play.py:28-152;
mode B: sims/8 virtual-loss waves, nomodel_self_play.py:59-140), move pick, re-root and
make_play.  Workload = BASELINE.json configs[2]: 19x19, conf.py default tower (20 blocks x
256 ch, random init), 800 sims/ply, 1024 concurrent games per GPU (weak scaling: configs[4]
is 8 x 1024).  `value` = simulations (leaves expanded + backed up) per second summed over
GPUs with every input resident in HBM; `e2e` = the same through the public batched API with
the per-ply records (packed board, policy target, value, move) copied to host memory and the
host-drawn noise / sampling inputs copied to the device inside the timed region.

--impl reference times the reference's CPU path: /root/reference is pure Python and absent on
the GPU box, so this runs the pinned C oracle port (oracle/go_oracle.c, kind "port") of the
same self-play loop with a free uniform evaluator on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE, SIMS, BATCH_A, ENERGY_B, BLOCKS, GAMES_PER_GPU = 19, 800, 100, 8, 20, 1024
FLOP_PER_CONV_POS = 2.0 * 17 * 17 * 256 * 256 * 9           # one 3x3 256->256 layer, one position


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]), src="measured (sustained)")
    except Exception:
        return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace('.', '').isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------ CPU arm
class _Uniform(object):
    """Free evaluator: policy 1/(S*S+1), value 0 — the CPU number excludes network cost."""
    name = "uniform"

    def predict_on_batch(self, X):
        import numpy as np
        X = np.asarray(X)
        n, S = X.shape[0], X.shape[1]
        return np.full((n, S * S + 1), 1.0 / (S * S + 1), np.float32), np.zeros((n, 1), np.float32)


def _cpu_worker(args):
    seed, mode, plies = args
    from oracle import game_loop as gl
    m = _Uniform()
    t0 = time.time()
    if mode == 'a':
        gd = gl.play_game(m, m, SIMS, 30, self_play=True, num_moves=plies, size=SIZE, mcts_batch_size=BATCH_A, rng=gl.SeededRng(seed))
    else:
        gd = gl.play_game_async("BEST_SYM", "BEST_SYM", ENERGY_B, 30, 0, self_play=True, num_moves=plies, size=SIZE,
                                conf_sims=SIMS, conf_energy=ENERGY_B, rng=gl.SeededRng(seed),
                                predict=lambda tag, b, sym: (lambda p, v: (p[0], v[0]))(*gl.sym_predict(m, b, sym)))
    return len(gd['moves']), time.time() - t0


def cpu_selfplay(mode, plies_per_game, rounds=1):
    """All host cores, one game per process (the reference's own parallelism, main_selfplay.py:26)."""
    import multiprocessing as mp
    from oracle import oracle as o
    o.build()
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(1000 + i, mode, plies_per_game) for i in range(cores * rounds)])
    wall = time.time() - t0
    plies = sum(r[0] for r in res)
    per = SIMS // (BATCH_A if mode == 'a' else ENERGY_B) * (BATCH_A if mode == 'a' else ENERGY_B)
    return dict(sims=plies * per, plies=plies, wall=wall, cores=cores)


CPU_PLIES_PER_STEP = 10      # uniform-evaluator side measurement: 10 plies of one game on every host core


class _CpuTower(object):
    """The network of the SAME config on the host cores: oracle/tower_ref.py (fp32 torch restatement of
    model.py:55-96) with all intra-op threads — the reference's CPU path includes evaluating its tower."""
    name = "cpu_tower"

    def __init__(self, blocks):
        from sejonggo_b200 import model
        self.params = model.init_params(SIZE, blocks, seed=0)
        self.evals = 0

    def predict_on_batch(self, X):
        import numpy as np
        import torch
        from oracle import tower_ref
        with torch.no_grad():
            p, v = tower_ref.forward(self.params, np.asarray(X, dtype=np.float32))
        self.evals += len(X)
        return p.numpy(), v.numpy()


def cpu_tower_selfplay(mode, blocks, warm, steps):
    """One game on the host: W untimed + K timed search steps (mode A: one 100-leaf simulate batch;
    mode B: one ENERGY=8 wave) through the oracle port with the CPU tower as evaluator."""
    import numpy as np
    import torch
    from oracle import oracle as o, game_loop as gl
    torch.set_num_threads(os.cpu_count() or 1)       # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
    m = _CpuTower(blocks)
    board, _ = o.game_init(SIZE)
    p, _ = m.predict_on_batch(board)
    tree = o.new_tree(p[0], board, noise=np.random.RandomState(0).dirichlet([0.03] * (SIZE * SIZE + 1)))
    rng = gl.SeededRng(0)
    per = BATCH_A if mode == 'a' else ENERGY_B

    def one():
        sym = rng.symmetry()
        ev = lambda b: gl.sym_predict(m, b, sym)
        if mode == 'a':
            return o.simulate(tree, np.copy(board), ev, BATCH_A, 1)
        return o.async_simulate2(tree, np.copy(board), ev, ENERGY_B, 1)

    for _ in range(warm):
        one()
    t0 = time.time()
    sims = sum(one() for _ in range(steps))
    wall = time.time() - t0
    return dict(sims=sims, wall=wall, cores=os.cpu_count() or 1, threads=torch.get_num_threads(), per=per)


def _cpu_baseline_dict(mode, blocks, warm, steps, with_uniform=True):
    r = cpu_tower_selfplay(mode, blocks, warm, steps)
    cb = dict(value=r['sims'] / r['wall'], unit="simulations/s", cores=r['threads'], kind="port",
              sample="1 game, %d timed %s of the same workload (19x19, %d-block tower evaluated in fp32 by torch on %d host threads, "
                     "oracle C port for rules/tree); the reference's real evaluator was a TF1.7 GPU process"
                     % (steps, "100-leaf simulate batches (mode A)" if mode == 'a' else "8-leaf waves (mode B)", blocks, r['threads']))
    if with_uniform:
        cpu_selfplay(mode, 1)
        c = cpu_selfplay(mode, CPU_PLIES_PER_STEP)
        cb["rules_tree_only"] = dict(value=c['sims'] / c['wall'], unit="simulations/s", cores=c['cores'],
                                     note="free uniform evaluator (network cost excluded), one game per core, C oracle port; "
                                          "the pure-Python reference measured 179 sims/s/core in the survey container")
    return cb, r


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, a.steps), a.warmup
    cb, r = _cpu_baseline_dict(a.mode, a.blocks, warm, steps, with_uniform=False)
    val = cb['value']
    print(json.dumps(dict(
        impl="reference", metric="selfplay_mcts_simulations_per_sec", value=val, unit="simulations/s", n_gpus=a.gpus,
        steps=steps, warmup=warm, ms_per_step=1e3 * r['wall'] / steps, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload="19x19 self-play, conf.py default tower (%d blocks x 256 ch, random init), %d sims/ply, mode %s, "
                             "1 game on all host cores" % (a.blocks, SIMS, a.mode.upper()),
                    moves_per_sec=val / SIMS, step="one %d-leaf search step of one game" % r['per']),
        cpu_baseline=cb, e2e=dict(value=val, unit="simulations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


# ------------------------------------------------------------------ GPU arm
def run_ours(a):
    import numpy as np
    import torch
    from sejonggo_b200 import dist as sd, model
    from sejonggo_b200.batched import BatchedGames, HostRng
    rank, world, local = sd.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    G = a.games
    batch = BATCH_A if a.mode == 'a' else ENERGY_B
    params = model.init_params(SIZE, a.blocks, seed=0 if rank == 0 else 1 + rank)
    if world > 1:
        sd.broadcast_params(params, src=0, device=dev)           # NCCL weight broadcast
    m = model.TowerModel("model_1", params=params, max_positions=a.max_positions)
    if a.match:
        # BASELINE.json configs[3] / SURVEY §8d config 4: evaluator.evaluate semantics — two weight sets, two trees
        # per game, no noise, temperature 0 from ply 0, 1600 sims/ply, one random symmetry per predict batch
        params2 = model.init_params(SIZE, a.blocks, seed=1)
        m2 = model.TowerModel("model_2", params=params2, max_positions=a.max_positions)
        arena = a.arena or 4 * (a.sims + batch)
        bg = BatchedGames((m, m2), G, size=SIZE, mode=a.mode, mcts_batch_size=BATCH_A, energy=ENERGY_B, mcts_simulations=a.sims,
                          stop_exploration=0, self_play=False, rng=HostRng(1234 + rank), arena_blocks=arena, device=local,
                          record_boards='packed')
    else:
        m2 = m
        arena = a.arena or 8 * (SIMS + batch)
        bg = BatchedGames((m, m), G, size=SIZE, mode=a.mode, mcts_batch_size=BATCH_A, energy=ENERGY_B, mcts_simulations=a.sims,
                          stop_exploration=30, self_play=True, rng=HostRng(1234 + rank), arena_blocks=arena, device=local,
                          record_boards='packed')
    e = bg.eng
    m.attach(e, 0)
    if a.match:
        m2.attach(e, 1)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()

    def timed(n_steps, record):
        sync_all()
        s0, p0, l0 = bg.sim_count, bg.plies_done, e.launch_count()
        h0, d0 = bg.h2d_bytes, bg.d2h_bytes
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n_steps):
            bg.step_ply(record=record)
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        cnt = torch.tensor([bg.sim_count - s0, bg.plies_done - p0], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)      # max over ranks, device timed
            torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
        return dict(ms=float(t.item()), sims=float(cnt[0].item()), plies=float(cnt[1].item()), launches=e.launch_count() - l0,
                    h2d=bg.h2d_bytes - h0, d2h=bg.d2h_bytes - d0)

    bg.start()
    for _ in range(a.warmup):
        bg.step_ply(record=False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    m.profile(e, 0, True)
    if a.match:
        m2.profile(e, 1, True)
    r = timed(a.steps, record=False)
    prof = m.profile_read(e, 0)
    m.profile(e, 0, False)
    if a.match:
        p2 = m2.profile_read(e, 1)
        m2.profile(e, 1, False)
        prof = {k: prof[k] + p2[k] for k in prof}
    clocks = sampler.stop() if rank == 0 else None
    r2 = timed(a.steps, record=True)                               # end-to-end through the public batched API
    e.check_errors()
    m.check(e, 0)
    records = sd.pack_records(bg.finish())
    gathered = sd.gather_records(records.to(dev) if world > 1 else records, dst=0, device=dev if world > 1 else None)
    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    peaks = measured_peaks()
    conv_flops = prof['positions'] * FLOP_PER_CONV_POS * 2 * a.blocks
    conv_s = prof['conv_ms'] * 1e-3
    achieved = conv_flops / conv_s / 1e12 if conv_s > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    out = dict(
        metric="selfplay_mcts_simulations_per_sec", value=r['sims'] / (r['ms'] * 1e-3), unit="simulations/s", n_gpus=world,
        steps=a.steps, warmup=a.warmup, ms_per_step=r['ms'] / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="bf16", data="synthetic",
        config=dict(workload=("19x19 %s, conf.py default tower (%d blocks x 256 ch, random init), %d sims/ply, %d concurrent "
                              "games per GPU, mode %s") % ("match play (evaluator.evaluate: two weight sets, two trees per game, "
                                                           "temperature 0, random symmetry per batch)" if a.match else "self-play",
                                                           a.blocks, a.sims, G, a.mode.upper()),
                    games_per_gpu=G, mode=a.mode, sims_per_ply=a.sims, moves_per_sec=r['plies'] / (r['ms'] * 1e-3),
                    l2="working set (node pool %.1f GB, activations %.1f GB) >> 126 MB L2; no flush needed" %
                       (G * bg.eng.T * arena * 6272 / 1e9, 3 * (a.max_positions * 18 + 1) * 18 * 512 / 1e9),
                    step="one ply of every game", parallelism="games sharded, %d rank(s)" % world,
                    records_gathered=None if gathered is None else [int(g.numel()) for g in gathered]),
        e2e=dict(value=r2['sims'] / (r2['ms'] * 1e-3), unit="simulations/s", h2d_bytes_per_step=r2['h2d'] / a.steps,
                 d2h_bytes_per_step=r2['d2h'] / a.steps, moves_per_sec=r2['plies'] / (r2['ms'] * 1e-3)),
        gpu_launches=r['launches'],
        clocks=clocks,
        roofline=dict(bound="tensor", kernel="k_conv3x3_pair", achieved=achieved, peak=peaks['tflops'], unit="TFLOP/s",
                      frac=achieved / peaks['tflops'], traffic=traffic, peak_source=peaks['src'],
                      launches=prof['conv_launches'], avg_launch_ms=prof['conv_ms'] / max(1, prof['conv_launches']),
                      share_of_step=prof['conv_ms'] / r['ms'], stem_ms=prof['stem_ms'], heads_ms=prof['heads_ms']),
    )
    if world == 1 and not a.no_cpu:
        out["cpu_baseline"], _ = _cpu_baseline_dict(a.mode, a.blocks, 1, 16)    # ~10-15 s of CPU work (bounded sample)
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="a", choices=["a", "b"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--sims", type=int, default=None)
    ap.add_argument("--match", action="store_true", help="SURVEY §8d config 4: match play, two models, 1600 sims/ply")
    ap.add_argument("--blocks", type=int, default=BLOCKS)
    ap.add_argument("--max-positions", type=int, default=16384)
    ap.add_argument("--arena", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.sims is None:
        a.sims = 1600 if a.match else SIMS
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

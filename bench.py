#!/usr/bin/env python
"""bench.py — self-play MCTS throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode a|b] [--match] [--full-games]

A "step" is one ply of EVERY concurrent game: root evaluation, new trees where needed,
MCTS_SIMULATIONS simulations (mode A: sims/100 batched top-N steps, self_play.py:28-152;
mode B: sims/8 virtual-loss waves, nomodel_self_play.py:59-140), move pick, re-root and
make_play.  Workload = BASELINE.json configs[2]: 19x19, conf.py default tower (20 blocks x
256 ch, random init), 800 sims/ply, 1024 concurrent games per GPU (weak scaling: configs[4]
is 8 x 1024).  `value` = simulations (leaves expanded + backed up) per second summed over
GPUs with every input resident in HBM; `e2e` = the same through the public batched API with
the per-ply records (packed board, policy target, value, move) copied to host memory and the
host-drawn noise / sampling inputs copied to the device inside the timed region.

The line also carries `other_modes`: the same measurement, short, for mode B (configs[2], 100 ENERGY-8 waves per
ply) and for match play (configs[3]: evaluator.evaluate, two networks, 1600 sims/ply), so that the driver times them
too; `per_rank`: every rank's own step time, conv time and SM clock (N > 1); and `cpu_baseline`, measured by a
SEPARATE process (the GPU arm never loads the oracle).

--impl reference times the reference's CPU path: /root/reference is pure Python and absent on
the GPU box, so this runs the pinned C oracle port (oracle/go_oracle.c, kind "port") of the
same self-play loop with the same tower evaluated in fp32 by torch on the host cores, in two shapes — one game
with every host thread in the tower, and the reference's own shape (main_selfplay.py:26: one game per process on
every core) — and reports the faster.

--full-games plays WHOLE games with slot refill (a slot starts its next game when one ends) and reports whole-run
simulations/s and games/h beside the steady-state figure.
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
        pw = sorted(float(r[3]) for r in self.rows if len(r) > 3 and r[3].replace('.', '').isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), power_w=pw[len(pw) // 2] if pw else None)


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
    """All host cores, one game per process (the reference's own parallelism, main_selfplay.py:26), FREE evaluator."""
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
    model.py:55-96) — the reference's CPU path includes evaluating its tower."""
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


def cpu_tower_selfplay(mode, blocks, warm, steps, threads=None):
    """One game on the host: W untimed + K timed search steps (mode A: one 100-leaf simulate batch;
    mode B: one ENERGY=8 wave) through the oracle port with the CPU tower as evaluator."""
    import numpy as np
    import torch
    from oracle import oracle as o, game_loop as gl
    torch.set_num_threads(threads or os.cpu_count() or 1)       # torchrun exports OMP_NUM_THREADS=1; the CPU arm sets its own
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


def _proc_worker(args):
    mode, blocks, steps, q = args
    r = cpu_tower_selfplay(mode, blocks, 0, steps, threads=1)
    return r['sims'], r['wall']


def cpu_tower_selfplay_per_core(mode, blocks, steps):
    """The reference's own shape (main_selfplay.py:26, N_GAME_PROCESS workers): one game per process on every host
    core, each evaluating its tower single-threaded.  Sum of simulations over the slowest process's wall time."""
    import multiprocessing as mp
    from oracle import oracle as o
    o.build()
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")            # (fork after this process has run OpenMP-threaded torch ops can hang in libgomp)
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        res = pool.map(_proc_worker, [(mode, blocks, steps, i) for i in range(cores)])
    wall = time.time() - t0
    return dict(sims=sum(r[0] for r in res), wall=max(r[1] for r in res), total_wall=wall, cores=cores)


def _cpu_baseline_dict(mode, blocks, warm, steps, with_uniform=True, per_core=False):
    r = cpu_tower_selfplay(mode, blocks, warm, steps)
    leaves = "100-leaf simulate batches (mode A)" if mode == 'a' else "8-leaf waves (mode B)"
    cb = dict(value=r['sims'] / r['wall'], unit="simulations/s", cores=r['threads'], kind="port", shape="one game, all threads in the tower",
              sample="1 game, %d timed %s of the same workload (19x19, %d-block tower evaluated in fp32 by torch on %d host threads, "
                     "oracle C port for rules/tree); the reference's real evaluator was a TF1.7 GPU process"
                     % (steps, leaves, blocks, r['threads']))
    if per_core:
        k = 1 if mode == 'a' else 4
        c = cpu_tower_selfplay_per_core(mode, blocks, k)
        alt = dict(value=c['sims'] / c['wall'], unit="simulations/s", cores=c['cores'], kind="port",
                   shape="one game per process on every core (main_selfplay.py:26), tower single-threaded per process",
                   sample="%d games in parallel, %d timed %s each" % (c['cores'], k, leaves))
        if alt['value'] > cb['value']:
            alt['sample'] += "; same workload (19x19, %d-block fp32 tower, oracle C port)" % blocks
            cb, alt = alt, cb
        cb['other_shape'] = dict(value=alt['value'], cores=alt['cores'], shape=alt['shape'])
    if with_uniform:
        cpu_selfplay(mode, 1)
        c = cpu_selfplay(mode, CPU_PLIES_PER_STEP)
        cb["rules_tree_only"] = dict(value=c['sims'] / c['wall'], unit="simulations/s", cores=c['cores'],
                                     note="free uniform evaluator (network cost excluded), one game per core, C oracle port; "
                                          "the pure-Python reference measured 179 sims/s/core in the survey container")
    return cb, r


def run_cpu_leg(a):
    """Child process of the GPU arm: prints the cpu_baseline dict (the GPU arm's own process never loads oracle/)."""
    cb, _ = _cpu_baseline_dict(a.mode, a.blocks, 1, 16)
    print(json.dumps(cb))


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, a.steps), a.warmup
    cb, r = _cpu_baseline_dict(a.mode, a.blocks, warm, steps, with_uniform=False, per_core=True)
    val = cb['value']
    step_sims = r['per'] * (cb['cores'] if 'per process' in cb['shape'] else 1)      # one search step of every game in flight
    print(json.dumps(dict(
        impl="reference", metric="selfplay_mcts_simulations_per_sec", value=val, unit="simulations/s", n_gpus=a.gpus,
        steps=steps, warmup=warm, ms_per_step=1e3 * step_sims / val, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload="19x19 self-play, conf.py default tower (%d blocks x 256 ch, random init), %d sims/ply, mode %s, "
                             "host cores only: %s" % (a.blocks, SIMS, a.mode.upper(), cb['shape']),
                    moves_per_sec=val / SIMS, step="one %d-leaf search step of every game in flight" % r['per']),
        cpu_baseline=cb, e2e=dict(value=val, unit="simulations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


# ------------------------------------------------------------------ GPU arm
def _pin_to_cores(local, world_local):
    """Each rank's host loop on its own slice of the cores (8 Python loops on one socket otherwise share them)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, world_local))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return None


class Arm(object):
    """One BatchedGames configuration on this rank, and its timed measurement."""

    def __init__(self, a, mode, match, rank, world, local):
        import torch
        from sejonggo_b200 import dist as sd, model
        from sejonggo_b200.batched import BatchedGames, HostRng
        self.a, self.mode, self.match, self.rank, self.world, self.local = a, mode, match, rank, world, local
        self.dev = torch.device("cuda", local)
        G = a.games
        batch = BATCH_A if mode == 'a' else ENERGY_B
        self.sims = 1600 if match else a.sims
        m = model.TowerModel("model_1", params=model.init_params(SIZE, a.blocks, seed=0), max_positions=a.max_positions) if rank == 0 else None
        self.blob_bytes = 0
        if world > 1:
            m = sd.broadcast_model(m, src=0, device=self.dev, max_positions=a.max_positions)    # NCCL weight-blob broadcast
            self.blob_bytes = sd.blob_bytes(m)
        self.models = [m]
        if match:
            # BASELINE.json configs[3] / SURVEY §8d config 4: evaluator.evaluate semantics — two weight sets, two trees
            # per game, no noise, temperature 0 from ply 0, 1600 sims/ply, one random symmetry per predict batch
            m2 = model.TowerModel("model_2", params=model.init_params(SIZE, a.blocks, seed=1), max_positions=a.max_positions)
            self.models.append(m2)
            self.arena = a.arena or 4 * (self.sims + batch)
            self.bg = BatchedGames((m, m2), G, size=SIZE, mode=mode, mcts_batch_size=BATCH_A, energy=ENERGY_B, mcts_simulations=self.sims,
                                   stop_exploration=0, self_play=False, rng=HostRng(1234 + rank), arena_blocks=self.arena, device=local,
                                   record_boards='packed')
        else:
            self.arena = a.arena or 8 * (SIMS + batch)
            self.bg = BatchedGames((m, m), G, size=SIZE, mode=mode, mcts_batch_size=BATCH_A, energy=ENERGY_B, mcts_simulations=self.sims,
                                   stop_exploration=30, self_play=True, rng=HostRng(1234 + rank), arena_blocks=self.arena, device=local,
                                   record_boards='packed')
        for i, mm in enumerate(self.models):
            mm.attach(self.bg.eng, i)

    def sync_all(self):
        import torch
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            torch.distributed.barrier()

    def timed(self, n_steps, record):
        import torch
        bg, e = self.bg, self.bg.eng
        self.sync_all()
        s0, p0, l0 = bg.sim_count, bg.plies_done, e.launch_count()
        h0, d0 = bg.h2d_bytes, bg.d2h_bytes
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n_steps):
            bg.step_ply(record=record)
        ev1.record()
        self.sync_all()
        own_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([own_ms], dtype=torch.float64, device=self.dev)
        cnt = torch.tensor([bg.sim_count - s0, bg.plies_done - p0], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)      # max over ranks, device timed
            torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
        return dict(ms=float(t.item()), own_ms=own_ms, sims=float(cnt[0].item()), plies=float(cnt[1].item()), launches=e.launch_count() - l0,
                    h2d=bg.h2d_bytes - h0, d2h=bg.d2h_bytes - d0)

    def measure(self, steps, warmup, e2e=True):
        bg, e = self.bg, self.bg.eng
        bg.start()
        for _ in range(warmup):
            bg.step_ply(record=False)
        sampler = ClockSampler(self.local)
        sampler.start()
        for i, mm in enumerate(self.models):
            mm.profile(e, i, True)
        r = self.timed(steps, record=False)
        prof = None
        for i, mm in enumerate(self.models):
            p = mm.profile_read(e, i)
            mm.profile(e, i, False)
            prof = p if prof is None else {k: prof[k] + p[k] for k in prof}
        clocks = sampler.stop()
        r2 = self.timed(steps, record=True) if e2e else None       # end-to-end through the public batched API
        e.check_errors()
        for i, mm in enumerate(self.models):
            mm.check(e, i)
        return r, r2, prof, clocks

    def workload(self):
        return ("19x19 %s, conf.py default tower (%d blocks x 256 ch, random init), %d sims/ply, %d concurrent games per GPU, mode %s"
                % ("match play (evaluator.evaluate: two weight sets, two trees per game, temperature 0, random symmetry per batch)"
                   if self.match else "self-play", self.a.blocks, self.sims, self.a.games, self.mode.upper()))

    def close(self):
        import gc
        import torch
        self.bg.eng.close()
        self.bg = None
        self.models = []
        gc.collect()
        torch.cuda.empty_cache()


def _roofline(prof, r, blocks):
    peaks = measured_peaks()
    conv_flops = prof['positions'] * FLOP_PER_CONV_POS * 2 * blocks
    conv_s = prof['conv_ms'] * 1e-3
    achieved = conv_flops / conv_s / 1e12 if conv_s > 0 else 0.0
    traffic, tsrc = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
            t = json.load(f)
        per_pos = t["dram_bytes_per_launch"] / t["n_positions"]
        pos_per_launch = prof['positions'] / max(1, prof['forwards'])
        traffic = per_pos * pos_per_launch
        tsrc = "ncu --set full at %d positions per launch (%s), scaled linearly to this run's %.0f positions per launch" \
               % (t["n_positions"], t.get("source", "profiles/"), pos_per_launch)
    except Exception:
        pass
    return dict(bound="tensor", kernel="k_conv3x3_pair", achieved=achieved, peak=peaks['tflops'], unit="TFLOP/s",
                frac=achieved / peaks['tflops'], traffic=traffic, traffic_source=tsrc, peak_source=peaks['src'],
                launches=prof['conv_launches'], avg_launch_ms=prof['conv_ms'] / max(1, prof['conv_launches']),
                share_of_step=prof['conv_ms'] / r['own_ms'], stem_ms=prof['stem_ms'], heads_ms=prof['heads_ms'])


def run_ours(a):
    import numpy as np
    import torch
    from sejonggo_b200 import dist as sd
    from sejonggo_b200.records import RecordStore, rows_from_game_data, row_words
    rank, world, local = sd.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pinned = _pin_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world))) if world > 1 else None
    arm = Arm(a, a.mode, a.match, rank, world, local)
    if a.full_games:
        return run_full_games(a, arm, rank, world)
    r, r2, prof, clocks = arm.measure(a.steps, a.warmup)
    # game records to rank 0: the rows of the recorded plies (board, move, value, policy_target), device to device
    bg = arm.bg
    games = bg.finish()
    store = RecordStore(SIZE, 16, dev)
    for gd in games:
        if gd['moves']:
            store.append_host_rows(rows_from_game_data(SIZE, gd['game_id'], gd, None))
    gathered = sd.gather_rows(store.take(), dst=0)
    per_rank = sd.gather_per_rank([r['own_ms'] / a.steps, prof['conv_ms'] / max(1, prof['conv_launches']), prof['conv_ms'] / a.steps,
                                   clocks['sm_mhz'] or 0.0, clocks.get('power_w') or 0.0, r2['own_ms'] / a.steps], device=dev)
    workload, G, arena, sims = arm.workload(), a.games, arm.arena, arm.sims
    T = bg.eng.T
    blob = arm.blob_bytes
    arm.close()
    others = {}
    if not a.no_sub and not a.match and a.mode == 'a':
        for key, mode, match in (("mode_b", 'b', False), ("match", 'a', True)):
            sub = Arm(a, mode, match, rank, world, local)
            sr, _, sprof, sclk = sub.measure(a.sub_steps, a.sub_warmup, e2e=False)
            rf = _roofline(sprof, sr, a.blocks)
            others[key] = dict(value=sr['sims'] / (sr['ms'] * 1e-3), unit="simulations/s", steps=a.sub_steps, warmup=a.sub_warmup,
                               ms_per_step=sr['ms'] / a.sub_steps, moves_per_sec=sr['plies'] / (sr['ms'] * 1e-3), workload=sub.workload(),
                               gpu_launches=sr['launches'], roofline_frac=rf['frac'], conv_tflops=rf['achieved'],
                               sm_mhz=sclk['sm_mhz'], reasons=sclk['reasons'])
            sub.close()
    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    out = dict(
        metric="selfplay_mcts_simulations_per_sec", value=r['sims'] / (r['ms'] * 1e-3), unit="simulations/s", n_gpus=world,
        steps=a.steps, warmup=a.warmup, ms_per_step=r['ms'] / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="bf16", data="synthetic",
        config=dict(workload=workload, games_per_gpu=G, mode=a.mode, sims_per_ply=sims, moves_per_sec=r['plies'] / (r['ms'] * 1e-3),
                    l2="working set (node pool %.1f GB, activations %.1f GB) >> 126 MB L2; no flush needed" %
                       (G * T * arena * 6272 / 1e9, 3 * (a.max_positions * 18 + 1) * 18 * 512 / 1e9),
                    step="one ply of every game", parallelism="games sharded, %d rank(s)" % world,
                    weights_broadcast_bytes=blob,
                    records_gathered_bytes=None if gathered is None else [int(g.numel()) * 4 for g in gathered],
                    record_row_bytes=row_words(SIZE) * 4, host_cores_per_rank=pinned),
        e2e=dict(value=r2['sims'] / (r2['ms'] * 1e-3), unit="simulations/s", h2d_bytes_per_step=r2['h2d'] / a.steps,
                 d2h_bytes_per_step=r2['d2h'] / a.steps, moves_per_sec=r2['plies'] / (r2['ms'] * 1e-3)),
        gpu_launches=r['launches'],
        clocks=clocks,
        roofline=_roofline(prof, r, a.blocks),
        other_modes=others,
    )
    if world > 1:
        out["per_rank"] = dict(columns=["ms_per_step", "conv_avg_launch_ms", "conv_ms_per_step", "sm_mhz_median", "power_w_median", "e2e_ms_per_step"],
                               rows=[[round(float(x), 4) for x in row] for row in per_rank],
                               slowest_rank=int(np.argmax(per_rank[:, 0])))
    if world == 1 and not a.no_cpu:
        try:                                  # the CPU baseline runs in its own process: this one never maps the oracle library
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu-leg", "--mode", a.mode, "--blocks", str(a.blocks)],
                                stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900,
                                env={k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"})
            out["cpu_baseline"] = json.loads(cp.stdout.strip().splitlines()[-1])
        except Exception as ex:
            out["cpu_baseline"] = dict(error=str(ex)[:200])
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_full_games(a, arm, rank, world):
    """Whole games with slot refill: `--total` games through `--games` slots.  A random-init network never resigns and
    rarely passes, so every game would run to the same cap and all slots would turn over in the same ply; to load the
    refill path the way real games do, each game gets its own length cap, uniform in [num_moves/3, num_moves]
    (seeded), so slots free up ply by ply."""
    import numpy as np
    import torch
    bg = arm.bg
    bg.n_total = a.total or 2 * a.games
    bg.num_moves = a.num_moves or 2 * SIZE * SIZE
    if a.resign is not None:
        bg.resign[0][:] = a.resign
        bg.resign[1][:] = a.resign
    lens = np.random.RandomState(99 + rank).randint(max(1, bg.num_moves // 3), bg.num_moves + 1, size=bg.n_total)
    bg.on_game_start = lambda gid: dict(num_moves=int(lens[gid]))
    ended = []
    bg.on_game_end = lambda gid, gd: ended.append(len(gd['moves']))
    bg.start()
    arm.sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    plies = 0
    while bg.step_ply(record=True):
        plies += 1
    ev1.record()
    arm.sync_all()
    ms = ev0.elapsed_time(ev1)
    bg.finish()
    whole_sims, total, lengths = bg.sim_count, bg.n_total, list(ended)
    st = bg.eng.pool_stats()
    steady = a.games * bg.sims * plies                      # what the same number of plies would do with every slot busy
    arm.sync_all()
    # the steady-state figure on the same box, same process: a few more plies of a full batch
    bg.n_total, bg.on_game_start = a.games, None
    bg.start()
    for _ in range(2):
        bg.step_ply(record=True)
    s0 = bg.sim_count
    arm.sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        bg.step_ply(record=True)
    e1.record()
    arm.sync_all()
    steady_rate = (bg.sim_count - s0) / (e0.elapsed_time(e1) * 1e-3)
    out = dict(metric="selfplay_whole_games", value=whole_sims / (ms * 1e-3), unit="simulations/s", n_gpus=world,
               games=len(lengths), games_per_hour=len(lengths) / (ms * 1e-3) * 3600, plies_stepped=plies, ms_total=ms,
               mean_active_fraction=whole_sims / max(1, steady), steady_state_sims_per_s=steady_rate,
               whole_run_over_steady_state=whole_sims / (ms * 1e-3) / steady_rate, mean_game_plies=sum(lengths) / max(1, len(lengths)),
               min_game_plies=min(lengths) if lengths else 0, max_game_plies=max(lengths) if lengths else 0,
               config=dict(workload=arm.workload(), total_games=total, slots=a.games, num_moves=bg.num_moves, resign=a.resign),
               pool=st, trees_dropped=bg.trees_dropped)
    if rank == 0:
        print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-leg"])
    ap.add_argument("--mode", default="a", choices=["a", "b"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--sims", type=int, default=None)
    ap.add_argument("--match", action="store_true", help="SURVEY §8d config 4: match play, two models, 1600 sims/ply")
    ap.add_argument("--blocks", type=int, default=BLOCKS)
    ap.add_argument("--max-positions", type=int, default=16384)
    ap.add_argument("--arena", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the mode-B / match sub-measurements")
    ap.add_argument("--sub-steps", type=int, default=1)
    ap.add_argument("--sub-warmup", type=int, default=3)
    ap.add_argument("--full-games", action="store_true")
    ap.add_argument("--total", type=int, default=0)
    ap.add_argument("--num-moves", type=int, default=0)
    ap.add_argument("--resign", type=float, default=None)
    a = ap.parse_args()
    if a.sims is None:
        a.sims = SIMS
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "cpu-leg":
        run_cpu_leg(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

"""What would removing the pad MACs buy?  Bounds, measured on the ablation build (python -m sejonggo_b200._build --ablate;
run with SGO_LIBRARY=sejonggo_b200/lib/libsejonggo_b200_ablate.so):

  full        the shipped conv layer, random activations, sustained for several seconds (power-capped clocks)
  ideal_tiles the same kernel launched over only 289/324 of its tiles (SGO_CONV_DEBUG=8): the time of an IDEAL kernel that
              issues no MMA for pad pixels / pad rows and pays nothing for skipping them (results are garbage: timing only)
  zero_act    the full tile count on all-zero activations: how much cheaper (power -> clocks) a MAC on a zero operand is,
              i.e. how much of the pad MACs' cost is pipe time rather than energy

    SGO_LIBRARY=... python tools/conv_pad_bound.py [positions] > gpurun_out/r02_conv_pad_bound.json
"""
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import torch

sys.path.insert(0, ".")
from sejonggo_b200.engine import Engine
from sejonggo_b200 import model


def sustained(fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    clk = []
    pr = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"],
                          stdout=subprocess.PIPE, text=True)
    threading.Thread(target=lambda: [clk.append(l.strip()) for l in pr.stdout], daemon=True).start()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.time()
    s.record()
    while time.time() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    t.record()
    torch.cuda.synchronize()
    pr.terminate()
    half = clk[len(clk) // 2:]
    mhz = sorted(float(x.split(",")[0]) for x in half if x)
    pw = sorted(float(x.split(",")[1]) for x in half if x)
    return s.elapsed_time(t) / n, (mhz[len(mhz) // 2] if mhz else None), (pw[len(pw) // 2] if pw else None)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
    S, W = 19, 17
    e = Engine(size=S, n_games=64, max_leaves=1, arena_blocks=2)
    m = model.TowerModel("v", size=S, n_blocks=1, seed=0, max_positions=n)
    m.attach(e, 0)
    g = torch.Generator(device="cuda").manual_seed(1)
    if os.environ.get("SGO_PADDED_LAYOUT"):          # a round-1 / early round-2 library: 18 x 18 slots per position, pads zero
        rows = n * (W + 1) + 1
        act = torch.zeros((rows, W + 1, 256), dtype=torch.bfloat16, device="cuda")
        act[1:].view(n, W + 1, W + 1, 256)[:, :W, :W] = torch.randn((n, W, W, 256), generator=g, device="cuda").to(torch.bfloat16)
    else:
        act = torch.randn((n * W * W, 256), generator=g, device="cuda").to(torch.bfloat16)
    zero = torch.zeros_like(act)
    flop = 2.0 * n * 289 * 256 * 256 * 9
    out = dict(positions=n, seconds_each=secs, library=os.environ.get("SGO_LIBRARY", "product build (no ablation switches)"))

    def load(t):
        for b in (0, 2):
            e._ck(e.lib.sgo_tower_act_copy(e.h, 0, b, n, C.c_void_p(t.data_ptr()), 1, e._stream()))

    fn = lambda: e._ck(e.lib.sgo_tower_debug_conv(e.h, 0, n, 1, 0, 1, 2, e._stream()))
    padded = bool(os.environ.get("SGO_PADDED_LAYOUT"))
    # bit 8 (padded libraries): 289/324 of the tiles = an ideal pad-free kernel; bit 16 (dense libraries): all lane masks zero = what the masks cost
    variants = [("full", act, 0), ("ideal_tiles" if padded else "no_masks", act, 8 if padded else 16), ("zero_act", zero, 0)]
    if not padded:      # where does the power go?  (ablated runs compute garbage: timing / clocks only)
        variants += [("no_B_loads", act, 4), ("no_A_loads", act, 2), ("no_epilogue_traffic", act, 1), ("no_skip_loads", act, 64)]
    variants.append(("full_again", act, 0))
    for name, data, dbg in variants:
        os.environ["SGO_CONV_DEBUG"] = str(dbg)
        load(data)
        ms, mhz, pw = sustained(fn, secs)
        out[name] = dict(ms=ms, useful_tflops=flop / (ms * 1e-3) / 1e12, sm_mhz=mhz, power_w=pw)
    os.environ["SGO_CONV_DEBUG"] = "0"
    f = out["full"]["ms"]
    if padded:
        out["bound"] = dict(ideal_speedup=f / out["ideal_tiles"]["ms"], zero_operand_speedup=f / out["zero_act"]["ms"],
                            note="ideal_speedup is the most ANY pad-free tiling could gain on this kernel at this power cap")
    else:
        out["bound"] = dict(mask_cost=f / out["no_masks"]["ms"], zero_operand_speedup=f / out["zero_act"]["ms"],
                            no_B_loads_speedup=f / out["no_B_loads"]["ms"], no_A_loads_speedup=f / out["no_A_loads"]["ms"],
                            no_epilogue_traffic_speedup=f / out["no_epilogue_traffic"]["ms"], no_skip_loads_speedup=f / out["no_skip_loads"]["ms"],
                            note="sustained, power-capped: a speed-up here is (mostly) the energy that part of the kernel costs")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""Bottleneck experiments on k_conv3x3_tc: drop operand loads / epilogue traffic (results are garbage,
timings tell which resource binds)."""
import ctypes as C, json, sys, torch
sys.path.insert(0, ".")
from sejonggo_b200.engine import Engine
from sejonggo_b200 import model


def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    t.record(); torch.cuda.synchronize()
    return s.elapsed_time(t) / iters * 1e-3

n = 8192
e = Engine(size=19, n_games=64, max_leaves=1, arena_blocks=2)
m = model.TowerModel("v", size=19, n_blocks=1, seed=0, max_positions=n)
m.attach(e, 0)
flop = 2.0 * n * 289 * 256 * 256 * 9
out = {}
for name, variant, skip in [("full+skip", 0, 2), ("full", 0, 15), ("noB", 1, 15), ("noA", 2, 15), ("noAB", 3, 15), ("noEpi", 4, 15),
                            ("noAB_noEpi", 7, 15), ("noB_noEpi", 5, 15)]:
    code = (variant << 4) | skip if variant else (skip if skip != 15 else -1)
    t = timed(lambda: e._ck(e.lib.sgo_tower_debug_conv(e.h, 0, n, 1, 0, 1, code, e._stream())))
    out[name] = dict(ms=t * 1e3, tflops=flop / t / 1e12)
m.check(e, 0)
print(json.dumps(out, indent=1))

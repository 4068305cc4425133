"""Timing ablations of k_conv3x3_pair (SGO_CONV_DEBUG bits: 1 no epilogue global traffic, 2 no A loads,
4 no B loads) — where does the time of one conv layer go?  Results of the ablated runs are garbage."""
import ctypes as C
import json
import os
import sys
import torch

sys.path.insert(0, ".")
from sejonggo_b200.engine import Engine
from sejonggo_b200 import model


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    S = 19
    e = Engine(size=S, n_games=64, max_leaves=1, arena_blocks=2)
    m = model.TowerModel("v", size=S, n_blocks=1, seed=0, max_positions=n)
    m.attach(e, 0)
    out = {}
    for name, dbg, skip in (("full+skip", 0, 2), ("full", 0, -1), ("noEpi", 1, 2), ("noA", 2, 2), ("noB", 4, 2), ("noAB", 6, 2),
                            ("noAB_noEpi", 7, 2), ("noStore", 32, 2), ("noSkipLoad", 64, 2), ("noStore_noSkipLoad", 96, 2), ("full+skip again", 0, 2)):
        os.environ["SGO_CONV_DEBUG"] = str(dbg)
        fn = lambda: e._ck(e.lib.sgo_tower_debug_conv(e.h, 0, n, 1, 0, 1, skip, e._stream()))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(100):
            fn()
        t.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(t) / 100
        out[name] = dict(ms=ms, tflops=2.0 * n * 289 * 256 * 256 * 9 / (ms * 1e-3) / 1e12)
        f = C.c_int32(0)
        e._ck(e.lib.sgo_tower_check_sync(e.h, 0, C.byref(f), e._stream()))
        out[name]["err"] = f.value
        torch.cuda.synchronize()
    os.environ["SGO_CONV_DEBUG"] = "0"
    # the roofline denominator on THIS box: cuBLAS bf16 8192^3 as a burst and sustained for ~3 s (how MEASURED_PEAKS.json was taken)
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.matmul(a, b); t.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(t))
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    iters = 3000
    for _ in range(iters):
        torch.matmul(a, b)
    t.record(); torch.cuda.synchronize()
    out["cublas_bf16_8192"] = dict(burst_tflops=2 * 8192 ** 3 / (best * 1e-3) / 1e12,
                                   sustained_tflops=2 * 8192 ** 3 * iters / (s.elapsed_time(t) * 1e-3) / 1e12,
                                   sustained_seconds=s.elapsed_time(t) * 1e-3)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""Tower micro-benchmark: one tcgen05 conv layer and the whole forward, CUDA-event timed."""
import ctypes as C
import json
import sys
import torch

sys.path.insert(0, ".")
from sejonggo_b200.engine import Engine
from sejonggo_b200 import model


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / iters * 1e-3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    S = 19
    e = Engine(size=S, n_games=n, max_leaves=1, arena_blocks=2)
    e.reset()
    e.random_playouts(seed=3, max_plies=120)
    m = model.TowerModel("bench", size=S, n_blocks=blocks, seed=0, max_positions=n)
    m.attach(e, 0)
    idx = torch.arange(n, dtype=torch.int32, device=e.device)
    pol = torch.empty((n, e.A), dtype=torch.float32, device=e.device)
    val = torch.empty((n,), dtype=torch.float32, device=e.device)
    t_conv = timed(lambda: e._ck(e.lib.sgo_tower_debug_conv(e.h, 0, n, 1, 0, 1, 2, e._stream())), iters=10)
    flop_layer = 2.0 * n * 289 * 256 * 256 * 9
    t_fwd = timed(lambda: e._ck(e.lib.sgo_tower_forward(e.h, 0, 0, C.c_void_p(idx.data_ptr()), n, C.c_void_p(0), 0,
                                                         C.c_void_p(pol.data_ptr()), C.c_void_p(val.data_ptr()), e._stream())), iters=3, warm=1)
    m.profile(e, 0, True)
    import subprocess, threading
    clk = []
    pr = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    threading.Thread(target=lambda: [clk.append(l.strip()) for l in pr.stdout], daemon=True).start()
    t_sus = timed(lambda: e._ck(e.lib.sgo_tower_forward(e.h, 0, 0, C.c_void_p(idx.data_ptr()), n, C.c_void_p(0), 0,
                                                         C.c_void_p(pol.data_ptr()), C.c_void_p(val.data_ptr()), e._stream())), iters=20, warm=0)
    pr.terminate()
    prof = m.profile_read(e, 0)
    m.check(e, 0)
    flop_fwd = n * (2.0 * 289 * 256 * 256 * 9 * 2 * blocks + 2.0 * 289 * 9 * 17 * 256 + 1.3e6)
    print(json.dumps(dict(n=n, blocks=blocks, conv_layer_s=t_conv, conv_layer_tflops=flop_layer / t_conv / 1e12,
                          forward_s=t_fwd, sustained_forward_s=t_sus, prof_per_forward_ms={k: (v / prof['forwards'] if k.endswith('_ms') else v) for k, v in prof.items()}, clocks=clk[::4], forward_tflops=flop_fwd / t_fwd / 1e12, evals_per_s=n / t_fwd)))


if __name__ == "__main__":
    main()

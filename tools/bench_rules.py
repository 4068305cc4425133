"""Micro-benchmark of the board kernels (HBM-bound rows of SURVEY §8d): CUDA-event timed."""
import json
import sys
import torch

sys.path.insert(0, ".")
from sejonggo_b200.engine import Engine


def timed(fn, iters=20, warm=3):
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
    out = {}
    S = 19
    for G in (4096, 65536):
        e = Engine(size=S, n_games=G, max_leaves=1, arena_blocks=2)
        e.reset()
        moves, npl = e.random_playouts(seed=1, max_plies=200)     # mid-game positions
        torch.cuda.synchronize()
        mv = torch.full((G,), S * S, dtype=torch.int32, device=e.device)  # pass keeps positions stable
        t_apply = timed(lambda: e.apply_moves(mv))
        t_legal = timed(lambda: e.legal_masks())
        t_score = timed(lambda: e.score())
        t_planes = timed(lambda: e.export_planes(0, 0, G, sym=4))
        out["G%d" % G] = dict(
            apply_s=t_apply, apply_GBs=G * 1444 / t_apply / 1e9,
            legal_s=t_legal, legal_GBs=G * 184 / t_legal / 1e9,
            score_s=t_score, score_GBs=G * 103 / t_score / 1e9,
            planes_f32_s=t_planes, planes_GBs=G * (722 + 361 * 17 * 4) / t_planes / 1e9)
        e.reset()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        moves, npl = e.random_playouts(seed=20260, max_plies=722)
        t.record()
        torch.cuda.synchronize()
        dt = s.elapsed_time(t) * 1e-3
        out["G%d" % G]["playout_s"] = dt
        out["G%d" % G]["playout_plies_per_s"] = float(npl.sum().item()) / dt
        e.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

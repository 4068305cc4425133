"""Where does a ply go outside the tower?  One 1,024-game mode-A ply with every host-visible phase timed exclusively
(synchronise, time, synchronise): root evaluation, new trees, the 8 native search steps (and inside them, from the
tower's own event profile: stem / convs / heads), pick, re-root, make_play, and the small host syncs.

    python tools/ply_breakdown.py [games] > gpurun_out/r02_ply_breakdown.json
"""
import json
import sys
import time
import collections
import torch

sys.path.insert(0, ".")
from sejonggo_b200 import model
from sejonggo_b200.batched import BatchedGames, HostRng


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    m = model.TowerModel("m", size=19, n_blocks=blocks, seed=0, max_positions=16384)
    bg = BatchedGames((m, m), G, size=19, mode='a', mcts_batch_size=100, mcts_simulations=800, stop_exploration=30, self_play=True,
                      rng=HostRng(7), record_boards='packed')
    e = bg.eng
    m.attach(e, 0)
    bg.start()
    for _ in range(3):
        bg.step_ply(record=False)
    acc = collections.OrderedDict()

    def wrap(obj, name, label):
        fn = getattr(obj, name)

        def timed(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize()
            d = acc.setdefault(label, [0, 0.0])
            d[0] += 1
            d[1] += (time.perf_counter() - t0) * 1e3
            return r

        setattr(obj, name, timed)

    wrap(bg, "_evaluate", "root evaluation (one forward of G positions)")
    wrap(e, "tree_new", "tree_new")
    wrap(e, "selfplay_step", "sgo_selfplay_step (select, gather, forwards, expand, backup)")
    wrap(bg, "_draw_syms_device", "symmetry draws (torch.randint on the device)")
    wrap(e, "pick", "pick")
    wrap(e, "reroot", "reroot (frees the discarded blocks)")
    wrap(e, "apply_moves", "make_play")
    wrap(e, "tree_valid", "tree_valid")
    wrap(e, "pool_stats", "pool_stats")
    wrap(e, "records_pack", "records_pack")
    m.profile(e, 0, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plies = 2
    for _ in range(plies):
        bg.step_ply(record=True)
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
    prof = m.profile_read(e, 0)
    out = dict(games=G, plies=plies, ms_per_ply=total / plies,
               phases={k: dict(calls=v[0] / plies, ms_per_ply=v[1] / plies) for k, v in acc.items()},
               tower_ms_per_ply=dict(stem=prof['stem_ms'] / plies, convs=prof['conv_ms'] / plies, heads=prof['heads_ms'] / plies,
                                     forwards=prof['forwards'] / plies))
    out["unaccounted_ms_per_ply"] = total / plies - sum(v[1] for v in acc.values()) / plies
    step = acc["sgo_selfplay_step (select, gather, forwards, expand, backup)"][1] / plies
    root = acc["root evaluation (one forward of G positions)"][1] / plies
    out["search_steps_minus_tower_ms_per_ply"] = step + root - (prof['stem_ms'] + prof['conv_ms'] + prof['heads_ms']) / plies
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

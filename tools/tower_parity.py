"""Measured numerics of the CUDA tower (bf16 tensor-core convs, fp32 heads) against the float64 restatement of
model.py:55-96 (oracle/tower_ref.py, TF32 off): max |dp|, max relative dp where p > 1e-3, max |d logit| (centred
log-probabilities), max |dv| and max |d atanh v|, on 19x19 mid-game positions for several weight sets.

    python tools/tower_parity.py [--out profiles/r02_tower_parity.json] [--n 256]

Used by tests/test_gpu_tower.py (which asserts the bounds) and committed under profiles/.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def positions(size, n, seed, max_plies=250):
    """n distinct plausible positions: device random playouts of varied length (k_random_playouts)."""
    from sejonggo_b200.engine import Engine
    e = Engine(size=size, n_games=n, max_leaves=1, arena_blocks=2)
    rs = np.random.RandomState(seed)
    lens = rs.randint(0, max_plies, size=n)
    e.reset()
    moves, nplies = e.random_playouts(seed=seed, max_plies=max_plies)
    moves = moves.cpu().numpy().astype(np.int32)
    e.reset()
    for t in range(max_plies):                       # replay each game up to its own length
        mv = np.where((t < lens) & (moves[:, t] >= 0), moves[:, t], -1).astype(np.int32)
        if (mv < 0).all():
            break
        e.apply_moves(mv)
    boards = e.export_boards().cpu().numpy()
    e.close()
    return boards


def weight_sets(size, blocks):
    """name -> params.  'default' = Keras initialisers as the reference builds the net (model.py:55-96; at 20 blocks
    its values saturate near +-1); 'random_bn' = every BN vector and bias randomised (exercises the folding; the
    spread is kept at 10% so that 41 layers do not blow up); 'peaked' = default trunk with the policy dense layer
    scaled up so that policies are far from uniform, and the value head scaled down out of tanh saturation."""
    from sejonggo_b200 import model
    out = {"default": model.init_params(size, blocks, seed=0)}
    rb = model.init_params(size, blocks, seed=1, randomize_bn=True, random_bias=True)
    for k, v in rb.items():
        if k == 'meta':
            continue
        if isinstance(v, dict):
            v['gamma'] = 1 + 0.1 * (v['gamma'] - 1); v['var'] = 1 + 0.1 * (v['var'] - 1)
            v['beta'] = 0.1 * v['beta']; v['mean'] = 0.1 * v['mean']
        elif v.dim() == 1:
            rb[k] = 0.1 * v
    out["random_bn"] = rb
    pk = model.init_params(size, blocks, seed=0)
    pk['pol_fc_w'] = pk['pol_fc_w'] * 2.5
    pk['val_fc2_w'] = pk['val_fc2_w'] * 0.25
    out["peaked"] = pk
    return out


def measure(params, boards, max_positions=256, chunk=64):
    from sejonggo_b200 import model
    from oracle import tower_ref
    m = model.TowerModel("t", params=params, max_positions=max_positions)
    pol, val = m.predict_on_batch(boards)
    m.check(m._host_engine, 0)
    m._host_engine.close()
    pol, val = pol.astype(np.float64), val.astype(np.float64).reshape(-1)
    rp, rv, rl, rpre = [], [], [], []
    for s in range(0, len(boards), chunk):
        a, b, c, d = tower_ref.forward(params, boards[s:s + chunk].astype(np.float64), device="cuda", dtype=torch.float64, raw=True)
        rp.append(a.cpu().numpy()); rv.append(b.cpu().numpy()); rl.append(c.cpu().numpy()); rpre.append(d.cpu().numpy())
    rp, rv, rl, rpre = np.concatenate(rp), np.concatenate(rv).reshape(-1), np.concatenate(rl), np.concatenate(rpre).reshape(-1)
    dp = np.abs(pol - rp)
    big = rp > 1e-3
    lg = np.log(np.maximum(pol, 1e-300))
    lg -= lg.mean(axis=1, keepdims=True)
    rlc = rl - rl.mean(axis=1, keepdims=True)
    keep = rp > 1e-6                                 # log of a float32 probability below ~1e-6 is rounding noise
    pre = np.arctanh(np.clip(val, -1 + 1e-7, 1 - 1e-7))
    return dict(n=int(len(boards)),
                max_abs_dp=float(dp.max()),
                max_rel_dp_where_p_gt_1e3=float((dp[big] / rp[big]).max()) if big.any() else 0.0,
                max_abs_dlogit=float(np.abs(lg - rlc)[keep].max()),
                max_abs_dv=float(np.abs(val - rv).max()),
                max_abs_dpre_tanh=float(np.abs(pre - rpre)[np.abs(rv) < 0.999].max()),
                policy_max_mean=float(rp.max(axis=1).mean()), policy_max_max=float(rp.max()),
                value_abs_mean=float(np.abs(rv).mean()), value_abs_max=float(np.abs(rv).max()),
                argmax_agree=float((pol.argmax(axis=1) == rp.argmax(axis=1)).mean()),
                sum_err=float(np.abs(pol.sum(axis=1) - 1).max()))


def run(n=256, size=19, blocks=20, seed=5):
    boards = positions(size, n, seed)
    res = {"config": dict(size=size, blocks=blocks, channels=256, positions=n,
                          reference="oracle/tower_ref.py in float64 on the GPU, cudnn/matmul TF32 disabled",
                          note="the reference network (TF1.7/Keras2.2.2, model.py:55-96) has no golden vectors: parity is "
                               "against this restatement (unpinned by the reference)")}
    for name, params in weight_sets(size, blocks).items():
        res[name] = measure(params, boards)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_tower_parity.json"))
    ap.add_argument("--n", type=int, default=256)
    a = ap.parse_args()
    r = run(a.n)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(r, f, indent=1)
    print(json.dumps(r, indent=1))

"""GPU parity of the tower kernels (csrc/tower.cu) against the torch restatement of model.py:55-96
(oracle/tower_ref.py) run in true fp32 / fp64 (TF32 off).  BASELINE.json's north_star states 1e-2 absolute on policy
and value for bf16 compute; at random init every policy entry is ~1/362, so that bound alone says little — the
full-size test below asserts much tighter figures (and relative ones), measured and recorded by
tools/tower_parity.py -> profiles/r02_tower_parity.json."""
import ctypes as C
import numpy as np
import pytest
import torch

from oracle import oracle as o
from oracle import tower_ref
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _positions(size, n, seed):
    """Plausible mid-game positions: random legal playouts from the oracle."""
    rs = np.random.RandomState(seed)
    boards = []
    for i in range(n):
        b, _ = o.game_init(size)
        for t in range(int(rs.randint(0, size * size // 2))):
            legal = np.nonzero(o.legal_moves(b)[:-1] == 0)[0]
            mv = size * size if len(legal) == 0 or rs.rand() < 0.05 else int(rs.choice(legal))
            o.make_play(mv % size, mv // size, b)
        boards.append(b)
    return np.concatenate(boards)


@pytest.mark.parametrize("S,n", [(19, 37), (19, 300), (9, 21), (5, 97)])
def test_conv_layer_in_isolation(S, n):
    """One tcgen05 conv layer (+bias, +skip, ReLU) vs torch conv2d on the same bf16 data.  The activations are dense
    ([n*W*W][256]): the image edges exist only as the disable-output-lane masks of the tap MMAs, so this is the test of
    those masks — every tile alignment occurs (tiles of 256 rows over images of W*W = 289 / 49 / 9 rows straddle
    positions at every offset), with non-zero neighbours on the far side of every edge and a ragged last tile."""
    from sejonggo_b200.engine import Engine
    from sejonggo_b200 import model
    W = S - 2
    params = model.init_params(S, 1, seed=3, randomize_bn=True, random_bias=True)
    m = model.TowerModel("t", params=params, max_positions=n + 3)
    e = Engine(size=S, n_games=4, max_leaves=1, arena_blocks=2)
    m.attach(e, 0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn((n, W, W, 256), generator=g).clamp_(-3, 3)
    skip = torch.randn((n, W, W, 256), generator=g)

    def to_dev(t):
        return t.to(torch.bfloat16).reshape(n * W * W, 256).contiguous().cuda()

    f = model.folded_arrays(params)
    for layer, use_skip in ((0, False), (1, True)):
        e._ck(e.lib.sgo_tower_act_copy(e.h, 0, 0, n, C.c_void_p(to_dev(x).data_ptr()), 1, e._stream()))
        e._ck(e.lib.sgo_tower_act_copy(e.h, 0, 2, n, C.c_void_p(to_dev(skip).data_ptr()), 1, e._stream()))
        e._ck(e.lib.sgo_tower_debug_conv(e.h, 0, n, layer, 0, 1, 2 if use_skip else -1, e._stream()))
        out = torch.empty((n * W * W, 256), dtype=torch.bfloat16, device="cuda")
        e._ck(e.lib.sgo_tower_act_copy(e.h, 0, 1, n, C.c_void_p(out.data_ptr()), 0, e._stream()))
        torch.cuda.synchronize()
        m.check(e, 0)
        got = out.cpu().view(n, W, W, 256).float()
        wk = f['conv_w'][layer].float()              # (kh,kw,out,in), bf16-rounded
        ref = torch.nn.functional.conv2d(x.to(torch.bfloat16).float().permute(0, 3, 1, 2),
                                         wk.permute(2, 3, 0, 1).contiguous(), f['conv_b'][layer], padding=1)
        if use_skip:
            ref = ref + skip.to(torch.bfloat16).float().permute(0, 3, 1, 2)
        ref = torch.relu(ref).permute(0, 2, 3, 1)
        err = (got - ref).abs()
        tol = 0.02 + 0.01 * ref.abs()                # bf16 output rounding (2^-8 relative) + accumulation order
        assert bool((err <= tol).all()), "layer %d: max err %g at %s" % (layer, err.max(), np.unravel_index(int(err.argmax()), err.shape))
        # the edges in particular (a wrong mask bit shows up as a full-size error on a border pixel)
        edge = torch.zeros((W, W), dtype=torch.bool)
        edge[0, :] = edge[-1, :] = edge[:, 0] = edge[:, -1] = True
        assert float(err[:, edge].max()) <= float(tol[:, edge].max())
    e.close()


# (size, blocks, positions) -> bounds on |dp|, relative dp where p > 1e-3, |d logit| (centred), |dv|: about 3x what was
# measured on a B200 (2 blocks: 2.6e-5 / 0.7% / 0.007 / 2.5e-3; 3 blocks 9x9: 1.4e-4 / 0.8% / 0.008 / 1.3e-3; 20 blocks:
# 1.5e-3 / 5.4% / 0.053 / 3e-4).  An all-zero or uniform output would miss every one of them by orders of magnitude.
SMALL_BOUNDS = {(19, 2, 24): (2e-4, 0.03, 0.03, 8e-3), (9, 3, 40): (6e-4, 0.03, 0.03, 5e-3), (19, 20, 12): (1e-2, 0.12, 0.13, 5e-3)}


@pytest.mark.parametrize("size,n_blocks,n", sorted(SMALL_BOUNDS))
def test_forward_vs_fp64_reference(size, n_blocks, n):
    from sejonggo_b200 import model
    params = model.init_params(size, n_blocks, seed=0, randomize_bn=(n_blocks != 20), random_bias=(n_blocks != 20))
    m = model.TowerModel("t", params=params, max_positions=32)
    boards = _positions(size, n, seed=size + n_blocks)
    pol, val = m.predict_on_batch(boards)
    m.check(m._host_engine, 0)
    rp, rv, rl, _ = tower_ref.forward(params, boards.astype(np.float64), device="cuda", dtype=torch.float64, raw=True)
    rp, rv, rl = rp.cpu().numpy(), rv.cpu().numpy(), rl.cpu().numpy()
    assert pol.shape == (n, size * size + 1) and val.shape == (n, 1)
    assert np.abs(pol.sum(axis=1) - 1).max() < 1e-4
    dp_max, rel_max, dl_max, dv_max = SMALL_BOUNDS[(size, n_blocks, n)]
    dp = np.abs(pol - rp)
    big, keep = rp > 1e-3, rp > 1e-6
    lg = np.log(np.maximum(pol.astype(np.float64), 1e-300))
    lg -= lg.mean(axis=1, keepdims=True)
    rlc = rl - rl.mean(axis=1, keepdims=True)
    assert dp.max() <= dp_max <= TOL, dp.max()
    assert (dp[big] / rp[big]).max() <= rel_max, (dp[big] / rp[big]).max()
    assert np.abs(lg - rlc)[keep].max() <= dl_max, np.abs(lg - rlc)[keep].max()
    assert np.abs(val - rv).max() <= dv_max <= TOL, np.abs(val - rv).max()


# Measured maxima over 256 positions (profiles/r02_tower_parity.json) and the bound asserted for each: the bound is the
# measurement plus headroom for box-to-box / position-set variation, NOT a design tolerance.  An ideal pipeline with
# the same storage (bf16 weights and activations, exact accumulation; emulated in float64 on the CPU) lands at the same
# figures (max |d logit| 0.07 / 0.12 / 0.17 on 12 positions), so the error is the format's, not the kernels'.
TOWER_BOUNDS = {
    #            |dp|    rel dp   |dlogit|  |dv|    |d pre-tanh|
    "default":   (1e-2,   0.12,    0.13,    5e-3,   0.09),      # the north star's case: Keras default init; its 1e-2 abs holds
    "random_bn": (4e-2,   0.22,    0.25,    1e-2,   0.012),
    "peaked":    (5e-2,   0.30,    0.28,    1.2e-2, 0.04),
}


def test_forward_20_blocks_bounds_vs_fp64():
    """The benchmarked network (20 blocks x 256 channels, 19x19) on 256 mid-game positions against the float64
    restatement (TF32 off): Keras default init, randomised BN/biases, and a 'peaked' weight set whose policies are far
    from uniform (mean max p 0.33, up to 0.90).  Compared: policy (absolute, and relative where p > 1e-3), the centred
    log-probabilities (= pre-softmax logits up to their mean), value and pre-tanh value, top move.  The measured maxima
    go to gpurun_out/r02_tower_parity.json (committed under profiles/)."""
    import json, os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import tower_parity
    res = tower_parity.run(n=256)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "r02_tower_parity.json"), "w") as f:
        json.dump(res, f, indent=1)
    assert res["peaked"]["policy_max_mean"] > 0.05              # the peaked set really is far from uniform (uniform = 0.0028)
    for name, (dp, rel, dl, dv, dpre) in TOWER_BOUNDS.items():
        r = res[name]
        assert r["max_abs_dp"] <= dp, (name, r)
        assert r["max_rel_dp_where_p_gt_1e3"] <= rel, (name, r)
        assert r["max_abs_dlogit"] <= dl, (name, r)
        assert r["max_abs_dv"] <= dv, (name, r)
        assert r["max_abs_dpre_tanh"] <= dpre, (name, r)
        assert r["argmax_agree"] >= 0.97 and r["sum_err"] < 1e-4, (name, r)
    # a kernel that returned a uniform policy / zero value would be nowhere near: the bounds bite
    assert res["default"]["policy_max_max"] > 10 * TOWER_BOUNDS["default"][0]
    assert res["peaked"]["value_abs_mean"] > 10 * TOWER_BOUNDS["peaked"][3]


def test_forward_symmetry_fusion():
    """random_symmetry_predict semantics (symmetry.py:127-132): board gather fused into the stem,
    the 'reverse' policy gather (same map, Q8) fused into the heads."""
    from sejonggo_b200 import model
    from sejonggo_b200.engine import Engine
    S, n = 9, 16
    params = model.init_params(S, 2, seed=5, randomize_bn=True, random_bias=True)
    m = model.TowerModel("t", params=params, max_positions=32)
    boards = _positions(S, n, seed=77)
    e = Engine(size=S, n_games=n, max_leaves=1, arena_blocks=2)
    e.import_boards(boards)
    idx = torch.arange(n, device=e.device)
    syms = torch.as_tensor(np.arange(n) % 8, dtype=torch.int32, device=e.device)
    pol, val = m.evaluate(e, 0, idx, syms, slot=0)
    m.check(e, 0)
    pol, val = pol.cpu().numpy(), val.cpu().numpy()
    for i in range(n):
        k = int(syms[i])
        sb = o.sym_board(k, boards[i:i + 1])
        rp, rv = tower_ref.forward(params, sb.astype(np.float32), device="cuda")
        rp = o.sym_policy(k, rp.cpu().numpy(), S)
        assert np.abs(pol[i] - rp[0]).max() <= TOL and abs(val[i] - float(rv[0, 0])) <= TOL, (i, k)
    # gather by index / chunking over max_positions
    perm = torch.as_tensor(np.random.RandomState(0).permutation(n), device=e.device)
    m2 = model.TowerModel("t2", params=params, max_positions=5)
    p2, v2 = m2.evaluate(e, 0, perm, syms[perm], slot=1)
    assert np.abs(p2.cpu().numpy() - pol[perm.cpu().numpy()]).max() < 1e-6
    e.close()


def test_self_play_with_tower_runs():
    """End-to-end: play_games with the CUDA tower as the evaluator (mode A and mode B)."""
    from sejonggo_b200 import model, self_play as sp, nomodel_self_play as nsp, predicting_queue_worker as pq
    from sejonggo_b200.conf import conf
    old = dict(conf)
    try:
        conf.update(SIZE=9, MCTS_BATCH_SIZE=8, ENERGY=8, MCTS_SIMULATIONS=16, KOMI=5.5)
        m = model.TowerModel("model_1", size=9, n_blocks=2, seed=1, max_positions=256)
        from sejonggo_b200.batched import HostRng
        games = sp.play_games(m, m, 6, 16, 3, self_play=True, num_moves=6, rng=HostRng(3))
        # a random-init net may pass twice in a row: those games end early (self_play.py:217-220)
        assert len(games) == 6 and all(len(g['moves']) == 6 or g['end_reason'] == 'BOTH_PASSED' for g in games)
        assert all(abs(mv['policy'].sum() - 1.0) < 0.3 for g in games for mv in g['moves'][1:])
        pq.register_models(best=m, latest=m)
        games = nsp.play_games_async("BEST_SYM", "BEST_SYM", 4, 8, 3, self_play=True, num_moves=5, exact_rng_order=False, rng=HostRng(4))
        assert len(games) == 4 and all(len(g['moves']) == 5 or g['end_reason'] == 'BOTH_PASSED' for g in games)
    finally:
        conf.clear()
        conf.update(old)

"""The tower restatement (oracle/tower_ref.py, torch) against a second, independent restatement written with plain
numpy loops in float64 straight from the Keras-2.2.2 layer definitions (NHWC cross-correlation, 'valid' / 'same'
padding, inference BatchNormalization, HWC flatten, Dense = x @ W + b).  The reference's own arithmetic
(TF 1.7) cannot run here, so this does not pin the oracle to the reference — it guards the restatement against
slips of orientation (kernel flip, HWC vs CHW flatten, padding side) that a single implementation could hide."""
import numpy as np
import torch

from oracle import tower_ref
from sejonggo_b200 import model

EPS = 1e-3


def _conv_nhwc(x, k, b, same):
    n, H, W, _ = x.shape
    kh, kw, ci, co = k.shape
    if same:
        ph, pw = kh // 2, kw // 2
        xp = np.zeros((n, H + 2 * ph, W + 2 * pw, ci))
        xp[:, ph:ph + H, pw:pw + W] = x
        oh, ow = H, W
    else:
        xp, oh, ow = x, H - kh + 1, W - kw + 1
    out = np.zeros((n, oh, ow, co))
    for y in range(oh):
        for xx in range(ow):
            patch = xp[:, y:y + kh, xx:xx + kw, :]                      # cross-correlation: no kernel flip
            out[:, y, xx, :] = np.tensordot(patch, k, axes=([1, 2, 3], [0, 1, 2])) + b
    return out


def _bn(x, p):
    return (x - p['mean']) / np.sqrt(p['var'] + EPS) * p['gamma'] + p['beta']


def _np_forward(params, boards):
    P = {k: ({kk: vv.double().numpy() for kk, vv in v.items()} if isinstance(v, dict) else v.double().numpy())
         for k, v in params.items() if k != 'meta'}
    x = np.asarray(boards, dtype=np.float64)
    x = np.maximum(_bn(_conv_nhwc(x, P['stem_k'], P['stem_b'], False), P['stem_bn']), 0)
    for i in range(params['meta']['n_blocks']):
        t = np.maximum(_bn(_conv_nhwc(x, P['res%d_k1' % i], P['res%d_b1' % i], True), P['res%d_bn1' % i]), 0)
        t = _bn(_conv_nhwc(t, P['res%d_k2' % i], P['res%d_b2' % i], True), P['res%d_bn2' % i])
        x = np.maximum(t + x, 0)
    n = x.shape[0]
    p = np.maximum(_bn(_conv_nhwc(x, P['pol_k'], P['pol_b'], True), P['pol_bn']), 0).reshape(n, -1)     # NHWC reshape = HWC order
    logits = p @ P['pol_fc_w'] + P['pol_fc_b']
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    policy = e / e.sum(axis=1, keepdims=True)
    v = np.maximum(_bn(_conv_nhwc(x, P['val_k'], P['val_b'], True), P['val_bn']), 0).reshape(n, -1)
    v = np.maximum(v @ P['val_fc1_w'] + P['val_fc1_b'], 0)
    return policy, np.tanh(v @ P['val_fc2_w'] + P['val_fc2_b'])


def test_torch_restatement_matches_numpy_restatement():
    for size, blocks, seed in ((5, 1, 1), (7, 2, 2)):
        params = model.init_params(size, blocks, seed=seed, randomize_bn=True, random_bias=True)
        rs = np.random.RandomState(seed)
        boards = (rs.rand(3, size, size, 17) < 0.3).astype(np.float32)
        boards[..., 16] = np.where(rs.rand(3, 1, 1) < 0.5, 1.0, -1.0)
        pt, vt = tower_ref.forward(params, boards, dtype=torch.float64)
        pn, vn = _np_forward(params, boards)
        assert np.abs(pt.double().numpy() - pn).max() < 1e-6
        assert np.abs(vt.double().numpy() - vn).max() < 1e-6
        assert pn.shape == (3, size * size + 1) and vn.shape == (3, 1)


def test_parameter_count_matches_keras_summary():
    """conf.py defaults (19x19, 20 blocks): 24,043,731 parameters the way Keras counts them."""
    p = model.init_params(19, 20, seed=0)
    assert model.n_params(p) == 24043731

"""TEST INFRASTRUCTURE ONLY — fp32 PyTorch restatement of the reference network
(model.py:37-46 residual_block, model.py:55-96 build_model).

Parity status: UNPINNED by the reference — its arithmetic lives in
tensorflow-gpu==1.7.0 + Keras==2.2.2 (requirements.txt:24,10), neither installable
here, and the reference holds no golden policy/value vectors.  This
file restates the published Keras semantics:
  Conv2D: cross-correlation, kernel (kh,kw,in,out), use_bias=True; first conv
          padding='valid' (Q11), tower convs 'same'; BatchNormalization(eps=1e-3)
          in inference mode: (x-mean)/sqrt(var+eps)*gamma+beta; Reshape flattens
          [H,W,C] in HWC order; Dense = x @ W + b; softmax / relu / tanh.
"""
import contextlib
import torch
import torch.nn.functional as F

BN_EPS = 1e-3


@contextlib.contextmanager
def _no_tf32():
    """On a GPU, PyTorch's cuDNN convolutions default to TF32 (10-bit mantissa) for float32 inputs: a checker that
    is supposed to be fp32 (or better) must switch that off for the duration of its forward."""
    c, m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = c, m


def _bn(x, p):
    return (x - p['mean'].view(1, -1, 1, 1)) / torch.sqrt(p['var'].view(1, -1, 1, 1) + BN_EPS) * p['gamma'].view(1, -1, 1, 1) \
        + p['beta'].view(1, -1, 1, 1)


def _conv(x, k, b, same):
    # Keras kernel (kh,kw,in,out) -> torch (out,in,kh,kw)
    return F.conv2d(x, k.permute(3, 2, 0, 1).contiguous(), b, padding=(k.shape[0] // 2) if same else 0)


def forward(params, boards, device="cpu", dtype=torch.float32, raw=False):
    """boards [n,S,S,17] (any numeric) -> (policy [n,A], value [n,1]) in float32 (float64 when dtype is float64);
    raw=True also returns the pre-softmax logits [n,A] and the pre-tanh value [n,1].  True fp32 / fp64
    arithmetic on any device (TF32 is disabled for the call)."""
    with _no_tf32():
        return _forward(params, boards, device, dtype, raw)


def _forward(params, boards, device, dtype, raw):
    P = {k: ({kk: vv.to(device=device, dtype=dtype) for kk, vv in v.items()} if isinstance(v, dict)
             else v.to(device=device, dtype=dtype)) for k, v in params.items() if k != 'meta'}
    x = torch.as_tensor(boards).to(device=device, dtype=dtype).permute(0, 3, 1, 2).contiguous()
    x = F.relu(_bn(_conv(x, P['stem_k'], P['stem_b'], False), P['stem_bn']))                  # model.py:57-61
    nb = params['meta']['n_blocks']
    for i in range(nb):                                                                      # model.py:37-46
        t = F.relu(_bn(_conv(x, P['res%d_k1' % i], P['res%d_b1' % i], True), P['res%d_bn1' % i]))
        t = _bn(_conv(t, P['res%d_k2' % i], P['res%d_b2' % i], True), P['res%d_bn2' % i])
        x = F.relu(t + x)
    n = x.shape[0]
    p = F.relu(_bn(_conv(x, P['pol_k'], P['pol_b'], True), P['pol_bn']))                      # model.py:72-80
    p = p.permute(0, 2, 3, 1).reshape(n, -1)                                                 # HWC flatten
    logits = p @ P['pol_fc_w'] + P['pol_fc_b']
    policy = torch.softmax(logits, dim=1)
    v = F.relu(_bn(_conv(x, P['val_k'], P['val_b'], True), P['val_bn']))                      # model.py:82-92
    v = v.permute(0, 2, 3, 1).reshape(n, -1)
    v = F.relu(v @ P['val_fc1_w'] + P['val_fc1_b'])
    pre = v @ P['val_fc2_w'] + P['val_fc2_b']
    value = torch.tanh(pre)
    if dtype != torch.float64:
        policy, value, logits, pre = policy.float(), value.float(), logits.float(), pre.float()
    if raw:
        return policy, value, logits, pre
    return policy, value

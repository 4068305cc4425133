"""The evaluator: model.py:55-96's residual policy/value tower as a GPU-resident object.

`TowerModel` satisfies the reference's duck-typed model protocol
(`.name`, `.predict_on_batch(X[n,S,S,17]) -> (policy[n,S*S+1], value[n,1])`,
self_play.py:70,187; symmetry.py:130) for legacy callers and parity tests, and the
engine-side protocol (`.evaluate(engine, which, idx, syms, slot)`) that keeps
boards, activations and outputs in HBM.  Weight I/O in Keras .h5 is out of scope
(SURVEY §2 row 6); weights come from `init_params` (Keras-2.2.2 default
initialisers: glorot-uniform kernels, zero biases, BN gamma=1 beta=0 mean=0 var=1)
or from a dict in the same layout.
"""
import ctypes as C
import math
import os
import numpy as np
import torch

from . import _abi
from .conf import conf
from .engine import Engine, EngineError

BN_EPS = 1e-3
CHANNELS = 256


def _glorot(gen, shape, fan_in, fan_out):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * limit


def _bn_default(c):
    return dict(gamma=torch.ones(c), beta=torch.zeros(c), mean=torch.zeros(c), var=torch.ones(c))


def _bn_random(gen, c):
    return dict(gamma=0.5 + torch.rand(c, generator=gen), beta=0.2 * torch.randn(c, generator=gen),
                mean=0.2 * torch.randn(c, generator=gen), var=0.5 + torch.rand(c, generator=gen))


def init_params(size=None, n_blocks=None, seed=0, randomize_bn=False, random_bias=False):
    """Random-init parameters in the reference's Keras layout (model.py:55-96)."""
    size = size or conf['SIZE']
    n_blocks = conf['N_RESIDUAL_BLOCKS'] if n_blocks is None else n_blocks
    g = torch.Generator().manual_seed(seed)
    C_, W = CHANNELS, size - 2
    bn = (lambda c: _bn_random(g, c)) if randomize_bn else _bn_default
    bias = (lambda c: 0.1 * torch.randn(c, generator=g)) if random_bias else (lambda c: torch.zeros(c))
    p = {'meta': dict(size=size, n_blocks=n_blocks, seed=seed)}
    p['stem_k'] = _glorot(g, (3, 3, 17, C_), 9 * 17, 9 * C_)
    p['stem_b'] = bias(C_)
    p['stem_bn'] = bn(C_)
    for i in range(n_blocks):
        for j in (1, 2):
            p['res%d_k%d' % (i, j)] = _glorot(g, (3, 3, C_, C_), 9 * C_, 9 * C_)
            p['res%d_b%d' % (i, j)] = bias(C_)
            p['res%d_bn%d' % (i, j)] = bn(C_)
    F = 2 * W * W
    A = size * size + 1
    p['pol_k'] = _glorot(g, (1, 1, C_, 2), C_, 2); p['pol_b'] = bias(2); p['pol_bn'] = bn(2)
    p['pol_fc_w'] = _glorot(g, (F, A), F, A); p['pol_fc_b'] = bias(A)
    p['val_k'] = _glorot(g, (1, 1, C_, 2), C_, 2); p['val_b'] = bias(2); p['val_bn'] = bn(2)
    p['val_fc1_w'] = _glorot(g, (F, 256), F, 256); p['val_fc1_b'] = bias(256)
    p['val_fc2_w'] = _glorot(g, (256, 1), 256, 1); p['val_fc2_b'] = bias(1)
    return p


def n_params(p):
    """Parameter count the way Keras reports it (BN has 4 vectors per channel)."""
    t = 0
    for k, v in p.items():
        if k == 'meta':
            continue
        t += sum(x.numel() for x in v.values()) if isinstance(v, dict) else v.numel()
    return t


def fold_bn(k, b, bn):
    """conv/dense followed by inference BatchNorm -> one affine op (fp32)."""
    s = bn['gamma'] / torch.sqrt(bn['var'] + BN_EPS)
    return k * s.view(*([1] * (k.dim() - 1)), -1), (b - bn['mean']) * s + bn['beta']


def folded_arrays(p):
    """Tensors in the layout sgo_tower_weights expects (include/sejonggo_b200.h)."""
    nb = p['meta']['n_blocks']
    out = {}
    k, b = fold_bn(p['stem_k'], p['stem_b'], p['stem_bn'])
    out['stem_w'], out['stem_b'] = k.contiguous(), b.contiguous()
    ws, bs = [], []
    for i in range(nb):
        for j in (1, 2):
            k, b = fold_bn(p['res%d_k%d' % (i, j)], p['res%d_b%d' % (i, j)], p['res%d_bn%d' % (i, j)])
            ws.append(k.permute(0, 1, 3, 2).contiguous())          # (kh,kw,in,out) -> (kh,kw,out,in)
            bs.append(b)
    out['conv_w'] = torch.stack(ws).to(torch.bfloat16).contiguous() if ws else torch.zeros((0,), dtype=torch.bfloat16)
    out['conv_b'] = torch.stack(bs).contiguous() if bs else torch.zeros((0,))
    k, b = fold_bn(p['pol_k'], p['pol_b'], p['pol_bn'])
    out['pol_conv_w'], out['pol_conv_b'] = k.reshape(CHANNELS, 2).contiguous(), b.contiguous()
    out['pol_fc_w'], out['pol_fc_b'] = p['pol_fc_w'].contiguous(), p['pol_fc_b'].contiguous()
    k, b = fold_bn(p['val_k'], p['val_b'], p['val_bn'])
    out['val_conv_w'], out['val_conv_b'] = k.reshape(CHANNELS, 2).contiguous(), b.contiguous()
    out['val_fc1_w'], out['val_fc1_b'] = p['val_fc1_w'].contiguous(), p['val_fc1_b'].contiguous()
    out['val_fc2_w'], out['val_fc2_b'] = p['val_fc2_w'].reshape(256).contiguous(), p['val_fc2_b'].contiguous()
    return out


# ---- the weights as ONE blob: what a rank needs to play (BN folded, tower convs in bf16), in the order of
# sgo_tower_weights.  SURVEY §8e: this is what rank 0 broadcasts over NCCL at a model change (48 MB at 20 blocks)
# instead of the reference's scp of a Keras .h5 (slave_coordinator.py:45-63).
BLOB_KEYS = ("stem_w", "stem_b", "conv_w", "conv_b", "pol_conv_w", "pol_conv_b", "pol_fc_w", "pol_fc_b",
             "val_conv_w", "val_conv_b", "val_fc1_w", "val_fc1_b", "val_fc2_w", "val_fc2_b")
BLOB_ALIGN = 256


def blob_layout(size, n_blocks):
    """-> ([(key, dtype, shape, byte offset)], total bytes); a function of the architecture only, so every rank
    computes it without communication."""
    C_, W, A = CHANNELS, size - 2, size * size + 1
    F = 2 * W * W
    shapes = dict(stem_w=(3, 3, 17, C_), stem_b=(C_,), conv_w=(2 * n_blocks, 3, 3, C_, C_), conv_b=(2 * n_blocks, C_),
                  pol_conv_w=(C_, 2), pol_conv_b=(2,), pol_fc_w=(F, A), pol_fc_b=(A,), val_conv_w=(C_, 2), val_conv_b=(2,),
                  val_fc1_w=(F, 256), val_fc1_b=(256,), val_fc2_w=(256,), val_fc2_b=(1,))
    out, o = [], 0
    for k in BLOB_KEYS:
        dt = torch.bfloat16 if k == "conv_w" else torch.float32
        n = int(np.prod(shapes[k])) * (2 if dt == torch.bfloat16 else 4)
        out.append((k, dt, shapes[k], o))
        o = (o + n + BLOB_ALIGN - 1) // BLOB_ALIGN * BLOB_ALIGN
    return out, o


def pack_blob(folded, size, n_blocks, device):
    layout, total = blob_layout(size, n_blocks)
    blob = torch.zeros(total, dtype=torch.uint8, device=device)
    for k, dt, shape, o in layout:
        t = folded[k].to(device=device, dtype=dt).contiguous().reshape(-1)
        if t.numel():
            blob[o:o + t.numel() * t.element_size()] = t.view(torch.uint8)
    return blob


def unpack_blob(blob, size, n_blocks):
    """Typed views into the blob (no copies): the dict folded_arrays() would give."""
    layout, total = blob_layout(size, n_blocks)
    assert blob.numel() == total and blob.dtype == torch.uint8
    out = {}
    for k, dt, shape, o in layout:
        n = int(np.prod(shape)) * (2 if dt == torch.bfloat16 else 4)
        out[k] = blob[o:o + n].view(dt).reshape(shape) if n else torch.zeros(shape, dtype=dt, device=blob.device)
    return out


class TowerModel(object):
    is_sgo_evaluator = True

    def __init__(self, name="model_1", params=None, size=None, n_blocks=None, seed=0, max_positions=8192, folded=None):
        """params: a Keras-layout dict (init_params / load_params); or folded=(dict from unpack_blob, size, n_blocks) for a
        model received as a blob (it can play, but has no fp32 parameters to save)."""
        self.name = name
        if folded is not None:
            self._folded, self.size, self.n_blocks = folded
            self.params = None
        else:
            self.params = params if params is not None else init_params(size, n_blocks, seed)
            self.size, self.n_blocks = self.params['meta']['size'], self.params['meta']['n_blocks']
            self._folded = None
        self.max_positions = max_positions
        self._host_engine = None

    def folded(self):
        if self._folded is None:
            self._folded = folded_arrays(self.params)
        return self._folded

    def blob(self, device):
        return pack_blob(self.folded(), self.size, self.n_blocks, device)

    # ---- engine-side protocol ----------------------------------------------------
    def attach(self, engine, slot=0, max_positions=None):
        # which model occupies a weight slot is recorded ON the engine (ids of dead engines get reused)
        slots = engine.__dict__.setdefault("_tower_slots", {})
        if slots.get(slot, (None, 0))[0] is self:
            return
        if engine.S != self.size:
            raise EngineError("model is for %dx%d boards, engine is %dx%d" % (self.size, self.size, engine.S, engine.S))
        dev = {k: v.to(engine.device).contiguous() for k, v in self.folded().items()}
        w = _abi.SgoTowerWeights(n_blocks=self.n_blocks, channels=CHANNELS, size=self.size)
        for k, v in dev.items():
            setattr(w, k, v.data_ptr())
        mp = max_positions or self.max_positions
        engine._ck(engine.lib.sgo_tower_load_weights(engine.h, slot, C.byref(w), int(mp), engine._stream()))
        slots[slot] = (self, mp)
        torch.cuda.synchronize(engine.device)

    def evaluate(self, engine, which, idx, syms=None, slot=0, out=None):
        """idx int64/int32 device tensor of positions (games / leaf slots); returns compact
        device tensors (policy [k,A] f32, value [k] f32)."""
        self.attach(engine, slot)
        mp = engine._tower_slots[slot][1]
        k = int(idx.numel())
        idx32 = idx.to(torch.int32).contiguous()
        sy = None if syms is None else syms.to(torch.int32).contiguous()
        policy = torch.empty((k, engine.A), dtype=torch.float32, device=engine.device)
        value = torch.empty((k,), dtype=torch.float32, device=engine.device)
        for s in range(0, k, mp):
            n = min(mp, k - s)
            engine._ck(engine.lib.sgo_tower_forward(
                engine.h, slot, which, C.c_void_p(idx32[s:].data_ptr()), n,
                C.c_void_p(0 if sy is None else sy[s:].data_ptr()), 0,
                C.c_void_p(policy[s:].data_ptr()), C.c_void_p(value[s:].data_ptr()), engine._stream()))
        return policy, value

    def check(self, engine, slot=0):
        f = C.c_int32(0)
        engine._ck(engine.lib.sgo_tower_check_sync(engine.h, slot, C.byref(f), engine._stream()))
        if f.value:
            raise EngineError("tower kernel error flags 0x%x (16 = mbarrier wait timed out)" % f.value)

    def profile(self, engine, slot=0, enable=True):
        engine._ck(engine.lib.sgo_tower_profile(engine.h, slot, int(enable)))

    def profile_read(self, engine, slot=0):
        out = (C.c_double * 6)()
        engine._ck(engine.lib.sgo_tower_profile_read_sync(engine.h, slot, out))
        return dict(stem_ms=out[0], conv_ms=out[1], heads_ms=out[2], conv_launches=int(out[3]), positions=int(out[4]),
                    forwards=int(out[5]))

    def save(self, path):
        """Keras model.save (evaluator.py:20, model.py:117-119) for the .npz weight files."""
        if self.params is None:
            raise EngineError("a model received as a folded blob has no fp32 parameters to save")
        save_params(path, self.params, self.name)

    # ---- reference protocol (self_play.py:70,187) ----------------------------------
    def predict_on_batch(self, X):
        X = np.asarray(X)
        n = X.shape[0]
        if self._host_engine is None or self._host_engine.G < n:
            self._host_engine = Engine(size=self.size, n_games=max(n, 64), trees_per_game=1, max_leaves=1, arena_blocks=2)
        e = self._host_engine
        e.import_boards(X.astype(np.int32))
        idx = torch.arange(n, dtype=torch.int32, device=e.device)
        p, v = self.evaluate(e, 0, idx, None, slot=0)
        return p.cpu().numpy(), v.cpu().numpy().reshape(n, 1)


# ---- weight files: the reference keeps Keras .h5 files in MODEL_DIR (model.py:96-160); here the same
# directory protocol with .npz files of the params dict ("model_<n>.npz", BEST_MODEL) -----------------
def save_params(path, params, name):
    flat = {'__name__': np.array(name), '__meta__': np.array([params['meta']['size'], params['meta']['n_blocks'],
                                                              params['meta'].get('seed', 0)], np.int64)}
    for k, v in params.items():
        if k == 'meta':
            continue
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat["%s/%s" % (k, kk)] = vv.numpy()
        else:
            flat[k] = v.numpy()
    with open(path, 'wb') as f:                  # a file object: numpy must not append ".npz" to BEST_MODEL-style names
        np.savez(f, **flat)


def load_params(path):
    z = np.load(path)
    size, n_blocks, seed = (int(x) for x in z['__meta__'])
    p = {'meta': dict(size=size, n_blocks=n_blocks, seed=seed)}
    for k in z.files:
        if k.startswith('__'):
            continue
        t = torch.from_numpy(np.array(z[k]))
        if '/' in k:
            a, b = k.split('/')
            p.setdefault(a, {})[b] = t
        else:
            p[k] = t
    return str(z['__name__']), p


def load_model_by_name(filename, **kw):
    """model.py:157: a TowerModel from MODEL_DIR/<filename>."""
    name, params = load_params(os.path.join(conf['MODEL_DIR'], filename))
    return TowerModel(name, params=params, **kw)


def load_latest_model(**kw):
    """model.py:122-141: the highest-numbered model_<n> file in MODEL_DIR."""
    index, best = -1, None
    for filename in os.listdir(conf['MODEL_DIR']):
        try:
            i = int(filename.split('.')[0].split('_')[-1])
        except ValueError:
            continue
        if i > index:
            index, best = i, filename
    if best is None:
        raise FileNotFoundError("no model_<n> file in %s" % conf['MODEL_DIR'])
    return load_model_by_name(best, **kw)


def load_best_model(**kw):
    """model.py:144-155: BEST_MODEL, or a fresh model_1 when there is none yet."""
    path = os.path.join(conf['MODEL_DIR'], conf['BEST_MODEL'])
    if os.path.isfile(path):
        return load_model_by_name(conf['BEST_MODEL'], **kw)
    os.makedirs(conf['MODEL_DIR'], exist_ok=True)
    m = build_model("model_1")
    m.save(os.path.join(conf['MODEL_DIR'], "model_1.npz"))
    m.save(path)
    return m


def build_model(name, size=None, n_blocks=None, seed=0):
    """model.py:55 build_model: a random-init tower with the reference's architecture."""
    return TowerModel(name, size=size, n_blocks=n_blocks, seed=seed)

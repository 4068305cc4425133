"""TEST INFRASTRUCTURE ONLY — deterministic injected evaluator for MCTS parity.

A pure integer hash of the position gives a float32 policy / value that is
bit-identical on every machine, so the reference (run in the build container),
the C oracle and the CUDA engine can be fed *identical evaluator outputs*
(BASELINE.json: "MCTS visit counts must be bit-exact ... given identical
evaluator outputs").
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(h):
    h = (h ^ (h >> np.uint64(33))) * np.uint64(0xFF51AFD7ED558CCD)
    h = (h ^ (h >> np.uint64(33))) * np.uint64(0xC4CEB9FE1A85EC53)
    return h ^ (h >> np.uint64(33))


def hash_rows(a):
    """uint64 hash of every row of a 2-D unsigned integer array (packed states, legality masks): lets a fixture
    pin thousands of plies without storing them."""
    a = np.ascontiguousarray(a)
    a = a.reshape(a.shape[0], -1).astype(np.uint64)
    w = _mix(np.arange(1, a.shape[1] + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    with np.errstate(over='ignore'):
        return _mix((_mix(a + w[None, :]) * w[None, :]).sum(axis=1, dtype=np.uint64) + np.uint64(a.shape[1]))


def board_key(boards):
    """boards [n,S,S,17] (any numeric dtype) -> uint64[n] position hash."""
    b = np.asarray(boards)
    n = b.shape[0]
    flat = (b.reshape(n, -1) != 0).astype(np.uint64)
    flat[:, 16::17] = (b.reshape(n, -1)[:, 16::17] > 0).astype(np.uint64)
    w = _mix(np.arange(1, flat.shape[1] + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    with np.errstate(over='ignore'):
        return _mix((flat * w[None, :]).sum(axis=1, dtype=np.uint64) + np.uint64(0x1234567))


def evaluate(boards, salt=0, sharp=False):
    """-> (policy float32[n,A] in (0,1], unnormalised like a raw net head would
    not be — the reference never renormalises (Q13) so any positive vector is a
    valid test input), value float32[n] in [-1,1]."""
    b = np.asarray(boards)
    n, S = b.shape[0], b.shape[1]
    A = S * S + 1
    key = board_key(b) + np.uint64(salt)
    with np.errstate(over='ignore'):
        a = _mix(key[:, None] * np.uint64(0x100000001B3) + np.arange(1, A + 1, dtype=np.uint64)[None, :])
    u = ((a >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.float32)
    policy = (u + np.float32(1)) / np.float32(65536.0)
    if sharp:
        policy = policy * policy * policy
    policy = (policy / np.float32(A)).astype(np.float32)
    v = (_mix(key + np.uint64(77)) >> np.uint64(20)) % np.uint64(2001)
    value = (v.astype(np.float32) - np.float32(1000)) / np.float32(1000)
    return policy.astype(np.float32), value.astype(np.float32)


def evaluate_uniform(boards):
    """SURVEY §8d config 1 (BASELINE.json configs[0]): policy 1/(S*S+1) float32, value 0.0 — the
    reference's no-network CPU case; every PUCT score ties, so it is the tie-break stress test."""
    b = np.asarray(boards)
    n, S = b.shape[0], b.shape[1]
    A = S * S + 1
    return np.full((n, A), np.float32(1.0) / np.float32(A), np.float32), np.zeros(n, np.float32)


def evaluate_kind(boards, salt=0, sharp=False, kind="fake"):
    return evaluate_uniform(boards) if kind == "uniform" else evaluate(boards, salt, sharp)


class FakeModel(object):
    """predict_on_batch-compatible wrapper (self_play.py:70,187 protocol)."""

    def __init__(self, name="fake_model", salt=0, sharp=False, kind="fake"):
        self.name = name
        self.salt = salt
        self.sharp = sharp
        self.kind = kind
        self.calls = []

    def predict_on_batch(self, X):
        p, v = evaluate_kind(X, self.salt, self.sharp, self.kind)
        self.calls.append(np.asarray(X).shape[0])
        return p, v.reshape(-1, 1)

"""np.random.choice(moves, size=1, p=N/total) (self_play.py:149) restated from one uniform draw: fp64 p, cumulative sum
in child order, cdf /= cdf[-1], searchsorted(u, side='right') — the algorithm a numpy-exact device pick has to follow
(DESIGN.md §7).  Checked against numpy itself over random visit-count vectors."""
import numpy as np


def choice_from_u(moves, counts, u):
    total = int(np.sum(counts))
    p = np.array([int(c) / float(total) for c in counts], dtype=np.float64)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return moves[int(np.searchsorted(cdf, u, side='right'))]


def test_choice_from_u_equals_numpy_choice():
    gen = np.random.RandomState(123)
    for trial in range(400):
        k = int(gen.randint(1, 60))
        moves = sorted(gen.choice(362, size=k, replace=False).tolist())
        counts = gen.randint(1, 200, size=k)
        seed = int(gen.randint(1 << 30))
        a, b = np.random.RandomState(seed), np.random.RandomState(seed)
        total = int(counts.sum())
        want = int(a.choice(moves, size=1, p=[int(c) / float(total) for c in counts])[0])
        got = choice_from_u(moves, counts, b.random_sample())
        assert got == want, trial
        assert a.random_sample() == b.random_sample()           # both consumed exactly one double

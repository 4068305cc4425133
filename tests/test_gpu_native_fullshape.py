"""GPU parity of the BENCHMARKED path at the BENCHMARKED shapes: the native per-step driver
(sgo_selfplay_step: select -> tower forward with the symmetry gathers fused -> expand -> backup, csrc/driver.cu)
at 19x19 against the C oracle (oracle/go_oracle.c + oracle/game_loop.py) tree for tree:

  config 3, mode A   800 sims/ply = 8 simulate batches of 100 leaves    (self_play.py:28-152)
  config 3, mode B   800 sims/ply = 100 waves of ENERGY 8               (nomodel_self_play.py:59-140)
  config 4, match    1600 sims/ply, two networks, two trees per game, temperature 0, no noise (evaluator.py:23-36)

A few concurrent games, several plies with tree reuse, a small real TowerModel as the evaluator.  The oracle replays
every game with the SAME evaluator outputs: it calls the tower on its own leaf boards (rows of a forward are bitwise
independent of their position in the batch: test_tower_full_batch_is_position_independent) with the symmetry ids
the engine drew, and the engine's recorded coin / Dirichlet / uniform draws.  Checked per ply: the whole searched
tree (structure, counts, value sums, priors — bitwise) before the move pick, the pick (temperature 0: tuple max;
temperature 1: numpy's choice from the recorded uniform), and at the end the game records.
"""
import numpy as np
import pytest
import torch

from oracle import game_loop as gl
from tests.treeio import engine_rows, oracle_rows_fast, rows_equal
from tests.test_cpu_choice import choice_from_u

pytestmark = pytest.mark.gpu
S = 19
A = S * S + 1


def _rec_rng(seed):
    from sejonggo_b200.batched import HostRng

    class RecRng(HostRng):
        """HostRng that remembers what it drew (per game: coin, noise, root symmetries; rngs[0]: the uniforms)."""

        def __init__(self, seed):
            HostRng.__init__(self, seed)
            self.coins, self.noises, self.syms, self.us = [], [], [], []

        def coin(self):
            self.coins.append(HostRng.coin(self))
            return self.coins[-1]

        def dirichlet(self, n):
            self.noises.append(HostRng.dirichlet(self, n))
            return self.noises[-1]

        def symmetry(self):
            self.syms.append(HostRng.symmetry(self))
            return self.syms[-1]

        def uniform(self):
            self.us.append(HostRng.uniform(self))
            return self.us[-1]

    return RecRng(seed)


class _Replay(object):
    """The oracle-side rng of one game, fed from the engine run."""

    def __init__(self, g, rec, dev_syms, mode, L, waves, picks, us, stop_exploration):
        self.g, self.rec, self.dev_syms, self.mode, self.L, self.waves = g, rec, dev_syms, mode, L, waves
        self.picks, self.us, self.stop = picks, us, stop_exploration
        self.noise_i = self.step = self.root_i = self.k = self.ply = 0
        self.ctx = None

    def coin(self):
        return self.rec.coins[0]

    def dirichlet(self, n):
        self.noise_i += 1
        return self.rec.noises[self.noise_i - 1]

    def begin_root(self, move_n):
        self.ctx, self.ply = 'root', move_n

    def begin_wave(self, w):
        self.ctx, self.k = 'wave', 0
        self.step = self.ply * self.waves + w

    def symmetry(self):
        if self.mode == 'a':                              # one draw per game per simulate batch (self_play.py:70)
            self.step += 1
            return int(self.dev_syms[self.step - 1][self.g])
        if self.ctx == 'root':                            # host draw per root request (predicting_queue_worker.py:88-92)
            self.root_i += 1
            return self.rec.syms[self.root_i - 1]
        self.k += 1                                       # device draw per leaf slot of the wave
        return int(self.dev_syms[self.step][self.g * self.L + self.k - 1])

    def choice(self, moves, ps):
        # the engine sampled on the device from one uniform; numpy's algorithm on the ORACLE's counts must agree
        ply = self.ply
        u = self.us[ply][self.g]
        cdf = np.cumsum(np.asarray(ps, dtype=np.float64))
        cdf /= cdf[-1]
        want = int(moves[int(np.searchsorted(cdf, u, side='right'))])
        assert want == self.picks[ply], "temperature-1 pick: game %d ply %d engine %d numpy %d" % (self.g, ply, self.picks[ply], want)
        return want


def _run(mode, self_play, sims, G, plies, stop_exploration, seed):
    from sejonggo_b200 import model
    from sejonggo_b200.batched import BatchedGames
    batch = 100 if mode == 'a' else 8
    m1 = model.TowerModel("model_1", size=S, n_blocks=2, seed=11, max_positions=512)
    m2 = m1 if self_play else model.TowerModel("model_2", size=S, n_blocks=2, seed=12, max_positions=512)
    rngs = [_rec_rng(seed + g) for g in range(G)]
    bg = BatchedGames((m1, m2), G, size=S, mode=mode, mcts_batch_size=100, energy=8, mcts_simulations=sims,
                      stop_exploration=stop_exploration, self_play=self_play, rngs=rngs, record_boards='full',
                      arena_blocks=2 * plies * (sims + batch))
    assert bg.fast and bg.native_step                     # the path bench.py times
    dev_syms, trees = [], []
    draw = bg._draw_syms_device

    def rec_draw(per_leaf=False):
        t = draw(per_leaf)
        dev_syms.append(t.cpu().numpy())
        return t

    bg._draw_syms_device = rec_draw

    def after_search(b, tree_sel):
        row = []
        for g in range(G):
            if tree_sel[g] < 0:
                row.append(None)
                continue
            blocks, meta, p64 = b.eng.download_tree(g * b.eng.T + int(tree_sel[g]))
            assert not meta['overflow']
            row.append(engine_rows(blocks, meta, p64, A))
        trees.append(row)

    bg.after_search = after_search
    bg.start()
    sims0 = bg.sim_count
    for _ in range(plies):
        bg.step_ply(record=True)
    bg.eng.check_errors()
    m1.check(bg.eng, 0)
    if m2 is not m1:
        m2.check(bg.eng, 1)
    games = bg.finish()
    return bg, games, rngs, dev_syms, trees, (m1, m2), bg.sim_count - sims0


def _check_against_oracle(mode, self_play, sims, G, plies, stop_exploration, seed):
    bg, games, rngs, dev_syms, trees, (m1, m2), n_sims = _run(mode, self_play, sims, G, plies, stop_exploration, seed)
    L = bg.eng.L
    waves = int(sims / 8)
    us = [rngs[0].us[i * G:(i + 1) * G] for i in range(len(rngs[0].us) // G)]
    assert n_sims > 0.9 * G * plies * sims                # double passes aside, every game ran its simulations
    total_nodes = 0
    for g in range(G):
        got = games[g]
        picks = [mv['move'][0] + S * mv['move'][1] for mv in got['moves']]
        rng = _Replay(g, rngs[g], dev_syms, mode, L, waves, picks, us, stop_exploration)
        seen = []

        def on_search(move_n, tree, current, g=g, seen=seen):
            want = oracle_rows_fast(tree)
            have = trees[move_n][g]
            assert have is not None, (g, move_n)
            ok, why = rows_equal(have, want)
            assert ok, "game %d ply %d: %s" % (g, move_n, why)
            seen.append(len(want))

        if mode == 'a':
            ref = gl.play_game(m1, m2, sims, stop_exploration, self_play=self_play, num_moves=plies, size=S,
                               mcts_batch_size=100, rng=rng, on_search=on_search)
        else:
            tags = {"BEST_SYM": m1}
            ref = gl.play_game_async("BEST_SYM", "BEST_SYM", 8, stop_exploration, 0, self_play=self_play, num_moves=plies, size=S,
                                     conf_sims=sims, conf_energy=8, rng=rng, names={"BEST_SYM": m1.name}, on_search=on_search,
                                     predict=lambda tag, b, sym: (lambda p, v: (p[0], v[0]))(*gl.sym_predict(tags[tag], b, sym)))
        assert len(seen) == len(got['moves']) >= 1
        total_nodes += sum(seen)
        assert len(got['moves']) == len(ref['moves'])
        for a, b in zip(got['moves'], ref['moves']):
            assert a['move'] == b['move'] and a['player'] == b['player'] and a['move_n'] == b['move_n']
            assert np.array_equal(a['board'], b['board'])
            assert np.array_equal(np.asarray(a['policy'], np.float64).view(np.uint64), np.asarray(b['policy'], np.float64).view(np.uint64))
            assert np.float32(a['value']).view(np.uint32) == np.float32(b['value']).view(np.uint32)
        assert got['result'] == ref['result'] and got['winner'] == ref['winner'] and got['winner_model'] == ref['winner_model']
        assert got['modelB_name'] == ref['modelB_name'] and got['modelW_name'] == ref['modelW_name']
    return total_nodes


def test_config3_mode_a_native_800_sims_tree_for_tree():
    """BASELINE configs[2] mode A: 8 x 100-leaf simulate batches per ply, Dirichlet-noised fp64 root, shared tree
    re-rooted across plies, temperature 1 then 0."""
    n = _check_against_oracle('a', True, 800, G=6, plies=4, stop_exploration=2, seed=100)
    assert n > 6 * 4 * 800 * 100                          # rows compared: > 800 expanded nodes x ~360 children per tree


def test_config3_mode_b_native_100_waves_tree_for_tree():
    """BASELINE configs[2] mode B: 100 ENERGY-8 waves per ply with one symmetry per predict request."""
    n = _check_against_oracle('b', True, 800, G=4, plies=3, stop_exploration=2, seed=200)
    assert n > 4 * 3 * 700 * 100


def test_config4_match_native_1600_sims_two_trees():
    """BASELINE configs[3]: evaluator.evaluate semantics — two weight sets, a tree per model, no noise,
    temperature 0 from ply 0, 16 simulate batches per ply, one symmetry per game per batch."""
    n = _check_against_oracle('a', False, 1600, G=4, plies=4, stop_exploration=0, seed=300)
    assert n > 4 * 4 * 1600 * 100

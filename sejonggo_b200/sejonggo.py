"""Mirror of the reference's GTP front-end sejonggo.py: `SejongGoEngine` (sejonggo.py:19-68, one game
with tree reuse across `play`/`genmove`) and `GTPEngine` (:70-160, the GTP text commands) on a
batch-of-one engine: board, tree and network stay in HBM, each `genmove` runs
int(MCTS_SIMULATIONS / MCTS_BATCH_SIZE) mode-A search steps (self_play.select_play).

    python -m sejonggo_b200.sejonggo            # GTP on stdin/stdout with load_best_model()
"""
import string
import sys

import numpy as np
import torch

from .conf import conf
from .batched import as_evaluator, HostRng
from .engine import Engine, EngineError

__version__ = "b200-0.1"
COLOR_TO_PLAYER = {'B': 1, 'W': -1, 'b': 1, 'w': -1, 'black': 1, 'white': -1}


class SejongGoEngine(object):
    def __init__(self, model, mcts_simulations, board=None, resign=None, temperature=0, add_noise=False, size=None,
                 mcts_batch_size=None, rng=None, use_symmetry=True, device=0):
        self.model = model
        self.ev = as_evaluator(model)
        self.mcts_simulations = mcts_simulations
        self.resign = resign
        self.temperature = temperature
        self.add_noise = add_noise
        self.S = size or conf['SIZE']
        self.A = self.S * self.S + 1
        self.batch = mcts_batch_size or conf['MCTS_BATCH_SIZE']
        self.rng = rng or HostRng()
        self.use_symmetry = use_symmetry
        self.move = 1
        steps = int(mcts_simulations / self.batch)
        self.eng = Engine(size=self.S, n_games=1, trees_per_game=1, max_leaves=self.batch,
                          arena_blocks=max(1024, 64 * (steps * self.batch + self.batch)), komi=conf['KOMI'], device=device)   # one game: the whole pool is its tree's (680 MB at 1,600 sims)
        self.eng.reset()
        self.eng.tree_reset()
        if board is not None:
            self.eng.import_boards(np.asarray(board).astype(np.int32))
        self.player = int(self.eng.export_packed(0)[0, -1].item())

    # the reference keeps the numpy board as an attribute; here it is exported on demand
    @property
    def board(self):
        return self.eng.export_boards().cpu().numpy()

    @property
    def mcts_tree(self):
        return bool(self.eng.tree_valid()[0].item())

    def set_temperature(self, temperature):
        self.temperature = temperature

    def clear(self):
        self.eng.reset()
        self.eng.tree_reset()
        self.move = 1
        self.player = 1

    def play(self, color, x, y, update_tree=True):
        """sejonggo.py:35-46: keep the subtree under the move if the tree has it, else drop the tree."""
        e = self.eng
        index = self.S * self.S if y == self.S else y * self.S + x
        if update_tree and self.mcts_tree:
            in_subtree = int(e.legal_masks()[0, index].item()) == 0       # children exist exactly for the legal moves (play.py:396)
            if in_subtree:
                e.reroot(np.array([index], np.int32))
            else:
                e.tree_reset()
        e.apply_moves(np.array([index], np.int32), None if color is None else np.array([color], np.int32))
        try:
            e.check_errors()
        except EngineError as ex:
            raise AssertionError(str(ex))                                 # play.py:233 asserts the point is empty
        self.move += 1
        self.player = int(e.export_packed(0)[0, -1].item())
        return self.board, self.player

    def _evaluate_root(self):
        idx = torch.zeros(1, dtype=torch.int64, device=self.eng.device)
        p, v = self.ev.evaluate(self.eng, 0, idx, None, slot=0)           # no symmetry at the root (sejonggo.py:49)
        return p, v

    def _search(self):
        e = self.eng
        for _ in range(int(self.mcts_simulations / self.batch)):         # self_play.py:128
            e.select_a(self.batch)
            idx, n = e.leaf_compact()
            policy = torch.zeros((e.L, e.A), dtype=torch.float32, device=e.device)
            value = torch.zeros((e.L,), dtype=torch.float32, device=e.device)
            if n:
                syms = None
                if self.use_symmetry:                                     # one draw per predict batch (symmetry.py:128)
                    syms = torch.full((n,), self.rng.symmetry(), dtype=torch.int32, device=e.device)
                p, v = self.ev.evaluate(e, 1, idx.long(), syms, slot=0)
                policy[idx.long()] = p
                value[idx.long()] = v
            e.expand(policy, value)
            e.backup_a()

    def genmove(self, color):
        """sejonggo.py:48-68 -> (x, y, policy_target, value, board, player)."""
        e = self.eng
        policy, value = self._evaluate_root()
        value_h = float(value[0].item())
        if self.resign and value_h <= self.resign:
            return 0, self.S + 1, policy[0].cpu().numpy(), value_h, self.board, self.player
        if not self.mcts_tree:
            noise = None
            if self.add_noise:
                noise = np.asarray(self.rng.dirichlet(self.A), np.float64).reshape(1, self.A)
            e.tree_new(policy, noise=noise, eps=conf['DIRICHLET_EPSILON'], force=True)
        self._search()
        prior, count, _ = e.child_stats(want=("prior", "count"))
        forced = None
        if self.temperature == 1:
            c = count[0].cpu().numpy()
            nz = np.nonzero(c)[0]
            forced = np.array([self.rng.choice([int(m) for m in nz], [int(c[m]) / float(c.sum()) for m in nz])], np.int32)
        index = int(e.pick(np.array([self.temperature], np.int32), None, forced)[0].item())
        x, y = index % self.S, index // self.S
        policy_target = prior[0].cpu().numpy()
        board, player = self.play(color, x, y)
        return x, y, policy_target, value_h, board, player


class GTPEngine(object):
    def __init__(self, model=None, mcts_simulations=None, **kw):
        self._komi = 0
        if model is None:
            from .model import load_best_model
            model = load_best_model()
        self.SIZE = kw.get('size') or conf['SIZE']
        self.sejong_engine = SejongGoEngine(model, mcts_simulations or conf['MCTS_SIMULATIONS'], **kw)
        self.player = self.sejong_engine.player

    @property
    def board(self):
        return self.sejong_engine.board

    def name(self):
        return "SejongGo - {} - {} simulations".format(self.sejong_engine.model.name, self.sejong_engine.mcts_simulations)

    def version(self):
        return __version__

    def protocol_version(self):
        return "2"

    def list_commands(self):
        return ""

    def boardsize(self, size):
        size = int(size)
        if size != self.SIZE:
            raise Exception("The board size in configuration is {0}x{0} but GTP asked to play {1}x{1}".format(self.SIZE, size))
        return ""

    def komi(self, komi):
        self._komi = komi
        return ""

    def parse_move(self, move):
        """sejonggo.py:103-119: GTP vertex -> (x, y) with y counted from the top; pass = (0, SIZE)."""
        if move.lower() == 'pass':
            return 0, self.SIZE
        x = string.ascii_uppercase.index(move[0].upper())
        if x >= 9:
            x -= 1                      # I is a skipped letter
        y = int(move[1:]) - 1
        return x, self.SIZE - y - 1

    def print_move(self, x, y):
        """sejonggo.py:121-129 (the reference prints a pass as column A of row 0 too; kept)."""
        y = self.SIZE - y - 1
        if x >= 8:
            x += 1
        return string.ascii_uppercase[x] + str(y + 1)

    def play(self, color, move):
        x, y = self.parse_move(move)
        _, self.player = self.sejong_engine.play(COLOR_TO_PLAYER[color], x, y)
        return ""

    def genmove(self, color):
        x, y, _, _, _, self.player = self.sejong_engine.genmove(COLOR_TO_PLAYER[color])
        return self.print_move(x, y)

    def clear_board(self):
        self.sejong_engine.clear()
        self.player = 1
        return ""

    def parse_command(self, line):
        tokens = line.strip().split(" ")
        method = getattr(self, tokens[0])
        result = method(*tokens[1:])
        if not result.strip():
            return "=\n\n"
        return "= " + result + "\n\n"


def main():
    engine = GTPEngine()
    for line in sys.stdin:
        for cmd in line.split("\n"):
            if not cmd.strip():
                continue
            if cmd.strip() == "quit":
                sys.stdout.write("=\n\n")
                return
            sys.stdout.write(engine.parse_command(cmd))
            sys.stdout.flush()


if __name__ == "__main__":
    main()

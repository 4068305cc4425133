"""The 8-GPU self-play launcher: the batched form of the reference's main_selfplay.main (main_selfplay.py:9-29)
plus what its slave coordinator did around it (slave_coordinator.py:45-82).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \\
        -m sejonggo_b200.main_selfplay [--games N] [--concurrent 1024] [--sims 800] [--mode b]

One process per GPU (the reference: one predicting worker per GPU + N_GAME_PROCESS game processes).  Per round of the
reference's `while True` loop:
  1. rank 0 loads the best model from MODEL_DIR; its name is compared with the last one played — "No new best model
     for self-playing. Stopping.." ends the run (main_selfplay.py:18-24);
  2. the network goes to every rank as ONE folded blob over NCCL (dist.broadcast_model);
  3. the still-unplayed games of range(N_GAMES) are sharded g mod world; every rank keeps `concurrent` games in
     HBM and refills a slot the moment its game ends (selfplay_worker.py:81-124), with the resignation calibration
     kept per rank (the reference keeps it per worker process);
  4. the record rows (board, move, value, policy_target per ply) stay in HBM and are gathered to rank 0 device to
     device every `--gather-every` plies; rank 0 alone writes SELF_PLAY_DIR/<model>/game_%05d/move_%03d/sample.*
     (sgfsave.save_self_play_data), so the training side sees the directory tree the reference's workers wrote.
"""
import argparse
import os
import sys

import numpy as np
import torch

from .conf import conf
from . import dist as sd, model as M
from .batched import BatchedGames, HostRng
from .records import games_from_rows
from .self_play import ResignationCalibrator
from .sgfsave import save_self_play_data


def _unplayed(model_name, n_games):
    root = os.path.join(conf['SELF_PLAY_DIR'], model_name)
    return [g for g in range(n_games) if not os.path.isdir(os.path.join(root, "game_%05d" % g))]


class RankSelfPlay(object):
    """One rank's share of a self-play round: BatchedGames with the device record store, plus the gather."""

    def __init__(self, model, games, rank, world, device, concurrent, sims, mode='b', size=None, seed=0, num_moves=None,
                 arena_blocks=None, gather_every=16, save=True):
        self.model, self.games, self.rank, self.world = model, list(games), rank, world
        self.size = size or conf['SIZE']
        self.gather_every, self.save, self.mode = gather_every, save, mode
        self.cal = ResignationCalibrator()
        self.saved, self.rows_gathered = [], []
        self.bg = None
        if self.games:
            self.bg = BatchedGames((model, model), min(concurrent, len(self.games)), size=self.size, mode=mode,
                                   mcts_batch_size=conf['MCTS_BATCH_SIZE'], energy=conf['ENERGY'], mcts_simulations=sims,
                                   stop_exploration=conf['STOP_EXPLORATION'], self_play=True, num_moves=num_moves,
                                   komi=conf['KOMI'], dirichlet_eps=conf['DIRICHLET_EPSILON'], rng=HostRng(seed + rank),
                                   arena_blocks=arena_blocks, device=device if isinstance(device, int) else 0,
                                   record_boards='device', n_total=len(self.games),
                                   on_game_start=self.cal.start, on_game_end=self.cal.end)
        self.dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)

    def _gather(self):
        from .records import row_words
        rows = self.bg.store.take() if self.bg is not None else torch.zeros((0, row_words(self.size)), dtype=torch.int32, device=self.dev)
        got = sd.gather_rows(rows, dst=0)
        if got is None:
            return
        for r, t in enumerate(got):
            self.rows_gathered.append(int(t.numel()) * 4)
            self._pending[r] = np.concatenate([self._pending[r], t.cpu().numpy().view(np.uint32).reshape(-1, t.shape[1])])
            done = games_from_rows(self._pending[r], self.size, names=(self.model.name, self.model.name), mode=self.mode)
            if done:
                keep = ~np.isin(self._pending[r][:, 0], np.array(sorted(done), np.uint32))
                self._pending[r] = self._pending[r][keep]
                for local_id, gd in sorted(done.items()):
                    game_no = self._game_no(r, local_id)
                    if gd['moves'] and self.save:
                        save_self_play_data(self.model.name, game_no, gd, size=self.size)
                    self.saved.append(game_no)

    def _game_no(self, rank, local_id):
        return self.all_games[rank::self.world][local_id]

    def run(self, all_games):
        """all_games: the round's global game numbers (every rank passes the same list); this rank plays all_games[rank::world]."""
        from .records import row_words
        self.all_games = list(all_games)
        self._pending = [np.zeros((0, row_words(self.size)), np.uint32) for _ in range(self.world)]
        if self.bg is not None:
            self.bg.start()
        plies = 0
        while True:
            more = self.bg.step_ply(record=True) if self.bg is not None else False
            plies += 1
            flag = torch.tensor([1.0 if more else 0.0], device=self.dev)
            if self.world > 1:
                torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX)    # ranks leave the loop together
            last = flag.item() == 0
            if last or plies % self.gather_every == 0:
                self._gather()
            if last:
                break
        if self.bg is not None:
            self.bg.finish()
        return sorted(self.saved)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=None, help="conf['N_GAMES']")
    ap.add_argument("--concurrent", type=int, default=None, help="games in HBM per GPU (conf['CONCURRENT_GAMES'])")
    ap.add_argument("--sims", type=int, default=None, help="conf['MCTS_SIMULATIONS']")
    ap.add_argument("--mode", default="b", choices=["a", "b"], help="b = main_selfplay.py's virtual-loss waves; a = self_play.py batches")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--blocks", type=int, default=None)
    ap.add_argument("--num-moves", type=int, default=None)
    ap.add_argument("--gather-every", type=int, default=16)
    ap.add_argument("--max-rounds", type=int, default=0, help="0 = until there is no new best model")
    ap.add_argument("--backend", default=None)
    a = ap.parse_args(argv)
    sys.setrecursionlimit(10000)
    rank, world, local = sd.init(a.backend)
    cuda = torch.cuda.is_available()
    dev = local if cuda else "cpu"
    if a.size:
        conf['SIZE'] = a.size
    if a.blocks is not None:
        conf['N_RESIDUAL_BLOCKS'] = a.blocks
    n_games = a.games if a.games is not None else conf['N_GAMES']
    sims = a.sims or conf['MCTS_SIMULATIONS']
    finished_best_model_name, rounds = None, 0
    while True:
        model = M.load_best_model(max_positions=16384) if rank == 0 else None
        model = sd.broadcast_model(model, src=0, device=torch.device("cuda", local) if cuda else None, max_positions=16384)
        if model.name == finished_best_model_name:
            if rank == 0:
                print("No new best model for self-playing. Stopping..")
            break
        finished_best_model_name = model.name
        if rank == 0:
            print("SELF-PLAYING BEST MODEL ", model.name)
        # every rank must see the same list: rank 0 scans the directory and shares it
        todo = [_unplayed(model.name, n_games) if rank == 0 else None]
        if world > 1:
            torch.distributed.broadcast_object_list(todo, src=0)
        games = todo[0]
        rs = RankSelfPlay(model, games[rank::world], rank, world, dev, a.concurrent or conf['CONCURRENT_GAMES'], sims, mode=a.mode,
                          size=conf['SIZE'], seed=1000 * rounds, num_moves=a.num_moves, gather_every=a.gather_every)
        saved = rs.run(games)
        if rank == 0:
            print("round %d: %d games saved, %.1f MB of records gathered" % (rounds, len(saved), sum(rs.rows_gathered) / 1e6))
        rounds += 1
        if a.max_rounds and rounds >= a.max_rounds:
            break
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

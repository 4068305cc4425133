"""BatchedGames — G concurrent games advanced in lock-step on one GPU.

This is the batched form of the reference's two game drivers:
  mode 'a'  self_play.play_game          (self_play.py:164-290; search self_play.py:28-152)
  mode 'b'  nomodel_self_play.play_game_async (nomodel_self_play.py:142-271; waves :59-82)
Every per-game decision (tree reuse, noise only on new trees, temperature switch,
pass/pass and 2*S*S termination, resignation, value sign, quirks Q12-Q17) follows
the reference; what changes is that the rules, the tree and (with a TowerModel)
the network all stay in HBM and one kernel launch serves all games.

Randomness: `rng` must provide coin()/dirichlet(n)/symmetry()/choice(moves, ps);
the default HostRng draws from numpy like the reference.  Parity tests inject
recorded draws.
"""
import numpy as np
import torch

from .engine import Engine, EngineError


class HostRng(object):
    def __init__(self, seed=None, alpha=0.03):
        self.r = np.random.RandomState(seed)
        self.alpha = alpha
        self.device_pick = True      # temperature-1 picks sampled on the device from u01

    def coin(self):
        return float(self.r.random_sample())

    def dirichlet(self, n):
        return self.r.dirichlet([self.alpha] * n)

    def symmetry(self):
        return int(self.r.randint(7))          # random.choice over the 7 SYMMETRIES (symmetry.py:117-128)

    def uniform(self):
        return float(self.r.random_sample())

    def choice(self, moves, ps):
        return int(self.r.choice(moves, size=1, p=ps)[0])


class HostModelAdapter(object):
    """Evaluator over the reference's duck-typed model protocol
    (`.predict_on_batch(X[n,S,S,17]) -> (policy[n,A], value[n,1])`, `.name`)."""

    def __init__(self, model):
        self.model = model
        self.name = getattr(model, "name", "model")

    def evaluate(self, engine, which, idx, syms, slot=0):
        """idx: int64 device tensor of position indices (games or leaf slots);
        syms: int32 device tensor [len(idx)] or None.  Returns device (policy [k,A], value [k])."""
        k = int(idx.numel())
        planes = engine.export_planes_indexed(which, idx, syms)         # only the requested positions leave the device
        p, v = self.model.predict_on_batch(planes.cpu().numpy())
        p = torch.as_tensor(np.ascontiguousarray(p, dtype=np.float32)).to(engine.device).reshape(k, engine.A)
        v = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float32)).to(engine.device).reshape(k)
        if syms is not None:
            p = engine.policy_unsym(p, syms=syms)
        return p, v


def as_evaluator(model):
    if hasattr(model, "evaluate") and hasattr(model, "is_sgo_evaluator"):
        return model
    return HostModelAdapter(model)


class BatchedGames(object):
    def __init__(self, models, n_games, size=19, mode='a', mcts_batch_size=100, energy=8, mcts_simulations=1600,
                 stop_exploration=30, self_play=False, num_moves=None, resign=(None, None), komi=5.5,
                 dirichlet_eps=0.25, use_symmetry=True, root_symmetry=None, rng=None, rngs=None, arena_blocks=None,
                 device=0, record_boards='full', engine=None, native_step=True, names=None, n_total=None,
                 on_game_start=None, on_game_end=None, rng_for_game=None):
        """models: (model1, model2); pass the same object twice for self-play.
        n_games: concurrent games (slots in HBM); n_total: games to play in all (default n_games) — with more
        games than slots, a slot starts its next game as soon as one ends.
        resign: (resign_model1, resign_model2), each None, a float, or an array[G] (per slot; on_game_start(game_id)
        may return a (r1, r2) pair for the game that is starting, e.g. from a calibration in progress, or a dict
        {'resign': (r1, r2), 'num_moves': n} to cap that game's length, or False to skip the id).
        on_game_end(game_id, game_data) is called as each game finishes.
        rngs: optional list of per-slot rng objects (parity runs); else one shared `rng`; rng_for_game(game_id)
        hands each starting game its own rng (parity runs with more games than slots)."""
        self.m1, self.m2 = models
        # names reported in game_data; given explicitly when a tag resolves to another model's network (Q21)
        self.names = names or (getattr(self.m1, "name", "model1"), getattr(self.m2, "name", "model2"))
        self.same_model = self.m1 is self.m2
        self.ev = [as_evaluator(self.m1), as_evaluator(self.m1) if self.same_model else as_evaluator(self.m2)]
        if self.same_model:
            self.ev[1] = self.ev[0]
        self.G, self.S, self.A = n_games, size, size * size + 1
        self.n_total = n_games if n_total is None else int(n_total)
        self.on_game_start, self.on_game_end, self.rng_for_game = on_game_start, on_game_end, rng_for_game
        self.mode = mode
        self.batch = mcts_batch_size if mode == 'a' else energy
        self.energy = energy
        self.sims = mcts_simulations
        self.stop_exploration = stop_exploration
        self.self_play = self_play
        self.num_moves = size * size * 2 if num_moves is None else num_moves
        self.komi = komi
        self.eps = dirichlet_eps
        self.use_symmetry = use_symmetry
        # mode A evaluates the root without symmetry (self_play.py:187); mode B's *_SYM tags
        # send the root through random_symmetry_predict too (predicting_queue_worker.py:88-92)
        self.root_symmetry = (mode == 'b') if root_symmetry is None else root_symmetry
        self.rngs = rngs if rngs is not None else [rng or HostRng()] * n_games
        self.record_boards = record_boards
        steps = int(self.sims / self.batch)
        if arena_blocks is None:
            # node blocks per tree ON AVERAGE (all trees share one pool): a ply adds `sims` blocks and re-rooting
            # keeps the played child's share r of them, so a tree settles near sims / (1 - r); 8x covers r = 0.875
            # for every game at once, and a single tree may take far more than its share
            arena_blocks = max(256, 8 * (steps * self.batch + self.batch))
        T = 1 if self_play else 2
        self.eng = engine or Engine(size=size, n_games=n_games, trees_per_game=T, max_leaves=self.batch,
                                    arena_blocks=arena_blocks, komi=komi, device=device)
        self.resign = [self._per_game(resign[0]), self._per_game(resign[1])]
        # device fast path: every evaluator is a GPU tower and no rng injects recorded picks (per-game rngs that
        # allow device-side sampling keep it: their coin / noise / root-symmetry draws stay per game)
        self.fast = all(hasattr(ev, "is_sgo_evaluator") for ev in self.ev) and \
            (rngs is None or all(getattr(r, "device_pick", False) for r in rngs)) and \
            (rng_for_game is None or getattr(rng_for_game, "device_pick", False))
        self.after_search = None      # parity hook: called as after_search(self, tree_sel) between the search and the move pick
        # with GPU towers the whole search step runs inside the library (sgo_selfplay_step)
        self.native_step = native_step and self.fast and all(hasattr(ev, "attach") for ev in self.ev)
        self.store = None             # records.RecordStore when record_boards == 'device' (rows stay in HBM for the gather)
        if record_boards == 'device':
            from .records import RecordStore
            self.store = RecordStore(size, max(4096, 64 * n_games), self.eng.device)
        self.trees_dropped = 0        # trees given up for lack of pool room (_ensure_pool_room)
        self.sim_count = 0            # leaves expanded + backed up (the north-star "simulations")
        self.eval_count = 0
        self.plies_done = 0

    def _per_game(self, r):
        if r is None:
            return np.full(self.G, np.nan)
        return np.broadcast_to(np.asarray(r, dtype=np.float64), (self.G,)).copy()

    # --------------------------------------------------------------- pieces
    def _evaluate(self, which, idx_list, model_of_game, syms_of_game):
        """Evaluate positions `idx_list` (numpy int64: games or leaf slots).  Returns dense device
        buffers policy [lim, A], value [lim] with rows of idx filled."""
        e = self.eng
        lim = e.G * e.L if which else e.G
        policy = torch.zeros((lim, e.A), dtype=torch.float32, device=e.device)
        value = torch.zeros((lim,), dtype=torch.float32, device=e.device)
        if len(idx_list) == 0:
            return policy, value
        games = idx_list // e.L if which else idx_list
        for mi in ((0,) if self.same_model else (0, 1)):
            sel = idx_list if self.same_model else idx_list[model_of_game[games] == mi]
            if len(sel) == 0:
                continue
            idx = torch.as_tensor(sel, dtype=torch.int64, device=e.device)
            syms = None
            if syms_of_game is not None:
                g = sel // e.L if which else sel
                syms = torch.as_tensor(syms_of_game[g], dtype=torch.int32, device=e.device)
            p, v = self.ev[mi].evaluate(e, which, idx, syms, slot=mi)
            policy[idx] = p
            value[idx] = v
            self.eval_count += len(sel)
        return policy, value

    def _search_a(self, tree_sel, active, model_of_game):
        e = self.eng
        for _ in range(int(self.sims / self.batch)):                    # self_play.py:128
            e.select_a(self.batch, tree_sel)
            counts = e.leaf_counts().cpu().numpy()
            slots = np.concatenate([g * e.L + np.arange(counts[g]) for g in range(self.G)]).astype(np.int64) \
                if counts.sum() else np.zeros(0, np.int64)
            syms = None
            if self.use_symmetry:                                       # one draw per game per batch (symmetry.py:128)
                syms = np.zeros(self.G, np.int32)
                for g in np.nonzero(active)[0]:
                    syms[g] = self.rngs[g].symmetry()
            policy, value = self._evaluate(1, slots, model_of_game, syms)
            e.expand(policy, value, tree_sel)
            e.backup_a(tree_sel)
            self.sim_count += int(counts.sum())

    def _search_b(self, tree_sel, active, model_of_game):
        e = self.eng
        for _ in range(int(self.sims / self.energy)):                   # nomodel_self_play.py:116
            restart = True
            prev = np.zeros(self.G, np.int64)
            while True:
                newly, stalled = e.select_b(self.batch, restart, tree_sel)
                restart = False
                if newly == 0:
                    break
                counts = e.leaf_counts().cpu().numpy().astype(np.int64)
                slots = np.concatenate([g * e.L + np.arange(prev[g], counts[g]) for g in range(self.G)]).astype(np.int64)
                prev = counts
                policy = torch.zeros((e.G * e.L, e.A), dtype=torch.float32, device=e.device)
                value = torch.zeros((e.G * e.L,), dtype=torch.float32, device=e.device)
                # put_predict_request: one request (and one symmetry draw) per leaf, in issue order
                for s in slots:
                    g = int(s // e.L)
                    sy = None
                    if self.use_symmetry:
                        sy = np.zeros(self.G, np.int32)
                        sy[g] = self.rngs[g].symmetry()
                    p1, v1 = self._evaluate(1, np.array([s], np.int64), model_of_game, sy)
                    policy[s] = p1[s]
                    value[s] = v1[s]
                e.expand(policy, value, tree_sel)
                self.sim_count += len(slots)
                if stalled == 0:
                    break
            e.backup_b(self.energy, tree_sel)

    def _search_b_fast(self, tree_sel, active, model_of_game):
        """Mode B with one batched evaluation per wave phase (throughput path).  Identical
        tree results to _search_b; only the order of host RNG symmetry draws across games differs."""
        e = self.eng
        for _ in range(int(self.sims / self.energy)):
            restart = True
            prev = np.zeros(self.G, np.int64)
            while True:
                newly, stalled = e.select_b(self.batch, restart, tree_sel)
                restart = False
                if newly == 0:
                    break
                counts = e.leaf_counts().cpu().numpy().astype(np.int64)
                slots = np.concatenate([g * e.L + np.arange(prev[g], counts[g]) for g in range(self.G)]).astype(np.int64)
                prev = counts
                syms = None
                if self.use_symmetry:
                    syms = np.array([self.rngs[g].symmetry() if active[g] else 0 for g in range(self.G)], np.int32)
                policy, value = self._evaluate(1, slots, model_of_game, syms)
                e.expand(policy, value, tree_sel)
                self.sim_count += len(slots)
                if stalled == 0:
                    break
            e.backup_b(self.energy, tree_sel)

    # ------------------------------------------------- device-resident search (TowerModel evaluators)
    def _eval_leaves_device(self, tree_sel_dev, cur_model_dev, syms_game):
        """Evaluate every leaf slot awaiting evaluation; returns dense per-slot buffers."""
        e = self.eng
        if getattr(self, "_pbuf", None) is None:
            self._pbuf = torch.zeros((e.G * e.L, e.A), dtype=torch.float32, device=e.device)
            self._vbuf = torch.zeros((e.G * e.L,), dtype=torch.float32, device=e.device)
        idx, n = e.leaf_compact()
        if n == 0:
            return 0
        games = torch.div(idx, e.L, rounding_mode='floor').long()
        if syms_game is None:
            syms = None
        elif syms_game.numel() == e.G * e.L:             # per leaf slot (mode B)
            syms = syms_game[idx.long()].contiguous()
        else:
            syms = syms_game[games].contiguous()
        groups = [(0, idx, syms)]
        if not self.same_model:
            mm = cur_model_dev[games]
            groups = [(mi, idx[mm == mi].contiguous(), None if syms is None else syms[mm == mi].contiguous()) for mi in (0, 1)]
        for mi, ix, sy in groups:
            if ix.numel() == 0:
                continue
            p, v = self.ev[mi].evaluate(e, 1, ix, sy, slot=mi)
            self._pbuf[ix.long()] = p
            self._vbuf[ix.long()] = v
        self.eval_count += n
        return n

    def _draw_syms_device(self, per_leaf=False):
        if not self.use_symmetry:
            return None
        r = self.rngs[0]
        seed = int(r.r.randint(1 << 31)) if hasattr(r, "r") else 0
        g = torch.Generator(device=self.eng.device).manual_seed(seed)
        n = self.G * self.eng.L if per_leaf else self.G
        return torch.randint(0, 7, (n,), generator=g, device=self.eng.device, dtype=torch.int32)

    def _search_native(self, tree_sel_dev, cur_model_dev):
        """The search of one ply through sgo_selfplay_step: the host only loops over steps/waves."""
        e = self.eng
        for mi in ((0,) if self.same_model else (0, 1)):
            self.ev[mi].attach(e, mi)
        mg = None if self.same_model else cur_model_dev.to(torch.int32)
        # self_play.py:128 int(sims / MCTS_BATCH_SIZE) simulate calls; nomodel_self_play.py:116 int(SIMS / conf ENERGY) waves (Q17)
        for _ in range(int(self.sims / (self.batch if self.mode == 'a' else self.energy))):
            if self.mode == 'a':      # one symmetry per game per simulate batch (self_play.py:70)
                n = e.selfplay_step('a', self.batch, self.energy, tree_sel_dev, mg, sym_game=self._draw_syms_device())
            else:                     # one symmetry per predict request = per leaf (predicting_queue_worker.py:88-92)
                n = e.selfplay_step('b', self.batch, self.energy, tree_sel_dev, mg, sym_leaf=self._draw_syms_device(per_leaf=True))
            self.sim_count += n
            self.eval_count += n

    def _search_a_device(self, tree_sel_dev, cur_model_dev):
        e = self.eng
        for _ in range(int(self.sims / self.batch)):
            e.select_a(self.batch, tree_sel_dev)
            n = self._eval_leaves_device(tree_sel_dev, cur_model_dev, self._draw_syms_device())
            e.expand(self._pbuf, self._vbuf, tree_sel_dev)
            e.backup_a(tree_sel_dev)
            self.sim_count += n

    def _search_b_device(self, tree_sel_dev, cur_model_dev):
        e = self.eng
        for _ in range(int(self.sims / self.energy)):
            restart = True
            syms_wave = self._draw_syms_device(per_leaf=True)               # one draw per leaf slot per wave
            while True:
                newly, stalled = e.select_b(self.batch, restart, tree_sel_dev)
                restart = False
                if newly == 0:
                    break
                n = self._eval_leaves_device(tree_sel_dev, cur_model_dev, syms_wave)
                e.expand(self._pbuf, self._vbuf, tree_sel_dev)
                self.sim_count += n
                if stalled == 0:
                    break
            e.backup_b(self.energy, tree_sel_dev)

    # ------------------------------------------------------------------ run
    def run(self, exact_rng_order=True):
        """Plays n_total games (default: one per slot) and returns their game_data in game order.  With
        n_total > n_games a slot starts its next game the ply after one ends (selfplay_worker.py:81-124:
        the reference's worker loop), so the batch stays full until the games run out."""
        self.start()
        while self.step_ply(exact_rng_order=exact_rng_order):
            pass
        return self.finish()

    def start(self):
        e, G = self.eng, self.G
        e.reset()
        e.tree_reset()
        self.active = np.zeros(G, bool)
        self.cur_model = np.zeros(G, np.int32)
        self.model1_isblack = np.ones(G, bool)
        self.skipped_last = np.zeros(G, bool)
        self.end_reason = np.array(["PLAYED ALL MOVES"] * G, dtype=object)
        self.player = np.ones(G, np.int32)          # move_data['player'] (lags one ply, self_play.py:236)
        self.moves_rec = [[] for _ in range(G)]
        self.move_n = np.zeros(G, np.int32)         # ply counter of the game in each slot
        self.num_moves_slot = np.full(G, self.num_moves, np.int32)     # ply cap of the game in each slot
        self.slot_game = np.full(G, -1, np.int64)   # id of the game a slot is playing
        self.next_game = 0
        self.results = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._begin_games(np.arange(G), fresh=True)

    def _begin_games(self, slots, fresh=False):
        """Start the next games in `slots` (ascending): coin for the colours (play.py:301-306), thresholds from
        on_game_start; a re-used slot gets game_init and its trees go back to the pool."""
        slots = [int(g) for g in slots]
        started = []
        for g in slots:
            gid, r = None, None
            while self.next_game < self.n_total:               # on_game_start may decline an id (a game another worker claimed)
                gid = self.next_game
                self.next_game += 1
                r = self.on_game_start(gid) if self.on_game_start is not None else None
                if r is not False:
                    break
                gid = None
            if gid is None:
                break
            started.append(g)
            self.slot_game[g] = gid
            if self.rng_for_game is not None:
                self.rngs[g] = self.rng_for_game(gid)
            # choose_first_player: cur_model[g] in {0,1} indexes (model1, model2)
            self.cur_model[g] = 0 if self.rngs[g].coin() < .5 else 1
            self.model1_isblack[g] = self.cur_model[g] == 0
            self.active[g] = True
            self.skipped_last[g] = False
            self.end_reason[g] = "PLAYED ALL MOVES"
            self.player[g] = 1
            self.moves_rec[g] = []
            self.move_n[g] = 0
            self.num_moves_slot[g] = self.num_moves
            if isinstance(r, dict):                 # per-game settings: {'resign': (r1, r2), 'num_moves': n}
                self.num_moves_slot[g] = r.get('num_moves', self.num_moves)
                r = r.get('resign')
            if r is not None:
                r1, r2 = r
                self.resign[0][g] = np.nan if r1 is None else r1
                self.resign[1][g] = np.nan if r2 is None else r2
        if started and not fresh:
            mask = np.zeros(self.G, np.int32)
            mask[started] = 1
            self.eng.games_restart(mask)
            self.h2d_bytes += mask.nbytes

    def _end_games(self, slots):
        """Score and package the games that just ended (before their slots are re-used)."""
        if len(slots) == 0:
            return
        sc = self.eng.score().cpu().numpy()
        for g in slots:
            gd = self._game_data(int(g), sc[g])
            gid = int(self.slot_game[g])
            if self.store is not None:
                from .records import footer_row
                self.store.append_host_rows(footer_row(self.S, gid, gd))
            self.results[gid] = gd
            self.slot_game[g] = -1
            self.moves_rec[g] = []
            if self.on_game_end is not None:
                self.on_game_end(gid, gd)

    def step_ply(self, exact_rng_order=True, record=True):
        """One ply of every active game.  Returns False when no game is left."""
        e, G, S, A = self.eng, self.G, self.S, self.A
        active, cur_model = self.active, self.cur_model
        # games that have played all their moves end here (the reference's `for move_n in range(num_moves)`)
        over = active & (self.move_n >= self.num_moves_slot)
        if over.any():
            active &= ~over
            self._end_games(np.nonzero(over)[0])
        # slots whose game has ended take the next game
        idle = np.nonzero(~active & (self.slot_game < 0))[0]
        if len(idle) and self.next_game < self.n_total:
            self._begin_games(idle)
        if not active.any():
            return False
        move_n = self.move_n
        temps = np.where(move_n >= self.stop_exploration, 0, 1).astype(np.int32)       # self_play.py:183-184
        tree_sel = np.where(active, 0 if self.self_play else cur_model, -1).astype(np.int32)
        act_idx = np.nonzero(active)[0].astype(np.int64)
        # root evaluation (self_play.py:187 / nomodel_self_play.py:165)
        rsyms = None
        if self.root_symmetry and self.use_symmetry:
            rsyms = np.zeros(G, np.int32)
            for g in act_idx:
                rsyms[g] = self.rngs[g].symmetry()
        policy, value = self._evaluate(0, act_idx, cur_model, rsyms)
        value_h = value.cpu().numpy()
        self.d2h_bytes += value_h.nbytes
        # resignation (self_play.py:190-193)
        thr = np.where(cur_model == 0, self.resign[0], self.resign[1])
        with np.errstate(invalid='ignore'):
            resigning = active & ~np.isnan(thr) & (thr != 0) & (value_h <= thr)
        if resigning.any():
            for g in np.nonzero(resigning)[0]:
                self.end_reason[g] = "resign"
            active &= ~resigning
            self._end_games(np.nonzero(resigning)[0])
            if not active.any():
                return self.next_game < self.n_total
        tree_sel = np.where(active, 0 if self.self_play else cur_model, -1).astype(np.int32)
        act_idx = np.nonzero(active)[0].astype(np.int64)
        self._ensure_pool_room(int(active.sum()))
        # new trees only where there is no reusable subtree (self_play.py:195-198)
        valid = e.tree_valid(tree_sel).cpu().numpy()
        need = active & (valid == 0)
        if need.any():
            noise = None
            if self.self_play:
                noise = np.zeros((G, A), np.float64)
                for g in np.nonzero(need)[0]:
                    noise[g] = self.rngs[g].dirichlet(A)
                self.h2d_bytes += noise.nbytes
            e.tree_new(policy, noise=noise, eps=self.eps, force=False, tree_sel=np.where(need, tree_sel, -1).astype(np.int32))
        # search
        if self.fast:
            ts_dev = e.dev(tree_sel, torch.int32)
            cm_dev = e.dev(cur_model, torch.int64)
            self.h2d_bytes += tree_sel.nbytes + cur_model.nbytes
            if self.native_step:
                self._search_native(ts_dev, cm_dev)
            elif self.mode == 'a':
                self._search_a_device(ts_dev, cm_dev)
            else:
                self._search_b_device(ts_dev, cm_dev)
        elif self.mode == 'a':
            self._search_a(tree_sel, active, cur_model)
        elif exact_rng_order:
            self._search_b(tree_sel, active, cur_model)
        else:
            self._search_b_fast(tree_sel, active, cur_model)
        if not self.native_step:
            e.check_errors()              # (the native step reports a failed allocation itself)
        if self.after_search is not None:
            self.after_search(self, tree_sel)
        # move pick (self_play.py:138-152)
        forced = None
        u01 = None
        prior_h = None
        packed_rec = record and self.record_boards == 'packed'
        device_rec = record and self.record_boards == 'device'
        if record and not packed_rec and not device_rec:
            prior, count, _ = e.child_stats(tree_sel, want=("prior", "count"))
            prior_h = prior.cpu().numpy()                       # policy_target = root priors (Q14)
            self.d2h_bytes += prior_h.nbytes
        explore = active & (temps == 1)
        if explore.any() and self.fast:
            u01 = np.array([self.rngs[0].uniform() for _ in range(G)], np.float64)    # device-side sampling
            self.h2d_bytes += u01.nbytes
        elif explore.any():
            if not record or packed_rec or device_rec:
                _, count, _ = e.child_stats(tree_sel, want=("count",))
            count_h = count.cpu().numpy()
            forced = np.full(G, -1, np.int32)
            for g in np.nonzero(explore)[0]:
                nz = np.nonzero(count_h[g])[0]
                total = int(count_h[g].sum())
                forced[g] = self.rngs[g].choice([int(m) for m in nz], [int(count_h[g][m]) / float(total) for m in nz])
        index_dev = e.pick(temps, u01, forced, tree_sel)
        index = index_dev.cpu().numpy()
        self.d2h_bytes += index.nbytes
        boards_h = None
        if record and self.record_boards == 'full':
            boards_h = e.export_boards().cpu().numpy()
        elif packed_rec:
            # one packed row per game: board words, move, value, policy_target (sgo_records_pack)
            rec_dev = e.records_pack(tree_sel, index_dev, value)
            if getattr(self, "_rec_pinned", None) is None:               # device -> pinned host, one copy per ply
                self._rec_pinned = torch.empty(tuple(rec_dev.shape), dtype=rec_dev.dtype, pin_memory=True)
            self._rec_pinned.copy_(rec_dev, non_blocking=True)
            torch.cuda.current_stream(e.device).synchronize()
            rec_h = self._rec_pinned.numpy().view(np.uint32)
            PW = e.packed_words
            boards_h = rec_h[:, :PW]
            prior_h = rec_h[:, PW + 3:PW + 3 + A].view(np.float32)
            self.d2h_bytes += rec_h.nbytes - boards_h.nbytes
        elif device_rec:
            # the rows stay in HBM (records.RecordStore) until they are gathered; the host keeps what the game flow needs
            self.store.append_plies(e.records_pack(tree_sel, index_dev, value), act_idx, self.slot_game[act_idx], move_n[act_idx])
        if boards_h is not None:
            self.d2h_bytes += boards_h.nbytes
        apply = np.full(G, -1, np.int32)
        ended = []
        for g in act_idx:
            idx = int(index[g])
            x, y = idx % S, idx // S
            if record:
                self.moves_rec[g].append(dict(
                    board=None if boards_h is None else (boards_h[g:g + 1].copy() if self.record_boards == 'full' else boards_h[g].copy()),
                    policy=None if prior_h is None else prior_h[g].copy(), value=value_h[g], move=(x, y), move_n=int(move_n[g]),
                    player=int(self.player[g])))
            if self.skipped_last[g] and y == S:
                self.end_reason[g] = "BOTH_PASSED"
                active[g] = False
                ended.append(g)
                continue
            self.skipped_last[g] = y == S
            apply[g] = idx
        # update trees, play the move, swap sides (self_play.py:223-238)
        e.reroot(apply)
        mover = np.where(apply >= 0, np.where(move_n % 2 == 0, 1, -1), 0)      # black moves on even plies
        e.apply_moves(apply)
        self.h2d_bytes += 2 * apply.nbytes
        moved = apply >= 0
        self.player = np.where(moved, mover, self.player).astype(np.int32)
        self.cur_model = np.where(moved, 1 - cur_model, cur_model).astype(np.int32)
        self.move_n = np.where(moved, move_n + 1, move_n).astype(np.int32)
        self.plies_done += int(moved.sum())
        self._end_games(ended)
        return bool(active.any()) or self.next_game < self.n_total

    def _ensure_pool_room(self, n_active):
        """The search of one ply allocates at most `sims` node blocks per game (+ a root).  The reference's trees live
        in unbounded host memory; here they share a pool, so before each ply the pool must hold that much.  When it
        does not, the games with the largest trees give theirs up and search this ply from a new tree — what the
        reference does whenever a tree is empty (self_play.py:195-198) — instead of any game running short mid-search.
        Counted in `trees_dropped`; never triggered while the pool is sized generously (the default)."""
        e = self.eng
        per = int(self.sims / (self.batch if self.mode == 'a' else self.energy)) * (self.batch if self.mode == 'a' else self.energy)
        need = n_active * (per + 1)
        free = e.pool_stats()['free']
        if free >= need:
            return
        sizes = e.tree_sizes().cpu().numpy().sum(axis=1)
        mask = np.zeros(self.G, np.int32)
        for g in np.argsort(-sizes, kind='stable'):
            if free >= need or sizes[g] == 0:
                break
            mask[g] = 1
            free += int(sizes[g])
            self.trees_dropped += 1
        if mask.any():
            e.tree_free(mask)
        if free < need:
            raise EngineError("MCTS node pool too small for one ply of %d games x %d simulations (capacity %d blocks): raise arena_blocks"
                              % (n_active, per, e.pool_stats()['capacity']))

    def finish(self):
        """Closes the games still in their slots (cut short by the caller) and returns every game's game_data
        in game order."""
        self.eng.check_errors()
        open_slots = np.nonzero(self.slot_game >= 0)[0]
        self.active[open_slots] = False
        self._end_games(open_slots)
        return [self.results[k] for k in sorted(self.results)]

    def _to_move(self):
        # plane 16 of every game (+1 black / -1 white): packed word PW-1
        pk = self.eng.export_packed(0)
        return pk[:, -1].cpu().numpy().astype(np.int32)

    def _game_data(self, g, sc):
        """game_data of the game in slot g (self_play.py:240-290); sc = its (winner, black, white) score row."""
        ps = {1: "B", 0: "D", -1: "W"}
        n1, n2 = self.names
        black, white = float(sc[1]), float(sc[2]) + self.komi
        winner = 1 if black > white else (0 if black == white else -1)
        if self.end_reason[g] == "resign":
            result = "%s+R" % ps[int(self.player[g])]
        else:
            result = "%s+%s" % (ps[winner], abs(black - white))
        isblack = bool(self.model1_isblack[g])
        modelB, modelW = (n1, n2) if isblack else (n2, n1)
        if winner == 0:
            winner_model = None
        elif self.mode == 'a':
            winner_model = n1 if (winner == 1) == isblack else n2       # self_play.py:261
        else:
            # nomodel_self_play.py:246-249 picks between the B/W names with the same test (reference quirk)
            winner_model = modelB if (winner == 1) == isblack else modelW
        return dict(moves=self.moves_rec[g], modelB_name=modelB, modelW_name=modelW,
                    winner={1: 1, -1: 0, 0: None}[winner], winner_model=winner_model, result=result,
                    resign_model1=None if np.isnan(self.resign[0][g]) else float(self.resign[0][g]),
                    resign_model2=None if np.isnan(self.resign[1][g]) else float(self.resign[1][g]),
                    end_reason=str(self.end_reason[g]), model1_isblack=isblack, game_id=int(self.slot_game[g]))

"""Engine — thin Python host over the C ABI (include/sejonggo_b200.h).

torch is used only for device buffers and streams; every kernel is ours and is
reached through ctypes.  There is no CPU path: constructing an Engine without a
CUDA device raises."""
import ctypes as C
import numpy as np
import torch

from . import _abi

NODEBLOCK_DTYPE = np.dtype([('prior', 'f4', 384), ('n', 'i4', 384), ('w', 'f4', 384), ('child', 'i4', 384),
                            ('exist', 'u4', 12), ('busy', 'u4', 12), ('parent_block', 'i4'), ('parent_slot', 'i4'),
                            ('pad', 'i4', 6)])
assert NODEBLOCK_DTYPE.itemsize == 6272
META_FIELDS = ('root', 'n_blocks', 'valid', 'root_f64', 'root_count', 'root_value', 'overflow', 'pad')

ERR_BITS = {1: "make_play on an occupied point (play.py:233 assert)", 2: "MCTS node pool exhausted (raise arena_blocks)",
            4: "selector found no child (reference would raise)", 8: "tree deeper than SGO_MAXDEPTH"}


class EngineError(RuntimeError):
    pass


class Engine(object):
    def __init__(self, size=19, n_games=1, trees_per_game=1, max_leaves=100, arena_blocks=2048, komi=5.5, device=0):
        if not torch.cuda.is_available():
            raise EngineError("sejonggo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _abi.load()
        self.S, self.A, self.G, self.T, self.L, self.NB = size, size * size + 1, n_games, trees_per_game, max_leaves, arena_blocks
        self.komi = float(komi)
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)            # make sure the primary context exists
        cfg = _abi.SgoConfig(device=device, size=size, n_games=n_games, trees_per_game=trees_per_game,
                             max_leaves=max_leaves, arena_blocks=arena_blocks, komi=komi)
        h = C.c_void_p()
        rc = self.lib.sgo_create(C.byref(cfg), C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.lib.sgo_last_error(h).decode() if h else "sgo_create failed"
            raise EngineError("sgo_create rc=%d: %s" % (rc, msg))
        self.packed_words = 16 * ((size * size + 31) // 32) + 1

    def close(self):
        if getattr(self, "h", None):
            self.lib.sgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc):
        if rc != 0:
            raise EngineError("rc=%d: %s" % (rc, self.lib.sgo_last_error(self.h).decode()))

    def dev(self, x, dtype):
        """numpy / list / tensor -> contiguous device tensor of `dtype`."""
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device)

    @staticmethod
    def _p(t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    def launch_count(self):
        return int(self.lib.sgo_launch_count(self.h))

    def check_errors(self):
        f = C.c_int32(0)
        self._ck(self.lib.sgo_check_errors_sync(self.h, self._stream(), C.byref(f)))
        if f.value:
            raise EngineError("; ".join(v for k, v in ERR_BITS.items() if f.value & k))

    # -------------------------------------------------------------- rules
    def reset(self, first=0, n=None):
        n = self.G - first if n is None else n
        self._ck(self.lib.sgo_games_reset(self.h, first, n, self._stream()))

    def apply_moves(self, moves, colors=None, first=0):
        m = self.dev(moves, torch.int32)
        c = self.dev(colors, torch.int32)
        self._ck(self.lib.sgo_apply_moves(self.h, first, m.numel(), self._p(m), self._p(c), self._stream()))

    def legal_masks(self, first=0, n=None):
        n = self.G - first if n is None else n
        out = torch.empty((n, self.A), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.sgo_legal_masks(self.h, first, n, self._p(out), self._stream()))
        return out

    def score(self, first=0, n=None):
        n = self.G - first if n is None else n
        out = torch.empty((n, 3), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_score(self.h, first, n, self._p(out), self._stream()))
        return out

    def import_boards(self, boards, first=0):
        b = self.dev(np.asarray(boards).reshape(-1, self.S, self.S, 17), torch.int32)
        self._ck(self.lib.sgo_import_boards(self.h, first, b.shape[0], self._p(b), self._stream()))

    def export_boards(self, first=0, n=None):
        n = self.G - first if n is None else n
        out = torch.empty((n, self.S, self.S, 17), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_export_boards(self.h, first, n, self._p(out), self._stream()))
        return out

    def export_packed(self, which=0, first=0, n=None):
        lim = self.G * self.L if which else self.G
        n = lim - first if n is None else n
        out = torch.empty((n, self.packed_words), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_export_packed(self.h, which, first, n, self._p(out), self._stream()))
        return out

    def random_playouts(self, seed, max_plies, first=0, n=None):
        n = self.G - first if n is None else n
        moves = torch.empty((n, max_plies), dtype=torch.int16, device=self.device)
        nplies = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_random_playouts(self.h, first, n, C.c_uint64(seed), max_plies, self._p(moves), self._p(nplies), self._stream()))
        return moves, nplies

    def export_planes(self, which=0, first=0, n=None, sym=0, syms=None, out=None):
        lim = self.G * self.L if which else self.G
        n = lim - first if n is None else n
        if out is None:
            out = torch.empty((n, self.S, self.S, 17), dtype=torch.float32, device=self.device)
        sy = self.dev(syms, torch.int32)
        self._ck(self.lib.sgo_export_planes(self.h, which, first, n, int(sym), self._p(sy), self._p(out), self._stream()))
        return out

    def export_planes_indexed(self, which, index, syms=None):
        """float32 [k][S][S][17] planes of the positions index[k] (games or leaf slots), per-position symmetry ids."""
        ix = self.dev(index, torch.int32)
        sy = self.dev(syms, torch.int32)
        k = int(ix.numel())
        out = torch.empty((k, self.S, self.S, 17), dtype=torch.float32, device=self.device)
        self._ck(self.lib.sgo_export_planes_indexed(self.h, which, self._p(ix), k, self._p(sy), self._p(out), self._stream()))
        return out

    def policy_unsym(self, policy, sym=0, syms=None):
        p = self.dev(policy, torch.float32).reshape(-1, self.A)
        out = torch.empty_like(p)
        sy = self.dev(syms, torch.int32)
        self._ck(self.lib.sgo_policy_unsym(self.h, p.shape[0], int(sym), self._p(sy), self._p(p), self._p(out), self._stream()))
        return out

    # --------------------------------------------------------------- tree
    def tree_reset(self):
        self._ck(self.lib.sgo_tree_reset(self.h, self._stream()))

    def tree_free(self, game_mask):
        m = self.dev(game_mask, torch.int32)
        assert m.numel() == self.G
        self._ck(self.lib.sgo_tree_free(self.h, self._p(m), self._stream()))

    def games_restart(self, game_mask):
        """game_init + drop the trees of the flagged games (a slot starting its next game)."""
        m = self.dev(game_mask, torch.int32)
        assert m.numel() == self.G
        self._ck(self.lib.sgo_games_restart(self.h, self._p(m), self._stream()))

    def tree_sizes(self):
        out = torch.empty((self.G, self.T), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_tree_sizes(self.h, self._p(out), self._stream()))
        return out

    def pool_stats(self):
        out = (C.c_int64 * 4)()
        self._ck(self.lib.sgo_pool_stats_sync(self.h, out, self._stream()))
        return dict(capacity=int(out[0]), free=int(out[1]), min_free=int(out[2]), failed_allocs=int(out[3]))

    def tree_new(self, policy, noise=None, eps=0.25, force=False, tree_sel=None):
        p = self.dev(policy, torch.float32)
        assert p.numel() == self.G * self.A
        nz = self.dev(noise, torch.float64)
        ts = self.dev(tree_sel, torch.int32)
        self._ck(self.lib.sgo_tree_new(self.h, self._p(ts), self._p(p), self._p(nz), float(eps), int(force), self._stream()))

    def select_a(self, batch, tree_sel=None):
        ts = self.dev(tree_sel, torch.int32)
        self._ck(self.lib.sgo_tree_select_a(self.h, self._p(ts), int(batch), self._stream()))

    def select_b(self, energy, restart=True, tree_sel=None):
        ts = self.dev(tree_sel, torch.int32)
        counts = (C.c_int32 * 2)()
        self._ck(self.lib.sgo_tree_select_b_sync(self.h, self._p(ts), int(energy), int(restart), counts, self._stream()))
        return counts[0], counts[1]

    def expand(self, policy, value, tree_sel=None):
        p = self.dev(policy, torch.float32)
        v = self.dev(value, torch.float32)
        assert p.numel() == self.G * self.L * self.A and v.numel() == self.G * self.L
        ts = self.dev(tree_sel, torch.int32)
        self._ck(self.lib.sgo_tree_expand(self.h, self._p(ts), self._p(p), self._p(v), self._stream()))

    def backup_a(self, tree_sel=None):
        ts = self.dev(tree_sel, torch.int32)
        self._ck(self.lib.sgo_tree_backup_a(self.h, self._p(ts), self._stream()))

    def backup_b(self, total_energy, tree_sel=None):
        ts = self.dev(tree_sel, torch.int32)
        self._ck(self.lib.sgo_tree_backup_b(self.h, self._p(ts), int(total_energy), self._stream()))

    def pick(self, temperature=None, u01=None, forced=None, tree_sel=None):
        t = self.dev(temperature, torch.int32)
        u = self.dev(u01, torch.float64)
        f = self.dev(forced, torch.int32)
        ts = self.dev(tree_sel, torch.int32)
        out = torch.empty((self.G,), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_tree_pick(self.h, self._p(ts), self._p(t), self._p(u), self._p(f), self._p(out), self._stream()))
        return out

    def reroot(self, moves):
        m = self.dev(moves, torch.int32)
        assert m.numel() == self.G
        self._ck(self.lib.sgo_tree_reroot(self.h, self._p(m), self._stream()))

    def child_stats(self, tree_sel=None, want=("prior", "count", "value")):
        ts = self.dev(tree_sel, torch.int32)
        pr = torch.empty((self.G, self.A), dtype=torch.float64, device=self.device) if "prior" in want else None
        ct = torch.empty((self.G, self.A), dtype=torch.int32, device=self.device) if "count" in want else None
        vl = torch.empty((self.G, self.A), dtype=torch.float32, device=self.device) if "value" in want else None
        self._ck(self.lib.sgo_tree_child_stats(self.h, self._p(ts), self._p(pr), self._p(ct), self._p(vl), self._stream()))
        return pr, ct, vl

    def leaf_counts(self):
        out = torch.empty((self.G,), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_leaf_counts(self.h, self._p(out), self._stream()))
        return out

    def leaf_compact(self):
        """-> (int32 device tensor of leaf slots awaiting evaluation, count)"""
        if getattr(self, "_compact_buf", None) is None:
            self._compact_buf = torch.empty((self.G * self.L,), dtype=torch.int32, device=self.device)
        n = C.c_int32(0)
        self._ck(self.lib.sgo_leaf_compact_sync(self.h, self._p(self._compact_buf), C.byref(n), self._stream()))
        return self._compact_buf[:n.value], n.value

    def selfplay_step(self, mode, leaves, total_energy=0, tree_sel=None, model_of_game=None, sym_game=None, sym_leaf=None):
        """One search step of every game inside the library (select -> evaluate on the loaded
        network slot(s) -> expand -> backup).  mode 'a' = self_play.simulate, 'b' = one
        async_simulate2 wave.  Returns the number of simulations performed."""
        ts = self.dev(tree_sel, torch.int32)
        mg = self.dev(model_of_game, torch.int32)
        sg = self.dev(sym_leaf if sym_leaf is not None else sym_game, torch.int32)     # per leaf slot [G*L] or per game [G]
        assert sg is None or sg.numel() == (self.G * self.L if sym_leaf is not None else self.G)
        n = C.c_int32(0)
        self._ck(self.lib.sgo_selfplay_step(self.h, 0 if mode in (0, 'a') else 1, self._p(ts), self._p(mg), int(leaves),
                                            int(total_energy), self._p(sg), int(sym_leaf is not None), C.byref(n), self._stream()))
        return n.value

    def record_words(self):
        return int(self.lib.sgo_record_words(self.h))

    def records_pack(self, tree_sel=None, moves=None, values=None):
        """uint32 [G][record_words()] packed move_data rows (board, move, value, valid, policy_target)."""
        ts = self.dev(tree_sel, torch.int32)
        mv = self.dev(moves, torch.int32)
        vl = self.dev(values, torch.float32)
        out = torch.empty((self.G, self.record_words()), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_records_pack(self.h, self._p(ts), self._p(mv), self._p(vl), self._p(out), self._stream()))
        return out

    def tree_valid(self, tree_sel=None):
        ts = self.dev(tree_sel, torch.int32)
        out = torch.empty((self.G,), dtype=torch.int32, device=self.device)
        self._ck(self.lib.sgo_tree_valid(self.h, self._p(ts), self._p(out), self._stream()))
        return out

    def download_tree(self, tree=0):
        """-> (blocks, meta, root_p64) in canonical form: block 0 is the root, the rest in breadth-first slot order,
        child / parent links index into `blocks` (the pool ids a tree occupies are an allocation detail)."""
        meta = np.zeros(8, dtype=np.int32)
        p64 = np.zeros(384, dtype=np.float64)
        self._ck(self.lib.sgo_tree_download_sync(self.h, tree, None, 0, meta.ctypes.data_as(C.c_void_p), None))
        nb = max(1, int(meta[1]))
        blocks = np.zeros(nb, dtype=NODEBLOCK_DTYPE)
        self._ck(self.lib.sgo_tree_download_sync(self.h, tree, blocks.ctypes.data_as(C.c_void_p), nb,
                                                 meta.ctypes.data_as(C.c_void_p), p64.ctypes.data_as(C.c_void_p)))
        m = dict(zip(META_FIELDS, meta.tolist()))
        m['root_value'] = float(meta[5:6].view(np.float32)[0])
        return blocks[:m['n_blocks']], m, p64

    def upload_tree(self, tree, blocks, meta, p64=None):
        blocks = np.ascontiguousarray(blocks, dtype=NODEBLOCK_DTYPE)
        mm = np.zeros(8, dtype=np.int32)
        for i, k in enumerate(META_FIELDS):
            if k == 'root_value':
                mm[i:i + 1] = np.array([meta.get(k, 0.0)], np.float32).view(np.int32)
            else:
                mm[i] = int(meta.get(k, 0))
        pp = None if p64 is None else np.ascontiguousarray(p64, dtype=np.float64)
        self._ck(self.lib.sgo_tree_upload_sync(self.h, tree, blocks.ctypes.data_as(C.c_void_p), len(blocks),
                                               mm.ctypes.data_as(C.c_void_p),
                                               None if pp is None else pp.ctypes.data_as(C.c_void_p)))

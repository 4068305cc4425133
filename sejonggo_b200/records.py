"""Game records as flat uint32 rows — the form in which they stay in HBM, cross NVLink and reach rank 0.

One row per ply = the reference's move_data (self_play.py:207-214) as sgo_records_pack writes it
(include/sejonggo_b200.h), prefixed with two tag words:

    [0] game id   [1] ply (move_n)   [2 .. 2+PW) packed board   [2+PW] move index   [3+PW] value (f32 bits)
    [4+PW] tree-valid flag   [5+PW .. 5+PW+A) policy_target = root priors as f32 (Q14)

and one FOOTER row per finished game (tag [1] = 0xFFFFFFFF) with the fields of game_data that are not per ply
(self_play.py:277-289): winner, number of plies, end reason, colours, the area score, the thresholds.  Rank 0 rebuilds
`game_data` dicts from the rows of any rank and writes them with sgfsave.save_self_play_data, replacing the
reference's zip-and-send of whole game directories (slave_coordinator.py:71-82).
"""
import numpy as np
import torch

FOOTER = 0xFFFFFFFF
END_REASONS = ["PLAYED ALL MOVES", "BOTH_PASSED", "resign"]
_WINNER_CODE = {None: 0, 0: 1, 1: 2}          # game_data['winner']: None draw, 0 white, 1 black
_WINNER_DECODE = {0: None, 1: 0, 2: 1}


def packed_words(size):
    return 16 * ((size * size + 31) // 32) + 1


def row_words(size):
    return 2 + packed_words(size) + 3 + size * size + 1


def _f32_bits(x):
    return int(np.array([np.nan if x is None else x], np.float32).view(np.uint32)[0])


def _bits_f32(w):
    v = float(np.array([w], np.uint32).view(np.float32)[0])
    return None if np.isnan(v) else v


def footer_row(size, game_id, game_data, black_points=None, white_points=None):
    row = np.zeros(row_words(size), np.uint32)
    row[0], row[1] = game_id, FOOTER
    row[2] = _WINNER_CODE[game_data['winner']]
    row[3] = len(game_data['moves'])
    row[4] = END_REASONS.index(game_data.get('end_reason', "PLAYED ALL MOVES"))
    row[5] = 1 if game_data.get('model1_isblack', True) else 0
    row[6], row[7] = _f32_bits(game_data.get('resign_model1')), _f32_bits(game_data.get('resign_model2'))
    # "B+12.5" / "W+R" / "D+0.0": the margin is |black - white| incl. komi; kept as the string's float
    res = game_data['result']
    row[8] = {"B": 2, "W": 0, "D": 1}[res[0]]
    row[9] = _f32_bits(None if res.endswith("+R") else float(res[2:]))
    return row


def rows_from_game_data(size, game_id, game_data, pack_board):
    """Host-side packer (tests, the host-evaluator path): the same rows the device store holds.
    pack_board(board[1,S,S,17]) -> uint32[PW] (Engine.export_packed's format)."""
    PW, A = packed_words(size), size * size + 1
    out = np.zeros((len(game_data['moves']) + 1, row_words(size)), np.uint32)
    for i, m in enumerate(game_data['moves']):
        r = out[i]
        r[0], r[1] = game_id, m['move_n']
        b = np.asarray(m['board'])
        r[2:2 + PW] = b.astype(np.uint32) if b.ndim == 1 else pack_board(b)
        x, y = m['move']
        r[2 + PW] = x + size * y
        r[3 + PW] = np.array([m['value']], np.float32).view(np.uint32)[0]
        r[4 + PW] = 1
        r[5 + PW:5 + PW + A] = np.asarray(m['policy'], np.float32).view(np.uint32)
    out[-1] = footer_row(size, game_id, game_data)
    return out


def games_from_rows(rows, size, names=("model_1", "model_1"), mode='a'):
    """uint32 [n][row_words] (any order, any mix of games) -> {game_id: game_data}; only games whose footer is
    present are returned (a game still in flight on its rank has rows but no footer yet)."""
    rows = np.asarray(rows, np.uint32).reshape(-1, row_words(size))
    PW, A = packed_words(size), size * size + 1
    plies, footers = {}, {}
    for r in rows:
        gid = int(r[0])
        if r[1] == FOOTER:
            footers[gid] = r
        else:
            plies.setdefault(gid, {})[int(r[1])] = r
    out = {}
    n1, n2 = names
    for gid, f in footers.items():
        got = plies.get(gid, {})
        n = int(f[3])
        if len(got) != n:
            raise ValueError("game %d: footer says %d plies, %d rows present" % (gid, n, len(got)))
        moves = []
        for k in range(n):
            r = got[k]
            idx = int(r[2 + PW])
            # move_data['player'] lags one ply (self_play.py:236): +1 at ply 0, then the mover of the previous ply
            player = 1 if k == 0 else (1 if (k - 1) % 2 == 0 else -1)
            moves.append(dict(board=r[2:2 + PW].copy(), policy=r[5 + PW:5 + PW + A].view(np.float32).copy(),
                              value=np.float32(r[3 + PW:4 + PW].view(np.float32)[0]), move=(idx % size, idx // size),
                              move_n=k, player=player))
        winner = _WINNER_DECODE[int(f[2])]
        isblack = bool(f[5])
        modelB, modelW = (n1, n2) if isblack else (n2, n1)
        if winner is None:
            winner_model = None
        elif mode == 'a':
            winner_model = n1 if (winner == 1) == isblack else n2
        else:
            winner_model = modelB if (winner == 1) == isblack else modelW
        margin = _bits_f32(int(f[9]))
        result = "%s+%s" % ({2: "B", 0: "W", 1: "D"}[int(f[8])], "R" if margin is None else margin)
        out[gid] = dict(moves=moves, modelB_name=modelB, modelW_name=modelW, winner=winner, winner_model=winner_model,
                        result=result, resign_model1=_bits_f32(int(f[6])), resign_model2=_bits_f32(int(f[7])),
                        end_reason=END_REASONS[int(f[4])], model1_isblack=isblack, game_id=gid)
    return out


class RecordStore(object):
    """Append-only row store on one device (HBM on a GPU rank).  The engine's per-ply record rows are copied in
    device-to-device; `take()` hands the filled part over (for the gather) and starts again."""

    def __init__(self, size, capacity_rows, device):
        self.size, self.RW = size, row_words(size)
        self.device = torch.device(device)
        self.buf = torch.zeros((capacity_rows, self.RW), dtype=torch.int32, device=self.device)
        self.n = 0

    def _room(self, k):
        if self.n + k > self.buf.shape[0]:
            grown = torch.zeros((max(2 * self.buf.shape[0], self.n + k), self.RW), dtype=torch.int32, device=self.device)
            grown[:self.n] = self.buf[:self.n]
            self.buf = grown

    def append_plies(self, rec_rows, slots, game_ids, move_n):
        """rec_rows: int32 [G][RW-2] device tensor from sgo_records_pack; slots: indices of the games that moved."""
        k = len(slots)
        if k == 0:
            return
        self._room(k)
        ix = torch.as_tensor(np.asarray(slots, np.int64), device=self.device)
        dst = self.buf[self.n:self.n + k]
        tags = np.stack([np.asarray(game_ids, np.int64), np.asarray(move_n, np.int64)], axis=1).astype(np.uint32).view(np.int32)
        dst[:, :2] = torch.as_tensor(tags, device=self.device)
        dst[:, 2:] = rec_rows.view(torch.int32)[ix]
        self.n += k

    def append_host_rows(self, rows):
        rows = np.ascontiguousarray(rows, np.uint32).reshape(-1, self.RW)
        self._room(len(rows))
        self.buf[self.n:self.n + len(rows)] = torch.as_tensor(rows.view(np.int32), device=self.device)
        self.n += len(rows)

    def take(self):
        out = self.buf[:self.n].clone()
        self.n = 0
        return out

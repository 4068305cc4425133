"""Game-record writer with the reference's layout (sgfsave.py:16-79):

    <SELF_PLAY_DIR>/<model>/game_%05d/move_%03d/sample.{h5|npz}
        board          float32 [1,S,S,17]   the position before the move
        policy_target  float32 [S*S+1]      the root children's PRIORS (Q14)
        value_target   float32 scalar       1 if winner == player else -1, where winner is 1/0/None and
                                            player is +1/-1, so white plies always get -1 (Q15, sgfsave.py:20,56)

h5py is not in this image, so samples are written as .npz with the same dataset names unless
h5py is importable; `npz_to_h5` converts a tree in place for train.py.  Boards recorded in the
engine's packed form (record_boards='packed') are expanded here.
"""
import os
import numpy as np

from .conf import conf

try:                                            # pragma: no cover - depends on the image
    import h5py
except Exception:
    h5py = None


def unpack_board(packed, size):
    """uint32 [16*ceil(S*S/32)+1] (sgo_export_packed) -> float32 [1,S,S,17]."""
    W = (size * size + 31) // 32
    p = np.asarray(packed, dtype=np.uint32)
    bits = ((p[:16 * W].reshape(16, W)[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1)
    planes = bits.reshape(16, W * 32)[:, :size * size].astype(np.float32)
    board = np.empty((1, size, size, 17), np.float32)
    board[0, :, :, :16] = planes.T.reshape(size, size, 16)
    board[0, :, :, 16] = np.float32(np.int32(p[16 * W]))
    return board


def sample_arrays(move_data, winner, size=None):
    board = move_data['board']
    if board is not None and np.asarray(board).ndim == 1:
        board = unpack_board(board, size or conf['SIZE'])
    value_target = 1 if winner == move_data['player'] else -1
    return dict(board=np.asarray(board, dtype=np.float32), policy_target=np.asarray(move_data['policy'], dtype=np.float32),
                value_target=np.array(value_target, dtype=np.float32))


def _write_sample(directory, arrays):
    if h5py is not None:
        with h5py.File(os.path.join(directory, 'sample.h5'), 'w') as f:
            for k, v in arrays.items():
                f.create_dataset(k, data=v, dtype=np.float32)
    else:
        np.savez(os.path.join(directory, 'sample.npz'), **arrays)


def _save(root, model_name, game_no, game_data, fmt, size):
    winner = game_data['winner']
    for move_data in game_data['moves']:
        move = move_data['move_n']
        directory = os.path.join(root, model_name, fmt % game_no, "move_%03d" % move)
        try:
            os.makedirs(directory)
        except OSError:
            while True:                          # the reference bumps the game number on a clash (sgfsave.py:61-69)
                game_no += 1
                directory = os.path.join(root, model_name, fmt % game_no, "move_%03d" % move)
                try:
                    os.makedirs(directory)
                    break
                except OSError:
                    pass
        _write_sample(directory, sample_arrays(move_data, winner, size))
    return game_no


def save_self_play_data(model_name, game_no, game_data, size=None):
    """sgfsave.py:49-79."""
    out = _save(conf['SELF_PLAY_DIR'], model_name, game_no, game_data, "game_%05d", size)
    if conf.get('SGF_ENABLED'):
        save_game_sgf(model_name, game_no, game_data, size=size)
    return out


def save_game_data(model_name, game_n, game_data, game_name="game", size=None):
    """sgfsave.py:40-46 (evaluation games, GAMES_DIR)."""
    out = _save(conf.get('GAMES_DIR', 'sp_eval_games'), model_name, game_n, game_data, game_name + "_%03d", size)
    if conf.get('SGF_ENABLED'):
        save_game_sgf(model_name, game_n, game_data, size=size)
    return out


def _sgf_escape(text):
    return str(text).replace("\\", "\\\\").replace("]", "\\]")


def real_board(board, size):
    """play.get_real_board (play.py:106-112): +1 black / -1 white / 0 empty from planes 0, 1 and the side to move."""
    b = np.asarray(board)
    if b.ndim == 1:
        b = unpack_board(b, size)
    own, opp, tm = b[0, :, :, 0], b[0, :, :, 1], int(b[0, 0, 0, 16])
    return ((own - opp) * tm).astype(np.int32)


def save_game_sgf(model_name, game_n, game_data, size=None):
    """sgfsave.py:130-167 without sgfmill: GAMES_DIR/<model>/game_%03d.sgf (the number is bumped while the file exists), an
    FF[4] game record with PB / PW / KM / RE and one node per ply: colour from move_data['player'] (which lags one ply in
    the reference, self_play.py:236 — reproduced), the move in SGF letters (column x, row y from the top; pass = empty),
    and the comment "Value <v>\n <board after the move>" taken, like the reference, from the NEXT ply's board (the last
    ply wraps around to the first)."""
    S = size or conf['SIZE']
    head = "(;FF[4]CA[UTF-8]GM[1]SZ[%d]PB[%s]PW[%s]KM[%s]RE[%s]" % (
        S, _sgf_escape(game_data['modelB_name']), _sgf_escape(game_data['modelW_name']), conf['KOMI'], _sgf_escape(game_data['result']))
    nodes = []
    moves = game_data['moves']
    for md in moves:
        color = 'B' if md['player'] == 1 else 'W'
        x, y = md['move']
        coord = "" if y == S else chr(97 + x) + chr(97 + y)
        nxt = moves[(md['move_n'] + 1) % len(moves)]['board']
        comment = "Value %s" % md['value']
        if nxt is not None:
            comment += "\n %s" % real_board(nxt, S)
        nodes.append(";%s[%s]C[%s]" % (color, coord, _sgf_escape(comment)))
    directory = os.path.join(conf.get('GAMES_DIR', 'sp_eval_games'), model_name)
    os.makedirs(directory, exist_ok=True)
    filename = os.path.join(directory, "game_%03d.sgf" % game_n)
    while os.path.isfile(filename):
        game_n += 1
        filename = os.path.join(directory, "game_%03d.sgf" % game_n)
    with open(filename, "w") as f:
        f.write(head + "".join(nodes) + ")\n")
    return filename


def npz_to_h5(root):
    """Convert every sample.npz under `root` to sample.h5 (needs h5py)."""
    if h5py is None:
        raise RuntimeError("h5py is not installed")
    n = 0
    for d, _, files in os.walk(root):
        if 'sample.npz' in files:
            z = np.load(os.path.join(d, 'sample.npz'))
            with h5py.File(os.path.join(d, 'sample.h5'), 'w') as f:
                for k in ('board', 'policy_target', 'value_target'):
                    f.create_dataset(k, data=z[k], dtype=np.float32)
            n += 1
    return n

"""TEST INFRASTRUCTURE ONLY — restatement of the reference's game drivers on top
of the C oracle primitives (oracle/go_oracle.c).

  mode A: self_play.py:123-290  (mcts_decision / select_play / play_game)
  mode B: nomodel_self_play.py:114-271 (select_play / play_game_async)

The reference draws from MT19937 in four places; here they are injected through
`rng` so the reference (fixtures), this oracle and the CUDA engine replay the
same decisions:
    rng.coin()            -> float      play.py:302  choose_first_player
    rng.dirichlet(A)      -> f64[A]     play.py:402
    rng.symmetry()        -> int 0..6   symmetry.py:128 (one per batch)
    rng.choice(moves, ps) -> int        self_play.py:149 / nomodel_self_play.py:136
"""
import numpy as np
from . import oracle as o


class ReplayRng(object):
    """Feeds recorded draws back (fixtures from oracle/gen_golden.py)."""

    def __init__(self, coin=(), noise=(), choice=(), symmetry=None):
        self._coin, self._noise, self._choice = list(coin), list(noise), list(choice)
        self._sym = None if symmetry is None else list(symmetry)

    def coin(self):
        return self._coin.pop(0)

    def dirichlet(self, n):
        return np.asarray(self._noise.pop(0), dtype=np.float64)

    def symmetry(self):
        return 0 if self._sym is None else self._sym.pop(0)

    def choice(self, moves, ps):
        return self._choice.pop(0)


class SeededRng(object):
    """Fresh draws (for oracle-vs-engine runs where both sides share one object
    type but separate instances with the same seed)."""

    def __init__(self, seed):
        self.r = np.random.RandomState(seed)

    def coin(self):
        return float(self.r.random_sample())

    def dirichlet(self, n, alpha=0.03):
        return self.r.dirichlet([alpha] * n)

    def symmetry(self):
        return int(self.r.randint(7))

    def choice(self, moves, ps):
        return int(moves[int(np.searchsorted(np.cumsum(ps), self.r.random_sample() * np.sum(ps)))
                         if len(moves) > 1 else 0])


def sym_predict(model, boards, sym):
    """symmetry.py:127-132 random_symmetry_predict with the symmetry id given."""
    boards = np.ascontiguousarray(boards, dtype=np.int32)
    S = boards.shape[-2]
    sb = o.sym_board(sym, boards)
    p, v = model.predict_on_batch(sb)
    return o.sym_policy(sym, np.asarray(p, dtype=np.float32), S), np.asarray(v, dtype=np.float32).reshape(-1)


def _pick(tree, temperature, rng):
    """self_play.py:138-152."""
    ch = tree.children()
    if temperature == 1:
        total = int(ch['counts'].sum())
        keep = ch['counts'] > 0
        moves = [int(m) for m in ch['moves'][keep]]
        ps = [int(c) / float(total) for c in ch['counts'][keep]]
        return rng.choice(moves, ps)
    return o.pick_t0(tree)


def _hook(rng, name, *args):
    """Optional context callbacks of a replaying rng (begin_root / begin_wave): parity runs that take
    their symmetry draws from a recorded engine run need to know which request a draw belongs to."""
    f = getattr(rng, name, None)
    if f is not None:
        f(*args)


def _game(models, size, stop_exploration, self_play, num_moves, resigns, rng, komi, root_eval, search, on_search=None):
    """Shared body of play_game (self_play.py:164-290) and play_game_async
    (nomodel_self_play.py:142-271).  on_search(move_n, tree, model) is called after the search of
    each ply, before the move pick (tests compare the whole tree there)."""
    S = size
    board, player = o.game_init(S)
    model1, model2 = models
    if rng.coin() < .5:                       # play.py:301-306
        current, other = model1, model2
    else:
        other, current = model1, model2
    tree, other_tree = None, None
    value = None
    model1_isblack = current is model1
    skipped_last = False
    temperature = 1
    end_reason = "PLAYED ALL MOVES"
    if num_moves is None:
        num_moves = S * S * 2
    moves = []
    for move_n in range(num_moves):
        if move_n == stop_exploration:
            temperature = 0
        _hook(rng, 'begin_root', move_n)
        policy, value = root_eval(current, board)
        resign = resigns[0] if current is model1 else resigns[1]
        if resign and value <= resign:
            end_reason = "resign"
            break
        if tree is None or tree.nchild == 0:
            noise = rng.dirichlet(S * S + 1) if self_play else None
            tree = o.new_tree(policy, board, noise=noise)
            if self_play:
                other_tree = tree
        search(current, board, tree)
        if on_search is not None:
            on_search(move_n, tree, current)
        index = _pick(tree, temperature, rng)
        x, y = (index % S, index // S)
        ch = tree.children()
        policy_target = np.zeros(S * S + 1)
        policy_target[ch['moves']] = ch['ps']
        moves.append(dict(board=np.copy(board), policy=policy_target, value=value, move=(x, y),
                          move_n=move_n, player=player))
        if skipped_last and y == S:
            end_reason = "BOTH_PASSED"
            break
        skipped_last = y == S
        if not self_play:
            if other_tree is not None and other_tree.child(index) is not None:
                other_tree = other_tree.child(index).detach()
            tree = tree.child(index).detach()
        else:
            tree = tree.child(index).detach()
            other_tree = tree
        # make_play returns the MOVER and the reference rebinds `player` to it
        # (self_play.py:236), so move_data['player'] lags one ply from ply 1 on.
        board, player = o.make_play(x, y, board)
        current, other = other, current
        tree, other_tree = other_tree, tree
    winner, black_points, white_points = o.get_winner(board, komi)
    ps = {1: "B", 0: "D", -1: "W"}
    if end_reason == "resign":
        result = "%s+R" % ps[player]
    else:
        result = "%s+%s" % (ps[winner], abs(black_points - white_points))
    if winner == 0:
        winner_model = None
    else:
        winner_model = model1 if (winner == 1) == model1_isblack else model2
    modelB, modelW = (model1, model2) if model1_isblack else (model2, model1)
    return dict(moves=moves, modelB=modelB, modelW=modelW, winner={1: 1, -1: 0, 0: None}[winner],
                winner_model=winner_model, result=result, resign_model1=resigns[0], resign_model2=resigns[1],
                end_reason=end_reason)


def play_game(model1, model2, mcts_simulations, stop_exploration, self_play=False, num_moves=None,
              resign_model1=None, resign_model2=None, size=19, mcts_batch_size=100, rng=None, komi=5.5, on_search=None):
    """self_play.py:164 (mode A)."""
    def root_eval(model, board):
        p, v = model.predict_on_batch(board)
        return np.asarray(p[0], dtype=np.float32), np.asarray(v[0], dtype=np.float32).reshape(-1)[0]

    def search(model, board, tree):
        op = int(board[0, 0, 0, 16])
        for _ in range(int(mcts_simulations / mcts_batch_size)):     # self_play.py:128
            sym = rng.symmetry()
            o.simulate(tree, np.copy(board), lambda b: sym_predict(model, b, sym), mcts_batch_size, op)

    gd = _game((model1, model2), size, stop_exploration, self_play, num_moves, (resign_model1, resign_model2),
               rng, komi, root_eval, search, on_search)
    gd['modelB_name'], gd['modelW_name'] = gd['modelB'].name, gd['modelW'].name
    gd['winner_model'] = None if gd['winner_model'] is None else gd['winner_model'].name
    return gd


def play_game_async(model1_indicator, model2_indicator, energy, stop_exploration, process_id, self_play=False,
                    num_moves=None, resign_model1=None, resign_model2=None, size=19, conf_sims=1600,
                    conf_energy=8, rng=None, komi=5.5, predict=None, names=None, on_search=None):
    """nomodel_self_play.py:142 (mode B).  `predict(indicator, boards, sym)` stands
    for put_predict_request; `*_SYM` indicators draw one symmetry per request."""
    class Tag(object):
        def __init__(self, t):
            self.t = t

    def is_sym(tag):
        return tag.t.endswith("_SYM")

    def one(tag, board):
        sym = rng.symmetry() if is_sym(tag) else 0
        p, v = predict(tag.t, board, sym)
        return np.asarray(p, dtype=np.float32).reshape(-1), np.float32(np.asarray(v).reshape(-1)[0])

    def search(tag, board, tree):
        op = int(board[0, 0, 0, 16])
        for w in range(int(conf_sims / conf_energy)):                 # nomodel_self_play.py:116
            _hook(rng, 'begin_wave', w)
            o.async_simulate2(tree, np.copy(board), lambda b: one(tag, b), energy, op, total_energy=conf_energy)

    t1 = Tag(model1_indicator)
    t2 = t1 if model2_indicator == model1_indicator else Tag(model2_indicator)
    gd = _game((t1, t2), size, stop_exploration, self_play, num_moves, (resign_model1, resign_model2),
               rng, komi, one, search, on_search)
    names = names or {}
    gd['modelB_name'] = names.get(gd['modelB'].t, gd['modelB'].t)
    gd['modelW_name'] = names.get(gd['modelW'].t, gd['modelW'].t)
    if gd['winner'] is None:
        gd['winner_model'] = None
    else:
        # nomodel_self_play.py:246-249 — note the reference's own expression
        model1_isblack = gd['modelB'] is t1
        gd['winner_model'] = gd['modelB_name'] if ((gd['winner'] == 1) == model1_isblack) else gd['modelW_name']
    return gd


def self_play(model, n_games, mcts_simulations, lottery, percent, allowed_error, stop_exploration, size=19,
              mcts_batch_size=100, rng=None, komi=5.5):
    """self_play.py:343-378 (and, with the directory handling stripped, model_self_play :293-340 and
    NoModelSelfPlayWorker.run selfplay_worker.py:76-112): n games one after the other with the resignation
    calibration.  lottery() stands for the `random()` of the per-game lottery; rng feeds play_game."""
    games_data = []
    current_resign = None
    min_values = []
    for game in range(n_games):
        resign = current_resign if lottery() > percent else None
        gd = play_game(model, model, mcts_simulations, stop_exploration, self_play=True, resign_model1=resign,
                       resign_model2=resign, size=size, mcts_batch_size=mcts_batch_size, rng=rng, komi=komi)
        if resign is None:
            mv = gd['moves'][::2] if gd['winner'] == 1 else gd['moves'][1::2]
            min_values.append(min(m['value'] for m in mv))
            idx = int(allowed_error * len(min_values))
            if idx > 0:
                current_resign = min_values[idx]
        games_data.append(gd)
    return games_data

"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the
UNMODIFIED reference (/root/reference) in the build container.

    python oracle/gen_golden.py            # regenerates every fixture

One subprocess per (board size, config) because the reference freezes
conf['SIZE'] / MCTS_BATCH_SIZE / ENERGY at import time (play.py:14,
self_play.py:22-23, nomodel_self_play.py:21).  The fixtures are what pins the C
oracle (oracle/go_oracle.c) and, through it, the CUDA engine.  Randomness the
reference draws from MT19937 (np.random.choice, np.random.dirichlet,
random.random, random.choice) is RECORDED here and INJECTED in the parity tests.
"""
import os
import sys
import json
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def _setup(size, overrides=None):
    sys.path.insert(0, ROOT)
    from oracle import ref_harness as rh
    ns = rh.load(size, overrides=overrides or {})
    return ns


def tree_rows(node, depth=0, rows=None):
    """Canonical pre-order serialisation of a reference dict tree:
    rows of (depth, move, count, value f32, mean f32, p f64, busy, expanded)."""
    import numpy as np
    if rows is None:
        rows = []
    for m, c in node['subtree'].items():
        rows.append((depth, m, c['count'], float(np.float32(c['value'])), float(np.float32(c['mean_value'])),
                     float(c['p']), c.get('virtual_loss', 0), 1 if c['subtree'] else 0))
        tree_rows(c, depth + 1, rows)
    return rows


def pack_rows(rows):
    import numpy as np
    a = np.zeros(len(rows), dtype=[('depth', 'i4'), ('move', 'i4'), ('count', 'i8'), ('value', 'f4'),
                                   ('mean', 'f4'), ('p', 'f8'), ('busy', 'i4'), ('expanded', 'i4')])
    for i, r in enumerate(rows):
        a[i] = r
    return a


# ------------------------------------------------------------------- rules
def gen_rules(size, ngames, seed):
    import numpy as np, random
    ns = _setup(size)
    from oracle import oracle as o          # only for the packing helper
    play = ns.play
    rng = random.Random(seed)
    S, A = size, size * size + 1
    all_moves, all_states, all_masks, results = [], [], [], []
    for g in range(ngames):
        board, _ = play.game_init()
        moves, states, masks = [], [], []
        passes = 0
        for t in range(2 * S * S):
            m = play.legal_moves(board)
            states.append(o.pack_board(np.ascontiguousarray(board, dtype=np.int32)))
            masks.append(m.astype(np.uint8))
            empties = [i for i in range(S * S) if board[0, i // S, i % S, 0] == 0 and board[0, i // S, i % S, 1] == 0]
            legal = [i for i in range(S * S) if m[i] == 0]
            r = rng.random()
            if r < 0.02 or not legal:
                mv = S * S
            elif r < 0.08 and empties:
                mv = rng.choice(empties)      # includes suicides / ko retakes: make_play executes them (Q5)
            else:
                mv = rng.choice(legal)
            x, y = (0, S) if mv == S * S else (mv % S, mv // S)
            play.make_play(x, y, board)
            moves.append(mv)
            passes = passes + 1 if mv == S * S else 0
            if passes == 2:
                break
        states.append(o.pack_board(np.ascontiguousarray(board, dtype=np.int32)))
        masks.append(play.legal_moves(board).astype(np.uint8))
        w, b, wh = play.get_winner(board)
        all_moves.append(np.array(moves, np.int32))
        all_states.append(np.array(states, np.uint32))
        all_masks.append(np.array(masks, np.uint8))
        results.append((w, float(b), float(wh)))
    offs = np.cumsum([0] + [len(m) for m in all_moves]).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "rules_s%d.npz" % size), size=size, komi=5.5,
                        move_offsets=offs, moves=np.concatenate(all_moves),
                        states=np.concatenate(all_states), masks=np.concatenate(all_masks),
                        results=np.array(results, np.float64))


def gen_rules_hashed(size, ngames, seed, tag):
    """Many long games from the reference (19x19: most run into the 2*S*S = 722-ply cap): the move lists, the final
    results, and a 64-bit hash of the packed state and of the legality mask at EVERY ply (fake_eval.hash_rows), so
    that the fixture stays small.  Move choice: uniform over the legal points, 1% passes, 3% "any empty point"
    (suicides and ko retakes, which make_play executes, Q5)."""
    import numpy as np, random
    ns = _setup(size)
    from oracle import oracle as o
    from oracle.fake_eval import hash_rows
    play = ns.play
    rng = random.Random(seed)
    S, A = size, size * size + 1
    all_moves, sh, mh, results = [], [], [], []
    for g in range(ngames):
        board, _ = play.game_init()
        moves, states, masks = [], [], []
        passes = 0
        pass_rate = 0.0 if g % 2 == 0 else 0.01       # even games never pass voluntarily: they run to the ply cap
        for t in range(2 * S * S):
            m = play.legal_moves(board)
            states.append(o.pack_board(np.ascontiguousarray(board, dtype=np.int32)))
            masks.append(m.astype(np.uint8))
            occ = board[0, :, :, 0] + board[0, :, :, 1]
            legal = np.nonzero(m[:S * S] == 0)[0]
            r = rng.random()
            if r < pass_rate or len(legal) == 0:
                mv = S * S
            elif r < pass_rate + 0.03:
                empties = np.nonzero(occ.reshape(-1) == 0)[0]
                mv = int(empties[rng.randrange(len(empties))])
            else:
                mv = int(legal[rng.randrange(len(legal))])
            x, y = (0, S) if mv == S * S else (mv % S, mv // S)
            play.make_play(x, y, board)
            moves.append(mv)
            passes = passes + 1 if mv == S * S else 0
            if passes == 2:
                break
        states.append(o.pack_board(np.ascontiguousarray(board, dtype=np.int32)))
        masks.append(play.legal_moves(board).astype(np.uint8))
        w, b, wh = play.get_winner(board)
        all_moves.append(np.array(moves, np.int32))
        sh.append(hash_rows(np.array(states, np.uint32)))
        mh.append(hash_rows(np.array(masks, np.uint8)))
        results.append((w, float(b), float(wh)))
        print("  game", g, len(moves), results[-1], flush=True)
    offs = np.cumsum([0] + [len(m) for m in all_moves]).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "ruleshash_%s.npz" % tag), size=size, komi=5.5, move_offsets=offs,
                        moves=np.concatenate(all_moves).astype(np.int16), state_hash=np.concatenate(sh), mask_hash=np.concatenate(mh),
                        results=np.array(results, np.float64))


# ---------------------------------------------------------------- symmetry
def gen_symmetry(size):
    import numpy as np
    ns = _setup(size)
    sym = ns.symmetry
    S, A = size, size * size + 1
    fwd = [sym._id, sym.left_diagonal, sym.vertical_axis, sym.horizontal_axis,
           sym.rotation_90, sym.rotation_180, sym.rotation_270, sym.right_diagonal]
    rev = [sym._id, sym.reverse_left_diagonal, sym.reverse_vertical_axis, sym.reverse_horizontal_axis,
           sym.reverse_rotation_90, sym.reverse_rotation_180, sym.reverse_rotation_270, sym.reverse_right_diagonal]
    btab = np.zeros((8, S * S), np.int32)
    ptab = np.zeros((8, A), np.int32)
    for k in range(8):
        idx = np.arange(S * S, dtype=np.int32).reshape(1, S, S, 1).repeat(17, axis=3)
        btab[k] = np.array(fwd[k](np.copy(idx)))[0, :, :, 0].reshape(-1)      # out[i] = in[btab[i]]
        pol = np.arange(A, dtype=np.float32).reshape(1, A)
        ptab[k] = np.array(rev[k](np.copy(pol)))[0].astype(np.int32)
    order = [f.__name__ for f, _ in sym.SYMMETRIES]
    np.savez_compressed(os.path.join(OUT, "symmetry_s%d.npz" % size), size=size, board_src=btab, policy_src=ptab,
                        symmetries=np.array(order))


# -------------------------------------------------------------------- mcts
def gen_mcts(size, mode, batch, steps, plies, seed, tag):
    """Search trace: a noised root, `steps` simulate/wave calls per ply, the full
    tree after every ply, then a prescribed pick + re-root (self_play.py:223-238)."""
    import numpy as np
    ns = _setup(size, overrides={'ENERGY': batch, 'MCTS_BATCH_SIZE': batch})
    from oracle.fake_eval import evaluate, FakeModel
    ns.symmetry.SYMMETRIES = ns.symmetry.SYMMETRIES[0:1]
    play = ns.play
    S, A = size, size * size + 1
    model = FakeModel(sharp=True, salt=seed)
    ns.evaluator = lambda ind, board: (lambda p, v: (p[0], v[0]))(*evaluate(board, seed, True))
    np.random.seed(seed)
    noise = np.random.dirichlet([0.03] * A)
    play.np.random.dirichlet = lambda a: noise
    board, _ = play.game_init()
    pol, _ = model.predict_on_batch(board)
    tree = play.new_tree(pol[0], board, add_noise=True)
    trees, picks, t0picks, newtree = [], [], [], []
    for ply in range(plies):
        op = board[0, 0, 0, -1]
        for i in range(steps):
            if mode == 'a':
                ns.self_play.simulate(tree, np.copy(board), model, batch, op)
            else:
                ns.nomodel_self_play.async_simulate2(tree, np.copy(board), "BEST", batch, op, 0)
        trees.append(pack_rows(tree_rows(tree)))
        _, _, t0 = max((d['count'], d['mean_value'], a) for a, d in tree['subtree'].items())
        t0picks.append(t0)
        sel = t0
        if ply % 3 == 1:                      # a recorded "temperature 1" style pick
            vis = [a for a, d in tree['subtree'].items() if d['count'] > 0]
            sel = vis[(ply * 7) % len(vis)]
        picks.append(sel)
        tree = tree['subtree'][sel]
        tree['parent'] = None
        x, y = (0, S) if sel == S * S else (sel % S, sel // S)
        play.make_play(x, y, board)
        nt = 0
        if not tree['subtree']:
            pol, _ = model.predict_on_batch(board)
            tree = play.new_tree(pol[0], board)
            nt = 1
        newtree.append(nt)
    offs = np.cumsum([0] + [len(t) for t in trees]).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "mcts_%s.npz" % tag), size=size, mode=mode, batch=batch, steps=steps,
                        seed=seed, noise=noise, tree_offsets=offs, trees=np.concatenate(trees),
                        picks=np.array(picks, np.int32), t0picks=np.array(t0picks, np.int32),
                        newtree=np.array(newtree, np.int32))


# --------------------------------------------------------------- game loop
def gen_game(size, mode, batch, sims, stop_exploration, self_play, num_moves, seed, tag, resign=None, evalkind="fake",
             live_sym=False):
    """Full play_game / play_game_async run with recorded RNG draws.  live_sym: all 7 SYMMETRIES stay in play and
    every random.choice of symmetry.random_symmetry_predict (symmetry.py:127-132) is recorded; in mode B the stub of
    put_predict_request then does what the predicting worker does for a *_SYM tag (predicting_queue_worker.py:88-92)
    on a COPY of the board — the real request crosses a process boundary, so the flip transforms' in-place writes
    (Q9) never reach the game process."""
    import numpy as np, random
    ns = _setup(size, overrides={'ENERGY': batch, 'MCTS_BATCH_SIZE': batch, 'MCTS_SIMULATIONS': sims})
    from oracle.fake_eval import evaluate_kind, FakeModel
    from oracle import oracle as o
    sym_rec = []
    if live_sym:
        all_syms = list(ns.symmetry.SYMMETRIES)
        assert len(all_syms) == 7

        def sym_choice(seq):
            k = random.randrange(len(seq))
            sym_rec.append(k)
            return seq[k]

        ns.symmetry.choice = sym_choice
    else:
        ns.symmetry.SYMMETRIES = ns.symmetry.SYMMETRIES[0:1]
    play = ns.play
    S, A = size, size * size + 1
    random.seed(seed)
    np.random.seed(seed)
    rec = dict(choice=[], noise=[], coin=[])
    real_choice = np.random.choice
    real_dir = np.random.dirichlet

    def choice(moves, size=1, p=None):
        r = real_choice(moves, size=size, p=p)
        rec['choice'].append(int(r[0]))
        return r

    def dirichlet(alpha):
        r = real_dir(alpha)
        rec['noise'].append(np.array(r))
        return r

    import play as _play_mod
    real_random = _play_mod.random

    def coin():
        r = real_random()
        rec['coin'].append(r)
        return r

    _play_mod.random = coin
    _play_mod.np.random.dirichlet = dirichlet
    m1 = FakeModel("model_1", salt=seed, sharp=True, kind=evalkind)
    m2 = FakeModel("model_2", salt=seed + 1, sharp=True, kind=evalkind)
    if mode == 'a':
        ns.self_play.np.random.choice = choice
        if self_play:
            m2 = m1
        gd = ns.self_play.play_game(m1, m2, sims, stop_exploration, self_play=self_play, num_moves=num_moves,
                                    resign_model1=resign, resign_model2=resign)
        calls = np.array(m1.calls + [-1] + (m2.calls if m2 is not m1 else []), np.int64)
    else:
        ns.nomodel_self_play.np.random.choice = choice
        salts = {"BEST_SYM": seed, "LATEST_SYM": seed + 1, "BEST": seed, "LATEST": seed + 1}
        ns.names = {"BEST_SYM": "model_1", "LATEST_SYM": "model_2", "BEST": "model_1", "LATEST": "model_2"}
        ns.evaluator = lambda ind, board: (lambda p, v: (p[0], v[0]))(*evaluate_kind(board, salts[ind], True, evalkind))
        if live_sym:
            # Q21: LATEST_SYM is served by the BEST network
            served = {"BEST_SYM": FakeModel("model_1", salt=seed, sharp=True, kind=evalkind),
                      "LATEST_SYM": FakeModel("model_1", salt=seed, sharp=True, kind=evalkind)}

            def sym_eval(ind, board):
                p, v = ns.symmetry.random_symmetry_predict(served[ind], np.copy(board))
                return p[0], v[0][0]

            ns.evaluator = sym_eval
        i1, i2 = ("BEST_SYM", "BEST_SYM") if self_play else ("BEST_SYM", "LATEST_SYM")
        gd = ns.nomodel_self_play.play_game_async(i1, i2, batch, stop_exploration, 0, self_play=self_play,
                                                  num_moves=num_moves, resign_model1=resign, resign_model2=resign)
        calls = np.zeros(0, np.int64)
    mv = gd['moves']
    np.savez_compressed(
        os.path.join(OUT, "game_%s.npz" % tag), size=size, mode=mode, batch=batch, sims=sims,
        stop_exploration=stop_exploration, self_play=int(self_play), num_moves=-1 if num_moves is None else num_moves,
        seed=seed, resign=np.nan if resign is None else resign, evalkind=evalkind,
        live_sym=int(live_sym), symmetry=np.array(sym_rec, np.int32),
        choice=np.array(rec['choice'], np.int32), noise=np.array(rec['noise'], np.float64).reshape(-1, A),
        coin=np.array(rec['coin'], np.float64),
        boards=np.array([o.pack_board(np.ascontiguousarray(m['board'], dtype=np.int32)) for m in mv], np.uint32),
        policy=np.array([m['policy'] for m in mv], np.float64),
        value=np.array([np.float32(m['value']) for m in mv], np.float32).reshape(-1),
        move=np.array([m['move'] for m in mv], np.int32), move_n=np.array([m['move_n'] for m in mv], np.int32),
        player=np.array([m['player'] for m in mv], np.int32),
        modelB_name=gd['modelB_name'], modelW_name=gd['modelW_name'],
        winner=-1 if gd['winner'] is None else gd['winner'], winner_model=str(gd['winner_model']),
        result=gd['result'], calls=calls)


# ------------------------------------------------- self_play with the resignation calibration
def gen_selfplay(size, batch, sims, n_games, seed, tag, percent, allowed_error, stop_exploration, komi=0.5):
    """self_play.self_play(model, n_games, sims) (self_play.py:343-378) with every random draw recorded: the
    lottery `random() > RESIGNATION_PERCENT`, the colour coins, the Dirichlet noise and the temperature-1 choices.
    Saved per game: the threshold it played with, its moves / values / result."""
    import numpy as np, random
    ns = _setup(size, overrides={'KOMI': komi, 'ENERGY': batch, 'MCTS_BATCH_SIZE': batch, 'MCTS_SIMULATIONS': sims,
                                 'RESIGNATION_PERCENT': percent, 'RESIGNATION_ALLOWED_ERROR': allowed_error,
                                 'STOP_EXPLORATION': stop_exploration})
    from oracle.fake_eval import FakeModel
    ns.symmetry.SYMMETRIES = ns.symmetry.SYMMETRIES[0:1]
    S, A = size, size * size + 1
    random.seed(seed)
    np.random.seed(seed)
    rec = dict(choice=[], noise=[], coin=[], lottery=[])
    real_choice, real_dir = np.random.choice, np.random.dirichlet

    def choice(moves, size=1, p=None):
        r = real_choice(moves, size=size, p=p)
        rec['choice'].append(int(r[0]))
        return r

    def dirichlet(alpha):
        r = real_dir(alpha)
        rec['noise'].append(np.array(r))
        return r

    import play as _play_mod
    real_random = _play_mod.random

    def coin():
        r = real_random()
        rec['coin'].append(r)
        return r

    def lottery():
        r = real_random()
        rec['lottery'].append(r)
        return r

    _play_mod.random = coin
    _play_mod.np.random.dirichlet = dirichlet
    ns.self_play.random = lottery
    ns.self_play.np.random.choice = choice
    ns.self_play.tqdm.tqdm = lambda it, desc=None: type("T", (), {"__iter__": lambda s: iter(it), "set_description": lambda s, d: None})()
    model = FakeModel("model_1", salt=seed, sharp=True)
    games = ns.self_play.self_play(model, n_games, sims)
    offs = np.cumsum([0] + [len(g['moves']) for g in games]).astype(np.int64)
    mv = [m for g in games for m in g['moves']]
    np.savez_compressed(
        os.path.join(OUT, "selfplay_%s.npz" % tag), size=size, batch=batch, sims=sims, n_games=n_games, seed=seed, komi=komi,
        percent=percent, allowed_error=allowed_error, stop_exploration=stop_exploration,
        lottery=np.array(rec['lottery'], np.float64), coin=np.array(rec['coin'], np.float64),
        noise=np.array(rec['noise'], np.float64).reshape(-1, A), choice=np.array(rec['choice'], np.int32),
        resign=np.array([np.nan if g['resign_model1'] is None else float(np.asarray(g['resign_model1']).reshape(-1)[0]) for g in games], np.float64),
        move_offsets=offs, move=np.array([m['move'] for m in mv], np.int32).reshape(-1, 2),
        value=np.array([np.float32(np.asarray(m['value']).reshape(-1)[0]) for m in mv], np.float32).reshape(-1),
        result=np.array([g['result'] for g in games]), winner=np.array([-1 if g['winner'] is None else g['winner'] for g in games], np.int32))
    print("  thresholds", [g['resign_model1'] for g in games], [g['result'] for g in games])


JOBS = [
    ("rules", dict(size=9, ngames=12, seed=1)),
    ("rules", dict(size=19, ngames=3, seed=2)),
    ("rules", dict(size=5, ngames=20, seed=3)),
    ("symmetry", dict(size=9)),
    ("symmetry", dict(size=19)),
    ("mcts", dict(size=9, mode='a', batch=8, steps=8, plies=10, seed=3, tag="a_s9")),
    ("mcts", dict(size=9, mode='b', batch=8, steps=8, plies=10, seed=4, tag="b_s9")),
    ("mcts", dict(size=5, mode='a', batch=100, steps=2, plies=30, seed=5, tag="a_s5_wide")),
    ("mcts", dict(size=3, mode='b', batch=8, steps=4, plies=18, seed=6, tag="b_s3_stall")),
    ("mcts", dict(size=19, mode='a', batch=100, steps=2, plies=2, seed=7, tag="a_s19")),
    ("mcts", dict(size=19, mode='b', batch=8, steps=12, plies=2, seed=8, tag="b_s19")),
    ("game", dict(size=9, mode='a', batch=8, sims=32, stop_exploration=6, self_play=True, num_moves=14, seed=11, tag="a_selfplay_s9")),
    ("game", dict(size=9, mode='a', batch=8, sims=32, stop_exploration=0, self_play=False, num_moves=10, seed=12, tag="a_eval_s9")),
    ("game", dict(size=5, mode='a', batch=8, sims=16, stop_exploration=4, self_play=True, num_moves=None, seed=13, tag="a_selfplay_s5_full")),
    ("game", dict(size=9, mode='b', batch=8, sims=32, stop_exploration=6, self_play=True, num_moves=14, seed=14, tag="b_selfplay_s9")),
    ("game", dict(size=9, mode='b', batch=8, sims=32, stop_exploration=0, self_play=False, num_moves=10, seed=15, tag="b_eval_s9")),
    ("game", dict(size=9, mode='a', batch=8, sims=16, stop_exploration=3, self_play=True, num_moves=None, seed=16, tag="a_resign_s9", resign=0.5)),
    # BASELINE.json configs[0] / SURVEY §8d config 1: 9x9, 64 sims/ply, uniform evaluator, seed 0, 30 plies, modes A and B
    ("game", dict(size=9, mode='a', batch=8, sims=64, stop_exploration=30, self_play=True, num_moves=30, seed=0, tag="a_config1_s9", evalkind="uniform")),
    ("game", dict(size=9, mode='b', batch=8, sims=64, stop_exploration=30, self_play=True, num_moves=30, seed=0, tag="b_config1_s9", evalkind="uniform")),
    # round 2: all 7 symmetries live (Q7/Q8 end to end from the reference), the resignation calibration, many long 19x19 games
    ("game", dict(size=9, mode='a', batch=8, sims=32, stop_exploration=5, self_play=True, num_moves=12, seed=21, tag="a_livesym_s9", live_sym=True)),
    ("game", dict(size=9, mode='b', batch=8, sims=32, stop_exploration=5, self_play=True, num_moves=12, seed=22, tag="b_livesym_s9", live_sym=True)),
    ("game", dict(size=7, mode='a', batch=8, sims=32, stop_exploration=0, self_play=False, num_moves=10, seed=23, tag="a_livesym_eval_s7", live_sym=True)),
    ("game", dict(size=7, mode='b', batch=8, sims=32, stop_exploration=0, self_play=False, num_moves=10, seed=24, tag="b_livesym_eval_s7", live_sym=True)),
    ("selfplay", dict(size=5, batch=4, sims=8, n_games=16, seed=31, tag="calib_s5", percent=0.5, allowed_error=0.34, stop_exploration=3)),
    ("ruleshash", dict(size=19, ngames=64, seed=41, tag="s19_64")),
]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        kind, kw = sys.argv[1], json.loads(sys.argv[2])
        {"rules": gen_rules, "symmetry": gen_symmetry, "mcts": gen_mcts, "game": gen_game, "selfplay": gen_selfplay,
         "ruleshash": gen_rules_hashed}[kind](**kw)
    else:
        os.makedirs(OUT, exist_ok=True)
        only = os.environ.get("GEN_ONLY")               # e.g. GEN_ONLY=livesym,calib,s19_64 regenerates the matching fixtures only
        for kind, kw in JOBS:
            if only and not any(t in str(kw.get("tag", "")) for t in only.split(",")):
                continue
            print(kind, kw, flush=True)
            subprocess.check_call([sys.executable, os.path.abspath(__file__), kind, json.dumps(kw)], cwd="/tmp")

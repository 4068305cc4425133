"""Batched stand-in for the reference's EvaluateWorker / NoModelEvaluateWorker
(evaluate_worker.py:26-162): ONE worker per GPU plays every still-unclaimed evaluation game
concurrently.  Kept from the reference: a game is claimed by creating
EVAL_DIR/<latest>/game_%03d (:66-73, :125-131; several workers / GPUs can share the directory),
the result is recorded by touching a file named after the winner model in it (:41-43), mode B
games are also saved as training data ("eval_game", :150-151), and `promote_best_model` reads
those files (evaluator.py:50-82)."""
import os
from pathlib import Path

from .conf import conf
from . import predicting_queue_worker as pq
from .self_play import play_games
from .nomodel_self_play import play_games_async
from .sgfsave import save_game_data


def save_eval_game(model_name, game_no, winner_model):
    path = os.path.join(conf['EVAL_DIR'], model_name, "game_%03d" % game_no, str(winner_model))
    Path(path).touch()


def _claim(latest_name, n_games, limit):
    claimed = []
    for game in range(n_games):
        if len(claimed) >= limit:
            break
        directory = os.path.join(conf['EVAL_DIR'], latest_name, "game_%03d" % game)
        if os.path.isdir(directory):
            continue
        try:
            os.makedirs(directory)
        except OSError:
            continue
        claimed.append(game)
    return claimed


def run_evaluation(best_model, latest_model, n_games=None, concurrent=None, mode='a', save_training_data=None,
                   reference_q21=False, **kw):
    """Plays the unclaimed games of best vs latest; returns (wins of latest, games played).
    mode 'a' = EvaluateWorker (self_play.play_game), 'b' = NoModelEvaluateWorker (play_game_async
    with the BEST_SYM / LATEST_SYM tags).  In the reference the LATEST_SYM tag is served by the BEST network (Q21),
    so its mode-B evaluation never plays the candidate; here LATEST_SYM goes to the latest model unless
    reference_q21=True asks for the reference's behaviour."""
    n_games = n_games or conf['EVALUATE_N_GAMES']
    concurrent = concurrent or conf['CONCURRENT_GAMES']
    if save_training_data is None:
        save_training_data = mode == 'b'
    if latest_model.name == best_model.name:
        return 0, 0                                            # "No new trained model" (:96)
    wins = total = 0
    while True:
        claimed = _claim(latest_model.name, n_games, concurrent)
        if not claimed:
            break
        if mode == 'a':
            games = play_games(best_model, latest_model, len(claimed), conf['MCTS_SIMULATIONS'], stop_exploration=0, **kw)
        else:
            pq.register_models(best=best_model, latest=latest_model)
            q21, pq.REFERENCE_Q21 = pq.REFERENCE_Q21, bool(reference_q21)
            try:
                games = play_games_async("BEST_SYM", "LATEST_SYM", len(claimed), conf['ENERGY'], 0, **kw)
            finally:
                pq.REFERENCE_Q21 = q21
        for game, gd in zip(claimed, games):
            winner_model = gd['winner_model']
            wins += winner_model == latest_model.name
            total += 1
            if save_training_data:
                save_game_data(latest_model.name, game, gd, game_name="eval_game", size=kw.get('size'))
            save_eval_game(latest_model.name, game, winner_model)
    return wins, total

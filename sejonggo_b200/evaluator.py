"""Mirror of the reference's evaluator.evaluate (evaluator.py:23-47): EVALUATE_N_GAMES
games of play_game(best, tested, MCTS_SIMULATIONS, stop_exploration=0), all played
concurrently on the engine; promotion bookkeeping (evaluator.py:50-85) is out of scope."""
from .conf import conf
from .self_play import play_games


def evaluate_games(best_model, tested_model, n_games=None, mcts_simulations=None, **kw):
    return play_games(best_model, tested_model, n_games or conf['EVALUATE_N_GAMES'],
                      mcts_simulations or conf['MCTS_SIMULATIONS'], stop_exploration=0, **kw)


def evaluate(best_model, tested_model, **kw):
    games = evaluate_games(best_model, tested_model, **kw)
    wins = sum(1 for g in games if g['winner_model'] == tested_model.name)
    return wins / float(len(games)) > conf['EVALUATE_MARGIN']

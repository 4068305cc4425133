"""Mirror of the reference's evaluator.py: `evaluate` (evaluator.py:23-47) plays EVALUATE_N_GAMES
games of play_game(best, tested, MCTS_SIMULATIONS, stop_exploration=0) — here all concurrently on
the engine — and the win-file bookkeeping around match play: `eval_statistic` (:50-64),
`promote_best_model` (:67-82), `clean_up_result` (:84-86).  scp sync (`sync_model`) is out of
scope; `on_promote` is the hook for it."""
import os
import shutil

from .conf import conf
from .self_play import play_games
from .sgfsave import save_game_data


def evaluate_games(best_model, tested_model, n_games=None, mcts_simulations=None, **kw):
    return play_games(best_model, tested_model, n_games or conf['EVALUATE_N_GAMES'],
                      mcts_simulations or conf['MCTS_SIMULATIONS'], stop_exploration=0, **kw)


def elect_model_as_best_model(model, self_play_games=0, **kw):
    """evaluator.py:17-20: (optionally) seed the new best model's self-play data, then save it as BEST_MODEL."""
    if self_play_games:
        from .self_play import self_play
        self_play(model, self_play_games, conf['MCTS_SIMULATIONS'], **kw)
    os.makedirs(conf['MODEL_DIR'], exist_ok=True)
    model.save(os.path.join(conf['MODEL_DIR'], conf['BEST_MODEL']))


def evaluate(best_model, tested_model, save_games=False, elect=False, **kw):
    games = evaluate_games(best_model, tested_model, **kw)
    wins = sum(1 for g in games if g['winner_model'] == tested_model.name)
    if save_games:
        for n, gd in enumerate(games):
            save_game_data(best_model.name, n, gd, size=kw.get('size'))
    if wins / float(len(games)) > conf['EVALUATE_MARGIN']:
        if elect:
            elect_model_as_best_model(tested_model)
        return True
    return False


def eval_statistic():
    """evaluator.py:50-64: per tested model, the share of EVAL_DIR/<model>/game_*/ directories holding a
    file named after the model (EvaluateWorker.save_eval_game touches <winner_model> in each)."""
    result = {}
    root = conf['EVAL_DIR']
    if not os.path.isdir(root):
        return result
    for model_name in os.listdir(root):
        model_dir = os.path.join(root, model_name)
        if not os.path.isdir(model_dir):
            continue
        wins = total = 0
        for game_dir in os.listdir(model_dir):
            if game_dir.startswith('game'):
                total += 1
                if os.path.isfile(os.path.join(model_dir, game_dir, model_name)):
                    wins += 1
        result[model_name] = wins / total if total != 0 else 0
    return result


def clean_up_result(result):
    for model_name in result.keys():
        shutil.rmtree(os.path.join(conf['EVAL_DIR'], model_name))


def promote_best_model(cleanup=True, on_promote=None):
    """evaluator.py:67-82: the first model whose win rate beats EVALUATE_MARGIN becomes BEST_MODEL."""
    result = eval_statistic()
    for model_name in result.keys():
        if result[model_name] > conf['EVALUATE_MARGIN']:
            ext = os.path.splitext(conf['BEST_MODEL'])[1]
            shutil.copyfile(os.path.join(conf['MODEL_DIR'], model_name + ext), os.path.join(conf['MODEL_DIR'], conf['BEST_MODEL']))
            if on_promote is not None:
                on_promote(model_name)                # the reference scp-syncs the new best model here
            if cleanup:
                clean_up_result(result)
            return True
    return False

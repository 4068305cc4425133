"""Build-container only (needs /root/reference): the conf keys the hot path reads keep the reference's defaults,
and the mirrored entry points keep the reference's positional signatures (SURVEY §8b game API)."""
import ast
import inspect
import os

import pytest

REF = os.environ.get("SEJONGGO_REFERENCE", "/root/reference")
pytestmark = pytest.mark.ref


def _ref_conf():
    tree = ast.parse(open(os.path.join(REF, "conf.py")).read())
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "conf":
            out = {}
            for k, v in zip(node.value.keys, node.value.values):
                try:
                    out[ast.literal_eval(k)] = ast.literal_eval(v)
                except Exception:
                    pass
            return out
    raise AssertionError("conf dict not found")


def _ref_signature(module, func):
    tree = ast.parse(open(os.path.join(REF, module)).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func:
            args = [a.arg for a in node.args.args]
            defaults = [ast.literal_eval(d) for d in node.args.defaults]
            return args, defaults
    raise AssertionError("%s.%s not found" % (module, func))


def test_conf_defaults_match_reference():
    from sejonggo_b200.conf import conf
    ref = _ref_conf()
    shared = ['N_RESIDUAL_BLOCKS', 'N_GAMES', 'MCTS_SIMULATIONS', 'ENERGY', 'SIZE', 'KOMI', 'STOP_EXPLORATION', 'MCTS_BATCH_SIZE',
              'DIRICHLET_ALPHA', 'DIRICHLET_EPSILON', 'RESIGNATION_PERCENT', 'RESIGNATION_ALLOWED_ERROR', 'EVALUATE_N_GAMES',
              'EVALUATE_MARGIN', 'MODEL_DIR', 'EVAL_DIR', 'GAMES_DIR']
    for k in shared:
        assert k in ref, k
        assert conf[k] == ref[k], (k, conf[k], ref[k])


@pytest.mark.parametrize("module,func,ours", [
    ("self_play.py", "play_game", "sejonggo_b200.self_play"),
    ("nomodel_self_play.py", "play_game_async", "sejonggo_b200.nomodel_self_play"),
    ("evaluator.py", "evaluate", "sejonggo_b200.evaluator"),
    ("play.py", "make_play", "sejonggo_b200.play"),
    ("play.py", "legal_moves", "sejonggo_b200.play"),
    ("play.py", "get_winner", "sejonggo_b200.play"),
    ("predicting_queue_worker.py", "put_predict_request", "sejonggo_b200.predicting_queue_worker"),
    ("predicting_queue_worker.py", "put_name_request", "sejonggo_b200.predicting_queue_worker"),
])
def test_signatures_match_reference(module, func, ours):
    import importlib
    args, defaults = _ref_signature(module, func)
    sig = inspect.signature(getattr(importlib.import_module(ours), func))
    params = [p for p in sig.parameters.values() if p.kind in (p.POSITIONAL_OR_KEYWORD,)]
    names = [p.name for p in params]
    assert names[:len(args)] == args, (names, args)                       # same positional order (extras may follow)
    ours_defaults = [p.default for p in params[:len(args)] if p.default is not inspect._empty]
    assert ours_defaults == defaults, (ours_defaults, defaults)

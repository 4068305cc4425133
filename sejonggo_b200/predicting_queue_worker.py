"""Client stubs of the reference's predicting_queue_worker.py:109-124 without the
queue daemon: the GPU engine batches natively, so a "request" is a direct call into
the registered models.  Tags: BEST | LATEST | BEST_SYM | LATEST_SYM (| *_NAME);
as in the reference LATEST_SYM is served by the BEST model (Q21, :91-92)."""
import numpy as np

from .symmetry import random_symmetry_predict

_models = {}


def register_models(best=None, latest=None):
    if best is not None:
        _models['BEST'] = best
    if latest is not None:
        _models['LATEST'] = latest


def model_for(indicator):
    base = 'BEST' if indicator in ('BEST', 'BEST_SYM', 'LATEST_SYM', 'BEST_NAME') else 'LATEST'
    if base not in _models:
        raise KeyError("no model registered for %s (call register_models)" % indicator)
    return _models[base]


def put_name_request(model_indicator):
    if model_indicator in ('BEST_SYM', 'BEST', 'BEST_NAME'):
        return _models['BEST'].name
    return _models['LATEST'].name          # LATEST / LATEST_SYM / LATEST_NAME (:110-113)


def put_predict_request(model_indicator, board, response_now=False):
    model = model_for(model_indicator)
    if model_indicator.endswith('_SYM'):
        p, v = random_symmetry_predict(model, np.asarray(board))
    else:
        p, v = model.predict_on_batch(np.asarray(board))
    return p[0], np.asarray(v).reshape(-1)[0]

"""Client stubs of the reference's predicting_queue_worker.py:109-124 without the
queue daemon: the GPU engine batches natively, so a "request" is a direct call into
the registered models.  Tags: BEST | LATEST | BEST_SYM | LATEST_SYM (| *_NAME);
as in the reference LATEST_SYM is served by the BEST model (Q21, :91-92)."""
import numpy as np

from .symmetry import random_symmetry_predict

_models = {}
# Q21: in the reference the LATEST_SYM tag is evaluated by the BEST network (predicting_queue_worker.py:91-92) while its
# *_NAME reports the latest model — NoModelEvaluateWorker therefore plays best against best and credits `latest`.
# True reproduces that (the fixtures recorded from the reference need it); evaluate_worker.run_evaluation switches it
# off for its own games unless told otherwise, so that the candidate network really is the one evaluated.
REFERENCE_Q21 = True


def register_models(best=None, latest=None):
    if best is not None:
        _models['BEST'] = best
    if latest is not None:
        _models['LATEST'] = latest


def model_for(indicator):
    best_tags = ('BEST', 'BEST_SYM', 'BEST_NAME') + (('LATEST_SYM',) if REFERENCE_Q21 else ())
    base = 'BEST' if indicator in best_tags else 'LATEST'
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

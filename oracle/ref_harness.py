"""TEST INFRASTRUCTURE ONLY — never imported by the product path (sejonggo_b200/).

Loads the *unmodified* reference (drsagitn/sejonggo, mounted read-only at
/root/reference) in this container so that golden fixtures can be generated
from the reference itself.  /root/reference does not exist on
the GPU box, so nothing under tests/ marked `gpu`, bench.py or smoke() calls
this module; only oracle/gen_golden.py and the container-only pin tests do.

Stubs (exactly the two SURVEY §8c lists):
  * `sgfsave`                  — real one needs sgfmill + h5py (absent)
  * `predicting_queue_worker`  — real one imports model.py -> TensorFlow 1.7
The reference freezes conf['SIZE'] at import (play.py:14), so one process can
hold one board size: call `load(size)` once per interpreter.
"""
import os
import sys
import types
import pickle
import tempfile
import collections

REF_ROOT = os.environ.get("SEJONGGO_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "play.py"))


class _SyncPool(object):
    """Synchronous stand-in for multiprocessing.Pool (simulation_workers.py:26-28).

    apply_async pickle-round-trips its arguments like the real Pool does, so
    make_play in the worker mutates a private copy of the board; tasks run at
    issue time, so results queue up in issue order."""

    def apply_async(self, fn, args=(), kwds=None, callback=None, error_callback=None):
        args = pickle.loads(pickle.dumps(args))
        res = fn(*args, **(kwds or {}))
        if callback is not None:
            callback(res)
        return types.SimpleNamespace(wait=lambda: None, get=lambda: res)

    def map(self, fn, items):
        return [fn(pickle.loads(pickle.dumps(i))) for i in items]

    def close(self):
        pass

    def join(self):
        pass


class _Fifo(object):
    """In-process replacement for multiprocessing.SimpleQueue; pickles on put
    like the pipe does so returned leaves are fresh objects."""

    def __init__(self):
        self.q = collections.deque()

    def put(self, item):
        self.q.append(pickle.dumps(item))

    def get(self):
        return pickle.loads(self.q.popleft())

    def empty(self):
        return not self.q


def load(size, komi=5.5, overrides=None):
    """Import the reference modules for one board size; returns a namespace."""
    assert available(), "reference not mounted at %s" % REF_ROOT
    assert "play" not in sys.modules, "reference already imported in this process"
    scratch = tempfile.mkdtemp(prefix="sgo_ref_")
    os.chdir(scratch)  # app_log.setup_logging() looks for logconfig.json in CWD
    sys.path.insert(0, REF_ROOT)
    from conf import conf
    conf["SIZE"] = size
    conf["KOMI"] = komi
    conf["THREAD_SIMULATION"] = False
    conf["SHOW_EACH_MOVE"] = False
    conf["SHOW_END_GAME"] = False
    conf["N_GAME_PROCESS"] = 1
    for k, v in (overrides or {}).items():
        conf[k] = v

    sgf = types.ModuleType("sgfsave")
    sgf.save_self_play_data = lambda *a, **k: None
    sgf.save_game_data = lambda *a, **k: None
    sgf.save_game_sgf = lambda *a, **k: None
    sys.modules["sgfsave"] = sgf

    ns = types.SimpleNamespace(conf=conf, evaluator=None, names={})
    pq = types.ModuleType("predicting_queue_worker")

    def put_predict_request(model_indicator, board, response_now=False):
        return ns.evaluator(model_indicator, board)

    def put_name_request(model_indicator):
        return ns.names.get(model_indicator, str(model_indicator))

    pq.put_predict_request = put_predict_request
    pq.put_name_request = put_name_request
    sys.modules["predicting_queue_worker"] = pq

    import play, symmetry, tree_util, self_play, simulation_workers, nomodel_self_play
    simulation_workers.process_pool = _SyncPool()
    simulation_workers.simulation_result_queue = {0: _Fifo()}
    ns.play, ns.symmetry, ns.tree_util = play, symmetry, tree_util
    ns.self_play, ns.simulation_workers = self_play, simulation_workers
    ns.nomodel_self_play = nomodel_self_play
    return ns

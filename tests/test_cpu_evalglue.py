"""Host logic around match play: weight files in MODEL_DIR, win-file statistics and
promotion (evaluator.py:50-86) — file bookkeeping only, no GPU."""
import os
import numpy as np
import pytest
import torch

from sejonggo_b200 import evaluator, evaluate_worker, model
from sejonggo_b200.conf import conf


@pytest.fixture
def dirs(tmp_path):
    old = dict(conf)
    conf.update(MODEL_DIR=str(tmp_path / "models"), EVAL_DIR=str(tmp_path / "eval"), GAMES_DIR=str(tmp_path / "eval"),
                SELF_PLAY_DIR=str(tmp_path / "sp"), SIZE=9, N_RESIDUAL_BLOCKS=1)
    os.makedirs(conf['MODEL_DIR'])
    yield tmp_path
    conf.clear()
    conf.update(old)


def test_params_roundtrip_and_model_dir_protocol(dirs):
    m3 = model.TowerModel("model_3", params=model.init_params(9, 1, seed=3, randomize_bn=True, random_bias=True))
    m10 = model.TowerModel("model_10", params=model.init_params(9, 1, seed=10))
    m3.save(os.path.join(conf['MODEL_DIR'], "model_3.npz"))
    m10.save(os.path.join(conf['MODEL_DIR'], "model_10.npz"))
    open(os.path.join(conf['MODEL_DIR'], "notes.txt"), "w").close()          # non-model files are skipped (model.py:126-133)
    latest = model.load_latest_model()
    assert latest.name == "model_10"                                           # numeric, not lexicographic, order
    back = model.load_model_by_name("model_3.npz")
    assert back.name == "model_3" and back.params['meta'] == m3.params['meta']
    for k, v in m3.params.items():
        if k == 'meta':
            continue
        if isinstance(v, dict):
            for kk in v:
                assert torch.equal(v[kk], back.params[k][kk]), (k, kk)
        else:
            assert torch.equal(v, back.params[k]), k
    # no best model yet -> model_1 is created and saved under both names (model.py:144-155)
    best = model.load_best_model()
    assert best.name == "model_1"
    assert os.path.isfile(os.path.join(conf['MODEL_DIR'], conf['BEST_MODEL']))
    assert model.load_best_model().name == "model_1"


def _touch_results(name, winners):
    for g, w in enumerate(winners):
        d = os.path.join(conf['EVAL_DIR'], name, "game_%03d" % g)
        os.makedirs(d)
        evaluate_worker.save_eval_game(name, g, w)


def test_eval_statistic_and_promotion(dirs):
    assert evaluator.eval_statistic() == {}
    assert evaluator.promote_best_model() is False
    model.TowerModel("model_2", params=model.init_params(9, 1, seed=2)).save(os.path.join(conf['MODEL_DIR'], "model_2.npz"))
    model.TowerModel("model_1", params=model.init_params(9, 1, seed=1)).save(os.path.join(conf['MODEL_DIR'], conf['BEST_MODEL']))
    _touch_results("model_2", ["model_2"] * 5 + ["model_1"] * 5)             # 50% <= margin 55%
    assert evaluator.eval_statistic() == {"model_2": 0.5}
    assert evaluator.promote_best_model() is False
    assert model.load_best_model().name == "model_1"
    _touch_results_more = ["model_2"] * 3
    for g, w in enumerate(_touch_results_more, start=10):
        d = os.path.join(conf['EVAL_DIR'], "model_2", "game_%03d" % g)
        os.makedirs(d)
        evaluate_worker.save_eval_game("model_2", g, w)
    assert abs(evaluator.eval_statistic()["model_2"] - 8 / 13) < 1e-12
    promoted = []
    assert evaluator.promote_best_model(cleanup=True, on_promote=promoted.append) is True
    assert promoted == ["model_2"]
    assert model.load_best_model().name == "model_2"
    assert not os.path.exists(os.path.join(conf['EVAL_DIR'], "model_2"))     # clean_up_result


def test_claiming_skips_existing_game_dirs(dirs):
    os.makedirs(os.path.join(conf['EVAL_DIR'], "model_2", "game_001"))
    assert evaluate_worker._claim("model_2", 5, 3) == [0, 2, 3]
    assert evaluate_worker._claim("model_2", 5, 10) == [4]
    assert evaluate_worker._claim("model_2", 5, 10) == []

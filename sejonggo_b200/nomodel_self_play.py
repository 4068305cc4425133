"""Mirror of the reference's nomodel_self_play.py (mode B: virtual-loss waves).
play_game_async keeps the reference signature (nomodel_self_play.py:142); model
"indicators" are resolved through predicting_queue_worker.register_models."""
from .conf import conf
from .batched import BatchedGames
from . import predicting_queue_worker as pq


class _Tagged(object):
    """A model seen through an indicator tag (name as put_name_request reports it)."""

    def __init__(self, indicator):
        self.indicator = indicator
        self.model = pq.model_for(indicator)
        self.name = pq.put_name_request(indicator)

    def predict_on_batch(self, X):
        return self.model.predict_on_batch(X)


def play_games_async(model1_indicator, model2_indicator, n_games, energy, stop_exploration, self_play=False,
                     num_moves=None, resign_model1=None, resign_model2=None, size=None, rng=None, rngs=None,
                     record_boards='full', arena_blocks=None, device=0, exact_rng_order=True, concurrent=None,
                     on_game_start=None, on_game_end=None, rng_for_game=None):
    m1 = _Tagged(model1_indicator)
    m2 = m1 if model2_indicator == model1_indicator else _Tagged(model2_indicator)
    names = (m1.name, m2.name)            # put_name_request: LATEST_SYM reports the latest model's name (:110-113)
    # an evaluator object can be shared when both tags resolve to one network
    if hasattr(m1.model, "is_sgo_evaluator"):
        m1 = m1.model
        m2 = m1 if model2_indicator == model1_indicator else pq.model_for(model2_indicator)
    sym = model1_indicator.endswith('_SYM')
    bg = BatchedGames((m1, m2), min(n_games, concurrent or n_games), size=size or conf['SIZE'], mode='b', energy=energy,
                      mcts_simulations=conf['MCTS_SIMULATIONS'],          # Q17: the argument is ignored (:116)
                      stop_exploration=stop_exploration, self_play=self_play, num_moves=num_moves,
                      resign=(resign_model1, resign_model2), komi=conf['KOMI'], dirichlet_eps=conf['DIRICHLET_EPSILON'],
                      use_symmetry=sym, rng=rng, rngs=rngs, arena_blocks=arena_blocks or conf['ARENA_BLOCKS'],
                      device=device, record_boards=record_boards, names=names, n_total=n_games,
                      on_game_start=on_game_start, on_game_end=on_game_end, rng_for_game=rng_for_game)
    bg.energy = conf['ENERGY']          # wave count and final back-props read conf (:116, :80)
    return bg.run(exact_rng_order=exact_rng_order)


def play_game_async(model1_indicator, model2_indicator, energy, stop_exploration, process_id, self_play=False,
                    num_moves=None, resign_model1=None, resign_model2=None, **kw):
    return play_games_async(model1_indicator, model2_indicator, 1, energy, stop_exploration, self_play,
                            num_moves, resign_model1, resign_model2, **kw)[0]

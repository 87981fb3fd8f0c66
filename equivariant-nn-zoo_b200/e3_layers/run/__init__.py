"""Score-based diffusion on top of the B200 model path (reference ``e3_layers/run/sde_utils.py`` and
``sde_sampling.py``): the VP-SDE, the score function wrapper, the score-matching loss and the
predictor-corrector sampler.  SURVEY 8f rank 2."""
from .sde_sampling import (EulerMaruyamaPredictor, LangevinCorrector, NoneCorrector, NonePredictor, get_corrector,
                           get_pc_sampler, get_predictor)
from .sde_utils import VPSDE, ExponentialMovingAverage, get_score_fn, get_sde_loss_fn, get_step_fn

__all__ = ["VPSDE", "ExponentialMovingAverage", "get_score_fn", "get_sde_loss_fn", "get_step_fn", "get_pc_sampler",
           "get_predictor", "get_corrector", "EulerMaruyamaPredictor", "NonePredictor", "LangevinCorrector", "NoneCorrector"]

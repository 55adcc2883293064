"""blackjax_b200 -- B200-native HMC/NUTS hot path behind BlackJAX's API.

``import blackjax_b200 as blackjax`` gives ``blackjax.hmc``, ``blackjax.nuts`` and
``blackjax.window_adaptation`` with the reference's ``init / build_kernel / __call__ ->
SamplingAlgorithm(init, step)`` surface (blackjax/__init__.py:70-80,104-112), natively batched over
chains and executed by hand-written sm_100a CUDA kernels through the C ABI in ``include/bjx.h``.
"""
import functools as _functools

from . import diagnostics, random, targets, util  # noqa: F401
from ._lib import BjxError  # noqa: F401
from .adaptation.window_adaptation import build_schedule, window_adaptation  # noqa: F401
from .adaptation.chees_adaptation import chees_adaptation  # noqa: F401
from .adaptation.meads_adaptation import meads_adaptation  # noqa: F401
from .base import AdaptationAlgorithm, AdaptationResults, GenerateSamplingAPI, SamplingAlgorithm  # noqa: F401
from . import mcmc  # noqa: F401
from .mcmc import hmc as _hmc
from .mcmc import nuts as _nuts
from .mcmc import dynamic_hmc as _dynamic_hmc
from .mcmc import ghmc as _ghmc
from .util import run_inference_algorithm, sample_hmc_native, sample_nuts_native  # noqa: F401

hmc = GenerateSamplingAPI(_hmc.as_top_level_api, _hmc.init, _hmc.build_kernel)     # blackjax/__init__.py:111
nuts = GenerateSamplingAPI(_nuts.as_top_level_api, _nuts.init, _nuts.build_kernel)  # blackjax/__init__.py:112
mhmc = GenerateSamplingAPI(                                                          # blackjax/__init__.py:145-151
    _functools.partial(_hmc.as_top_level_api, build_proposal=_hmc.multinomial_hmc_proposal),
    _hmc.init,
    _functools.partial(_hmc.build_kernel, build_proposal=_hmc.multinomial_hmc_proposal),
)
multinomial_hmc = mhmc  # backward-compatible alias (:152)
dhmc = GenerateSamplingAPI(_dynamic_hmc.as_top_level_api, _dynamic_hmc.init, _dynamic_hmc.build_kernel)  # :117
dynamic_hmc = dhmc  # backward-compatible alias (:118)
dmhmc = GenerateSamplingAPI(                                                         # blackjax/__init__.py:154-162
    _functools.partial(_dynamic_hmc.as_top_level_api, build_proposal=_hmc.multinomial_hmc_proposal),
    _dynamic_hmc.init,
    _functools.partial(_dynamic_hmc.build_kernel, build_proposal=_hmc.multinomial_hmc_proposal),
)

ghmc = GenerateSamplingAPI(_ghmc.as_top_level_api, _ghmc.init, _ghmc.build_kernel)   # blackjax/__init__.py:139

__version__ = "0.1.0"

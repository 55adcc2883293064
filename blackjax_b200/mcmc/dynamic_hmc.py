"""Dynamic HMC: HMC whose number of integration steps is drawn anew for every transition -- and, batched, for every
chain (blackjax/mcmc/dynamic_hmc.py).  The transition itself is the fused HMC kernel with a per-chain trajectory
length (``bjx_set_integration_steps``); the default step-count draw ``jax.random.randint(key, (), 1, 10)`` runs in
``bjx_prng_randint``."""
from typing import Callable, NamedTuple

import torch

from .. import random as jr
from . import hmc
from .hmc import HMCInfo, HMCState, hmc_proposal, multinomial_hmc_proposal
from .integrators import velocity_verlet

__all__ = ["DynamicHMCState", "init", "build_kernel", "as_top_level_api", "hmc_proposal"]


class DynamicHMCState(NamedTuple):
    """blackjax/mcmc/dynamic_hmc.py:40-51; ``random_generator_arg`` holds one key per chain, uint32 [n_chains, 2]."""
    position: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor
    random_generator_arg: torch.Tensor


def init(position, logdensity_fn, random_generator_arg):
    """blackjax/mcmc/dynamic_hmc.py:54-60."""
    st = hmc.init(position, logdensity_fn)
    return DynamicHMCState(st.position, st.logdensity, st.logdensity_grad, random_generator_arg)


def _next_random_arg(keys):
    return jr.split(keys)[..., 1, :]            # jax.random.split(key)[1]   dynamic_hmc.py:65


def _integration_steps(keys, *params):
    return jr.randint(keys, (), 1, 10)          # jax.random.randint(key, (), 1, 10)   dynamic_hmc.py:66


def build_kernel(integrator=velocity_verlet, divergence_threshold: float = 1000,
                 next_random_arg_fn: Callable = _next_random_arg, integration_steps_fn: Callable = _integration_steps,
                 build_proposal=hmc_proposal, **kw):
    """blackjax/mcmc/dynamic_hmc.py:62-130.  ``integration_steps_fn(random_generator_arg, *params)`` returns an int32
    tensor [n_chains] (or a Python int); ``next_random_arg_fn`` maps the [n_chains, 2] keys to the next ones."""
    hmc_base = hmc.build_kernel(integrator, divergence_threshold, build_proposal, **kw)

    def kernel(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, integration_steps_params=()):
        num_integration_steps = integration_steps_fn(state.random_generator_arg, *integration_steps_params)
        hmc_state = HMCState(state.position, state.logdensity, state.logdensity_grad)
        new, info = hmc_base(rng_key, hmc_state, logdensity_fn, step_size, inverse_mass_matrix, num_integration_steps)
        next_random_arg = next_random_arg_fn(state.random_generator_arg)
        return DynamicHMCState(new.position, new.logdensity, new.logdensity_grad, next_random_arg), info

    return kernel


def as_top_level_api(logdensity_fn, step_size, inverse_mass_matrix, *, divergence_threshold: int = 1000,
                     integrator=velocity_verlet, next_random_arg_fn: Callable = _next_random_arg,
                     integration_steps_fn: Callable = _integration_steps, integration_steps_params=(),
                     build_proposal=hmc_proposal, **kw):
    """blackjax/mcmc/dynamic_hmc.py:133-214."""
    from ..base import SamplingAlgorithm
    kernel = build_kernel(integrator, divergence_threshold, next_random_arg_fn, integration_steps_fn, build_proposal, **kw)

    def init_fn(position, rng_key=None):
        # the per-chain keys of the step-count sequence: [n_chains, 2], or one key that is split over the chains
        arg = rng_key
        if arg is not None and arg.ndim == 1:
            arg = jr.split(arg, position.shape[0])
        return init(position, logdensity_fn, arg)

    def step_fn(rng_key, state):
        return kernel(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, integration_steps_params)

    return SamplingAlgorithm(init_fn, step_fn)

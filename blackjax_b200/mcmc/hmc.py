"""HMC behind BlackJAX's ``init / build_kernel / as_top_level_api`` surface (blackjax/mcmc/hmc.py).

Intentional surface difference: the chain axis is explicit -- ``position`` is a
``[n_chains, dim]`` CUDA tensor and ``rng_key`` is either one raw key ``uint32[2]`` (split into one key
per chain, the reference's step-major pattern ``vmap(kernel)(split(key, C), states)``,
docs/examples/howto_sample_multiple_chains.md:116-129) or per-chain keys ``uint32[C, 2]``.
``logdensity_fn`` is a :mod:`blackjax_b200.targets` descriptor.
"""
from typing import NamedTuple, Optional

import torch

from .. import random as bjx_random
from .._engine import get_engine
from ..base import build_sampling_algorithm

__all__ = ["HMCState", "HMCInfo", "init", "build_kernel", "as_top_level_api", "hmc_proposal",
           "multinomial_hmc_proposal"]

# proposal selectors for build_kernel(build_proposal=...): the endpoint proposal of hmc.py:115-178 (default) and the
# multinomial proposal of hmc.py:181-248; each maps to one fused transition kernel
hmc_proposal = "hmc_proposal"
multinomial_hmc_proposal = "multinomial_hmc_proposal"

from . import integrators
from .integrators import velocity_verlet


class HMCState(NamedTuple):
    """blackjax/mcmc/hmc.py:38-49, batched: position [C,D], logdensity [C], logdensity_grad [C,D]."""

    position: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor


class IntegratorState(NamedTuple):
    """blackjax/mcmc/integrators.py:43-53."""

    position: torch.Tensor
    momentum: torch.Tensor
    logdensity: Optional[torch.Tensor]
    logdensity_grad: Optional[torch.Tensor]


class HMCInfo(NamedTuple):
    """blackjax/mcmc/hmc.py:52-87.  D-sized fields (momentum, proposal) are materialised only with
    ``build_kernel(..., full_info=True)``; otherwise they are None."""

    momentum: Optional[torch.Tensor]
    acceptance_rate: torch.Tensor
    is_accepted: torch.Tensor
    is_divergent: torch.Tensor
    energy: torch.Tensor
    proposal: Optional[IntegratorState]
    num_integration_steps: int


def init(position, logdensity_fn):
    """blackjax/mcmc/hmc.py:90-92."""
    eng = get_engine(position, logdensity_fn)
    position = position.contiguous()
    logp, grad = eng.init_state(position)
    return HMCState(position, logp, grad)


def build_kernel(integrator=velocity_verlet, divergence_threshold: float = 1000, build_proposal=None,
                 full_info: bool = False, inplace: bool = False, chain_offset: int = 0):
    """blackjax/mcmc/hmc.py:251-314.  ``inplace=True`` overwrites the input state's tensors (no
    allocation; the returned state aliases the input).  ``chain_offset``: global index of this process's first chain
    when one shared ``rng_key`` is split across GPUs."""
    coefficients = integrators.as_coefficients(integrator)
    if build_proposal not in (None, hmc_proposal, multinomial_hmc_proposal):
        raise NotImplementedError("build_proposal must be hmc.hmc_proposal or hmc.multinomial_hmc_proposal "
                                  "(each is one fused transition kernel)")
    multinomial = build_proposal == multinomial_hmc_proposal

    def kernel(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, num_integration_steps):
        q, logp, g = state
        eng = get_engine(q, logdensity_fn, divergence_threshold=divergence_threshold)
        eng.set_integrator(coefficients)
        eng.ensure_metric(inverse_mass_matrix)
        keys = rng_key  # one key [2]: chain c uses split(rng_key, n_global)[chain_offset + c], derived in-kernel
        C, dev = eng.C, eng.device
        fields = dict(acceptance_rate=torch.empty(C, dtype=torch.float32, device=dev),
                      is_accepted=torch.empty(C, dtype=torch.uint8, device=dev),
                      is_divergent=torch.empty(C, dtype=torch.uint8, device=dev),
                      energy=torch.empty(C, dtype=torch.float32, device=dev))
        if full_info:
            fields.update(momentum=torch.empty_like(q), proposal_position=torch.empty_like(q),
                          proposal_momentum=torch.empty_like(q))
        out = (q, logp, g) if inplace else None
        qo, lo, go = eng.hmc_step(keys, q, logp, g, step_size, num_integration_steps, out=out, info_fields=fields,
                                  multinomial=multinomial, chain_offset=chain_offset)
        proposal = None
        if full_info:
            proposal = IntegratorState(fields["proposal_position"], fields["proposal_momentum"], None, None)
        info = HMCInfo(fields.get("momentum"), fields["acceptance_rate"], fields["is_accepted"].bool(),
                       fields["is_divergent"].bool(), fields["energy"], proposal, num_integration_steps)
        return HMCState(qo, lo, go), info

    return kernel


def as_top_level_api(logdensity_fn, step_size, inverse_mass_matrix, num_integration_steps, *,
                     divergence_threshold: int = 1000, integrator=velocity_verlet, **kw):
    """blackjax/mcmc/hmc.py:317-414."""
    kernel = build_kernel(integrator, divergence_threshold, **kw)
    return build_sampling_algorithm(kernel, init, logdensity_fn,
                                    kernel_args=(step_size, inverse_mass_matrix, num_integration_steps))

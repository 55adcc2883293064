"""Generalized HMC behind BlackJAX's ``init / build_kernel / as_top_level_api`` surface (blackjax/mcmc/ghmc.py).

Persistent momentum (partial refresh with weight ``alpha``), ONE velocity-Verlet step per transition and a
non-reversible slice acceptance (translation ``delta``): one fused kernel launch per transition (``bjx_ghmc_step``).
Batched like :mod:`blackjax_b200.mcmc.hmc`: ``position`` is ``[n_chains, dim]``, ``rng_key`` one raw key or one per chain.
``momentum_inverse_scale`` in its 1-D form is an inverse SCALE, squared into the inverse mass matrix (ghmc.py:64-84);
a 2-D array or a ``LowRankMetric`` is an inverse mass matrix as in ``hmc``.  ``noise_fn`` maps the per-chain noise keys to per-chain values (see ``build_kernel``).
"""
import ctypes as C
from typing import NamedTuple

import torch

from .. import random as bjx_random
from .._engine import as_keys, get_engine
from .._lib import check, lib, ptr
from ..base import build_sampling_algorithm
from .hmc import HMCInfo, IntegratorState
from .integrators import velocity_verlet
from . import integrators

__all__ = ["GHMCState", "init", "build_kernel", "as_top_level_api", "update_momentum"]


class GHMCState(NamedTuple):
    """blackjax/mcmc/ghmc.py:30-47, batched: position, momentum, logdensity_grad [C,D]; logdensity, slice [C]."""

    position: torch.Tensor
    momentum: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor
    slice: torch.Tensor


def init(position, logdensity_fn, rng_key, chain_offset: int = 0):
    """blackjax/mcmc/ghmc.py:50-62: momentum ~ N(0, I) and slice ~ U(-1, 1) from ``split(rng_key)`` per chain."""
    eng = get_engine(position, logdensity_fn)
    position = position.contiguous()
    logp, grad = eng.init_state(position)
    keys = as_keys(rng_key, eng.C, eng.device)
    if keys.ndim == 1:
        n_global = max(eng.C + int(chain_offset), 1)
        keys = bjx_random.split(keys, n_global)[int(chain_offset):int(chain_offset) + eng.C]
    ks = bjx_random.split(keys, 2).view(torch.int32)
    momentum = bjx_random.normal(ks[:, 0].contiguous(), (eng.D,))
    u = bjx_random.uniform(ks[:, 1].contiguous())
    sl = torch.maximum(torch.full_like(u, -1.0), u * 2.0 + (-1.0))      # jax.random.uniform(minval=-1, maxval=1)
    return GHMCState(position, momentum, logp, grad, sl)


def _metric_from_momentum_inverse_scale(x):
    """ghmc.py:64-84: rich metrics pass through, the 1-D / scalar form is squared."""
    from .metrics import LowRankMetric
    if isinstance(x, LowRankMetric) or (isinstance(x, torch.Tensor) and x.ndim >= 2):
        return x
    return x * x


def _param(v, eng, name):
    """scalar -> (float, None); tensor [C] -> (0.0, device array)"""
    if isinstance(v, torch.Tensor) and v.ndim >= 1:
        t = v.to(device=eng.device, dtype=torch.float32).contiguous()
        if t.shape != (eng.C,):
            raise ValueError(f"per-chain {name} must have shape ({eng.C},)")
        return 0.0, t
    return float(v), None


def build_kernel(noise_fn=None, divergence_threshold: float = 1000, integrator=velocity_verlet, full_info: bool = False,
                 inplace: bool = False, chain_offset: int = 0):
    """blackjax/mcmc/ghmc.py:87-189.  ``noise_fn`` (ghmc.py:90,172): a callable mapping the per-chain noise keys
    ``key_noise = split(rng_key)[1]`` (uint32 [C, 2] on the device) to the per-chain noise values float32 [C] added to the
    slice translation, e.g. ``lambda k: 0.1 * blackjax.random.normal(k)``; None = the reference default (0)."""
    coefficients = integrators.as_coefficients(integrator)

    def kernel(rng_key, state, logdensity_fn, step_size, momentum_inverse_scale, alpha, delta, *, _rows=None):
        q, p, logp, g, sl = state
        eng = get_engine(q, logdensity_fn, divergence_threshold=divergence_threshold)
        eng.set_integrator(coefficients)
        if _rows is None:
            eng.ensure_metric(_metric_from_momentum_inverse_scale(momentum_inverse_scale))
        keys = as_keys(rng_key, eng.C, eng.device)
        eng._key_mode(keys, chain_offset)
        if not inplace:
            q, p, logp, g, sl = (t.clone() for t in (q, p, logp, g, sl))
        Cn, dev = eng.C, eng.device
        fields = dict(acceptance_rate=torch.empty(Cn, dtype=torch.float32, device=dev),
                      is_accepted=torch.empty(Cn, dtype=torch.uint8, device=dev),
                      is_divergent=torch.empty(Cn, dtype=torch.uint8, device=dev),
                      energy=torch.empty(Cn, dtype=torch.float32, device=dev))
        if full_info:
            fields.update(momentum=torch.empty_like(q), proposal_position=torch.empty_like(q),
                          proposal_momentum=torch.empty_like(q))
        info = eng._info(fields)
        noise = None
        if noise_fn is not None:
            # the chain keys of this transition (given, or split(step_key, C_global)[offset + c]), then ghmc.py:169
            chain_keys = keys
            if keys.ndim == 1:
                chain_keys = bjx_random.split(keys, Cn + int(chain_offset))[int(chain_offset):int(chain_offset) + Cn]
            noise = noise_fn(bjx_random.split(chain_keys.contiguous(), 2).view(torch.int32)[:, 1].contiguous()
                             .view(torch.uint32))
            noise = torch.as_tensor(noise, dtype=torch.float32, device=dev).expand(Cn).contiguous()
        check(lib().bjx_set_ghmc_noise(eng.h, ptr(noise)), eng.h)
        eng._ghmc_noise_keepalive = noise   # the launch is asynchronous
        if _rows is not None:   # MEADS: per-fold device parameters (step_size, alpha, delta: [K]; imm, msqrt: [K, D])
            eps_d, a_d, d_d, imm_rows, msqrt_rows, group, skip = _rows
            check(lib().bjx_ghmc_step(eng.h, ptr(keys), ptr(q), ptr(p), ptr(logp), ptr(g), ptr(sl), 0.0, ptr(eps_d), 0.0,
                                      ptr(a_d), 0.0, ptr(d_d), ptr(imm_rows), ptr(msqrt_rows), int(group), int(skip[0]),
                                      int(skip[1]), C.byref(info)), eng.h)
        else:
            eps, eps_d = eng._eps(step_size)
            a, a_d = _param(alpha, eng, "alpha")
            d, d_d = _param(delta, eng, "delta")
            check(lib().bjx_ghmc_step(eng.h, ptr(keys), ptr(q), ptr(p), ptr(logp), ptr(g), ptr(sl), eps, ptr(eps_d), a,
                                      ptr(a_d), d, ptr(d_d), None, None, 1, 0, 0, C.byref(info)), eng.h)
        proposal = None
        if full_info:
            proposal = IntegratorState(fields["proposal_position"], fields["proposal_momentum"], None, None)
        hinfo = HMCInfo(fields.get("momentum"), fields["acceptance_rate"], fields["is_accepted"].bool(),
                        fields["is_divergent"].bool(), fields["energy"], proposal, 1)
        return GHMCState(q, p, logp, g, sl), hinfo

    return kernel


def update_momentum(rng_key, state, alpha, momentum_generator):
    """blackjax/mcmc/ghmc.py:192-213 (host-level helper; the transition kernel fuses it)."""
    fresh = momentum_generator(rng_key, state.position)
    a = torch.as_tensor(alpha, dtype=torch.float32, device=state.momentum.device)
    return state.momentum * torch.sqrt(1.0 - a) + torch.sqrt(a) * fresh


def as_top_level_api(logdensity_fn, step_size, momentum_inverse_scale, alpha, delta, *, divergence_threshold: int = 1000,
                     noise_fn=None, **kw):
    """blackjax/mcmc/ghmc.py:216-317."""
    kernel = build_kernel(noise_fn, divergence_threshold, **kw)
    return build_sampling_algorithm(kernel, init, logdensity_fn,
                                    kernel_args=(step_size, momentum_inverse_scale, alpha, delta),
                                    pass_rng_key_to_init=True)

"""ChEES-HMC warm-up: cross-chain adaptation of the step size and the trajectory length of jittered HMC.

Mirrors ``blackjax.chees_adaptation`` (blackjax/adaptation/chees_adaptation.py:574-1025) for its default configuration
(``mass_matrix_estimation=None``: identity metric, base-2 Halton jitter).  Every warm-up step is one HMC transition of all
chains with ``ceil(jitter(i) * T / eps)`` leapfrog steps followed by ``bjx_chees_update`` (libbjx): block statistics of the
proposals, two all-gathers when the chains are sharded over GPUs, dual averaging on the harmonic-mean acceptance
(:341-360) and Adam on ``log T`` along the ChEES gradient (:362-511), all on the device -- no host round trip per step.

``optim`` stands in for the reference's ``optax.adam(learning_rate, b1, b2)`` (optax is not available here): pass
``blackjax_b200.adaptation.chees_adaptation.adam(learning_rate, b1=0.9, b2=0.999)``.
"""
import math
from typing import NamedTuple

import torch

from .. import random as bjx_random
from .._lib import check, lib, ptr
from ..base import AdaptationAlgorithm, AdaptationResults
from ..mcmc import dynamic_hmc, hmc

OPTIMAL_TARGET_ACCEPTANCE_RATE = 0.651   # chees_adaptation.py:20


class adam(NamedTuple):
    """optax.adam(learning_rate, b1, b2) (eps = 1e-8), the optimiser the reference's examples pass as ``optim``."""
    learning_rate: float
    b1: float = 0.9
    b2: float = 0.999


def halton_sequence(i: int, max_bits: int = 10) -> float:
    """blackjax/mcmc/dynamic_hmc.py:205-215."""
    return sum((((int(i) + 1) >> k) & 1) * (0.5 / (1 << k)) for k in range(int(max_bits)))


def chees_adaptation(logdensity_fn, num_chains: int, *, jitter_amount: float = 1.0,
                     target_acceptance_rate: float = OPTIMAL_TARGET_ACCEPTANCE_RATE, decay_rate: float = 0.5,
                     max_leapfrog_steps: int = 1000, mass_matrix_estimation=None, process_group=None):
    """``num_chains`` is the number of chains of THIS process (``positions.shape[0]``); with a ``process_group`` the
    statistics are pooled over the chains of all its ranks."""
    if mass_matrix_estimation is not None:
        raise NotImplementedError("chees_adaptation is built for mass_matrix_estimation=None (identity metric)")

    def run(rng_key, positions, step_size: float, optim: adam, num_steps: int = 1000, *, max_sampling_steps: int = 1000):
        from .._comm import nccl_comm
        from .._engine import get_engine
        positions = positions.contiguous()
        C, D = positions.shape
        if C != num_chains:
            raise ValueError("initial `positions` leading dimension must be equal to the `num_chains`")
        dev = positions.device
        comm, n_ranks, rank = nccl_comm(dev, process_group)
        max_bits = int(math.ceil(math.log2(num_steps + max_sampling_steps)))
        kernel = hmc.build_kernel(full_info=True, chain_offset=rank * C)
        state = hmc.init(positions, logdensity_fn)
        eng = get_engine(positions, logdensity_fn)
        L_ = lib()
        st = torch.empty(L_.bjx_chees_state_floats(C, D, n_ranks), dtype=torch.float32, device=dev)
        eps_c = torch.empty(C, dtype=torch.float32, device=dev)
        steps_c = torch.empty(C, dtype=torch.int32, device=dev)
        hist = torch.empty(num_steps, 4, dtype=torch.float32, device=dev)
        imm = torch.ones(D, dtype=torch.float32, device=dev)
        check(L_.bjx_chees_init(eng.h, ptr(st), float(step_size), max_bits, float(jitter_amount), ptr(eps_c), ptr(steps_c)), eng.h)
        keys = bjx_random.split(rng_key.to(dev), num_steps)
        for t in range(num_steps):
            init_q = state.position
            state, info = kernel(keys[t], state, logdensity_fn, eps_c, imm, steps_c)
            check(L_.bjx_chees_update(eng.h, comm, n_ranks, ptr(st), ptr(init_q), ptr(info.proposal.position),
                                      ptr(info.proposal.momentum), ptr(info.acceptance_rate),
                                      ptr(info.is_divergent.to(torch.uint8)), float(optim.learning_rate), float(optim.b1),
                                      float(optim.b2), float(target_acceptance_rate), float(decay_rate),
                                      int(max_leapfrog_steps), ptr(eps_c), ptr(steps_c), ptr(hist)), eng.h)
        import ctypes as C_
        out = (C_.c_float * 2)()
        check(L_.bjx_chees_final(eng.h, ptr(st), out), eng.h)          # the one host read of the warm-up
        step, n_lf = float(out[0]), float(out[1])
        jitter = lambda i: halton_sequence(i, max_bits) * jitter_amount + (1.0 - jitter_amount)
        parameters = {
            "step_size": step,
            "inverse_mass_matrix": imm,
            "next_random_arg_fn": lambda i: i + 1,
            "integration_steps_fn": lambda i, n: int(math.ceil(jitter(i) * n)),
            "integration_steps_params": (n_lf,),
        }
        last = dynamic_hmc.DynamicHMCState(state.position, state.logdensity, state.logdensity_grad, num_steps)
        return AdaptationResults(last, parameters), hist.cpu()

    return AdaptationAlgorithm(run)

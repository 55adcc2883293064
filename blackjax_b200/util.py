"""Driver loop mirroring ``blackjax.util.run_inference_algorithm`` (blackjax/util.py:150-213)."""
import torch

from . import random as bjx_random


def sample_hmc_native(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, num_integration_steps, num_steps,
                      *, multinomial=False, thin=1, keep_history=True, chain_offset=0):
    """``run_inference_algorithm`` for HMC / multinomial HMC without a Python-level loop: libbjx ``bjx_hmc_sample``
    derives ``split(rng_key, num_steps)`` on the device and enqueues every transition back to back (no host sync), so
    small problems run at kernel-launch rate.  Same draws as ``num_steps`` calls of ``hmc.step(split(rng_key, T)[t], .)``.
    Returns (final HMCState, positions [num_steps // thin, C, D] or None, acceptance rates [num_steps, C])."""
    from ._engine import get_engine
    from ._lib import check, lib, ptr
    from .mcmc.hmc import HMCState
    q, logp, g = (t.clone() for t in state)
    eng = get_engine(q, logdensity_fn)
    eng.ensure_metric(inverse_mass_matrix)
    key = rng_key.to(q.device).contiguous()
    if key.ndim != 1:
        raise ValueError("sample_hmc_native takes ONE rng_key of shape [2]")
    eng._key_mode(key, chain_offset)
    C, D = q.shape
    hist = torch.empty(num_steps // thin, C, D, dtype=torch.float32, device=q.device) if keep_history else None
    acc = torch.empty(num_steps, C, dtype=torch.float32, device=q.device)
    eps, eps_dev = eng._eps(step_size)
    check(lib().bjx_hmc_sample(eng.h, ptr(key), ptr(q), ptr(logp), ptr(g), eps, ptr(eps_dev), int(num_integration_steps),
                               int(num_steps), int(bool(multinomial)), ptr(hist), int(thin), ptr(acc)), eng.h)
    return HMCState(q, logp, g), hist, acc


def sample_nuts_native(rng_key, state, logdensity_fn, step_size, inverse_mass_matrix, num_steps, *, max_num_doublings=10,
                       thin=1, keep_history=True, chain_offset=0):
    """``run_inference_algorithm`` for NUTS without a Python-level loop (libbjx ``bjx_nuts_sample``): same draws as
    ``num_steps`` calls of ``nuts.step(split(rng_key, T)[t], .)``.  Returns (final HMCState, positions
    [num_steps // thin, C, D] or None, acceptance rates [num_steps, C], tree sizes int32 [num_steps, C])."""
    from ._engine import get_engine
    from ._lib import check, lib, ptr
    from .mcmc.hmc import HMCState
    q, logp, g = (t.clone() for t in state)
    eng = get_engine(q, logdensity_fn, max_tree_depth=max(10, max_num_doublings))
    eng.ensure_metric(inverse_mass_matrix)
    key = rng_key.to(q.device).contiguous()
    if key.ndim != 1:
        raise ValueError("sample_nuts_native takes ONE rng_key of shape [2]")
    eng._key_mode(key, chain_offset)
    C, D = q.shape
    hist = torch.empty(num_steps // thin, C, D, dtype=torch.float32, device=q.device) if keep_history else None
    acc = torch.empty(num_steps, C, dtype=torch.float32, device=q.device)
    n_int = torch.empty(num_steps, C, dtype=torch.int32, device=q.device)
    eps, eps_dev = eng._eps(step_size)
    check(lib().bjx_nuts_sample(eng.h, ptr(key), ptr(q), ptr(logp), ptr(g), eps, ptr(eps_dev), int(max_num_doublings),
                                int(num_steps), ptr(hist), int(thin), ptr(acc), ptr(n_int)), eng.h)
    return HMCState(q, logp, g), hist, acc, n_int


def run_inference_algorithm(rng_key, inference_algorithm, num_steps, initial_state=None, initial_position=None,
                            transform=lambda state, info: (state, info), collect=True):
    """``keys = split(rng_key, num_steps)`` then ``num_steps`` calls of ``step`` (util.py:200-211).
    Each step key is split into one key per chain (step-major schedule).  ``transform`` picks what is kept;
    with ``collect=False`` only the final state is returned."""
    if (initial_state is None) == (initial_position is None):
        raise ValueError("Either `initial_state` or `initial_position` must be specified, but not both.")
    if initial_state is None:
        initial_state = inference_algorithm.init(initial_position)
    keys = bjx_random.split(rng_key, num_steps)
    state = initial_state
    history = []
    for t in range(num_steps):
        state, info = inference_algorithm.step(keys[t], state)
        if collect:
            history.append(transform(state, info))
    return state, history

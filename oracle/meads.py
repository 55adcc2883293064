"""Restatement of BlackJAX's MEADS warm-up for generalized HMC (default path: diagonal momentum scale per fold).

TEST INFRASTRUCTURE (see oracle/__init__.py).  float32.

Follows blackjax/adaptation/meads_adaptation.py:
* maximum_eigenvalue          :787-817 (lambda_sq / lambda from the n x n Gram matrix X X^T)
* base.compute_parameters     :97-152, base.init :154-170 (all chains, replicated per fold)
* meads_adaptation.one_step   :497-690 with low_rank_rank=None: per-fold scales, step size from the PREVIOUS fold's
                              preconditioned gradients (roll by one), damping from the fold's own preconditioned, centred
                              positions, one GHMC step of every chain with its fold's parameters, the fold ``t mod K``
                              frozen, a shuffle of all chains every K steps
* meads_adaptation.run        :692-784 (key schedule, final parameters = fold means)
``jax.random.permutation`` is JAX's (jax >= 0.9, pyproject.toml:34; absent from /root/reference): the sort-based shuffle
of jax/_src/random.py ``_shuffle`` -- ceil(3 ln n / ln(2^32 - 1)) rounds of (key, subkey = split(key); stable sort of the
running permutation by 32 random bits drawn from subkey).  Restated from the published source; PARITY UNPINNED like the
rest of the PRNG layer (oracle/prng.py).
The oracle has no golden vectors of its own for MEADS (the reference's tests, tests/adaptation/test_meads.py, check
shapes and the properties replayed in tests/test_oracle_kat.py: replicated init, multiplier linearity, frozen fold).
"""
from typing import NamedTuple

import numpy as np

from . import prng
from .ghmc import GHMCState, ghmc_kernel, init as ghmc_init
from .hmc import F


def maximum_eigenvalue(X):
    """meads_adaptation.py:805-817 on an [n, d] matrix."""
    X = np.asarray(X, F)
    n = X.shape[0]
    S = (X @ X.T).astype(F)
    diag = np.diag(S)
    lam = F(np.sum(diag, dtype=F) / F(n))
    lam_sq = F((np.sum(S * S, dtype=F) - np.sum(diag * diag, dtype=F)) / F(n * (n - 1)))
    return F(lam_sq / lam)


def permutation(key, n):
    """jax.random.permutation(key, n) (see the module docstring)."""
    x = np.arange(n)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    key = np.asarray(key, np.uint32)
    for _ in range(rounds):
        ks = prng.split(key, 2)
        key, sub = ks[0], ks[1]
        bits = prng.random_bits(sub, (n,))
        x = x[np.argsort(bits, kind="stable")]
    return x


class MEADSState(NamedTuple):
    current_iteration: int
    step_size: np.ndarray        # [K]
    position_sigma: np.ndarray   # [K, D]
    alpha: np.ndarray            # [K]
    delta: np.ndarray            # [K]


def compute_parameters(positions, grads, current_iteration, step_size_multiplier=0.5, damping_slowdown=1.0):
    """base.compute_parameters (:97-152): one set of chains for both statistics."""
    mean = positions.mean(axis=0, dtype=F)
    sd = positions.std(axis=0, dtype=F).astype(F)
    normalized = ((positions - mean) / sd).astype(F)
    eps = F(min(F(step_size_multiplier) / np.sqrt(maximum_eigenvalue((grads * sd).astype(F))), F(1.0)))
    gamma = F(max(F(1.0) / np.sqrt(maximum_eigenvalue(normalized)), F(damping_slowdown) / F(F(current_iteration + 1) * eps)))
    alpha = F(F(1.0) - np.exp(F(-2.0) * eps * gamma))
    return eps, sd, alpha, F(alpha / F(2.0))


def meads_init(positions, grads, num_folds, step_size_multiplier=0.5, damping_slowdown=1.0):
    eps, sd, alpha, delta = compute_parameters(positions, grads, 0, step_size_multiplier, damping_slowdown)
    return MEADSState(0, np.full(num_folds, eps, F), np.repeat(sd[None], num_folds, 0), np.full(num_folds, alpha, F),
                      np.full(num_folds, delta, F))


def fold_parameters(position, grad, t, num_folds, step_size_multiplier=0.5, damping_slowdown=1.0):
    """The statistics half of one_step (:507-585): (step_size_rolled [K], scales_rolled [K, D], alphas [K], deltas [K])."""
    C, D = position.shape
    n = C // num_folds
    fq = position.reshape(num_folds, n, D)
    fg = grad.reshape(num_folds, n, D)
    scales = fq.std(axis=1, dtype=F).astype(F)
    own = np.array([min(F(step_size_multiplier) / np.sqrt(maximum_eigenvalue((fg[k] * scales[k]).astype(F))), F(1.0))
                    for k in range(num_folds)], F)
    eps_rolled = np.roll(own, 1)
    scales_rolled = np.roll(scales, 1, axis=0)
    alphas = np.zeros(num_folds, F)
    for k in range(num_folds):
        pk = (fq[k] / scales[k]).astype(F)
        pk = (pk - pk.mean(axis=0, dtype=F)).astype(F)
        gamma = F(max(F(1.0) / np.sqrt(maximum_eigenvalue(pk)), F(damping_slowdown) / F(F(t + 1) * eps_rolled[k])))
        alphas[k] = F(F(1.0) - np.exp(F(-2.0) * eps_rolled[k] * gamma))
    return eps_rolled, scales_rolled, alphas, (alphas / F(2.0)).astype(F)


def one_step(rng_key, state, adapt, target, num_folds, step_size_multiplier=0.5, damping_slowdown=1.0):
    """meads_adaptation.one_step (:497-690).  Returns (new GHMCState, new MEADSState, HMCInfo)."""
    C, D = state.position.shape
    n = C // num_folds
    t = adapt.current_iteration
    keys = prng.split(np.asarray(rng_key, np.uint32), C + 1)
    chain_keys, shuffle_key = keys[:C], keys[C]
    eps_r, scales_r, alphas, deltas = fold_parameters(state.position, state.logdensity_grad, t, num_folds,
                                                      step_size_multiplier, damping_slowdown)
    new, info = ghmc_kernel(chain_keys, state, target, np.repeat(eps_r, n), np.repeat(scales_r, n, axis=0),
                            np.repeat(alphas, n), np.repeat(deltas, n))
    if num_folds > 1:
        skipped = np.repeat(np.arange(num_folds) == (t % num_folds), n)
        new = GHMCState(*[np.where(skipped.reshape((C,) + (1,) * (a.ndim - 1)), b, a) for a, b in zip(new, state)])
    adapt = MEADSState(t + 1, eps_r, scales_r, alphas, deltas)
    if num_folds > 1 and (t + 1) % num_folds == 0:
        perm = permutation(shuffle_key, C)
        new = GHMCState(*[a[perm] for a in new])
    return new, adapt, info


def meads_run(target, rng_key, positions, num_steps, num_folds=4, step_size_multiplier=0.5, damping_slowdown=1.0, trace=None):
    """meads_adaptation(...).run (:692-784).  Returns (last GHMCState, parameters dict, final MEADSState)."""
    positions = np.asarray(positions, F)
    C = positions.shape[0]
    ks = prng.split(np.asarray(rng_key, np.uint32), 2)
    key_init, key_adapt = ks[0], ks[1]
    state = ghmc_init(positions, target, prng.split(key_init, C))
    adapt = meads_init(positions, state.logdensity_grad, num_folds, step_size_multiplier, damping_slowdown)
    keys = prng.split(key_adapt, num_steps)
    for t in range(num_steps):
        state, adapt, _ = one_step(keys[t], state, adapt, target, num_folds, step_size_multiplier, damping_slowdown)
        if trace is not None:
            trace.append((state, adapt))
    params = {"step_size": F(adapt.step_size.mean(dtype=F)), "momentum_inverse_scale": adapt.position_sigma.mean(axis=0, dtype=F),
              "alpha": F(adapt.alpha.mean(dtype=F)), "delta": F(adapt.delta.mean(dtype=F))}
    return state, params, adapt

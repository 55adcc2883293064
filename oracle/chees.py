"""Restatement of BlackJAX's ChEES-HMC warm-up (default path: identity metric, Halton jitter).

TEST INFRASTRUCTURE (see oracle/__init__.py).  float32.

Follows blackjax/adaptation/chees_adaptation.py:
* state / init             :28-58, :513-523
* compute_parameters       :307-511 (harmonic-mean dual averaging, weighted proposal mean, the ChEES gradient
                           jitter * T * (|dx'|^2 - |dx|^2) <dx', p'>, Adam on log T with the +-0.35 clip, moving averages)
* run loop                 :737-1025 with mass_matrix_estimation=None (jitter_gn, integration_steps_fn, one_step, final
                           parameters)
* Halton sequence          blackjax/mcmc/dynamic_hmc.py:205-215
* dual averaging           blackjax/optimizers/dual_averaging.py:87-129 (gradient passed directly)
* Adam                     optax.adam(learning_rate, b1, b2, eps=1e-8): scale_by_adam with bias correction, then -lr
                           (optax is a third-party dependency of the reference, absent from /root/reference; algorithm
                           as published: Kingma & Ba 2015, optax/_src/transform.py scale_by_adam)
"""
from typing import NamedTuple

import numpy as np

from . import prng
from .adaptation import DAState, da_init
from .hmc import F, hmc_kernel, init as hmc_init

OPTIMAL_TARGET_ACCEPTANCE_RATE = 0.651
LOG_UPDATE_CLIP = F(0.35)
EPS_FLOAT = F(1e-20)


def halton(i, max_bits=10):
    """(i+1)-th element of the base-2 Halton sequence (dynamic_hmc.py:205-215); exact in float32."""
    i = int(i)
    s = F(0.0)
    for k in range(int(max_bits)):
        s = F(s + F(((i + 1) // (2 ** k)) % 2) * F(0.5 / 2 ** k))
    return s


def da_update_grad(s, gradient, t0=10, gamma=0.05, kappa=0.75):
    """dual_averaging.py:101-123 with the gradient given directly."""
    log_step, avg_log_step, step, avg_error, mu = s
    reg_step = F(step + t0)
    eta_t = F(F(step) ** F(-kappa))
    avg_error = F((F(1.0) - (F(1.0) / reg_step)) * avg_error + F(gradient) / reg_step)
    log_x = F(mu - (np.sqrt(F(step)) / F(gamma)) * avg_error)
    log_x_avg = F(eta_t * log_step + (F(1.0) - eta_t) * avg_log_step)
    return DAState(log_x, log_x_avg, step + 1, avg_error, mu)


class Adam(NamedTuple):
    count: int
    mu: np.float32
    nu: np.float32


def adam_update(g, s, lr, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam on a scalar parameter: returns (update, new state); update = -lr * m_hat / (sqrt(v_hat) + eps)."""
    g = F(g)
    mu = F(F(b1) * s.mu + (F(1.0) - F(b1)) * g)
    nu = F(F(b2) * s.nu + (F(1.0) - F(b2)) * g * g)
    count = s.count + 1
    with np.errstate(divide="ignore", invalid="ignore"):
        mu_hat = F(mu / (F(1.0) - F(F(b1) ** F(count))))
        nu_hat = F(nu / (F(1.0) - F(F(b2) ** F(count))))
        upd = F(F(-lr) * (mu_hat / (np.sqrt(nu_hat) + F(eps))))
    return upd, Adam(count, mu, nu)


class ChEESState(NamedTuple):
    step_size: np.float32
    log_step_size_ma: np.float32
    trajectory_length: np.float32
    log_trajectory_length_ma: np.float32
    da: DAState
    optim: Adam
    random_generator_arg: int
    step: int


def chees_init(step_size):
    """chees_adaptation.py:513-523."""
    x = F(step_size)
    return ChEESState(x, F(0.0), x, F(0.0), da_init(x), Adam(0, F(0.0), F(0.0)), 0, 1)


def jitter(i, jitter_amount, max_bits):
    return F(halton(i, max_bits) * F(jitter_amount) + (F(1.0) - F(jitter_amount)))


def chees_update(s, prop_q, prop_p, init_q, acc, is_div, *, lr, b1=0.9, b2=0.999, target=OPTIMAL_TARGET_ACCEPTANCE_RATE,
                 decay_rate=0.5, max_leapfrog_steps=1000, jitter_amount=1.0, max_bits=10):
    """compute_parameters (chees_adaptation.py:307-511) with inverse_mass_matrix = ones (the whitening is a no-op)."""
    acc = np.asarray(acc, F)
    nd = ~np.asarray(is_div, bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        hm = F(1.0) / (np.sum(np.where(nd, F(1.0) / acc, F(0.0)), dtype=F) / F(nd.sum()))
    hm = hm if np.isfinite(hm) else F(0.0)
    da_ = da_update_grad(s.da, F(target) - hm)
    eps_ = np.exp(da_.log_step_size).astype(F)
    if np.isfinite(eps_):
        new_eps, new_da, new_log_eps = eps_, da_, da_.log_step_size
    else:
        new_eps, new_da, new_log_eps = s.step_size, s.da, s.da.log_step_size
    uw = F(F(s.step) ** F(-decay_rate))
    new_log_eps_ma = F((F(1.0) - uw) * s.log_step_size_ma + uw * new_log_eps)
    # weighted_empirical_mean :239-247 and nanmean
    w = np.where(nd, acc, F(0.0)).astype(F)
    fin = np.isfinite(prop_q)
    xs = np.where(fin, prop_q, F(0.0)).astype(F)
    w_ = np.where(fin.all(axis=-1), w, F(0.0)).astype(F)
    pmean = (np.sum(w_[:, None] * xs, axis=0, dtype=F) / (np.sum(w_, dtype=F) + EPS_FLOAT)).astype(F)
    imean = np.nanmean(init_q, axis=0, dtype=F).astype(F)
    pc = (prop_q - pmean).astype(F)
    ic = (init_q - imean).astype(F)
    dots = ((np.sum(pc * pc, axis=1, dtype=F) - np.sum(ic * ic, axis=1, dtype=F)) * np.sum(pc * prop_p, axis=1, dtype=F)).astype(F)
    tg = (jitter(s.random_generator_arg, jitter_amount, max_bits) * s.trajectory_length * dots).astype(F)
    with np.errstate(invalid="ignore"):
        grad = F(np.sum(np.where(nd, acc * tg, F(0.0)), dtype=F) / np.sum(np.where(nd, acc + EPS_FLOAT, F(0.0)), dtype=F))
    log_T = np.log(s.trajectory_length).astype(F)
    upd, optim_ = adam_update(grad, s.optim, lr, b1, b2)
    upd = F(np.clip(upd, -LOG_UPDATE_CLIP, LOG_UPDATE_CLIP))
    log_T_ = F(log_T + upd)
    if np.isfinite(log_T_):
        new_log_T, new_optim = log_T_, optim_
    else:
        new_log_T, new_optim = log_T, s.optim
    new_log_T_ma = F((F(1.0) - uw) * s.log_trajectory_length_ma + uw * new_log_T)
    new_T = np.exp(new_log_T_ma).astype(F)
    new_T = F(np.clip(new_T, new_eps, F(max_leapfrog_steps) * new_eps))
    return ChEESState(new_eps, new_log_eps_ma, new_T, new_log_T_ma, new_da, new_optim, s.random_generator_arg + 1, s.step + 1)


def integration_steps(i, num_leapfrog_steps, jitter_amount, max_bits):
    """integration_steps_fn :775-779."""
    return int(np.ceil(F(jitter(i, jitter_amount, max_bits) * F(num_leapfrog_steps))))


def chees_run(target, rng_key, positions, step_size, *, lr, b1=0.9, b2=0.999, num_steps=1000, max_sampling_steps=1000,
              target_acceptance_rate=OPTIMAL_TARGET_ACCEPTANCE_RATE, decay_rate=0.5, max_leapfrog_steps=1000,
              jitter_amount=1.0, trace=None):
    """chees_adaptation(...).run (mass_matrix_estimation=None).  Returns (last HMCState, step_size, num_leapfrog_steps,
    final ChEESState); ``trace`` (list) receives the adaptation state after every step."""
    positions = np.asarray(positions, F)
    C, D = positions.shape
    max_bits = int(np.ceil(np.log2(num_steps + max_sampling_steps)))
    imm = np.ones(D, F)
    state = hmc_init(positions, target)
    s = chees_init(step_size)
    keys_step = prng.split(np.asarray(rng_key, np.uint32), num_steps)
    for t in range(num_steps):
        L = integration_steps(s.random_generator_arg, F(s.trajectory_length / s.step_size), jitter_amount, max_bits)
        keys = prng.split(keys_step[t], C)
        init_q = state.position
        state, info = hmc_kernel(keys, state, target, s.step_size, imm, L)
        s = chees_update(s, info.proposal[0], info.proposal[1], init_q, info.acceptance_rate, info.is_divergent, lr=lr, b1=b1,
                         b2=b2, target=target_acceptance_rate, decay_rate=decay_rate, max_leapfrog_steps=max_leapfrog_steps,
                         jitter_amount=jitter_amount, max_bits=max_bits)
        if trace is not None:
            trace.append((s, L))
    eps = np.exp(s.log_step_size_ma).astype(F)
    n_lf = np.exp(F(s.log_trajectory_length_ma - s.log_step_size_ma)).astype(F)
    return state, eps, n_lf, s

"""Restatement of blackjax.diagnostics.potential_scale_reduction (blackjax/diagnostics.py:39-89).  TEST INFRASTRUCTURE."""
import numpy as np


def potential_scale_reduction(x, chain_axis=0, sample_axis=1):
    x = np.asarray(x, np.float64)
    assert x.shape[chain_axis] > 1
    n = x.shape[sample_axis]
    per_chain_mean = x.mean(axis=sample_axis, keepdims=True)
    per_chain_var = x.var(axis=sample_axis, ddof=1, keepdims=True)
    between = n * per_chain_mean.var(axis=chain_axis, ddof=1, keepdims=True)
    within = per_chain_var.mean(axis=chain_axis, keepdims=True)
    estimator = (n - 1) / n * within + between / n
    return np.sqrt(estimator / within).squeeze()


def effective_sample_size(x, chain_axis=0, sample_axis=1):
    """Restatement of blackjax.diagnostics.effective_sample_size (blackjax/diagnostics.py:159-305): Stan-style ESS with
    Geyer's initial positive and initial monotone sequence estimators on the chain-averaged autocovariance.

    Written as explicit loops over the lag pairs instead of the reference's scans/scatter so that the two places where
    the reference indexes one past the last pair are spelled out: JAX drops an out-of-bounds scatter (``.at[].set``,
    diagnostics.py:273) and clamps an out-of-bounds gather (``rho_hat_even[indices]``, :273,:297)."""
    x = np.moveaxis(np.asarray(x, np.float64), (chain_axis, sample_axis), (0, 1))
    C, T = x.shape[:2]
    assert T > 1, f"The input array must have at least 2 samples, got only {T}."
    ev = x.shape[2:]
    x = x.reshape(C, T, -1)
    E = x.shape[2]
    has_var = np.any(x != x[:, :1], axis=(0, 1))                       # :203-208
    m = x.mean(axis=1, keepdims=True)                                  # :210
    xc = x - m
    # linear (zero-padded) autocovariance for every lag, divided by T (:212-220), then averaged over chains (:221)
    n = 1
    while n < 2 * T:
        n *= 2
    f = np.fft.rfft(xc, n=n, axis=1)
    acov = np.fft.irfft(f * np.conj(f), n=n, axis=1)[:, :T] / T
    macov = acov.mean(axis=0)                                          # [T, E]
    ess = np.zeros(E)
    for e in range(E):
        a = macov[:, e].astype(np.float32)                             # the reference computes in float32
        f32 = np.float32
        var0 = a[0] * f32(T) / f32(T - 1.0)                            # :222-226
        degenerate = np.isfinite(var0) and ((not has_var[e]) or var0 <= 0.0)   # :227-229
        wvar = var0 * f32(T - 1.0) / f32(T)                            # :230
        if C > 1:
            wvar = wvar + f32(m[:, 0, e].var(ddof=1))                  # :231-237
        if degenerate:
            wvar = f32(1.0)                                            # :238-240
        T_even = T - T % 2
        rho = np.ones(T_even, np.float32)
        rho[1:] = f32(1.0) - (var0 - a[1:T_even]) / wvar               # :243-253
        even, odd = rho[0::2].copy(), rho[1::2].copy()                 # :255-257
        K = len(even)
        # Geyer's initial positive sequence (:259-270)
        mask = np.zeros(K, bool)
        carry, max_t = True, 0
        for k in range(K):
            carry = carry and bool(even[k] + odd[k] > 0.0)
            mask[k] = carry
            if carry:
                max_t = k
        idx = max_t + 1
        idx_get = min(idx, K - 1)                                      # out-of-bounds gather clamps
        odd = np.where(mask, odd, f32(0.0))                            # :271
        mask_even = mask.copy()
        if idx < K:                                                    # out-of-bounds scatter is dropped
            mask_even[idx] = bool(even[idx_get] > 0)                   # :273
        even = np.where(mask_even, even, f32(0.0))                     # :274
        # Geyer's initial monotone sequence (:277-289)
        s = even + odd
        carry_v = s[0]
        even_f, odd_f = even.copy(), odd.copy()
        for k in range(K):
            upd = s[k] > carry_v
            nxt = carry_v if upd else s[k]
            carry_v = nxt
            if upd:
                even_f[k] = nxt / f32(2.0)
                odd_f[k] = nxt / f32(2.0)
        ess_raw = C * T                                                # :292
        tau = f32(-1.0) + f32(2.0) * np.sum(even_f + odd_f, dtype=np.float32) - even_f[idx_get]   # :293-297
        tau = max(tau, f32(1.0 / np.log10(ess_raw)))                   # :299
        ess[e] = 0.0 if degenerate else ess_raw / tau                  # :300-301
    return ess.reshape(ev) if ev else float(ess[0])

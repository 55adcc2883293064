"""GPU parity tests of generalized HMC and the MEADS warm-up against oracle/ghmc.py and
oracle/meads.py, through the C ABI (bjx_ghmc_step, bjx_meads_update, bjx_maximum_eigenvalue, bjx_permutation,
bjx_gather_rows)."""
import numpy as np
import pytest
import torch

import blackjax_b200 as bj
from blackjax_b200 import _engine, targets as T
from blackjax_b200._lib import check, lib, ptr
from blackjax_b200.adaptation import meads_adaptation as mm
from oracle import ghmc as oghmc
from oracle import meads as omeads
from oracle import prng as oprng
from oracle import targets as otargets

pytestmark = pytest.mark.gpu
F = np.float32
DEV = "cuda:0"


def tk(keys_np):
    return torch.from_numpy(np.ascontiguousarray(keys_np).view(np.int32)).to(DEV).view(torch.uint32)


def tf(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def close(a, b, rtol, floor):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    bound = rtol * np.maximum(np.abs(b), floor)
    bad = np.abs(a - b) > bound
    assert not bad.any(), (f"{bad.sum()} of {bad.size} off; worst {np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)):.3e} "
                           f"(rtol {rtol:.1e}, floor {floor:.1e})")


def _targets(kind, D):
    if kind == "diag":
        s = np.logspace(-0.3, 0.6, D)
        return T.DiagGaussian(s), otargets.DiagGaussian(s)
    return T.Funnel(D), otargets.Funnel(D)


@pytest.mark.parametrize("kind,D,C", [("diag", 16, 512), ("diag", 200, 256), ("funnel", 10, 512)])
def test_ghmc_transitions_match_oracle(kind, D, C):
    """Five free-running GHMC transitions with per-chain step size, alpha, delta and a shared momentum scale: state (q, p,
    logp, grad, slice) and info agree with the oracle; a chain may differ in its accept decision only when log|slice| sits
    within 1e-4 of the energy difference (a float tie) -- such chains are dropped from the later comparisons."""
    tgt, otgt = _targets(kind, D)
    rng = np.random.default_rng(3)
    q0 = (rng.standard_normal((C, D)) * 0.7).astype(F)
    scale = np.exp(0.3 * rng.standard_normal(D)).astype(F)
    eps = (0.05 + 0.1 * rng.random(C)).astype(F)
    alpha = (0.1 + 0.8 * rng.random(C)).astype(F)
    delta = (alpha / 2).astype(F)
    key0 = oprng.split(oprng.key(11), C)
    ost = oghmc.init(q0, otgt, key0)
    st = bj.ghmc.init(tf(q0), tgt, tk(key0))
    for name, a, b in zip(st._fields, st, ost):   # logdensity: float32 sums in a different order (funnel: exp terms)
        close(npy(a), b, 1e-4 if name.startswith("logdensity") else 1e-6, 1.0 if name == "logdensity" else 1e-3)
    kernel = bj.ghmc.build_kernel(full_info=True)
    ok = np.ones(C, bool)
    step_keys = oprng.split(oprng.key(12), 5)
    for t in range(5):
        keys = oprng.split(step_keys[t], C)
        margins = []
        ost, oinfo = oghmc.ghmc_kernel(keys, ost, otgt, eps, scale, alpha, delta, margins=margins)
        st, info = kernel(tk(keys), st, tgt, tf(eps), tf(scale), tf(alpha), tf(delta))
        acc = npy(info.is_accepted)
        flip = acc != oinfo.is_accepted
        live = flip & ok    # chains already off their oracle twin no longer say anything
        assert (margins[0][live] < 1e-4 * np.maximum(1.0, np.abs(oinfo.energy[live]))).all(), margins[0][live]
        ok &= ~flip
        close(npy(info.acceptance_rate)[ok], oinfo.acceptance_rate[ok], 2e-4, 1e-2)
        close(npy(info.energy)[ok], oinfo.energy[ok], 1e-4, 1.0)
        close(npy(info.momentum)[ok], oinfo.momentum[ok], 1e-4, 1e-2)
        close(npy(info.proposal.position)[ok], oinfo.proposal[0][ok], 1e-4, 1e-2)
        # slice' = slice * exp(-delta_energy) inherits the ABSOLUTE error of the energy difference (float32 sums of D terms
        # of size |H| ~ D) and, being persistent, compounds it over the five transitions
        tol = {"logdensity": (1e-4, 1.0), "slice": (5e-5 * max(D, 100), 1e-2)}
        for name, a, b in zip(st._fields, st, ost):
            close(npy(a)[ok], np.asarray(b)[ok], *tol.get(name, (2e-4, 1e-2)))
        np.testing.assert_array_equal(npy(info.is_divergent)[ok], oinfo.is_divergent[ok])
    assert ok.mean() > 0.97


def test_ghmc_noise_fn_matches_oracle():
    """ghmc.py:90,172: ``noise_fn(key_noise)`` added to the slice translation; key_noise = split(rng_key)[1] per chain.
    Per-chain keys and a shared step key (per-chain keys derived from the global chain index) give the same draws."""
    D, C = 12, 256
    tgt, otgt = _targets("diag", D)
    rng = np.random.default_rng(5)
    q0 = (rng.standard_normal((C, D)) * 0.7).astype(F)
    key0 = oprng.split(oprng.key(21), C)
    ost = oghmc.init(q0, otgt, key0)
    st = bj.ghmc.init(tf(q0), tgt, tk(key0))
    kernel = bj.ghmc.build_kernel(noise_fn=lambda k: 0.4 * bj.random.normal(k) + 0.05)
    onoise = lambda k: (F(0.4) * oprng.normal(k) + F(0.05)).astype(F)
    scale = np.ones(D, F)
    sk = oprng.key(22)
    keys = oprng.split(sk, C)
    onew, oinfo = oghmc.ghmc_kernel(keys, ost, otgt, F(0.1), scale, F(0.5), F(0.3), noise_fn=onoise)
    for k_dev in (tk(keys), tk(sk)):      # explicit per-chain keys | one step key
        new, info = kernel(k_dev, st, tgt, 0.1, tf(scale), 0.5, 0.3)
        same = npy(info.is_accepted) == oinfo.is_accepted
        assert same.mean() > 0.98
        close(npy(new.slice)[same], onew.slice[same], 2e-4, 1e-2)
        close(npy(new.position)[same], onew.position[same], 1e-4, 1e-2)
    # and the noise does change the chain: without it the slice variables differ
    plain, _ = bj.ghmc.build_kernel()(tk(keys), st, tgt, 0.1, tf(scale), 0.5, 0.3)
    assert float((plain.slice - new.slice).abs().max()) > 0.1


def test_ghmc_skipped_chains_keep_their_state():
    D, C = 16, 256
    tgt, _ = _targets("diag", D)
    q0 = tf(np.random.default_rng(0).standard_normal((C, D)))
    st = bj.ghmc.init(q0, tgt, bj.random.key(1, DEV))
    eng = _engine.get_engine(st.position, tgt)
    eng.ensure_metric(torch.ones(D, device=DEV))
    before = [x.clone() for x in st]
    key = bj.random.key(2, DEV)
    eng._key_mode(key, 0)
    info = eng._info({})
    import ctypes as C_
    check(lib().bjx_ghmc_step(eng.h, ptr(key), ptr(st.position), ptr(st.momentum), ptr(st.logdensity), ptr(st.logdensity_grad),
                              ptr(st.slice), 0.2, None, 0.5, None, 0.25, None, None, None, 1, 64, 128, C_.byref(info)), eng.h)
    for a, b in zip(st, before):
        assert torch.equal(a[64:128], b[64:128])
        assert not torch.equal(a[:64], b[:64])


def test_ghmc_reference_univariate_normal_case():
    """tests/mcmc/test_sampling.py:1160-1172: ghmc(step_size=1, momentum_inverse_scale=1, alpha=0.8, delta=2) on a
    N(1, 2^2) target; mean and std within the reference's tolerance (1e-1 relative), here pooled over 256 chains."""
    C = 256
    tgt = T.DiagGaussian(np.array([2.0]), mean=np.array([1.0]))
    alg = bj.ghmc(tgt, step_size=1.0, momentum_inverse_scale=torch.ones(1, device=DEV), alpha=0.8, delta=2.0, inplace=True)
    st = alg.init(torch.ones(C, 1, device=DEV), bj.random.key(3, DEV))
    keys = bj.random.split(bj.random.key(4, DEV), 1200)
    draws = []
    for i in range(1200):
        st, _ = alg.step(keys[i], st)
        if i >= 200:
            draws.append(st.position.clone())
    x = torch.stack(draws).reshape(-1)
    np.testing.assert_allclose(float(x.mean()), 1.0, rtol=1e-1)
    np.testing.assert_allclose(float(x.std()), 2.0, rtol=1e-1)


@pytest.mark.parametrize("n,d", [(64, 8), (1000, 100), (333, 70), (4096, 512)])
def test_maximum_eigenvalue_matches_oracle(n, d):
    rng = np.random.default_rng(n)
    X = (rng.standard_normal((n, d)) * np.logspace(-1, 1, d)).astype(F)
    got = float(mm.maximum_eigenvalue(tf(X)))
    want = float(omeads.maximum_eigenvalue(X)) if n <= 1000 else None
    if want is None:   # float64 reference of the same estimator (the n x n Gram matrix in float32 loses digits here)
        S = X.astype(np.float64) @ X.astype(np.float64).T
        dg = np.diag(S)
        want = ((S * S).sum() - (dg * dg).sum()) / (n * (n - 1)) / (dg.sum() / n)
    assert abs(got / want - 1) < 2e-4, (got, want)


@pytest.mark.parametrize("C,D,K", [(512, 24, 4), (128, 2, 4), (2048, 200, 8), (96, 40, 1)])
def test_meads_fold_parameters_match_oracle(C, D, K):
    """bjx_meads_update vs the statistics half of one_step (meads_adaptation.py:507-585): step size, alpha, delta (2e-4
    relative: different but fixed summation orders, D x D instead of n x n Gram matrix) and the rolled scales (1e-5)."""
    rng = np.random.default_rng(C + D)
    q = (rng.standard_normal((C, D)) * np.logspace(-0.5, 0.5, D) + rng.standard_normal(D)).astype(F)
    g = (-q / np.logspace(-1, 1, D) + 0.1 * rng.standard_normal((C, D))).astype(F)
    tgt = T.StdNormal(D)
    eng = _engine.get_engine(tf(q), tgt)
    for t in (0, 7, 400):
        f = mm._Folds(eng, K)
        f.update(tf(q), tf(g), t, 0.5, 1.0)
        eps, sig, al, de = omeads.fold_parameters(q, g, t, K)
        close(npy(f.step_size), eps, 2e-4, 1e-6)
        close(npy(f.alpha), al, 2e-4, 1e-6)
        close(npy(f.delta), de, 2e-4, 1e-6)
        close(npy(f.sigma), sig, 1e-5, 1e-6)
        close(npy(f.imm), sig * sig, 1e-5, 1e-12)
        close(npy(f.msqrt), 1 / sig, 1e-5, 1e-12)


@pytest.mark.parametrize("n", [1, 2, 10, 128, 1625, 1626, 5000, 70000])
def test_permutation_is_bit_exact(n):
    eng = _engine.get_engine(torch.zeros(8, 4, device=DEV), T.StdNormal(4))
    L = lib()
    key = oprng.fold_in(oprng.key(7), 3)
    perm = torch.empty(n, dtype=torch.int32, device=DEV)
    scratch = torch.empty(L.bjx_permutation_scratch_bytes(n), dtype=torch.uint8, device=DEV)
    check(L.bjx_permutation(eng.h, ptr(tk(key)), -1, n, ptr(perm), ptr(scratch)), eng.h)
    np.testing.assert_array_equal(npy(perm), omeads.permutation(key, n))
    # fold_index form: key' = fold_in(key, i) = split(key, i + 1)[i]
    check(L.bjx_permutation(eng.h, ptr(tk(oprng.key(7))), 3, n, ptr(perm), ptr(scratch)), eng.h)
    np.testing.assert_array_equal(npy(perm), omeads.permutation(key, n))


def test_meads_run_matches_oracle_free_run():
    """meads_adaptation(...).run against oracle.meads_run: 10 warm-up steps (two shuffles) of 128 chains in 4 folds.
    Frozen folds and the permutation are exact; states agree to 1e-3 for every chain whose accept decisions agreed."""
    C, D, K, steps = 128, 6, 4, 10
    s = np.logspace(-0.3, 0.5, D)
    tgt, otgt = T.DiagGaussian(s), otargets.DiagGaussian(s)
    q0 = (np.random.default_rng(5).standard_normal((C, D)) + 1.0).astype(F)
    trace = []
    ost, oparams, oad = omeads.meads_run(otgt, oprng.key(21), q0, steps, num_folds=K, trace=trace)
    snaps = []
    warm = bj.meads_adaptation(tgt, num_chains=C, num_folds=K,
                               adaptation_info_fn=lambda st, info, ad: snaps.append((st.position.clone(), ad)))
    (last, params), _ = warm.run(bj.random.key(21, DEV), tf(q0), num_steps=steps)
    n = C // K
    np.testing.assert_array_equal(npy(snaps[0][0])[:n], q0[:n])                          # fold 0 frozen at step 0
    np.testing.assert_array_equal(npy(snaps[1][0])[n:2 * n], npy(snaps[0][0])[n:2 * n])  # fold 1 frozen at step 1
    agree = 0
    for t in range(steps):
        dq = np.abs(npy(snaps[t][0]) - trace[t][0].position).max(axis=1)
        agree = (dq < 1e-3).mean()
        close(npy(snaps[t][1].step_size), trace[t][1].step_size, 1e-3 if agree == 1.0 else 5e-2, 1e-6)
    assert agree > 0.9, agree
    assert abs(float(params["step_size"]) / float(oparams["step_size"]) - 1) < 5e-2
    assert abs(float(params["alpha"]) / float(oparams["alpha"]) - 1) < 5e-2
    close(npy(params["momentum_inverse_scale"]), oparams["momentum_inverse_scale"], 5e-2, 1e-3)


def test_meads_reference_convergence_case():
    """The recipe of tests/mcmc/test_sampling.py:606-690 on a built-in target (the reference's linear-regression posterior
    is not a libbjx target): 128 chains, 4 folds, 1000 warm-up steps, then 100 GHMC steps with the adapted parameters;
    positive finite fold step sizes throughout and posterior moments recovered."""
    C, D, K = 128, 4, 4
    mean = np.array([3.0, -1.0, 0.5, 10.0])
    s = np.array([0.05, 1.0, 3.0, 0.3])
    tgt = T.DiagGaussian(s, mean=mean)
    rng = np.random.default_rng(9)
    q0 = (mean + 1.0 + rng.standard_normal((C, D))).astype(F)
    eps_hist = []
    warm = bj.meads_adaptation(tgt, num_chains=C, num_folds=K, adaptation_info_fn=lambda st, info, ad: eps_hist.append(ad.step_size))
    (last, params), _ = warm.run(bj.random.key(1, DEV), tf(q0), num_steps=1000)
    e = torch.stack(eps_hist)
    assert e.shape == (1000, K) and bool(torch.isfinite(e).all()) and bool((e > 0).all())
    assert not torch.allclose(e[-1, 0], e[-1, 1])                 # folds develop their own parameters
    alg = bj.ghmc(tgt, **params)
    st = last
    keys = bj.random.split(bj.random.key(2, DEV), 400)
    draws = []
    for k in keys:
        st, _ = alg.step(k, st)
        draws.append(st.position)
    x = torch.stack(draws[100:]).reshape(-1, D)
    assert (np.abs(npy(x.mean(0)) - mean) < 0.15 * s).all(), npy(x.mean(0)) - mean
    np.testing.assert_allclose(npy(x.std(0)), s, rtol=0.15)

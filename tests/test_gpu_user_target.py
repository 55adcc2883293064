"""User-defined targets (BJX_TARGET_USER, include/bjx_user_target.h): the slot of BlackJAX's arbitrary
``logdensity_fn`` callable (blackjax/mcmc/hmc.py:91, nuts.py:133) filled by a plug-in compiled from the user's fused
``value_and_grad``.

* plumbing: the diagonal Gaussian written as a plug-in must reproduce the built-in target BIT FOR BIT through every
  kernel of the path (same arithmetic, different translation unit / shared library / launch route), at every row size
  class;
* a model that is not built in -- the regression posterior of the reference's own sampling tests
  (tests/mcmc/test_sampling.py:103-111) -- against the numpy oracle, teacher-forced, at the 1e-5 of the other parity
  tests, for HMC / multinomial HMC / NUTS, diagonal / dense / per-chain metrics and a general integrator;
* the reference's statistical test on that posterior (test_sampling.py:322-379: window adaptation, then sampling,
  mean scale = 1 +- 0.1, mean coefficient = 3 +- 0.1).
"""
import numpy as np
import pytest
import torch

import blackjax_b200 as bj
from blackjax_b200 import plugin, targets as T
from oracle import hmc as ohmc
from oracle import nuts as onuts
from oracle import prng as oprng
from oracle import targets as otargets
from test_gpu_parity import DEV, F, close, npy, tf, tk

pytestmark = pytest.mark.gpu


def user_diag(inv_var_scale, dim, **opts):
    s = np.asarray(inv_var_scale, np.float64)
    inv_var = (1.0 / (s * s)).astype(F)
    return T.UserTarget(dim, plugin.read_example("diag_gaussian"), inv_var, name="diag_gaussian", **opts)


# ---------------------------------------------------------------------------------------------------------------------
# the plug-in route reproduces the built-in target bit for bit
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D, opts", [(100, {}), (18, dict(dense_metric=False, general_integrators=False)),
                                     (70, dict(dense_metric=False, general_integrators=False)),
                                     (256, dict(dense_metric=False, general_integrators=False)),
                                     (512, dict(dense_metric=False, general_integrators=False)),
                                     (1024, dict(dense_metric=False, general_integrators=False))])
def test_user_diag_gaussian_bit_identical_to_builtin(D, opts):
    rs = np.random.default_rng(D)
    s = np.exp(rs.uniform(-1, 1, D))
    builtin, user = T.DiagGaussian(s), user_diag(s, D, **opts)
    C = 67
    q = tf(rs.standard_normal((C, D)))
    imm = tf(np.exp(rs.uniform(-0.5, 0.5, D)))
    keys = tk(oprng.split(oprng.key(D), C))
    out = {}
    for name, tgt in (("builtin", builtin), ("user", user)):
        st = bj.hmc.init(q, tgt)
        h_new, h_info = bj.hmc.build_kernel(full_info=True)(keys, st, tgt, 0.11, imm, 7)
        m_new, m_info = bj.mhmc.build_kernel()(keys, st, tgt, 0.11, imm, 7)
        n_new, n_info = bj.nuts.build_kernel(full_info=True)(keys, st, tgt, 0.13, imm, 6)
        out[name] = [st.logdensity, st.logdensity_grad, h_new.position, h_new.logdensity, h_new.logdensity_grad,
                     h_info.energy, h_info.acceptance_rate, h_info.is_accepted, h_info.proposal.position,
                     m_new.position, m_new.logdensity, m_info.acceptance_rate,
                     n_new.position, n_new.logdensity, n_new.logdensity_grad, n_info.energy, n_info.acceptance_rate,
                     n_info.num_integration_steps, n_info.is_turning, n_info.trajectory_leftmost_state.position]
    torch.cuda.synchronize()
    for a, b in zip(out["builtin"], out["user"]):
        assert torch.equal(a, b)
    assert npy(out["user"][17]).max() > 3   # the trees did grow


def test_user_diag_gaussian_dense_metric_general_integrator_and_native_samplers():
    # the other template axes of the plug-in (small dense metric, coefficient-table integrators) and the native
    # multi-step samplers (bjx_hmc_sample / bjx_nuts_sample: the decoupled persistent kernel) against the built-in target
    D, C = 100, 96
    rs = np.random.default_rng(5)
    s = np.exp(rs.uniform(-1, 1, D))
    builtin, user = T.DiagGaussian(s), user_diag(s, D)
    A = rs.standard_normal((D, D))
    imm_dense = tf(A @ A.T / D + np.eye(D))
    imm = tf(np.exp(rs.uniform(-0.5, 0.5, D)))
    q = tf(rs.standard_normal((C, D)))
    keys = tk(oprng.split(oprng.key(9), C))
    res = {}
    for name, tgt in (("builtin", builtin), ("user", user)):
        st = bj.hmc.init(q, tgt)
        a, _ = bj.hmc.build_kernel()(keys, st, tgt, 0.1, imm_dense, 5)
        b, bi = bj.nuts.build_kernel()(keys, st, tgt, 0.1, imm_dense, 5)
        c, _ = bj.hmc.build_kernel(integrator=bj.mcmc.integrators.mclachlan)(keys, st, tgt, 0.2, imm, 4)
        d, di = bj.nuts.build_kernel(integrator=bj.mcmc.integrators.yoshida)(keys, st, tgt, 0.2, imm, 5)
        k = bj.random.key(3, DEV)
        _, hist_h, _ = bj.sample_hmc_native(k, st, tgt, 0.1, imm, 6, 5)
        fin_n = bj.sample_nuts_native(k, st, tgt, 0.15, imm, 4, max_num_doublings=6)
        res[name] = [a.position, b.position, bi.num_integration_steps, c.position, d.position, di.num_integration_steps,
                     hist_h, fin_n[0].position]
    torch.cuda.synchronize()
    for x, y in zip(res["builtin"], res["user"]):
        assert torch.equal(x, y)


# ---------------------------------------------------------------------------------------------------------------------
# a model that is not built in: the reference's regression posterior, against the oracle
# ---------------------------------------------------------------------------------------------------------------------
def regression_problem(K, N=1000, seed=0):
    """x ~ N(0,1), y = x . beta + N(0,1) (tests/mcmc/test_sampling.py:326-328 with beta = 3 for K = 1)."""
    rs = np.random.default_rng(seed)
    x = rs.standard_normal((N, K)).astype(F)
    beta = np.array([3.0, -1.0, 0.5, 2.0, -2.0, 1.0, 0.25, -0.5, 1.5, -1.5, 0.75, 3.0, -3.0, 0.1, -0.1, 1.0])[:K]
    y = (x @ beta + rs.standard_normal(N)).astype(F)
    return x, y, beta


def regression_start(C, K, beta, rs, spread=0.03):
    q = np.empty((C, 1 + K), F)
    q[:, 0] = spread * rs.standard_normal(C)
    q[:, 1:] = beta + spread * rs.standard_normal((C, K))
    return q


@pytest.mark.parametrize("K", [1, 3, 5, 16])
def test_user_regression_value_and_grad_matches_oracle(K):
    x, y, beta = regression_problem(K)
    tgt, otgt = T.LinearRegression(x, y), otargets.LinearRegression(x, y)
    rs = np.random.default_rng(K)
    q = regression_start(53, K, beta, rs, spread=0.3)
    st = bj.hmc.init(tf(q), tgt)
    lp, g = otgt(q)
    close(npy(st.logdensity), lp, rtol=3e-6, scale=np.max(np.abs(lp)))
    close(npy(st.logdensity_grad), g, rtol=1e-5, scale=np.max(np.abs(g)))
    # and against central differences of the float64 density (the oracle's gradient is hand-derived too)
    h = 1e-4
    for i in range(1 + K):
        e = np.zeros(1 + K)
        e[i] = h
        fd = (otgt.logp64(q.astype(np.float64) + e) - otgt.logp64(q.astype(np.float64) - e)) / (2 * h)
        np.testing.assert_allclose(npy(st.logdensity_grad)[:, i], fd, rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("K, imm_kind, L, algo", [(1, "diag", 12, "hmc"), (1, "dense", 12, "hmc"), (3, "diag", 8, "hmc"),
                                                  (5, "per_chain", 8, "hmc"), (16, "diag", 6, "hmc"),
                                                  (1, "diag", 10, "mhmc"), (5, "diag", 6, "mhmc")])
def test_user_regression_hmc_matches_oracle(K, imm_kind, L, algo):
    from oracle import adaptation as oadapt
    x, y, beta = regression_problem(K)
    tgt, otgt = T.LinearRegression(x, y), otargets.LinearRegression(x, y)
    D, C = 1 + K, 48
    rs = np.random.default_rng(100 + K)
    q = regression_start(C, K, beta, rs)
    # posterior standard deviations are ~ 1/sqrt(N): a metric of that size makes eps = O(1) stable
    if imm_kind == "diag":
        imm = (1e-3 * np.exp(rs.uniform(-0.3, 0.3, D))).astype(F)
    elif imm_kind == "dense":
        A = rs.standard_normal((D, D))
        imm = (1e-3 * (A @ A.T / D + np.eye(D))).astype(F)
    else:
        imm = (1e-3 * np.exp(rs.uniform(-0.3, 0.3, (C, D)))).astype(F)
    ometric = oadapt._PerChainDiag(imm) if imm_kind == "per_chain" else ohmc.Metric(imm)
    keys = oprng.split(oprng.key(K), C)
    eps = F(0.35)
    ostate = ohmc.init(q, otgt)
    state = bj.hmc.init(tf(q), tgt)
    if algo == "hmc":
        onew, oinfo = ohmc.hmc_kernel(keys, ostate, otgt, eps, ometric, L)
        new, info = bj.hmc.build_kernel(full_info=True)(tk(keys), state, tgt, float(eps), tf(imm), L)
        torch.cuda.synchronize()
        close(npy(info.momentum), oinfo.momentum, rtol=3e-6)
        # the log-density is a sum of N = 1000 terms of size ~1: energies carry that sum's rounding
        close(npy(info.energy), oinfo.energy, rtol=1e-5, scale=np.max(np.abs(oinfo.energy)))
        close(npy(info.proposal.position), oinfo.proposal[0])
        # the momentum integrates L gradients, each a 1000-term float32 sum evaluated in another order than numpy's
        # (lane-strided partial sums + shuffle tree vs BLAS / pairwise): measured 3e-5 .. 5e-5 of the row scale
        close(npy(info.proposal.momentum), oinfo.proposal[1], rtol=1e-4)
        u = oprng.uniform(oprng.split(keys, 2)[:, 1])
        acc = npy(info.is_accepted)
        # exp() of an energy difference known to ~1e-5 * |H| ~ 1e-2: decisions may differ inside that band
        tie = np.abs(u - oinfo.acceptance_rate) < 2e-2
        assert ((acc == oinfo.is_accepted) | tie).all()
        assert 0.3 < acc.mean() <= 1.0
        same = acc == oinfo.is_accepted
    else:
        onew, oinfo = ohmc.mhmc_kernel(keys, ostate, otgt, eps, ometric, L)
        new, info = bj.mhmc.build_kernel()(tk(keys), state, tgt, float(eps), tf(imm), L)
        torch.cuda.synchronize()
        same = np.all(np.isclose(npy(new.position), onew.position, rtol=1e-4, atol=1e-6), axis=1)
        assert same.mean() >= 0.85   # progressive draws compare uniforms with exp(energy differences) (see above)
    assert same.mean() > 0.8
    close(npy(new.position)[same], onew.position[same])
    close(npy(new.logdensity)[same], onew.logdensity[same], rtol=1e-5, scale=np.max(np.abs(onew.logdensity)))


@pytest.mark.parametrize("K, imm_kind", [(1, "diag"), (1, "dense"), (3, "diag"), (7, "diag")])
def test_user_regression_nuts_matches_oracle(K, imm_kind):
    x, y, beta = regression_problem(K)
    tgt, otgt = T.LinearRegression(x, y), otargets.LinearRegression(x, y)
    D, C = 1 + K, 64
    rs = np.random.default_rng(200 + K)
    q = regression_start(C, K, beta, rs)
    if imm_kind == "diag":
        imm = (1e-3 * np.exp(rs.uniform(-0.3, 0.3, D))).astype(F)
    else:
        A = rs.standard_normal((D, D))
        imm = (1e-3 * (A @ A.T / D + np.eye(D))).astype(F)
    keys = oprng.split(oprng.key(K + 50), C)
    onew, oinfo = onuts.nuts_kernel(keys, ohmc.init(q, otgt), otgt, F(0.3), imm, 8)
    new, info = bj.nuts.build_kernel(full_info=True)(tk(keys), bj.nuts.init(tf(q), tgt), tgt, 0.3, tf(imm), 8)
    torch.cuda.synchronize()
    close(npy(info.momentum), oinfo.momentum, rtol=3e-6)
    n_dev, n_ref = npy(info.num_integration_steps), oinfo.num_integration_steps
    same = ((n_dev == n_ref) & (npy(info.is_turning) == oinfo.is_turning) & (npy(info.is_divergent) == oinfo.is_divergent)
            & np.all(np.isclose(npy(new.position), onew.position, rtol=1e-4, atol=1e-6), axis=1))
    # the multinomial draws compare uniforms with exp(energy differences) whose rounding is ~1e-5 x |H| ~ 1e-2 here
    # (H is a sum of 1000 O(1) terms), so a few per cent of the chains may take another leaf of the same tree
    assert same.mean() >= 0.8, same.mean()
    shape_same = (n_dev == n_ref) & (npy(info.is_turning) == oinfo.is_turning)
    assert shape_same.mean() >= 0.95, shape_same.mean()
    assert n_dev.mean() > 2.5
    close(npy(new.logdensity)[same], onew.logdensity[same], rtol=1e-5, scale=np.max(np.abs(onew.logdensity)))
    close(npy(info.trajectory_rightmost_state.position)[shape_same], oinfo.trajectory_rightmost_state[0][shape_same],
          rtol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own statistical test on this posterior
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["nuts", "hmc"])
def test_window_adaptation_on_the_reference_regression_posterior(algo):
    """tests/mcmc/test_sampling.py:322-379 (HMC with 90 steps / NUTS, initial position {log_scale: 0, coefs: 4},
    1000 warm-up steps there; 256 chains and 300 + 150 steps here): mean scale 1 +- 0.1, mean coefficient 3 +- 0.1."""
    rs = np.random.default_rng(42)
    x = rs.standard_normal((1000, 1)).astype(F)
    y = (3 * x[:, 0] + rs.standard_normal(1000)).astype(F)
    tgt = T.LinearRegression(x, y)
    C = 256
    q0 = torch.tensor([[0.0, 4.0]], device=DEV).repeat(C, 1).contiguous()
    algorithm = bj.nuts if algo == "nuts" else bj.hmc
    extra = {} if algo == "nuts" else dict(num_integration_steps=30)
    warmup = bj.window_adaptation(algorithm, tgt, **extra)
    (state, params), _ = warmup.run(bj.random.key(1, DEV), q0, 300)
    alg = algorithm(tgt, **params)
    keys = bj.random.split(bj.random.key(2, DEV), 150)
    ssum = torch.zeros(2, dtype=torch.float64, device=DEV)
    for t in range(150):
        state, info = alg.step(keys[t], state)
        ssum += torch.stack([state.position[:, 0].double().exp().mean(), state.position[:, 1].double().mean()])
    scale_mean, coef_mean = (ssum / 150).cpu().numpy()
    np.testing.assert_allclose(scale_mean, 1.0, atol=1e-1)
    np.testing.assert_allclose(coef_mean, 3.0, atol=1e-1)
    # sharper: the posterior mean of the coefficient is the least-squares solution shrunk by the N(0, 5^2) prior
    ols = float(x[:, 0] @ y / (x[:, 0] @ x[:, 0]))
    assert abs(coef_mean - ols) < 0.01


def test_user_target_failure_modes():
    # rows beyond the warp kernels, a dense metric on the tensor-core path: loud errors, no fallback
    with pytest.raises(ValueError):
        T.UserTarget(20000, plugin.read_example("diag_gaussian"), np.ones(20000, F))
    tgt = user_diag(np.ones(256), 256, dense_metric=False, general_integrators=False)
    q = torch.zeros(8, 256, device=DEV)
    st = bj.hmc.init(q, tgt)
    with pytest.raises(bj.BjxError):   # the plug-in was built without the coefficient-table integrators
        bj.hmc.build_kernel(integrator=bj.mcmc.integrators.mclachlan)(bj.random.key(0, DEV), st, tgt, 0.1,
                                                                        torch.ones(256, device=DEV), 3)
    with pytest.raises(bj.BjxError):   # dense metric beyond 128 dims is the tensor-core path: Gaussian targets only
        bj.hmc.build_kernel()(bj.random.key(0, DEV), st, tgt, 0.1, torch.eye(256, device=DEV), 3)


# ---------------------------------------------------------------------------------------------------------------------
# a neighbour-coupled model (every element needs elements other lanes hold): row staging in the warp's scratch, at every
# row size class
# ---------------------------------------------------------------------------------------------------------------------
ROSEN_OPTS = dict(dense_metric=False, general_integrators=False)


@pytest.mark.parametrize("D", [5, 100, 70, 256, 1024])
def test_user_rosenbrock_value_and_grad_matches_oracle(D):
    tgt, otgt = T.Rosenbrock(D, **ROSEN_OPTS), otargets.Rosenbrock(D)
    rs = np.random.default_rng(D)
    q = (0.7 + 0.4 * rs.standard_normal((41, D))).astype(F)
    st = bj.hmc.init(tf(q), tgt)
    lp, g = otgt(q)
    close(npy(st.logdensity), lp, rtol=1e-5, scale=np.max(np.abs(lp)))
    close(npy(st.logdensity_grad), g, rtol=1e-5, scale=np.max(np.abs(g)))


@pytest.mark.parametrize("D, algo", [(256, "hmc"), (70, "hmc"), (100, "nuts"), (5, "nuts"), (1024, "hmc")])
def test_user_rosenbrock_transitions_match_oracle(D, algo):
    tgt, otgt = T.Rosenbrock(D, **ROSEN_OPTS), otargets.Rosenbrock(D)
    rs = np.random.default_rng(300 + D)
    C = 48
    q = (0.7 + 0.3 * rs.standard_normal((C, D))).astype(F)
    imm = np.exp(rs.uniform(-0.2, 0.2, D)).astype(F)
    keys = oprng.split(oprng.key(D), C)
    if algo == "hmc":
        onew, oinfo = ohmc.hmc_kernel(keys, ohmc.init(q, otgt), otgt, F(0.05), imm, 8)
        new, info = bj.hmc.build_kernel(full_info=True)(tk(keys), bj.hmc.init(tf(q), tgt), tgt, 0.05, tf(imm), 8)
        torch.cuda.synchronize()
        close(npy(info.proposal.position), oinfo.proposal[0])
        close(npy(info.proposal.momentum), oinfo.proposal[1], rtol=2e-5)
        close(npy(info.energy), oinfo.energy, rtol=1e-5, scale=np.max(np.abs(oinfo.energy)) + 1)
        u = oprng.uniform(oprng.split(keys, 2)[:, 1])
        acc = npy(info.is_accepted)
        assert ((acc == oinfo.is_accepted) | (np.abs(u - oinfo.acceptance_rate) < 1e-4)).all()
        same = acc == oinfo.is_accepted
        close(npy(new.position)[same], onew.position[same])
    else:
        eps = 0.2 if D >= 100 else 0.35   # trees of ~10-15 leaves (longer ones amplify float32 rounding on this nonlinear model)
        onew, oinfo = onuts.nuts_kernel(keys, ohmc.init(q, otgt), otgt, F(eps), imm, 7)
        new, info = bj.nuts.build_kernel(full_info=True)(tk(keys), bj.nuts.init(tf(q), tgt), tgt, eps, tf(imm), 7)
        torch.cuda.synchronize()
        same = ((npy(info.num_integration_steps) == oinfo.num_integration_steps)
                & (npy(info.is_turning) == oinfo.is_turning)
                & np.all(np.isclose(npy(new.position), onew.position, rtol=1e-4, atol=1e-5), axis=1))
        assert same.mean() >= 0.9, same.mean()
        assert npy(info.num_integration_steps).mean() > 3
        close(npy(new.logdensity)[same], onew.logdensity[same], rtol=1e-5, scale=np.max(np.abs(onew.logdensity)) + 1)


# ---------------------------------------------------------------------------------------------------------------------
# rows beyond a warp (1024 < dim <= 18432): the CTA-level contract bjx_user::BigModel
# ---------------------------------------------------------------------------------------------------------------------
def test_user_big_row_diag_gaussian_bit_identical_to_builtin():
    from blackjax_b200 import _engine
    D, C = 2048, 21
    rs = np.random.default_rng(3)
    s = np.exp(rs.uniform(-1, 1, D))
    builtin = T.DiagGaussian(s)
    user = T.UserTarget(D, plugin.read_example("diag_gaussian_big"), (1.0 / (s * s)).astype(F), name="diag_gaussian_big")
    q = tf(rs.standard_normal((C, D)))
    imm = tf(np.exp(rs.uniform(-0.5, 0.5, D)))
    keys = tk(oprng.split(oprng.key(4), C))
    out = {}
    for name, tgt in (("builtin", builtin), ("user", user)):
        st = bj.hmc.init(q, tgt)
        new, info = bj.hmc.build_kernel(full_info=True)(keys, st, tgt, 0.05, imm, 6)
        eng = _engine.get_engine(q, tgt)
        eng.ensure_metric(imm)
        p = eng.sample_momentum(keys)
        ql, lp, g = q.clone(), st.logdensity.clone(), st.logdensity_grad.clone()
        eng.leapfrog_(ql, p, lp, g, 0.05, 3)
        out[name] = [st.logdensity, st.logdensity_grad, new.position, new.logdensity, info.energy, info.acceptance_rate,
                     info.is_accepted, info.proposal.momentum, ql, p, lp, g]
    torch.cuda.synchronize()
    for a, b in zip(out["builtin"], out["user"]):
        assert torch.equal(a, b)
    st = bj.hmc.init(q, user)
    with pytest.raises(bj.BjxError):     # NUTS is not built for this size class (built-in targets neither)
        bj.nuts.build_kernel()(keys, st, user, 0.05, imm, 3)


def test_user_big_row_hier_logit_matches_builtin_and_oracle():
    from blackjax_b200 import _engine
    D, C = 1504, 10
    G = D - 4
    x, bits = T.HierLogit.synthetic_data(G, seed=1)
    builtin, otgt = T.HierLogit(x, bits), otargets.HierLogit(x, bits)
    theta = np.concatenate([np.asarray([G, 0, 0, 0], F), x.reshape(-1).astype(F), bits.astype(F)])
    user = T.UserTarget(D, plugin.read_example("hier_logit_big"), theta, name="hier_logit_big")
    rs = np.random.default_rng(6)
    q = (0.3 * rs.standard_normal((C, D))).astype(F)
    imm = np.exp(rs.uniform(-0.3, 0.3, D)).astype(F)
    keys = oprng.split(oprng.key(2), C)
    res = {}
    for name, tgt in (("builtin", builtin), ("user", user)):
        eng = _engine.Engine(DEV, C, D, tgt)
        eng.set_metric(tf(imm))
        dq = tf(q)
        lp, g = eng.init_state(dq)
        p = eng.sample_momentum(tk(keys))
        lp0, g0 = lp.clone(), g.clone()
        eng.leapfrog_(dq, p, lp, g, 0.01, 4)
        res[name] = [lp0, g0, dq, p, lp, g]
    torch.cuda.synchronize()
    for a, b in zip(res["builtin"], res["user"]):   # the same one-chain-per-CTA arithmetic and summation order: equal to rounding
        close(npy(a), npy(b), rtol=2e-6, scale=float(b.abs().max()) + 1.0)
    # whole transition through the plug-in's k_big_hmc against the oracle
    onew, oinfo = ohmc.hmc_kernel(keys, ohmc.init(q, otgt), otgt, F(0.01), ohmc.Metric(imm), 6)
    new, info = bj.hmc.build_kernel(full_info=True)(tk(keys), bj.hmc.init(tf(q), user), user, 0.01, tf(imm), 6)
    torch.cuda.synchronize()
    close(npy(info.proposal.position), oinfo.proposal[0], rtol=1e-5)
    close(npy(info.energy), oinfo.energy, rtol=1e-5, scale=np.max(np.abs(oinfo.energy)) + D)
    u = oprng.uniform(oprng.split(keys, 2)[:, 1])
    acc = npy(info.is_accepted)
    assert ((acc == oinfo.is_accepted) | (np.abs(u - oinfo.acceptance_rate) < 1e-3)).all()
    same = acc == oinfo.is_accepted
    close(npy(new.position)[same], onew.position[same], rtol=1e-5)

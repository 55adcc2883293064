"""Pins the CPU oracle against every known-answer test the reference's own suite holds for the
HMC/NUTS path, plus the JAX PRNG known answers.  CPU only."""
import numpy as np
import pytest

from oracle import adaptation, hmc, nuts, prng, targets

F = np.float32


# ---- JAX PRNG known answers (jax 0.10.0 defaults: threefry2x32, partitionable) -------------
def test_prng_split_key0():
    got = prng.split(prng.key(0))
    assert got.tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


def test_prng_fold_in_equals_split_child():
    k = prng.key(1234)
    assert (prng.fold_in(k, 3) == prng.split(k, 5)[3]).all()


def test_prng_normal_known_answers():
    assert prng.normal(prng.key(42)) == F(-0.028304616)
    assert abs(float(prng.normal(prng.key(0))) - 1.622642) < 2e-6


def test_prng_uniform_range_and_bits():
    u = prng.uniform(prng.key(7), (4096,))
    assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()
    b = prng.random_bits(prng.key(7), (8, 4))
    o0, o1 = prng.threefry2x32(0, 7, 0, np.arange(32, dtype=np.uint32))
    assert (b.ravel() == (o0 ^ o1)).all()


def test_erfinv_matches_scipy():
    from scipy.special import erfinv
    x = np.linspace(-0.999999, 0.999999, 20001).astype(F)
    ref = erfinv(x.astype(np.float64))
    got = prng.erfinv_f32(x).astype(np.float64)
    assert np.max(np.abs(got - ref) / np.maximum(1e-3, np.abs(ref))) < 5e-6


# ---- tests/mcmc/test_integrators.py:74-103 golden end state ---------------------------------
COV6 = np.array(
    [[5.9959664, 1.1494889, -1.0420643, -0.6328479, -0.20363973, 2.1600752],
     [1.1494889, 1.3504763, -0.3601517, -0.98311526, 1.1569028, -1.4185406],
     [-1.0420643, -0.3601517, 6.3011055, -2.0662997, -0.10126236, 1.2898219],
     [-0.6328479, -0.98311526, -2.0662997, 4.82699, -2.575554, 2.5724294],
     [-0.20363973, 1.1569028, -0.10126236, -2.575554, 3.35319, -2.9411654],
     [2.1600752, -1.4185406, 1.2898219, 2.5724294, -2.9411654, 6.3740206]])
Q6_INIT = np.array([[0.0, 1.0, 2.0, 3.0, 1.0, 1.0]], F)
P6_INIT = np.array([[0.53288144, 0.25310317, 1.3788314, -0.13486017, -0.59082425, 1.2088736]], F)
Q6_END = np.array([0.38887993, 0.85231394, 2.7879136, 3.0339851, 0.5856687, 1.9291426])
P6_END = np.array([0.46576163, 0.23854092, 1.2518811, -0.35647452, -0.742138, 1.2552949])


def test_velocity_verlet_mvn_golden():
    t = targets.DenseGaussian(np.linalg.inv(COV6))
    m = hmc.Metric(COV6.astype(F))
    lp, g = t(Q6_INIT)
    q1, p1, lp1, _ = hmc.static_integration(t, m, Q6_INIT, P6_INIT, lp, g, F(0.005), 16)
    np.testing.assert_allclose(q1[0], Q6_END, atol=2e-6)
    np.testing.assert_allclose(p1[0], P6_END, atol=2e-6)
    e0 = -lp + m.kinetic_energy(P6_INIT)
    e1 = -lp1 + m.kinetic_energy(p1)
    assert abs(float(e0[0] - e1[0])) < 1e-4          # reference precision for velocity_verlet


@pytest.mark.parametrize("coeffs", [hmc.VELOCITY_VERLET, hmc.MCLACHLAN, hmc.YOSHIDA, hmc.OMELYAN])
def test_integrators_analytic(coeffs):
    # free fall: U = g*x, q(1)=0.5?  reference examples use harmonic oscillator & free fall
    t = targets.StdNormal(1)                         # harmonic oscillator: logp = -x^2/2
    m = hmc.Metric(np.ones(1, F))
    q = np.zeros((1, 1), F)
    p = np.ones((1, 1), F)
    lp, g = t(q)
    q1, p1, _, _ = hmc.static_integration(t, m, q, p, lp, g, F(0.01), 100, coeffs)
    assert abs(float(q1[0, 0]) - np.sin(1.0)) < 1e-2
    assert abs(float(p1[0, 0]) - np.cos(1.0)) < 1e-2


# ---- tests/mcmc/test_metrics.py:124-142,158-179 momentum identities ---------------------------
def test_momentum_identity_diag():
    m = hmc.Metric(np.array([0.25], F))
    p = m.sample_momentum(prng.key(0)[None], 1)
    assert p[0, 0] == F(2.0) * prng.normal(prng.key(0))
    assert m.kinetic_energy(p)[0] == F(0.5) * (F(0.25) * p[0, 0]) * p[0, 0]


def test_momentum_identity_dense():
    imm = np.array([[2 / 3, 0.5], [0.5, 3 / 4]], F)
    m = hmc.Metric(imm)
    L = np.linalg.cholesky(imm.astype(np.float64))
    z = prng.normal(prng.key(0), (2,)).astype(np.float64)
    expected = np.linalg.solve(L.T, z)               # L^-T z
    p = m.sample_momentum(prng.key(0)[None], 2)
    np.testing.assert_allclose(p[0], expected, rtol=2e-6)
    np.testing.assert_allclose(m.kinetic_energy(p)[0], 0.5 * expected @ imm.astype(np.float64) @ expected, rtol=1e-5)


def test_metric_wrong_ndim():
    with pytest.raises(ValueError, match="wrong number of dimensions"):
        hmc.Metric(np.ones((2, 2, 2), F))


# ---- tests/mcmc/test_uturn.py:13-43 ------------------------------------------------------------
@pytest.mark.parametrize("idxs, expected", [((3, 2), False), ((3, 3), True), ((0, 0), False),
                                            ((0, 1), True), ((1, 3), True)])
def test_iterative_uturn_table(idxs, expected):
    m = hmc.Metric(np.ones(1, F))
    ck_p = np.array([1.0, 2.0, 3.0, -2.0], F)[:, None]
    ck_s = np.array([2.0, 4.0, 4.0, -1.0], F)[:, None]
    got = nuts.is_iterative_turning(m, ck_p, ck_s, idxs[0], idxs[1], np.array([3.0], F), np.array([1.0], F))
    assert got == expected


def test_leaf_idx_to_ckpt_idxs():
    # termination.py:77-82 docstring examples: idx_max 6->2, 7->2, 13->2 ; num_subtrees 6->0, 7->3, 13->1
    assert [nuts.leaf_idx_to_ckpt_idxs(n)[1] for n in (6, 7, 13)] == [2, 2, 2]
    assert [nuts.leaf_idx_to_ckpt_idxs(n) for n in (6, 7, 13)] == [(3, 2), (0, 2), (2, 2)]


# ---- tests/mcmc/test_trajectory.py:20-74 sub-tree divergence -----------------------------------
@pytest.mark.parametrize("step_size, should_diverge", [(0.0001, False), (1000, True)])
def test_subtree_divergence(step_size, should_diverge):
    t = targets.NormLogpdf(1)
    m = hmc.Metric(np.ones(1, F))
    k = prng.key(0)[None]
    q = np.ones((1, 1), F)
    p = m.sample_momentum(k, 1)
    lp, g = t(q)
    h0 = -lp + m.kinetic_energy(p)
    out = nuts.subtree(t, m, k, q, p, lp, g, np.array([1], np.int32), np.zeros((1, 10, 1), F),
                       np.zeros((1, 10, 1), F), 100, F(step_size), h0, np.array([True]))
    assert bool(out["is_div"][0]) is should_diverge


# ---- tests/mcmc/test_trajectory.py:193-260 tree-doubling outcomes --------------------------------
@pytest.mark.parametrize("step_size, diverge, turn, doublings",
                         [(1e-10, False, False, 10), (1.0, False, True, 2), (1e5, True, True, 1)])
def test_dynamic_expansion_outcomes(step_size, diverge, turn, doublings):
    t = targets.StdNormal(1)
    m = hmc.Metric(np.ones(1, F))
    k = prng.key(0)[None]
    p0 = m.sample_momentum(k, 1)
    q0 = np.zeros((1, 1), F)
    lp, g = t(q0)
    _, info = nuts.nuts_kernel(None, (q0, lp, g), t, F(step_size), m, 10, momentum=p0, key_integrator=k)
    assert bool(info.is_divergent[0]) == diverge
    assert bool(info.is_turning[0]) == turn
    assert int(info.num_trajectory_expansions[0]) == doublings


# ---- tests/mcmc/test_trajectory.py:76-191 progressive == recursive ------------------------------
def test_progressive_equals_recursive():
    t = targets.Banana()
    m = hmc.Metric(np.array([[1.0, 0.5], [0.5, 1.25]], F))
    rng_key = prng.key(23133)
    rs = np.random.default_rng(5)
    for i in range(50):
        sub = prng.fold_in(rng_key, i)
        k6 = prng.split(sub, 6)
        direction = int(rs.choice([-1, 1]))
        depth = int(rs.integers(2, 5))
        q = prng.normal(k6[4], (2,))[None]
        p = prng.normal(k6[5], (2,))[None]
        eps = F(abs(float(prng.normal(k6[3]))) * 0.1)
        lp, g = t(q)
        h0 = -lp + m.kinetic_energy(p)
        out = nuts.subtree(t, m, k6[0][None], q, p, lp, g, np.array([direction], np.int32),
                           np.zeros((1, depth, 2), F), np.zeros((1, depth, 2), F), 2 ** depth, eps, h0,
                           np.array([True]))
        _, prop1, tr1, div1, turn1 = nuts.recursive_subtree(t, m, k6[0], (q[0], p[0], lp[0], g[0]),
                                                            direction, depth, eps, h0[0])
        assert bool(out["is_div"][0]) == div1
        assert bool(out["has_term"][0]) == turn1
        left0, right0 = (out["first"], out["last"]) if direction > 0 else (out["last"], out["first"])
        for a, b in zip(left0, tr1["left"]):
            np.testing.assert_allclose(np.asarray(a)[0], b, rtol=1e-5, atol=1e-6)
        for a, b in zip(right0, tr1["right"]):
            np.testing.assert_allclose(np.asarray(a)[0], b, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out["p_sum"][0], tr1["p_sum"], rtol=1e-5, atol=1e-6)
        assert int(out["n"][0]) == tr1["n"]
        np.testing.assert_allclose(out["prop"]["weight"][0], prop1["weight"], rtol=1e-5)
        np.testing.assert_allclose(out["prop"]["slpa"][0], prop1["slpa"], rtol=1e-5)


# ---- tests/adaptation/test_adaptation.py:27-49 ----------------------------------------------------
@pytest.mark.parametrize("num_steps, expected", [
    (19, [(0, False)] * 19),
    (100, [(0, False)] * 15 + [(1, False)] * 74 + [(1, True)] + [(0, False)] * 10),
    (200, [(0, False)] * 75 + [(1, False)] * 24 + [(1, True)] + [(1, False)] * 49 + [(1, True)]
     + [(0, False)] * 50)])
def test_adaptation_schedule(num_steps, expected):
    s = adaptation.build_schedule(num_steps)
    assert len(s) == num_steps and s == expected


def test_welford_recovers_covariance():
    # tests/adaptation/test_mass_matrix.py:14-44 (rtol 1e-1)
    rs = np.random.default_rng(0)
    cov = np.array([[1.0, 0.3], [0.3, 2.0]])
    x = rs.multivariate_normal([0, 0], cov, 5000).astype(F)
    w = adaptation.welford_init(2, diagonal=False)
    wd = adaptation.welford_init(2, diagonal=True)
    for xi in x:
        w = adaptation.welford_update(w, xi)
        wd = adaptation.welford_update(wd, xi)
    np.testing.assert_allclose(w.m2 / (w.n - 1), cov, rtol=1e-1, atol=5e-2)
    np.testing.assert_allclose(wd.m2 / (wd.n - 1), np.diag(cov), rtol=1e-1)
    # CGL merge of two halves equals the sequential estimate
    a = adaptation.welford_init(2, True)
    b = adaptation.welford_init(2, True)
    for xi in x[:2000]:
        a = adaptation.welford_update(a, xi)
    for xi in x[2000:]:
        b = adaptation.welford_update(b, xi)
    ab = adaptation.cgl_merge(a, b)
    np.testing.assert_allclose(ab.m2, wd.m2, rtol=1e-3)
    np.testing.assert_allclose(ab.mean, wd.mean, atol=1e-4)


def test_dual_averaging_mean_pool_step_counter():
    # tests/adaptation/test_meta_builders_e2e.py:1443-1700 semantics: one update per warm-up step
    s = adaptation.da_init(1.0)
    s2 = adaptation.da_update(s, 0.5)
    assert s2.step == 2 and s.step == 1
    assert s2.log_step_size_avg == F(1.0) * s.log_step_size      # eta_1 = 1 -> pre-update iterate
    assert abs(float(s.mu) - np.log(10.0)) < 1e-6


# ---- statistical end-to-end (tests/mcmc/test_sampling.py:1055-1119,1174-1186), 10% tolerance -------
def test_univariate_normal_hmc_and_nuts():
    t = targets.DiagGaussian(2.0, 1)

    class Shifted:
        dim = 1

        def __call__(self, q):
            lp, g = t(q - F(1.0))
            return lp, g

    tgt = Shifted()
    C = 64
    key = prng.key(12)
    q = np.ones((C, 1), F)
    for name, kern, kw in (("hmc", hmc.hmc_kernel, dict(num_integration_steps=30, step_size=F(3.9))),
                           ("nuts", nuts.nuts_kernel, dict(step_size=F(1.0)))):
        st = hmc.init(q, tgt)
        draws = []
        keys = prng.split(key, 400)
        for i in range(400):
            ck = prng.split(keys[i], C)
            if name == "hmc":
                st, _ = kern(ck, st, tgt, kw["step_size"], np.ones(1, F), kw["num_integration_steps"])
            else:
                st, _ = kern(ck, st, tgt, kw["step_size"], np.ones(1, F))
            if i >= 100:
                draws.append(st.position[:, 0].copy())
        d = np.concatenate(draws)
        assert abs(d.mean() - 1.0) < 0.1, name
        assert abs(d.var() - 4.0) < 0.4, name


# ---- tests/mcmc/test_multinomial_hmc.py:36-80 (statistical + diagnostics) --------------------------------------
def test_multinomial_hmc_oracle():
    t = targets.StdNormal(1)
    C = 128
    st = hmc.init(np.zeros((C, 1), F), t)
    draws = []
    keys = prng.split(prng.key(0), 120)
    for i in range(120):
        st, info = hmc.mhmc_kernel(prng.split(keys[i], C), st, t, F(0.5), np.ones(1, F), 20)
        assert info.is_accepted.all()
        if i >= 20:
            draws.append(st.position[:, 0].copy())
    d = np.concatenate(draws)
    assert abs(d.mean()) < 0.3 and abs(d.std() - 1.0) < 0.3                      # :54-55
    _, info = hmc.mhmc_kernel(prng.split(prng.key(1), 4), hmc.init(np.ones((4, 1), F), t), t, F(1000.0), np.ones(1, F), 100)
    assert info.is_divergent.all()                                               # :57-68
    _, info = hmc.mhmc_kernel(prng.split(prng.key(2), 4), hmc.init(np.zeros((4, 1), F), t), t, F(0.1), np.ones(1, F), 10)
    assert (info.acceptance_rate > 0.5).all()                                    # :70-80


# ---- committed golden fixtures (tests/golden/, generated by the oracle: see make_golden.py) -----------------------
def test_oracle_reproduces_committed_golden():
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fresh = mg.cases()
    stored = np.load(os.path.join(here, "golden", "hmc_nuts_golden.npz"))
    assert sorted(fresh) == sorted(stored.files)
    for k in stored.files:
        a, b = fresh[k], stored[k]
        if a.dtype.kind in "iub":
            assert np.array_equal(a, b), k
        else:
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6, err_msg=k)   # BLAS/libm may differ across hosts


def test_hier_logit_oracle_gradient_matches_finite_differences():
    # BASELINE config 5 target (builder-defined): the hand-derived gradient against central differences in float64
    from blackjax_b200.targets import HierLogit
    x, bits = HierLogit.synthetic_data(12, seed=1)
    t = targets.HierLogit(x, bits)
    y = ((bits[:, None] >> np.arange(8, dtype=np.uint8)) & 1).astype(np.float64)

    def logp64(q):
        mu, lt, b0, b1, a = q[0], q[1], q[2], q[3], q[4:]
        eta = a[:, None] + b0 * x[:, :, 0] + b1 * x[:, :, 1]
        ll = np.sum(y * eta - np.logaddexp(0.0, eta))
        return (-0.005 * mu ** 2 - 0.5 * lt ** 2 - 0.08 * (b0 ** 2 + b1 ** 2)
                + np.sum(-0.5 * (a - mu) ** 2 * np.exp(-2 * lt) - lt) + ll)

    q = np.random.default_rng(0).standard_normal(16) * 0.3
    lp, g = t(q[None].astype(F))
    fd = np.array([(logp64(q + 1e-6 * e) - logp64(q - 1e-6 * e)) / 2e-6 for e in np.eye(16)])
    np.testing.assert_allclose(g[0], fd, rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(lp[0], logp64(q.astype(F).astype(np.float64)), rtol=1e-5)


# ---- tests/test_diagnostics.py:54-124 (effective sample size) -----------------------------------------------------
@pytest.mark.parametrize("num_chains", [1, 2, 10])
def test_ess_iid_draws_close_to_total(num_chains):
    from oracle import diagnostics as odiag
    x = np.random.default_rng(32).standard_normal((num_chains, 5000, 3))
    ess = odiag.effective_sample_size(x)
    assert ess.shape == (3,)
    np.testing.assert_allclose(ess, num_chains * 5000, rtol=0.1)          # the reference allows rtol=10
    # axis handling (tests/test_diagnostics.py:300-312)
    np.testing.assert_allclose(odiag.effective_sample_size(np.moveaxis(x, 0, 1), chain_axis=1, sample_axis=0), ess)


@pytest.mark.parametrize("num_chains", [1, 2])
def test_ess_zero_for_numerically_degenerate_chains(num_chains):
    from oracle import diagnostics as odiag
    T_ = 2000
    r = np.random.default_rng(32).standard_normal((num_chains, T_))
    samples = np.stack([np.zeros((num_chains, T_)), np.broadcast_to(np.arange(num_chains)[:, None], (num_chains, T_)),
                        1e-30 * r, r], axis=-1)
    ess = odiag.effective_sample_size(samples)
    np.testing.assert_array_equal(ess[:3], np.zeros(3))
    assert ess[3] > 0


def test_ess_antithetic_chain_exceeds_draw_count_and_ar1_matches_theory():
    from oracle import diagnostics as odiag
    assert odiag.effective_sample_size(np.tile(np.array([-1.0, 1.0]), 1000)[None, :]) > 2000
    rs = np.random.default_rng(3)
    phi, C, T_ = 0.9, 8, 4000
    e = rs.standard_normal((C, T_))
    x = np.zeros((C, T_))
    for t in range(1, T_):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    np.testing.assert_allclose(odiag.effective_sample_size(x), C * T_ * (1 - phi) / (1 + phi), rtol=0.15)


def test_randint_range_and_uniformity():
    # jax.random.randint contract used by dynamic_hmc.py:66: int32 in [minval, maxval), (near) uniform
    from oracle import prng as oprng
    k = oprng.split(oprng.key(3), 20000)
    r = oprng.randint(k, (), 1, 10)
    assert r.dtype == np.int32 and r.min() == 1 and r.max() == 9
    np.testing.assert_allclose(np.bincount(r)[1:] / len(r), 1 / 9, atol=0.01)
    assert np.all(oprng.randint(k[:10], (4,), 5, 5) == 5)          # maxval <= minval returns minval
    big = oprng.randint(k[:100], (3,), -(1 << 30), 1 << 30)
    assert big.min() < -(1 << 28) and big.max() > (1 << 28)


# ---- size-independent properties of the restated integrators (docs: integrators.py:62-152 are symplectic,
#      time-reversible maps; tests/mcmc/test_integrators.py checks energy conservation of the same schemes) ----------
@pytest.mark.parametrize("name", ["VELOCITY_VERLET", "MCLACHLAN", "YOSHIDA", "OMELYAN"])
def test_integrators_are_time_reversible_and_conserve_energy(name):
    from oracle import hmc as ohmc, targets as otargets
    rs = np.random.default_rng(4)
    D, C, L = 40, 16, 25
    t = otargets.DiagGaussian(np.exp(rs.uniform(-0.5, 0.5, D)))
    metric = ohmc.Metric(np.exp(rs.uniform(-0.3, 0.3, D)).astype(np.float32))
    q0 = rs.standard_normal((C, D)).astype(np.float32)
    p0 = rs.standard_normal((C, D)).astype(np.float32)
    logp0, g0 = t(q0)
    eps = np.float32(0.05)
    coef = getattr(ohmc, name)
    q1, p1, logp1, g1 = ohmc.static_integration(t, metric, q0, p0, logp0, g0, eps, L, coef)
    # flip the momentum, integrate back, flip again: the start state comes back up to float32 rounding
    q2, p2, _, _ = ohmc.static_integration(t, metric, q1, -p1, logp1, g1, eps, L, coef)
    np.testing.assert_allclose(q2, q0, rtol=0, atol=2e-4)
    np.testing.assert_allclose(-p2, p0, rtol=0, atol=2e-4)
    e0 = -logp0 + metric.kinetic_energy(p0)
    e1 = -logp1 + metric.kinetic_energy(p1)
    tol = 2e-2 if name == "VELOCITY_VERLET" else 5e-3          # second order vs the higher-order / tuned schemes
    assert np.max(np.abs(e1 - e0)) < tol


def test_chees_oracle_reference_test_problem():
    """oracle/chees.py on the reference's own ChEES test problem (tests/adaptation/test_adaptation.py:77-140: 2-D normal,
    std (1, 10), step 0.1, adam(0.5, b1=0, b2=0.95), target acceptance 0.75): the Halton sequence is the base-2 radical
    inverse, the adapted trajectory covers the wide direction and jittered HMC at the adapted parameters accepts near the
    target (the reference asserts the harmonic mean within 0.1 ... it has no stored vectors for this path)."""
    from oracle import chees as ochees
    from oracle import hmc as ohmc_
    assert [float(ochees.halton(i, 11)) for i in range(7)] == [0.5, 0.25, 0.75, 0.125, 0.625, 0.375, 0.875]
    tgt = targets.DiagGaussian(np.array([1.0, 10.0]))
    C = 64
    q = np.random.default_rng(0).standard_normal((C, 2)).astype(np.float32)
    st, eps, nlf, s = ochees.chees_run(tgt, prng.key(346), q, 0.1, lr=0.5, b1=0.0, b2=0.95, num_steps=400,
                                       target_acceptance_rate=0.75)
    assert 0.5 < float(eps) < 2.5 and 8.0 < float(eps) * float(nlf) < 40.0      # ~ a quarter period of the std-10 direction
    inv = []
    keys = prng.split(prng.key(7), 60)
    for t in range(60):
        L = ochees.integration_steps(400 + t, nlf, 1.0, 11)
        st, info = ohmc_.hmc_kernel(prng.split(keys[t], C), st, tgt, eps, np.ones(2, np.float32), L)
        inv.append(np.mean(1.0 / info.acceptance_rate))
    assert abs(1.0 / np.mean(inv) - 0.75) < 0.12


def test_low_rank_metric_oracle_matches_the_dense_formula():
    """oracle LowRankMetric (metrics.py:349-467) against the explicit matrix M^-1 = D (I + U (Lambda - I) U^T) D: velocity,
    kinetic energy, momentum covariance M, and the lam = 1 reduction to a diagonal metric with scale sigma (the reference's
    own statement in the docstring :368-369)."""
    rs = np.random.default_rng(0)
    D, k, C = 24, 4, 16
    U, _ = np.linalg.qr(rs.standard_normal((D, k)))
    sigma = np.exp(rs.uniform(-1, 1, D))
    lam = np.array([9.0, 0.25, 4.0, 1.5])
    m = hmc.LowRankMetric(sigma, U, lam)
    Minv = np.diag(sigma) @ (np.eye(D) + U @ np.diag(lam - 1) @ U.T) @ np.diag(sigma)
    p = rs.standard_normal((C, D)).astype(np.float32)
    np.testing.assert_allclose(m.velocity(p), p @ Minv.T, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(m.kinetic_energy(p), 0.5 * np.einsum("ci,ij,cj->c", p, Minv, p), rtol=2e-5)
    z = m.sample_momentum(prng.split(prng.key(0), 40000), D)
    M = np.linalg.inv(Minv)
    assert np.abs(np.cov(z.T) - M).max() < 0.05 * np.abs(M).max()
    ones = hmc.LowRankMetric(sigma, U, np.ones(k))
    diag = hmc.Metric((sigma ** 2).astype(np.float32))
    np.testing.assert_allclose(ones.velocity(p), diag.velocity(p), rtol=1e-5)
    assert (ones.is_turning(p, -p, p) == diag.is_turning(p, -p, p)).all()


def test_ghmc_oracle_samples_the_reference_univariate_case():
    """tests/mcmc/test_sampling.py:1160-1172 of the reference on the restatement: ghmc(step_size=1, scale 1, alpha=0.8,
    delta=2) on N(1, 2^2); mean and std to the reference's 1e-1 relative tolerance (pooled over 64 chains)."""
    from oracle import ghmc as oghmc
    C = 64
    tgt = targets.DiagGaussian(np.array([2.0]), mean=np.array([1.0]))
    st = oghmc.init(np.ones((C, 1), np.float32), tgt, prng.split(prng.key(3), C))
    assert (np.abs(st.slice) <= 1).all() and st.momentum.shape == (C, 1)
    keys = prng.split(prng.key(4), 1500)
    draws = []
    for t in range(1500):
        st, info = oghmc.ghmc_kernel(prng.split(keys[t], C), st, tgt, 1.0, np.ones(1, np.float32), 0.8, 2.0)
        assert (np.abs(st.slice) <= 1).all()
        if t >= 300:
            draws.append(st.position)
    x = np.concatenate(draws).ravel()
    np.testing.assert_allclose(x.mean(), 1.0, rtol=1e-1)
    np.testing.assert_allclose(x.std(), 2.0, rtol=1e-1)


def test_meads_oracle_properties_of_the_reference_tests():
    """The checkable statements of tests/adaptation/test_meads.py and tests/mcmc/test_sampling.py:606-690 replayed on the
    restatement: init replicates one parameter set over the folds (:44-56), the step size is linear in the multiplier
    (:85-97), damping_slowdown raises alpha early on (:99-117), the fold t mod K is frozen at step t (sampling:640-662),
    the maximum-eigenvalue estimator recovers a dominant eigenvalue, and the shuffle is a permutation."""
    from oracle import meads as om
    rs = np.random.default_rng(0)
    C, D, K = 64, 3, 4
    q = rs.standard_normal((C, D)).astype(np.float32)
    tgt = targets.DiagGaussian(np.array([1.0, 2.0, 0.5]))
    _, g = tgt(q)
    s1 = om.meads_init(q, g, K)
    assert (s1.step_size == s1.step_size[0]).all() and (s1.alpha == s1.alpha[0]).all()
    assert (s1.position_sigma == s1.position_sigma[0]).all() and s1.current_iteration == 0
    far = 10.0 * q                                           # gradients large enough that min(., 1) does not clip
    _, gf = tgt(far)
    a = om.compute_parameters(far, gf, 0, step_size_multiplier=0.25)[0]
    b = om.compute_parameters(far, gf, 0, step_size_multiplier=0.5)[0]
    np.testing.assert_allclose(b, 2 * a, rtol=1e-5)
    lo = om.compute_parameters(q, g, 0, damping_slowdown=1.0)[2]
    hi = om.compute_parameters(q, g, 0, damping_slowdown=5.0)[2]
    assert hi >= lo
    X = (rs.standard_normal((400, 6)) * np.array([10, 1, 1, 1, 1, 1])).astype(np.float32)
    assert abs(om.maximum_eigenvalue(X) / 100 - 1) < 0.2              # the dominant eigenvalue of the second-moment matrix
    for n in (1, 7, 2000):
        np.testing.assert_array_equal(np.sort(om.permutation(prng.key(n), n)), np.arange(n))
    trace = []
    om.meads_run(tgt, prng.key(2), q, 3, num_folds=K, trace=trace)
    n = C // K
    np.testing.assert_array_equal(trace[0][0].position[:n], q[:n])
    np.testing.assert_array_equal(trace[1][0].position[n:2 * n], trace[0][0].position[n:2 * n])
    np.testing.assert_array_equal(trace[2][0].position[2 * n:3 * n], trace[1][0].position[2 * n:3 * n])
    assert not np.array_equal(trace[0][0].position[n:], q[n:])
    assert trace[2][1].step_size.shape == (K,) and (trace[2][1].step_size > 0).all()


def test_linear_regression_oracle_gradient_matches_float64_differences():
    """oracle/targets.py LinearRegression (tests/mcmc/test_sampling.py:103-111 regression_logprob): the hand-derived
    float32 value_and_grad against the float64 restatement of the reference's expression and its central differences."""
    from oracle import targets as otargets
    rs = np.random.default_rng(3)
    x = rs.standard_normal((300, 4)).astype(np.float32)
    y = (x @ np.array([3.0, -1.0, 0.5, 2.0]) + rs.standard_normal(300)).astype(np.float32)
    t = otargets.LinearRegression(x, y)
    q = np.concatenate([0.3 * rs.standard_normal((9, 1)), np.array([3.0, -1.0, 0.5, 2.0]) + 0.2 * rs.standard_normal((9, 4))],
                       axis=1).astype(np.float32)
    lp, g = t(q)
    np.testing.assert_allclose(lp, t.logp64(q), rtol=2e-6)
    h = 1e-5
    for i in range(5):
        e = np.zeros(5)
        e[i] = h
        fd = (t.logp64(q.astype(np.float64) + e) - t.logp64(q.astype(np.float64) - e)) / (2 * h)
        np.testing.assert_allclose(g[:, i], fd, rtol=1e-4, atol=1e-3)
    # the reference's own call shape: x [N, 1], scalar coefficient
    t1 = otargets.LinearRegression(x[:, :1], y)
    assert t1.dim == 2 and t1(np.array([[0.0, 4.0]], np.float32))[1].shape == (1, 2)

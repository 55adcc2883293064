"""The C/pthreads oracle twin (CPU timing stand-in) agrees with the numpy oracle."""
import numpy as np
import pytest

from oracle import cport, hmc, prng, targets

F = np.float32


@pytest.mark.parametrize("kind, D", [("diag", 100), ("diag", 37), ("funnel", 16)])
def test_c_twin_matches_numpy_oracle(kind, D):
    rs = np.random.default_rng(0)
    C, L, eps = 64, 10, F(0.1)
    if kind == "diag":
        t = targets.DiagGaussian(np.exp(rs.uniform(-1, 1, D)))
        inv_var, k = t.inv_var, 0
    else:
        t = targets.Funnel(D)
        inv_var, k = np.ones(D, F), 1
    imm = np.exp(rs.uniform(-0.5, 0.5, D)).astype(F)
    q = (0.5 * rs.standard_normal((C, D))).astype(F)
    keys = prng.split(prng.key(3), C)
    st = hmc.init(q, t)
    new, info = hmc.hmc_kernel(keys, st, t, eps, imm, L)
    qc, lc, gc = q.copy(), st.logdensity.copy(), st.logdensity_grad.copy()
    acc, ok = cport.hmc_step(k, inv_var, imm, keys, qc, lc, gc, eps, L)
    np.testing.assert_allclose(acc, info.acceptance_rate, rtol=2e-4, atol=2e-5)
    same = ok == info.is_accepted
    assert same.mean() > 0.98
    np.testing.assert_allclose(qc[same], new.position[same], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gc[same], new.logdensity_grad[same], rtol=1e-5, atol=1e-5)


def test_c_twin_dense_matches_numpy_oracle():
    D, C, L, eps = 48, 40, 6, F(0.1)
    cov, prec = targets.correlated_gaussian(D, seed=2, lo=-0.5, hi=0.5)
    t = targets.DenseGaussian(prec)
    rs = np.random.default_rng(1)
    q = (0.5 * rs.standard_normal((C, D))).astype(F)
    keys = prng.split(prng.key(4), C)
    st = hmc.init(q, t)
    new, info = hmc.hmc_kernel(keys, st, t, eps, cov, L)
    qc, lc, gc = q.copy(), st.logdensity.copy(), st.logdensity_grad.copy()
    acc, ok = cport.hmc_dense_step(prec, cov, keys, qc, lc, gc, eps, L)
    np.testing.assert_allclose(acc, info.acceptance_rate, rtol=1e-3, atol=1e-4)
    same = ok == info.is_accepted
    assert same.mean() > 0.95
    np.testing.assert_allclose(qc[same], new.position[same], rtol=1e-4, atol=1e-5)


def test_c_twin_hier_logit_matches_numpy_oracle():
    """BASELINE config 5's target in the C twin (exact expf / log1pf) against oracle/targets.py HierLogit."""
    from blackjax_b200.targets import HierLogit
    G, C, L, eps = 60, 24, 6, F(0.01)
    D = 4 + G
    x, bits = HierLogit.synthetic_data(G, seed=1)
    t = targets.HierLogit(x, bits)
    rs = np.random.default_rng(2)
    q = np.empty((C, D), F)
    q[:, :4] = [0.5, np.log(0.7), 1.0, -0.5]
    q[:, 4:] = 0.5 + 0.7 * rs.standard_normal((C, G))
    imm = np.exp(rs.uniform(-0.3, 0.3, D)).astype(F)
    keys = prng.split(prng.key(5), C)
    st = hmc.init(q, t)
    new, info = hmc.hmc_kernel(keys, st, t, eps, imm, L)
    qc, lc, gc = q.copy(), st.logdensity.copy(), st.logdensity_grad.copy()
    acc, ok = cport.hmc_hier_step(x, bits, imm, keys, qc, lc, gc, eps, L)
    np.testing.assert_allclose(acc, info.acceptance_rate, rtol=2e-3, atol=2e-4)
    same = ok == info.is_accepted
    assert same.mean() > 0.9
    np.testing.assert_allclose(qc[same], new.position[same], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gc[same], new.logdensity_grad[same], rtol=1e-4, atol=1e-4)

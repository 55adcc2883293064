"""Generates tests/golden/hmc_nuts_golden.npz: small seeded (state, key) -> transition outputs.

Provenance: produced by the restated CPU oracle (oracle/), NOT by live JAX -- the reference cannot be imported in
this image (jax 0.10.0 is not installable).  The fixtures pin the oracle against drift and give
the GPU parity tests a stored target; the oracle itself is pinned on the reference's known-answer tests
(tests/test_oracle_kat.py).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hmc, nuts, prng, targets  # noqa: E402

F = np.float32


def cases():
    rs = np.random.default_rng(2026)
    out = {}
    # case A: HMC, BASELINE config 1 shape per chain (100-D isotropic Gaussian, diag mass, L=10, eps=0.2), 8 chains
    C, D = 8, 100
    q = rs.standard_normal((C, D)).astype(F)
    keys = prng.split(prng.key(1), C)
    t = targets.StdNormal(D)
    new, info = hmc.hmc_kernel(keys, hmc.init(q, t), t, F(0.2), np.ones(D, F), 10)
    out.update(A_q=q, A_keys=keys, A_pos=new.position, A_logp=new.logdensity, A_acc=info.acceptance_rate,
               A_accepted=info.is_accepted, A_energy=info.energy, A_momentum=info.momentum)
    # case B: NUTS, Neal's funnel D=16, eps=0.2, 8 chains
    C, D = 8, 16
    q = (0.1 * rs.standard_normal((C, D))).astype(F)
    keys = prng.split(prng.key(2), C)
    t = targets.Funnel(D)
    new, info = nuts.nuts_kernel(keys, hmc.init(q, t), t, F(0.2), np.ones(D, F), 8)
    out.update(B_q=q, B_keys=keys, B_pos=new.position, B_n=info.num_integration_steps,
               B_depth=info.num_trajectory_expansions, B_turn=info.is_turning, B_div=info.is_divergent,
               B_acc=info.acceptance_rate, B_energy=info.energy)
    # case C: multinomial HMC, diagonal Gaussian D=12 with a diagonal metric, L=7
    C, D = 8, 12
    s = np.exp(rs.uniform(-0.5, 0.5, D))
    imm = np.exp(rs.uniform(-0.5, 0.5, D)).astype(F)
    q = rs.standard_normal((C, D)).astype(F)
    keys = prng.split(prng.key(3), C)
    t = targets.DiagGaussian(s)
    new, info = hmc.mhmc_kernel(keys, hmc.init(q, t), t, F(0.15), imm, 7)
    out.update(C_q=q, C_keys=keys, C_scale=s, C_imm=imm, C_pos=new.position, C_acc=info.acceptance_rate,
               C_energy=info.energy)
    # case D: NUTS with a dense metric on the banana (tests/mcmc/test_trajectory.py:79-84 setting)
    C = 8
    imm = np.array([[1.0, 0.5], [0.5, 1.25]], F)
    q = rs.standard_normal((C, 2)).astype(F)
    keys = prng.split(prng.key(4), C)
    t = targets.Banana()
    new, info = nuts.nuts_kernel(keys, hmc.init(q, t), t, F(0.1), imm, 6)
    out.update(D_q=q, D_keys=keys, D_imm=imm, D_pos=new.position, D_n=info.num_integration_steps, D_acc=info.acceptance_rate)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hmc_nuts_golden.npz")
    np.savez_compressed(path, **cases())
    print("wrote", path, os.path.getsize(path), "bytes")

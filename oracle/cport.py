"""ctypes wrapper of the C/pthreads oracle twin (oracle/c/oracle_hmc.c).  TEST INFRASTRUCTURE:
CPU timing stand-in for bench.py; cross-checked against the numpy oracle in tests/test_oracle_c.py."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "liboracle_hmc.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        l = C.CDLL(_SO)
        l.oracle_hmc_step.restype = C.c_longlong
        l.oracle_hmc_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        l.oracle_hmc_dense_step.restype = C.c_longlong
        l.oracle_hmc_dense_step.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        l.oracle_hmc_step_hier.restype = C.c_longlong
        l.oracle_hmc_step_hier.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        l.oracle_num_threads.restype = C.c_int
        _lib = l
    return _lib


def num_threads():
    return lib().oracle_num_threads()


def hmc_step(kind, inv_var, imm, keys, q, logp, g, eps, L, n_threads=0):
    """In-place HMC transition on float32 arrays q[C,D], logp[C], g[C,D]; returns (acc_rate, accepted)."""
    Cn, D = q.shape
    for a in (q, logp, g):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    inv_var = np.ascontiguousarray(inv_var, np.float32)
    imm = np.ascontiguousarray(imm, np.float32)
    keys = np.ascontiguousarray(keys, np.uint32)
    acc = np.empty(Cn, np.float32)
    ok = np.empty(Cn, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().oracle_hmc_step(Cn, D, int(kind), p(inv_var), p(imm), p(keys), p(q), p(logp), p(g), float(eps), int(L),
                          p(acc), p(ok), int(n_threads))
    return acc, ok.astype(bool)


def hmc_hier_step(covariates, outcome_bits, imm, keys, q, logp, g, eps, L, n_threads=0):
    """In-place HMC transition on the hierarchical logistic regression (oracle/targets.py HierLogit; D = 4 + G)."""
    Cn, D = q.shape
    for a in (q, logp, g):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    x = np.ascontiguousarray(covariates, np.float32)
    y = np.ascontiguousarray(outcome_bits, np.uint8)
    assert x.shape == (D - 4, 8, 2) and y.shape == (D - 4,)
    imm = np.ascontiguousarray(imm, np.float32)
    keys = np.ascontiguousarray(keys, np.uint32)
    acc = np.empty(Cn, np.float32)
    ok = np.empty(Cn, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().oracle_hmc_step_hier(Cn, D, p(x), p(y), p(imm), p(keys), p(q), p(logp), p(g), float(eps), int(L), p(acc), p(ok),
                               int(n_threads))
    return acc, ok.astype(bool)


def dense_mass_sqrt(imm):
    """L^-T with L = chol(M^-1)  (metrics.py:712-715), float64 then cast."""
    Lc = np.linalg.cholesky(np.asarray(imm, np.float64))
    return np.ascontiguousarray(np.linalg.solve(Lc.T, np.eye(Lc.shape[0])), np.float32)


def hmc_dense_step(prec, imm, keys, q, logp, g, eps, L, n_threads=0, msqrt=None):
    """In-place HMC transition, dense Gaussian target (precision ``prec``) and dense inverse mass matrix ``imm``."""
    Cn, D = q.shape
    for a in (q, logp, g):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    prec = np.ascontiguousarray(prec, np.float32)
    imm = np.ascontiguousarray(imm, np.float32)
    msqrt = dense_mass_sqrt(imm) if msqrt is None else np.ascontiguousarray(msqrt, np.float32)
    keys = np.ascontiguousarray(keys, np.uint32)
    acc = np.empty(Cn, np.float32)
    ok = np.empty(Cn, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().oracle_hmc_dense_step(Cn, D, p(prec), p(imm), p(msqrt), p(keys), p(q), p(logp), p(g), float(eps), int(L),
                                p(acc), p(ok), int(n_threads))
    return acc, ok.astype(bool)

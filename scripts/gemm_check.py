"""Diagnostics for the hand-written tcgen05 product (bjx_gemm.cu) through the C ABI: velocity (exact split + product),
fused leapfrog (planes emitted by the epilogue) against float64 numpy, then kernel timings.
usage: python scripts/gemm_check.py [quick|time]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackjax_b200 import _engine, targets as T  # noqa: E402

DEV = "cuda:0"
F = np.float32


def tf(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)


def problem(D, C, seed=0):
    rs = np.random.default_rng(seed)
    A = rs.standard_normal((D, D))
    cov = (A @ A.T / D + np.eye(D))
    prec = np.linalg.inv(cov)
    return cov.astype(F), prec.astype(F), (0.5 * rs.standard_normal((C, D))).astype(F), rs.standard_normal((C, D)).astype(F)


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def check(D, C, L=3, eps=0.05):
    cov, prec, q, p = problem(D, C)
    tgt = T.DenseGaussian(prec)
    eng = _engine.Engine(DEV, C, D, tgt)
    eng.set_metric(tf(cov))
    v = eng.velocity(tf(p)).cpu().numpy()
    torch.cuda.synchronize()
    e_v = relerr(v, p.astype(np.float64) @ cov.astype(np.float64))
    dq, dp = tf(q), tf(p)
    logp, g = eng.init_state(dq)
    e_g = relerr(g.cpu().numpy(), -(q.astype(np.float64) @ prec.astype(np.float64)))
    eng.leapfrog_(dq, dp, logp, g, eps, L)
    torch.cuda.synchronize()
    q64, p64 = q.astype(np.float64), p.astype(np.float64)
    c64, P64 = cov.astype(np.float64), prec.astype(np.float64)
    g64 = -(q64 @ P64)
    for _ in range(L):
        p64 = p64 + 0.5 * eps * g64
        q64 = q64 + eps * (p64 @ c64)
        g64 = -(q64 @ P64)
        p64 = p64 + 0.5 * eps * g64
    e_q, e_p, e_gr = relerr(dq.cpu().numpy(), q64), relerr(dp.cpu().numpy(), p64), relerr(g.cpu().numpy(), g64)
    e_lp = relerr(logp.cpu().numpy(), -0.5 * np.einsum("ci,ij,cj->c", q64, P64, q64))
    print(f"D={D:5d} C={C:6d} L={L}: velocity {e_v:.2e} grad {e_g:.2e} | leapfrog q {e_q:.2e} p {e_p:.2e} g {e_gr:.2e} logp {e_lp:.2e}",
          flush=True)
    eng.close()
    return max(e_v, e_g, e_q, e_p, e_gr)


def timing(C=65536, D=1024, L=10):
    from oracle import targets as otargets
    cov, prec = otargets.correlated_gaussian(D, seed=0)
    tgt = T.DenseGaussian(prec)
    eng = _engine.Engine(DEV, C, D, tgt)
    eng.set_metric(tf(cov))
    g_ = torch.Generator(device=DEV).manual_seed(0)
    q = 0.1 * torch.randn(C, D, device=DEV, generator=g_)
    p = torch.randn(C, D, device=DEV, generator=g_)
    logp, g = eng.init_state(q)
    for n, reps in ((1, 3), (L, 3)):
        eng.leapfrog_(q, p, logp, g, 0.5, n)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            eng.leapfrog_(q, p, logp, g, 0.5, n)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        print(f"leapfrog x{n}: {ms:.3f} ms  ({ms / n:.3f} ms/step, {2 * 3 * 2 * C * D * D * n / ms / 1e9:.0f} fp16 TFLOP/s incl. row kernels)", flush=True)
    v = eng.velocity(p)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        v = eng.velocity(p)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print(f"velocity (split + product): {ms:.3f} ms -> {2 * C * D * D / ms / 1e9:.0f} f32-equivalent TFLOP/s", flush=True)
    eng.close()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    if mode == "loop":   # only the fused leapfrog loop at the config-2 shape
        from oracle import targets as otargets
        C, D, L = 65536, 1024, 8
        cov, prec = otargets.correlated_gaussian(D, seed=0)
        eng = _engine.Engine(DEV, C, D, T.DenseGaussian(prec))
        eng.set_metric(tf(cov))
        g_ = torch.Generator(device=DEV).manual_seed(0)
        q = 0.1 * torch.randn(C, D, device=DEV, generator=g_)
        p = torch.randn(C, D, device=DEV, generator=g_)
        logp, g = eng.init_state(q)
        eng.leapfrog_(q, p, logp, g, 0.01, L)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            eng.leapfrog_(q, p, logp, g, 0.01, L)
        t1.record(); torch.cuda.synchronize()
        if int(os.environ.get("BJX_GEMM_DEBUG", "0")) & 16:
            import ctypes
            from blackjax_b200 import _lib
            L_ = _lib.lib()
            buf = (ctypes.c_ulonglong * 16)()
            L_.bjx_debug_gemm_counters(buf, 1)
            eng.leapfrog_(q, p, logp, g, 0.01, L)
            L_.bjx_debug_gemm_counters(buf, 0)
            c = list(buf)
            nm, ne = max(c[9], 1), max(c[10], 1)
            print(f"  issuer (per CTA-launch, cycles): wait tmem_empty {c[0] / nm:.0f}  wait full {c[1] / nm:.0f}  total {c[2] / nm:.0f}")
            print(f"  producer: wait empty {c[3] / ne:.0f}")
            print(f"  epilogue warp 0: wait tmem_full {c[4] / ne:.0f}  wait cin {c[5] / ne:.0f}  wait store-read {c[6] / ne:.0f}  "
                  f"work {c[7] / ne:.0f}  total {c[8] / ne:.0f}")
        print(f"variant {os.environ.get('BJX_GEMM_VARIANT', '0')} debug {os.environ.get('BJX_GEMM_DEBUG', '0')}: "
              f"{t0.elapsed_time(t1) / 3 / L:.4f} ms per leapfrog step", flush=True)
        sys.exit(0)
    if mode == "prof":   # short run for ncu: only the config-2 shape
        timing(L=3)
        sys.exit(0)
    t = time.time()
    worst = 0.0
    for D, C in ((256, 100), (256, 300), (512, 37), (132, 21), (1024, 64), (1024, 600), (384, 8203)):
        worst = max(worst, check(D, C))
    print(f"worst {worst:.2e}  ({time.time() - t:.1f} s)")
    if mode == "time":
        timing()

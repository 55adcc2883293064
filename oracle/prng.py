"""Restatement of the JAX PRNG pieces BlackJAX's HMC/NUTS path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows the published algorithm of
``jax 0.10.0`` (``jax/_src/prng.py``, ``jax/_src/random.py``; pinned by the
reference's uv.lock:1309-1310) in its default configuration: ``threefry2x32``
implementation with ``jax_threefry_partitionable=True``.

Reference call sites this replaces: ``jax.random.split`` blackjax/mcmc/hmc.py:299,
nuts.py:133, trajectory.py:646, util.py:203; ``fold_in`` trajectory.py:216,321,645;
``bernoulli`` proposal.py:123,156,226, trajectory.py:650; ``normal`` util.py:90.

Keys are ``uint32[..., 2]`` arrays; every function is vectorised over leading axes.
"""
import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = _U32(0x1BD11BDA)


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011), as used by JAX."""
    k0 = np.asarray(k0, _U32)
    k1 = np.asarray(k1, _U32)
    x0 = np.asarray(x0, _U32)
    x1 = np.asarray(x1, _U32)
    with np.errstate(over="ignore"):
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def key(seed):
    """``jax.random.key(seed)`` / ``PRNGKey(seed)`` raw data for 0 <= seed < 2**32."""
    return np.array([0, seed], dtype=_U32)


def split(keys, num=2):
    """``jax.random.split``: child i = threefry(key, (0, i)).  keys [...,2] -> [...,num,2]."""
    keys = np.asarray(keys, _U32)
    idx = np.arange(num, dtype=_U32)
    k0 = keys[..., 0:1]
    k1 = keys[..., 1:2]
    o0, o1 = threefry2x32(k0, k1, np.zeros_like(idx), idx)
    return np.stack([o0, o1], axis=-1)


def fold_in(keys, data):
    """``jax.random.fold_in(key, data)`` = threefry(key, (0, data)); data broadcasts."""
    keys = np.asarray(keys, _U32)
    data = np.asarray(data).astype(_U32)
    o0, o1 = threefry2x32(keys[..., 0], keys[..., 1], np.zeros_like(data), data)
    return np.stack(np.broadcast_arrays(o0, o1), axis=-1)


def random_bits(keys, shape=()):
    """32 random bits per element: out0 ^ out1 of threefry(key, (hi(i), lo(i))), i row-major."""
    keys = np.asarray(keys, _U32)
    n = int(np.prod(shape)) if len(shape) else 1
    lin = np.arange(n, dtype=np.uint64)
    hi = (lin >> np.uint64(32)).astype(_U32)
    lo = (lin & np.uint64(0xFFFFFFFF)).astype(_U32)
    o0, o1 = threefry2x32(keys[..., 0:1], keys[..., 1:2], hi, lo)
    bits = o0 ^ o1
    return bits.reshape(keys.shape[:-1] + tuple(shape))


def uniform(keys, shape=(), minval=0.0, maxval=1.0):
    """``jax.random.uniform`` float32: mantissa trick, then scale, then max(lo, .)."""
    bits = random_bits(keys, shape)
    fb = (bits >> _U32(9)) | _U32(0x3F800000)
    f = fb.view(np.float32) - np.float32(1.0)
    lo = np.float32(minval)
    hi = np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


def randint(keys, shape, minval, maxval):
    """``jax.random.randint(key, shape, minval, maxval)`` int32 (jax/_src/random.py ``_randint``, restated from memory of
    jax 0.10 -- PARITY UNPINNED like the rest of the PRNG spec): ``k1, k2 = split(key)``; two 32-bit draws; the 64-bit
    value ``higher * 2**32 + lower`` is reduced modulo ``span = maxval - minval`` with the multiplier
    ``2**32 % span`` computed as ``((2**16 % span) ** 2) % span``; all arithmetic wraps in uint32."""
    keys = np.asarray(keys, _U32)
    ks = split(keys, 2)
    higher = random_bits(ks[..., 0, :], shape).astype(np.uint64)
    lower = random_bits(ks[..., 1, :], shape).astype(np.uint64)
    span = np.uint64((int(maxval) - int(minval)) & 0xFFFFFFFF) if maxval > minval else np.uint64(1)
    m32 = np.uint64(0xFFFFFFFF)
    mult = np.uint64(1 << 16) % span
    mult = ((mult * mult) & m32) % span
    off = ((((higher % span) * mult) & m32) + (lower % span)) & m32
    off = off % span
    return ((np.uint64(int(minval) & 0xFFFFFFFF) + off) & m32).astype(_U32).view(np.int32)


def bernoulli(keys, p=np.float32(0.5)):
    """``jax.random.bernoulli(key, p)`` = uniform(key, shape(p)) < p (p: per-key scalar)."""
    p = np.asarray(p, np.float32)
    return uniform(keys) < p


# Giles' single-precision erfinv polynomial as expanded by XLA (xla/client/lib/math.cc).
_ERFINV_LT5 = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
               0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
_ERFINV_GE5 = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
               0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)


def erfinv_f32(x):
    x = np.asarray(x, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-(x * x)).astype(np.float32)
        lt = w < np.float32(5.0)
        ww = np.where(lt, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        p = np.where(lt, np.float32(_ERFINV_LT5[0]), np.float32(_ERFINV_GE5[0])).astype(np.float32)
        for a, b in zip(_ERFINV_LT5[1:], _ERFINV_GE5[1:]):
            c = np.where(lt, np.float32(a), np.float32(b)).astype(np.float32)
            p = (c + p * ww).astype(np.float32)
        r = (p * x).astype(np.float32)
        r = np.where(np.abs(x) == np.float32(1.0), x * np.float32(np.inf), r)
    return r.astype(np.float32)


_NORMAL_LO = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)


def normal(keys, shape=()):
    """``jax.random.normal`` float32 = sqrt(2) * erfinv(uniform(key, lo=nextafter(-1,0), hi=1))."""
    u = uniform(keys, shape, minval=_NORMAL_LO, maxval=1.0)
    return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)

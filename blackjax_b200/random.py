"""jax.random key utilities restated on the device (threefry2x32, partitionable mode).

Keys are raw ``uint32[..., 2]`` tensors (``jax.random.key_data`` layout).  All derivations run in
libbjx's PRNG kernels (csrc/bjx_misc.cu), bit-exact with ``jax.random.split`` / ``fold_in``
(reference call sites: blackjax/util.py:200-203, staged_adaptation.py:868,920, hmc.py:299).
"""
import torch

from ._lib import check, lib, ptr


def _dev(device=None):
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def key(seed, device=None):
    """``jax.random.key(seed)`` raw data: (0, seed) for 0 <= seed < 2**32."""
    if not 0 <= int(seed) < 2 ** 32:
        raise ValueError("seed must be in [0, 2**32)")
    s = int(seed)
    s = s - (1 << 32) if s >= (1 << 31) else s
    return torch.tensor([0, s], dtype=torch.int32, device=_dev(device)).view(torch.uint32)


def _on_current_stream(device):
    """The handle-less PRNG entry points run on the stream registered here: torch's current stream of ``device``
    (side streams do not synchronise with the legacy default stream)."""
    check(lib().bjx_set_default_stream(torch.cuda.current_stream(device).cuda_stream))


def _flat(keys):
    if keys.dtype not in (torch.uint32, torch.int32) or keys.shape[-1] != 2:
        raise TypeError("keys must be uint32[..., 2]")
    if not keys.is_cuda:
        raise TypeError("keys must live on a CUDA device")
    k = keys.contiguous()
    return k, k.numel() // 2


def split(keys, num=2):
    """``jax.random.split(key, num)``: [..., 2] -> [..., num, 2]."""
    k, n = _flat(keys)
    out = torch.empty(tuple(k.shape[:-1]) + (int(num), 2), dtype=torch.uint32, device=k.device)
    with torch.cuda.device(k.device):
        _on_current_stream(k.device)
        check(lib().bjx_prng_split(None, ptr(k), n, int(num), ptr(out)))
    return out


def fold_in(keys, data):
    """``jax.random.fold_in(key, data)`` for an integer ``data`` shared by all keys."""
    k, n = _flat(keys)
    out = torch.empty_like(k, dtype=torch.uint32)
    with torch.cuda.device(k.device):
        _on_current_stream(k.device)
        check(lib().bjx_prng_fold_in(None, ptr(k), n, int(data) & 0xFFFFFFFF, ptr(out)))
    return out


def _draw(fn, keys, shape, dtype):
    k, n = _flat(keys)
    per = 1
    for s in shape:
        per *= int(s)
    out = torch.empty(tuple(k.shape[:-1]) + tuple(shape), dtype=dtype, device=k.device)
    with torch.cuda.device(k.device):
        _on_current_stream(k.device)
        check(fn(None, ptr(k), n, per, ptr(out)))
    return out


def bits(keys, shape=()):
    return _draw(lib().bjx_prng_random_bits, keys, shape, torch.uint32)


def uniform(keys, shape=()):
    """``jax.random.uniform(key, shape)`` float32 in [0, 1)."""
    return _draw(lib().bjx_prng_uniform, keys, shape, torch.float32)


def normal(keys, shape=()):
    """``jax.random.normal(key, shape)`` float32."""
    return _draw(lib().bjx_prng_normal, keys, shape, torch.float32)


def randint(keys, shape, minval, maxval):
    """``jax.random.randint(key, shape, minval, maxval)`` int32 in [minval, maxval)."""
    k, n = _flat(keys)
    per = 1
    for s_ in shape:
        per *= int(s_)
    out = torch.empty(tuple(k.shape[:-1]) + tuple(shape), dtype=torch.int32, device=k.device)
    with torch.cuda.device(k.device):
        _on_current_stream(k.device)
        check(lib().bjx_prng_randint(None, ptr(k), n, per, int(minval), int(maxval), ptr(out)))
    return out

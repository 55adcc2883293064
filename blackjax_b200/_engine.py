"""Engine: one libbjx handle per (device, stream, n_chains, dim, target, depth).  Thin, typed calls
into the C ABI; all arrays are torch CUDA tensors used purely as device memory."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr

import collections

# Engines own device workspaces (NUTS: (9 + 2*depth) rows of [C,D]; dense path: 4 rows + operand planes), so the cache is a
# small LRU: the least recently used engine leaves the cache when a new shape/target comes in and is destroyed (handle
# and workspaces freed) as soon as nobody else holds it.
_ENGINES = collections.OrderedDict()
_MAX_ENGINES = 8


def _f32(t, shape=None, name="array"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (blackjax_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


def as_keys(keys, n, device):
    """uint32 [n,2] per-chain keys, or ONE key uint32 [2] (per-chain keys are then derived in-kernel), on ``device``."""
    if not isinstance(keys, torch.Tensor):
        raise TypeError("rng_key must be a torch tensor of raw uint32 key data")
    if keys.dtype not in (torch.uint32, torch.int32):
        raise TypeError(f"rng_key must be uint32 (or int32 bit patterns), got {keys.dtype}")
    keys = keys.to(device).contiguous()
    if tuple(keys.shape) not in ((n, 2), (2,)):
        raise ValueError(f"rng_key has shape {tuple(keys.shape)}, expected ({n}, 2) or (2,)")
    return keys


class Engine:
    def __init__(self, device, n_chains, dim, target, max_tree_depth=10, divergence_threshold=1000.0, stream=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.BjxError("blackjax_b200 runs on CUDA devices only (no CPU fallback)")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)
        self.C, self.D = int(n_chains), int(dim)
        self.target = target
        self.max_tree_depth = int(max_tree_depth)
        self.stream = torch.cuda.current_stream(self.device) if stream is None else stream
        cfg = _lib.Config()
        cfg.device = self.index
        cfg.n_chains, cfg.dim = self.C, self.D
        cfg.max_tree_depth = self.max_tree_depth
        cfg.divergence_threshold = float(divergence_threshold)
        cfg.stream = self.stream.cuda_stream
        cfg.target = target.desc(self.device)
        h = C.c_void_p()
        check(lib().bjx_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._imm = None      # keeps the caller's inverse mass matrix alive
        self._imm_key = None

    def close(self):
        """Destroy the libbjx handle (synchronises its stream and frees its workspaces)."""
        if getattr(self, "h", None):
            lib().bjx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- metric ---------------------------------------------------------------------------------
    def set_metric(self, inverse_mass_matrix):
        """1-D => diagonal, 2-D [D,D] => dense, 2-D [C,D] => per-chain diagonal, 3-D [C,D,D] => per-chain dense
        (blackjax/mcmc/metrics.py:701-729 semantics incl. the ValueError; the per-chain layouts are what jax.vmap
        over chains gives the reference)."""
        imm = inverse_mass_matrix
        from .mcmc.metrics import LowRankMetric
        if isinstance(imm, LowRankMetric):  # metrics.gaussian_euclidean_low_rank (metrics.py:349-467)
            sg, U, lam = (t.to(self.device, torch.float32).contiguous() for t in imm)
            if sg.shape[0] != self.D:
                raise ValueError(f"low-rank metric has {sg.shape[0]} dimensions, expected {self.D}")
            check(lib().bjx_set_metric_low_rank(self.h, ptr(sg), ptr(U), ptr(lam), int(lam.shape[0])), self.h)
            self._imm = (sg, U, lam)
            return imm
        if not isinstance(imm, torch.Tensor):
            imm = torch.as_tensor(imm, dtype=torch.float32)
        imm = imm.to(self.device, torch.float32).contiguous()
        if imm.ndim == 1:
            if imm.shape[0] != self.D:
                raise ValueError(f"inverse_mass_matrix has {imm.shape[0]} entries, expected {self.D}")
            kind = _lib.METRIC_DIAG
        elif imm.ndim == 2 and tuple(imm.shape) == (self.D, self.D):
            kind = _lib.METRIC_DENSE
        elif imm.ndim == 2 and tuple(imm.shape) == (self.C, self.D):
            kind = _lib.METRIC_DIAG_PER_CHAIN
        elif imm.ndim == 3 and tuple(imm.shape) == (self.C, self.D, self.D):
            kind = _lib.METRIC_DENSE_PER_CHAIN   # one dense matrix per chain (vmapped dense adaptation), dim <= 64
        else:
            raise ValueError(
                "The mass matrix has the wrong number of dimensions:"
                f" expected 1 or 2, got {imm.ndim}.")
        self._imm = imm
        check(lib().bjx_set_metric(self.h, kind, ptr(imm)), self.h)
        return imm

    def ensure_metric(self, inverse_mass_matrix):
        """Install the metric unless this exact tensor CONTENT is already installed: the cache key is (storage pointer,
        shape, torch's in-place version counter), so an in-place update of the caller's tensor re-derives
        mass_matrix_sqrt and the dense operand planes instead of silently keeping stale ones."""
        imm = inverse_mass_matrix
        from .mcmc.metrics import LowRankMetric
        if isinstance(imm, LowRankMetric):
            key = tuple((t.data_ptr(), tuple(t.shape), t._version, str(t.device)) for t in imm)
        else:
            key = ((imm.data_ptr(), tuple(imm.shape), imm._version, str(imm.device)) if isinstance(imm, torch.Tensor)
                   else None)
        if key is None or key != self._imm_key:
            installed = self.set_metric(imm)
            # a converted copy (dtype / device / layout) is owned by the engine: key on the caller's tensor all the same
            self._imm_key = key
            return installed
        return self._imm

    def _key_mode(self, keys, chain_offset=0):
        """Select per-chain keys [C,2] or a shared step key [2] (+ global chain offset) for the next transition."""
        mode = (1 if keys.ndim == 1 else 0, int(chain_offset))
        if getattr(self, "_keymode", None) != mode:
            check(lib().bjx_set_key_mode(self.h, mode[0], mode[1]), self.h)
            self._keymode = mode

    def set_integrator(self, coefficients):
        """integrators.py:62-152 coefficient table (tuple of floats); cached per engine."""
        coef = tuple(float(c) for c in coefficients)
        if getattr(self, "_coef", (0.5, 1.0, 0.5)) != coef:
            arr = (C.c_float * len(coef))(*coef)
            check(lib().bjx_set_integrator(self.h, arr, len(coef)), self.h)
            self._coef = coef

    def mass_matrix_sqrt(self):
        out = C.c_void_p()
        check(lib().bjx_get_mass_matrix_sqrt(self.h, C.byref(out)), self.h)
        return out.value

    # -- building blocks ----------------------------------------------------------------------------
    def init_state(self, q):
        q = _f32(q, (self.C, self.D), "position")
        logp = torch.empty(self.C, dtype=torch.float32, device=self.device)
        g = torch.empty_like(q)
        check(lib().bjx_init_state(self.h, ptr(q), ptr(logp), ptr(g)), self.h)
        return logp, g

    def sample_momentum(self, keys, chain_offset=0):
        keys = as_keys(keys, self.C, self.device)
        self._key_mode(keys, chain_offset)
        p = torch.empty(self.C, self.D, dtype=torch.float32, device=self.device)
        check(lib().bjx_sample_momentum(self.h, ptr(keys), ptr(p)), self.h)
        return p

    def leapfrog_(self, q, p, logp, g, step_size, n_steps):
        """n_steps velocity-Verlet steps IN PLACE on (q, p, logp, g)."""
        eps, eps_dev = self._eps(step_size)
        check(lib().bjx_leapfrog(self.h, ptr(_f32(q, (self.C, self.D))), ptr(_f32(p, (self.C, self.D))),
                                 ptr(_f32(logp, (self.C,))), ptr(_f32(g, (self.C, self.D))), eps, ptr(eps_dev),
                                 int(n_steps)), self.h)

    def velocity(self, p):
        """linear_map(M^-1, p) for every chain (dim > 128)."""
        v = torch.empty(self.C, self.D, dtype=torch.float32, device=self.device)
        check(lib().bjx_metric_velocity(self.h, ptr(_f32(p, (self.C, self.D))), ptr(v)), self.h)
        return v

    def energy(self, p, logp):
        e = torch.empty(self.C, dtype=torch.float32, device=self.device)
        check(lib().bjx_energy(self.h, ptr(_f32(p, (self.C, self.D))), ptr(_f32(logp, (self.C,))), ptr(e)), self.h)
        return e

    def is_turning(self, pl, pr, ps):
        out = torch.empty(self.C, dtype=torch.uint8, device=self.device)
        check(lib().bjx_is_turning(self.h, ptr(_f32(pl, (self.C, self.D))), ptr(_f32(pr, (self.C, self.D))),
                                   ptr(_f32(ps, (self.C, self.D))), ptr(out)), self.h)
        return out.bool()

    def _eps(self, step_size):
        if isinstance(step_size, torch.Tensor) and step_size.ndim >= 1:
            return 0.0, _f32(step_size.to(self.device), (self.C,), "step_size")
        return float(step_size), None

    def _info(self, fields):
        info = _lib.Info()
        for k, v in fields.items():
            setattr(info, k, ptr(v))
        return info

    # -- transitions -----------------------------------------------------------------------------------
    def hmc_step(self, keys, q, logp, g, step_size, num_integration_steps, out=None, info_fields=None,
                 multinomial=False, chain_offset=0):
        keys = as_keys(keys, self.C, self.device)
        self._key_mode(keys, chain_offset)
        q = _f32(q, (self.C, self.D), "position")
        g = _f32(g, (self.C, self.D), "logdensity_grad")
        logp = _f32(logp, (self.C,), "logdensity")
        qo, lo, go = out if out is not None else (torch.empty_like(q), torch.empty_like(logp), torch.empty_like(g))
        eps, eps_dev = self._eps(step_size)
        info = self._info(info_fields or {})
        fn = lib().bjx_mhmc_step if multinomial else lib().bjx_hmc_step
        steps_dev = None
        if isinstance(num_integration_steps, torch.Tensor):  # dynamic HMC: one trajectory length per chain
            steps_dev = num_integration_steps.to(device=self.device, dtype=torch.int32).contiguous()
            if steps_dev.shape != (self.C,):
                raise ValueError(f"per-chain num_integration_steps must have shape ({self.C},)")
            num_integration_steps = 1
        check(lib().bjx_set_integration_steps(self.h, ptr(steps_dev)), self.h)
        try:
            check(fn(self.h, ptr(keys), ptr(q), ptr(logp), ptr(g), ptr(qo), ptr(lo), ptr(go), eps,
                     ptr(eps_dev), int(num_integration_steps), C.byref(info)), self.h)
        finally:
            if steps_dev is not None:
                check(lib().bjx_set_integration_steps(self.h, None), self.h)
                self._steps_keepalive = steps_dev  # the launch is asynchronous: keep the array until the next call
        return qo, lo, go

    def nuts_step(self, keys, q, logp, g, step_size, max_num_doublings, out=None, info_fields=None,
                  momentum=None, key_integrator=None, chain_offset=0):
        if max_num_doublings > self.max_tree_depth:
            raise ValueError("max_num_doublings exceeds the engine's max_tree_depth")
        q = _f32(q, (self.C, self.D), "position")
        g = _f32(g, (self.C, self.D), "logdensity_grad")
        logp = _f32(logp, (self.C,), "logdensity")
        if key_integrator is not None:
            key_integrator = as_keys(key_integrator, self.C, self.device)
            momentum = _f32(momentum, (self.C, self.D), "momentum")
            keys = None
        else:
            keys = as_keys(keys, self.C, self.device)
            self._key_mode(keys, chain_offset)
        qo, lo, go = out if out is not None else (torch.empty_like(q), torch.empty_like(logp), torch.empty_like(g))
        eps, eps_dev = self._eps(step_size)
        info = self._info(info_fields or {})
        check(lib().bjx_nuts_step(self.h, ptr(keys), ptr(q), ptr(logp), ptr(g), ptr(qo), ptr(lo), ptr(go), eps,
                                  ptr(eps_dev), int(max_num_doublings), C.byref(info), ptr(momentum),
                                  ptr(key_integrator)), self.h)
        if key_integrator is not None:
            self._keymode = None  # per-call override: re-establish the key mode on the next ordinary call
        return qo, lo, go

    def nuts_last_stats(self):
        a, b = C.c_int64(), C.c_int64()
        check(lib().bjx_nuts_last_stats(self.h, C.byref(a), C.byref(b)), self.h)
        return a.value, b.value

    def synchronize(self):
        check(lib().bjx_synchronize(self.h), self.h)


def get_engine(position, target, max_tree_depth=10, divergence_threshold=1000.0):
    """Engine cache keyed by (device, stream, shape, target identity, depth, threshold)."""
    if not isinstance(position, torch.Tensor) or not position.is_cuda:
        raise TypeError("position must be a CUDA tensor of shape [n_chains, dim]")
    if position.ndim != 2:
        raise ValueError("position must have shape [n_chains, dim] (the chain axis is explicit in blackjax_b200)")
    dev = position.device
    stream = torch.cuda.current_stream(dev)
    key = (dev.index, stream.cuda_stream, position.shape[0], position.shape[1], id(target), int(max_tree_depth),
           float(divergence_threshold))
    eng = _ENGINES.get(key)
    if eng is None:
        if position.shape[1] != target.dim:
            raise ValueError(f"position has dim {position.shape[1]} but the target has dim {target.dim}")
        eng = Engine(dev, position.shape[0], position.shape[1], target, max_tree_depth, divergence_threshold, stream)
        _ENGINES[key] = eng
        while len(_ENGINES) > _MAX_ENGINES:
            # drop the least recently used engine from the cache only: callers may still hold it (its handle and
            # workspaces are freed by Engine.__del__ once the last reference goes)
            _ENGINES.popitem(last=False)
    else:
        _ENGINES.move_to_end(key)
    return eng

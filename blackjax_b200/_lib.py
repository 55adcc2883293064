"""ctypes binding of libbjx.so (include/bjx.h).  The product path has NO CPU fallback: if the
CUDA library is missing this module raises at import of the first op."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbjx.so")

TARGET_DIAG_GAUSSIAN, TARGET_FUNNEL, TARGET_DENSE_GAUSSIAN, TARGET_BANANA, TARGET_HIER_LOGIT, TARGET_USER = 0, 1, 2, 3, 4, 5
METRIC_DIAG, METRIC_DENSE, METRIC_DIAG_PER_CHAIN, METRIC_LOW_RANK, METRIC_DENSE_PER_CHAIN = 0, 1, 2, 3, 4

_f32p = C.c_void_p  # device pointers travel as integers


class TargetDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("inv_var", C.c_void_p), ("mean", C.c_void_p),
                ("precision", C.c_void_p), ("logp_offset", C.c_float), ("data_x", C.c_void_p), ("data_y", C.c_void_p),
                ("n_groups", C.c_int32), ("n_user_params", C.c_int32), ("user_params", C.c_void_p),
                ("user_plugin", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_chains", C.c_int32), ("dim", C.c_int32),
                ("max_tree_depth", C.c_int32), ("divergence_threshold", C.c_float), ("stream", C.c_void_p),
                ("target", TargetDesc)]


class Info(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "acceptance_rate", "is_accepted", "is_divergent", "is_turning", "energy", "num_integration_steps",
        "num_trajectory_expansions", "momentum", "proposal_position", "proposal_momentum", "left_position",
        "left_momentum", "right_position", "right_momentum")]


class BjxError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "bjx_version": (C.c_int, []),
    "bjx_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "bjx_destroy": (C.c_int, [C.c_void_p]),
    "bjx_last_error": (C.c_char_p, [C.c_void_p]),
    "bjx_set_target": (C.c_int, [C.c_void_p, C.POINTER(TargetDesc)]),
    "bjx_plugin_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "bjx_plugin_abi": (C.c_int, []),
    "bjx_set_integrator": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int32]),
    "bjx_set_key_mode": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint32]),
    "bjx_set_ghmc_noise": (C.c_int, [C.c_void_p, _f32p]),
    "bjx_set_integration_steps": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bjx_synchronize": (C.c_int, [C.c_void_p]),
    "bjx_set_metric": (C.c_int, [C.c_void_p, C.c_int32, _f32p]),
    "bjx_set_metric_low_rank": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, C.c_int32]),
    "bjx_get_mass_matrix_sqrt": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "bjx_init_state": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p]),
    "bjx_sample_momentum": (C.c_int, [C.c_void_p, _f32p, _f32p]),
    "bjx_leapfrog": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, C.c_int32]),
    "bjx_metric_velocity": (C.c_int, [C.c_void_p, _f32p, _f32p]),
    "bjx_energy": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p]),
    "bjx_is_turning": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p]),
    "bjx_hmc_step": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p,
                               C.c_int32, C.POINTER(Info)]),
    "bjx_mhmc_step": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p,
                                C.c_int32, C.POINTER(Info)]),
    "bjx_hmc_sample": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, C.c_int32, C.c_int32, C.c_int32,
                                 _f32p, C.c_int32, _f32p]),
    "bjx_nuts_step": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p,
                                C.c_int32, C.POINTER(Info), _f32p, _f32p]),
    "bjx_nuts_sample": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, C.c_int32, C.c_int32, _f32p,
                                  C.c_int32, _f32p, C.c_void_p]),
    "bjx_adapt_shared_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _f32p, C.c_char_p, C.c_int32, _f32p, _f32p, _f32p,
                                       _f32p, _f32p, _f32p, C.c_float, C.c_int32, C.c_int32, _f32p, _f32p, C.c_void_p,
                                       C.c_void_p]),
    "bjx_nuts_last_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bjx_prng_split": (C.c_int, [C.c_void_p, _f32p, C.c_int64, C.c_int32, _f32p]),
    "bjx_prng_fold_in": (C.c_int, [C.c_void_p, _f32p, C.c_int64, C.c_uint32, _f32p]),
    "bjx_prng_random_bits": (C.c_int, [C.c_void_p, _f32p, C.c_int64, C.c_int64, _f32p]),
    "bjx_prng_uniform": (C.c_int, [C.c_void_p, _f32p, C.c_int64, C.c_int64, _f32p]),
    "bjx_prng_normal": (C.c_int, [C.c_void_p, _f32p, C.c_int64, C.c_int64, _f32p]),
    "bjx_prng_randint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "bjx_da_init": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p]),
    "bjx_da_update": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_float, _f32p]),
    "bjx_da_reset": (C.c_int, [C.c_void_p, _f32p, _f32p]),
    "bjx_da_final": (C.c_int, [C.c_void_p, _f32p, _f32p]),
    "bjx_welford_update": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, C.c_int32]),
    "bjx_welford_final": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int32, _f32p]),
    "bjx_welford_dense_update": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, C.c_int32]),
    "bjx_welford_dense_final": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int32, _f32p]),
    "bjx_pooled_stats": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p]),
    "bjx_pooled_stats_dense": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p]),
    "bjx_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "bjx_nccl_comm_init_rank": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "bjx_nccl_comm_destroy": (C.c_int, [C.c_void_p]),
    "bjx_allgather_stats": (C.c_int, [C.c_void_p, C.c_void_p, _f32p, C.c_int64, _f32p]),
    "bjx_adapt_shared_state_floats": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "bjx_adapt_shared_init": (C.c_int, [C.c_void_p, _f32p, C.c_float, _f32p, _f32p]),
    "bjx_adapt_shared_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _f32p, _f32p, _f32p, C.c_int32, C.c_int32,
                                          C.c_float, _f32p, _f32p, _f32p]),
    "bjx_adapt_shared_final": (C.c_int, [C.c_void_p, _f32p, _f32p]),
    "bjx_set_default_stream": (C.c_int, [C.c_void_p]),
    "bjx_chees_state_floats": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "bjx_chees_init": (C.c_int, [C.c_void_p, _f32p, C.c_float, C.c_int32, C.c_float, _f32p, C.c_void_p]),
    "bjx_chees_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p,
                                   C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, _f32p, C.c_void_p, _f32p]),
    "bjx_chees_final": (C.c_int, [C.c_void_p, _f32p, C.POINTER(C.c_float)]),
    "bjx_ghmc_step": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, C.c_float, _f32p,
                                C.c_float, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Info)]),
    "bjx_meads_state_floats": (C.c_size_t, [C.c_int32, C.c_int32]),
    "bjx_meads_scratch_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "bjx_meads_update": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_float, C.c_float, _f32p, _f32p]),
    "bjx_maximum_eigenvalue_scratch_floats": (C.c_size_t, [C.c_int64, C.c_int32]),
    "bjx_maximum_eigenvalue": (C.c_int, [C.c_void_p, _f32p, C.c_int64, C.c_int32, _f32p, _f32p]),
    "bjx_permutation_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "bjx_permutation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "bjx_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, _f32p, _f32p, C.c_int64, C.c_int32]),
    "bjx_potential_scale_reduction": (C.c_int, [C.c_void_p, _f32p, C.c_int32, _f32p, _f32p]),
    "bjx_ess_scratch_floats": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "bjx_effective_sample_size": (C.c_int, [C.c_void_p, _f32p, C.c_int32, _f32p, _f32p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load libbjx.so once; fail loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BjxError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C blackjax_b200/csrc`).  blackjax_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, handle=None):
    if rc != 0:
        msg = lib().bjx_last_error(handle)
        raise BjxError(f"libbjx error {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()

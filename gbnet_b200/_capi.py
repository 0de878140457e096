"""ctypes binding of the C ABI declared in include/gbnet_b200.h.

This is the only place the shared library is loaded.  There is no fallback: if
``libgbnet_b200.so`` is missing and cannot be built, importing fails; if no CUDA device is
present, every compute call raises ``GbnError`` (GBN_ERR_CUDA).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

GBN_OK, GBN_ERR_INVALID, GBN_ERR_CUDA, GBN_ERR_UNSUPPORTED, GBN_ERR_COMM = 0, -1, -2, -3, -4
KERNEL_AUTO, KERNEL_STRICT, KERNEL_FAST = 0, 1, 2
COMM_ID_BYTES = 128


class GbnError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"gbnet_b200 error {status}: {message}")
        self.status = status
        self.message = message


class Network(C.Structure):
    _fields_ = [("n_edge", C.c_uint32), ("src", C.POINTER(C.c_uint32)), ("trg", C.POINTER(C.c_uint32)),
                ("mor", C.POINTER(C.c_int32)), ("n_active", C.c_uint32), ("active_tf", C.POINTER(C.c_uint32))]


class Evidence(C.Structure):
    _fields_ = [("n_datasets", C.c_uint32), ("n_evid", C.c_uint32), ("evid_trg", C.POINTER(C.c_uint32)),
                ("evid_val", C.POINTER(C.c_int32))]


class Params(C.Structure):
    _fields_ = [("sprior", C.c_double * 9),
                ("z_alpha", C.c_double), ("z_beta", C.c_double), ("z0_alpha", C.c_double), ("z0_beta", C.c_double),
                ("t_alpha", C.c_double), ("t_beta", C.c_double),
                ("n_graphs", C.c_uint32),
                ("noise_listen_children", C.c_int32), ("comp_yprob", C.c_int32), ("const_params", C.c_int32),
                ("t_focus", C.c_double), ("t_lmargin", C.c_double), ("t_hmargin", C.c_double),
                ("zn_focus", C.c_double), ("zn_lmargin", C.c_double), ("zn_hmargin", C.c_double),
                ("seed", C.c_uint64), ("device", C.c_int32), ("kernel", C.c_int32),
                ("chain_offset", C.c_uint32), ("n_graphs_local", C.c_uint32)]


class Dims(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_tf", "n_trg", "n_edge", "n_nodes", "n_datasets", "n_graphs",
                                          "n_graphs_local", "chain_offset", "n_stats", "simulation")]


class TopologyDims(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_tf", "n_trg", "n_edge", "n_edge_unique", "n_slots", "n_groups",
                                          "max_indeg", "max_outdeg")]


# every symbol include/gbnet_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
_vp, _u32, _u64, _dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_double
SYMBOLS = {
    "gbn_params_default": (None, [_P(Params)]),
    "gbn_last_error": (C.c_char_p, []),
    "gbn_abi_version": (C.c_int, []),
    "gbn_device_count": (C.c_int, []),
    "gbn_model_create": (C.c_int, [_P(Network), _P(Evidence), _P(Params), _P(_vp)]),
    "gbn_model_destroy": (None, [_vp]),
    "gbn_model_dims": (C.c_int, [_vp, _P(Dims)]),
    "gbn_model_kernel": (C.c_int, [_vp]),
    "gbn_model_path": (C.c_int, [_vp]),
    "gbn_model_run_protocol": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _P(_u32), _P(_dbl)]),
    "gbn_fisher_counts": (C.c_int, [C.c_int, _P(Network), _u32, _u32, _P(_u32), _P(C.c_int8), _P(_u32), _P(_u32), _P(_u32)]),
    "gbn_fisher_last_error": (C.c_char_p, []),
    "gbn_model_tf_ids": (C.c_int, [_vp, _P(_u32)]),
    "gbn_model_trg_ids": (C.c_int, [_vp, _P(_u32)]),
    "gbn_model_sample": (C.c_int, [_vp, _u32, _u32]),
    "gbn_model_sample_n": (C.c_int, [_vp, _u32]),
    "gbn_model_burn_stats": (C.c_int, [_vp]),
    "gbn_model_chain_length": (C.c_int, [_vp, _P(_dbl)]),
    "gbn_model_max_gelman_rubin": (C.c_int, [_vp, _P(_dbl)]),
    "gbn_model_gelman_rubin": (C.c_int, [_vp, _u32, _P(_dbl)]),
    "gbn_model_posterior_means": (C.c_int, [_vp, _u32, C.c_char, _P(_dbl), _P(_u32)]),
    "gbn_model_posterior_sdevs": (C.c_int, [_vp, _u32, C.c_char, _P(_dbl), _P(_u32)]),
    "gbn_model_print_stats": (C.c_int, [_vp]),
    "gbn_set_signal": (None, [C.c_int]),
    "gbn_install_sigint_handler": (None, []),
    "gbn_model_set_evidence": (C.c_int, [_vp, _P(Evidence)]),
    "gbn_comm_unique_id": (C.c_int, [_vp]),
    "gbn_model_attach_comm": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "gbn_model_get_state": (C.c_int, [_vp, _u32, _P(_dbl)]),
    "gbn_model_set_state": (C.c_int, [_vp, _u32, _P(_dbl)]),
    "gbn_model_eval_conditionals": (C.c_int, [_vp, _u32, _P(_dbl)]),
    "gbn_model_target_likelihoods": (C.c_int, [_vp, _u32, _P(_dbl)]),
    "gbn_model_joint_logprob": (C.c_int, [_vp, _u32, _P(_dbl)]),
    "gbn_model_inject_draws": (C.c_int, [_vp, _u32, _P(_dbl), _P(_dbl), _P(_dbl)]),
    "gbn_model_export_draws": (C.c_int, [_vp, _u32, _u32, _u32, _P(_dbl), _P(_dbl), _P(_dbl)]),
    "gbn_model_node_moments": (C.c_int, [_vp, _u32, _P(_dbl), _P(_dbl)]),
    "gbn_model_checkpoint_size": (C.c_int, [_vp, _P(C.c_size_t)]),
    "gbn_model_checkpoint_save": (C.c_int, [_vp, _vp, C.c_size_t]),
    "gbn_model_checkpoint_load": (C.c_int, [_vp, _vp, C.c_size_t]),
    "gbn_philox_kat": (C.c_int, [C.c_int, _P(_u32), _P(_u32), _P(_u32)]),
    "gbn_topology_probe": (C.c_int, [_P(Network), _P(Evidence), C.c_int, _P(TopologyDims), _P(_u32), _P(_u32), _P(_u32),
                                     _P(_u32), _P(_u32), _P(_u32), _P(_u32), _P(C.c_uint8), _P(_dbl)]),
    "gbn_model_last_device_ms": (C.c_int, [_vp, _P(_dbl)]),
    "gbn_model_set_phase_timing": (C.c_int, [_vp, C.c_int]),
    "gbn_model_set_stream_groups": (C.c_int, [_vp, _u32]),
    "gbn_model_phase_ms": (C.c_int, [_vp, _P(_dbl), _P(_u64)]),
    "gbn_model_launch_count": (C.c_int, [_vp, _P(_u64)]),
    "gbn_model_sweep_count": (C.c_int, [_vp, _P(_u64)]),
    "gbn_model_timer_start": (C.c_int, [_vp]),
    "gbn_model_timer_stop": (C.c_int, [_vp, _P(_dbl)]),
    "gbn_model_debug_counters": (C.c_int, [_vp, C.c_int, _P(_u64)]),
    "gbn_model_synchronize": (C.c_int, [_vp]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the sources are newer and nvcc is present) the C-ABI library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    try:
        _build.build_native()
    except _build.CompileError:
        raise                 # the sources do not compile: never fall back to a stale library
    except Exception as exc:  # no nvcc here (a box with only the prebuilt .so)
        if not os.path.exists(path):
            raise ImportError(f"libgbnet_b200.so is missing and could not be built ({exc}); "
                              "gbnet_b200 has no CPU fallback") from exc
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)            # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int):
    if status != GBN_OK:
        raise GbnError(status, load().gbn_last_error().decode())


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_P(_dbl))


def u32ptr(a: np.ndarray):
    return a.ctypes.data_as(_P(_u32))


def i32ptr(a: np.ndarray):
    return a.ctypes.data_as(_P(C.c_int32))

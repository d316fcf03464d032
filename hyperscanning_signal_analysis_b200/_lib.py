"""ctypes binding of libhs_b200.so (the C ABI declared in include/hs_b200.h).

There is deliberately no CPU fallback: if the shared library is missing or no
CUDA device is visible, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhs_b200.so")

_lib = None
_lock = threading.Lock()

c_dp = C.c_void_p      # device / host pointers are passed as integers
c_i64 = C.c_int64
c_int = C.c_int
c_dbl = C.c_double
c_sz = C.c_size_t

# name -> (restype, argtypes); must list EVERY symbol of include/hs_b200.h
SIGNATURES = {
    "hs_last_error": (C.c_char_p, []),
    "hs_version": (c_int, []),
    "hs_launch_count": (C.c_longlong, []),
    "hs_timing_enable": (None, [c_int]),
    "hs_timing_last_k5_ms": (c_int, [C.POINTER(C.c_double)]),
    "hs_measure_dfma_tflops": (c_int, [C.POINTER(C.c_double), c_dp, c_int]),
    "hs_lagcov_f64": (c_int, [c_dp, c_dp, c_i64, c_int, c_int, c_int, c_int, c_int, c_dp, c_dp]),
    "hs_yw_assemble_f64": (c_int, [c_dp, c_int, c_int, c_int, c_dp, c_dp, c_dp]),
    "hs_yw_ws_bytes": (c_sz, [c_int, c_int, c_int]),
    "hs_yw_solve_f64": (c_int, [c_dp, c_int, c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "hs_mvar_criterion_f64": (c_int, [c_dp, c_int, c_int, c_int, c_int, c_int, c_dp, c_dp, c_dp, c_dp]),
    "hs_ztable_f64": (c_int, [c_dp, c_int, c_int, c_dbl, c_dp, c_dp]),
    "hs_transfer_ws_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "hs_transfer_ws_flag_offset": (c_sz, [c_int, c_int, c_int, c_int]),
    "hs_transfer_set_kernel": (c_int, [c_int]),
    "hs_transfer_dtf_f64": (c_int, [c_dp, c_dp, c_int, c_dbl, c_int, c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "hs_spectra_f64": (c_int, [c_dp, c_dp, c_int, c_int, c_int, c_dp, c_dp]),
    "hs_gpdc_f64": (c_int, [c_dp, c_dp, c_int, c_int, c_int, c_dp, c_dp]),
    "hs_partial_coherence_f64": (c_int, [c_dp, c_int, c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "hs_mvar_ffdtf_ws_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "hs_mvar_ffdtf_f64": (c_int, [c_dp, c_dp, c_i64, c_int, c_int, c_int, c_int, c_dp, c_int, c_dbl, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "hs_plan_create": (c_int, [C.POINTER(C.c_void_p), c_int, c_int, c_int, c_int, c_int, c_i64]),
    "hs_plan_destroy": (None, [C.c_void_p]),
    "hs_plan_mvar_ffdtf_host": (c_int, [C.c_void_p, c_dp, c_i64, c_dp, c_int, c_dp, c_dbl, c_dp, c_dp]),
    "hs_plan_host_result": (c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(c_sz)]),
    "hs_set_compute_sm_limit": (c_int, [c_int]),
    "hs_gather_push_f64": (c_int, [c_dp, c_i64, c_dp, C.POINTER(C.c_void_p), c_int, c_int, c_dp]),
    "hs_gather_push_ce": (c_int, [c_dp, c_i64, C.POINTER(C.c_void_p), c_int, c_dp]),
    "hs_ipc_alloc": (c_int, [C.POINTER(C.c_void_p), c_sz, C.c_char_p]),
    "hs_ipc_open": (c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "hs_ipc_close": (c_int, [c_dp]),
    "hs_ipc_free": (c_int, [c_dp]),
    "hs_filtfilt_ws_bytes": (c_sz, [c_int, c_i64]),
    "hs_iir_filtfilt_f64": (c_int, [c_dp, c_int, c_i64, c_i64, c_i64, c_dp, c_dp, c_int, c_int, c_int, c_dp, c_dp]),
    "hs_iir_lfilter_f64": (c_int, [c_dp, c_int, c_i64, c_i64, c_dp, c_dp, c_int, c_int, c_dp, c_i64, c_dp, c_dp]),
    "hs_fir_filter_f64": (c_int, [c_dp, c_int, c_i64, c_i64, c_int, c_i64, c_dp, c_int, c_dp, c_i64, c_i64, c_dp]),
    "hs_fir_decimate_f64": (c_int, [c_dp, c_int, c_i64, c_i64, c_int, c_dp, c_int, c_dp, c_i64, c_dp]),
    "hs_hilbert_ws_bytes": (c_sz, [c_int, c_i64]),
    "hs_hilbert_f64": (c_int, [c_dp, c_int, c_i64, c_i64, c_i64, c_dp, c_i64, c_dp, c_dp, c_dp]),
    "hs_mt_psd_set_path": (c_int, [c_int]),
    "hs_mt_psd_ws_bytes": (c_sz, [c_int, c_i64, c_int]),
    "hs_mt_psd_f64": (c_int, [c_dp, c_int, c_i64, c_dp, c_dp, c_int, c_int, c_int, c_dp, c_dp, c_dp]),
}


class HsError(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed just to load it)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise HsError(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                "This package has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().hs_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise HsError(f"{what} failed (code {rc}): {last_error()}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise HsError("no CUDA device visible: hyperscanning_signal_analysis_b200 runs on a B200 only (no CPU fallback)")
    return torch


def launch_count():
    return int(load().hs_launch_count())

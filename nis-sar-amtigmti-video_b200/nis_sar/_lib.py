"""ctypes binding of libnis_sar.so (declared in include/nis_sar.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present,
every compute entry point raises.  PyTorch is used only as the owner of device buffers and
streams; pointers cross the C ABI as plain integers.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libnis_sar.so")

NIS_OK = 0


class NisError(RuntimeError):
    pass


class EchoParams(C.Structure):
    _fields_ = [("c", C.c_double), ("fc", C.c_double), ("k_rate", C.c_double), ("t_p", C.c_double),
                ("t_start", C.c_double), ("dt_fast", C.c_double),
                ("per_target_velocity", C.c_int32), ("samples_per_thread", C.c_int32)]


class CsaParams(C.Structure):
    _fields_ = [("c", C.c_double), ("lambda_", C.c_double), ("kr", C.c_double), ("fs", C.c_double),
                ("prf", C.c_double), ("vr", C.c_double), ("r_ref", C.c_double), ("t_start", C.c_double)]


class RdaParams(C.Structure):
    _fields_ = [("c", C.c_double), ("lambda_", C.c_double), ("t_p", C.c_double), ("kr", C.c_double), ("fs", C.c_double),
                ("prf", C.c_double), ("vr", C.c_double), ("range_grp", C.c_double)]


class TdbpParams(C.Structure):
    _fields_ = [("c", C.c_double), ("fc", C.c_double), ("k_rate", C.c_double), ("t_p", C.c_double), ("fs", C.c_double),
                ("t_start", C.c_double), ("scene_size", C.c_double), ("n_samples", C.c_int32), ("nx", C.c_int32),
                ("ny", C.c_int32), ("reserved", C.c_int32)]


class GmtiResult(C.Structure):
    _fields_ = [("det_count", C.c_uint32), ("peak_idx", C.c_uint32), ("max_mag_sq", C.c_double)]


# name -> (restype, argtypes); the test-suite checks that every symbol of nis_sar.h is here and exported
_P = C.c_void_p
SIGNATURES = {
    "nis_version": (C.c_int, []),
    "nis_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "nis_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "nis_ctx_destroy": (C.c_int, [_P]),
    "nis_ctx_launch_count": (C.c_uint64, [_P]),
    "nis_echo_accumulate": (C.c_int, [_P, C.POINTER(EchoParams), _P, _P, _P, _P, _P, _P, _P,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, _P]),
    "nis_echo_spotlight": (C.c_int, [_P, C.POINTER(EchoParams), _P, _P, _P, _P, _P, _P, _P,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, _P, C.c_int32, _P]),
    "nis_tdbp_plan_create": (C.c_int, [_P, C.POINTER(TdbpParams), C.POINTER(_P)]),
    "nis_tdbp_plan_destroy": (C.c_int, [_P]),
    "nis_tdbp_range_compress": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P]),
    "nis_tdbp_backproject": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_double, _P, _P, C.c_int32,
                                       _P]),
    "nis_csa_plan_create": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(CsaParams), C.POINTER(_P)]),
    "nis_csa_plan_destroy": (C.c_int, [_P]),
    "nis_csa_size_class": (C.c_int, [C.c_int32, C.c_int32]),
    "nis_csa_axes": (C.c_int, [_P, _P, _P]),
    "nis_csa_focus": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    "nis_csa_plan_set_profiling": (C.c_int, [_P, C.c_int32]),
    "nis_csa_stage_times": (C.c_int, [_P, C.c_int32, _P]),
    "nis_rda_supported": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(RdaParams)]),
    "nis_rda_plan_create": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(RdaParams), C.POINTER(_P)]),
    "nis_rda_plan_destroy": (C.c_int, [_P]),
    "nis_rda_axes": (C.c_int, [_P, _P, _P, _P]),
    "nis_rda_focus": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "nis_viewer_products": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nis_region_stats": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "nis_region_select": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "nis_peer_alloc": (C.c_int, [C.c_uint64, C.POINTER(_P), _P]),
    "nis_peer_free": (C.c_int, [_P]),
    "nis_peer_open": (C.c_int, [_P, C.POINTER(_P)]),
    "nis_peer_close": (C.c_int, [_P]),
    "nis_peer_native_atomics": (C.c_int, [C.c_int32, C.c_int32]),
    "nis_comm_unique_id": (C.c_int, [_P]),
    "nis_comm_init": (C.c_int, [C.c_int32, C.c_int32, _P, C.POINTER(_P)]),
    "nis_comm_destroy": (C.c_int, [_P]),
    "nis_echo_reduce": (C.c_int, [_P, _P, C.c_uint64, C.c_int32, _P]),
    "nis_slc_exchange": (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
    "nis_power_sum": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "nis_power_max": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "nis_noise_add": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64,
                                C.c_int32, _P]),
    "nis_gmti_workspace_bytes": (C.c_uint64, [C.c_uint64]),
    "nis_gmti_fused": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_double, C.c_double,
                                 _P, _P, _P, _P, _P, _P, _P, _P, C.c_uint32, _P, _P, C.c_uint64, _P, _P]),
    "nis_gmti_balance_sum": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P]),
    "nis_narrow_c128_to_c32": (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
    "nis_widen_c32_to_c128": (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
    "nis_d2h_widen": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int32, _P]),
    "nis_h2d_narrow": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int32, _P]),
    "nis_transpose_c32": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
}

_lib = None
_lock = threading.Lock()


def load():
    """dlopen the library (works without a GPU: the CUDA runtime is linked statically and is
    only initialised by nis_ctx_create)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise NisError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().nis_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != NIS_OK:
        raise NisError(f"{what} failed (code {rc}): {last_error()}")


_ctx = {}


_ctx_lock = threading.Lock()


def context(device_index: int):
    """One nis_ctx per device, created on first use (the lock is held across the creation: two threads asking for the
    same device get the same context)."""
    lib = load()
    with _ctx_lock:
        h = _ctx.get(device_index)
        if h is None:
            out = _P()
            check(lib.nis_ctx_create(int(device_index), C.byref(out)), "nis_ctx_create")
            _ctx[device_index] = h = out
    return h


def launch_count(device_index: int = 0) -> int:
    h = _ctx.get(device_index)
    return int(load().nis_ctx_launch_count(h)) if h is not None else 0

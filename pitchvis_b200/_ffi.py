"""ctypes declarations for libpvqt.so (include/pvqt.h).

The library is the product; there is no Python or CPU fallback.  Importing this module
when the shared library has not been built raises ImportError with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libpvqt.so")

(PVQT_OK, PVQT_ABOVE_NYQUIST, PVQT_WINDOW_EXCEEDS_NFFT, PVQT_PANIC, PVQT_BAD_LENGTH, PVQT_INVALID_ARGUMENT,
 PVQT_UNSUPPORTED, PVQT_CUDA_ERROR, PVQT_OUT_OF_MEMORY) = range(9)


class PvqtParams(C.Structure):
    _fields_ = [
        ("sr", C.c_float),
        ("n_fft", C.c_uint64),
        ("min_freq", C.c_float),
        ("octaves", C.c_uint32),
        ("buckets_per_octave", C.c_uint32),
        ("sparsity_quantile", C.c_float),
        ("quality", C.c_float),
        ("gamma", C.c_float),
    ]


class PvqtError(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("highest_frequency", C.c_float),
        ("nyquist_frequency", C.c_float),
        ("window_length", C.c_float),
        ("n_fft", C.c_uint64),
        ("cuda_error", C.c_int32),
    ]


class PvqtFilterParams(C.Structure):
    _fields_ = [
        ("freq", C.c_float),
        ("window_length", C.c_float),
        ("sr_downscaling_factor", C.c_uint64),
        ("minimum_needed_window_size", C.c_uint64),
    ]


class PvqtCsrView(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("nnz", C.c_int64),
        ("indptr", C.POINTER(C.c_int32)),
        ("indices", C.POINTER(C.c_int32)),
        ("data", C.POINTER(C.c_float)),
    ]


# name -> (restype, argtypes); every symbol include/pvqt.h declares
_FP = C.POINTER(C.c_float)
_VP = C.c_void_p
_SZ = C.c_size_t
SIGNATURES = {
    "pvqt_abi_version": (C.c_int, []),
    "pvqt_last_error_string": (C.c_char_p, []),
    "pvqt_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pvqt_default_params": (C.c_int, [C.POINTER(PvqtParams)]),
    "pvqt_params_n_buckets": (_SZ, [C.POINTER(PvqtParams)]),
    "pvqt_filter_bank_params": (C.c_int, [C.POINTER(PvqtParams), _VP, _SZ, C.POINTER(PvqtError)]),
    "pvqt_kernel_create": (C.c_int, [C.POINTER(PvqtParams), C.POINTER(_VP), C.POINTER(PvqtError)]),
    "pvqt_kernel_destroy": (None, [_VP]),
    "pvqt_kernel_n_buckets": (_SZ, [_VP]),
    "pvqt_kernel_delay_seconds": (C.c_double, [_VP]),
    "pvqt_kernel_num_window_groups": (_SZ, [_VP]),
    "pvqt_kernel_group_window": (C.c_int, [_VP, _SZ, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pvqt_kernel_group_csr": (C.c_int, [_VP, _SZ, C.c_int, C.POINTER(PvqtCsrView)]),
    "pvqt_create": (C.c_int, [C.POINTER(PvqtParams), C.c_int, C.POINTER(_VP), C.POINTER(PvqtError)]),
    "pvqt_destroy": (None, [_VP]),
    "pvqt_get_params": (C.c_int, [_VP, C.POINTER(PvqtParams)]),
    "pvqt_n_buckets": (_SZ, [_VP]),
    "pvqt_n_fft": (_SZ, [_VP]),
    "pvqt_delay_seconds": (C.c_double, [_VP]),
    "pvqt_num_window_groups": (_SZ, [_VP]),
    "pvqt_group_window": (C.c_int, [_VP, _SZ, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pvqt_group_csr": (C.c_int, [_VP, _SZ, C.c_int, C.POINTER(PvqtCsrView)]),
    "pvqt_device": (C.c_int, [_VP]),
    "pvqt_first_sample_used": (_SZ, [_VP]),
    "pvqt_calc_instant_db": (C.c_int, [_VP, _FP, _SZ, _FP]),
    "pvqt_calc_batch_db": (C.c_int, [_VP, _FP, _SZ, _SZ, _SZ, _FP]),
    "pvqt_calc_frames_db": (C.c_int, [_VP, _FP, _SZ, _FP]),
    "pvqt_calc_streams_db": (C.c_int, [_VP, _FP, _SZ, _SZ, _SZ, _SZ, _SZ, _FP]),
    "pvqt_frames_in": (_SZ, [_VP, _SZ, _SZ]),
    "pvqt_calc_db_device": (C.c_int, [_VP, _VP, _SZ, _SZ, _SZ, _SZ, _VP, _VP, _VP]),
    "pvqt_fft_device": (C.c_int, [_VP, _VP, _SZ, _SZ, _SZ, _SZ, _VP, _VP]),
    "pvqt_spec_stride": (_SZ, [_VP]),
    "pvqt_group_columns": (C.c_int, [_VP, _SZ, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32)]),
    "pvqt_dev_alloc": (C.c_int, [C.c_int, _SZ, C.POINTER(_VP)]),
    "pvqt_dev_free": (C.c_int, [C.c_int, _VP]),
    "pvqt_host_alloc_pinned": (C.c_int, [_SZ, C.POINTER(_VP)]),
    "pvqt_host_free_pinned": (C.c_int, [_VP]),
    "pvqt_memcpy_h2d": (C.c_int, [_VP, _VP, _VP, _SZ, C.c_int]),
    "pvqt_memcpy_d2h": (C.c_int, [_VP, _VP, _VP, _SZ, C.c_int]),
    "pvqt_dev_memset": (C.c_int, [_VP, _VP, C.c_int, _SZ]),
    "pvqt_synchronize": (C.c_int, [_VP]),
    "pvqt_event_create": (C.c_int, [_VP, C.POINTER(_VP)]),
    "pvqt_event_destroy": (C.c_int, [_VP, _VP]),
    "pvqt_event_record": (C.c_int, [_VP, _VP]),
    "pvqt_event_elapsed_ms": (C.c_int, [_VP, _VP, _VP, C.POINTER(C.c_float)]),
    "pvqt_launch_count": (C.c_uint64, [_VP]),
    "pvqt_set_profiling": (C.c_int, [_VP, C.c_int]),
    "pvqt_get_profile": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "pvqt_shard_range": (C.c_int, [_SZ, _SZ, _SZ, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "pvqt_frame_range_samples": (C.c_int, [_SZ, _SZ, _SZ, _SZ, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "pvqt_multi_create": (C.c_int, [C.POINTER(PvqtParams), C.c_int, C.POINTER(C.c_int), C.POINTER(_VP),
                                    C.POINTER(PvqtError)]),
    "pvqt_multi_destroy": (None, [_VP]),
    "pvqt_multi_num_devices": (C.c_int, [_VP]),
    "pvqt_multi_handle": (_VP, [_VP, C.c_int]),
    "pvqt_multi_calc_batch_db": (C.c_int, [_VP, _FP, _SZ, _SZ, _SZ, _FP]),
    "pvqt_multi_calc_streams_db": (C.c_int, [_VP, _FP, _SZ, _SZ, _SZ, _SZ, _SZ, _FP]),
}

_lib = None


def load() -> C.CDLL:
    """Load libpvqt.so and declare every prototype.  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pitchvis_b200.build` "
            "(nvcc, sm_100a).  pitchvis_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().pvqt_last_error_string().decode("utf-8", "replace")

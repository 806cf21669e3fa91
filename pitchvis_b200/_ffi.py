"""ctypes declarations for libpvqt.so (include/pvqt.h).

The library is the product; there is no Python or CPU fallback.  Importing this module
when the shared library has not been built raises ImportError with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# PVQT_LIB: another build of the same library (kernel experiments: scripts/build_variant.sh); never a fallback
LIB_PATH = os.environ.get("PVQT_LIB") or os.path.join(_PKG, "lib", "libpvqt.so")

PROFILE_KINDS = 8  # PVQT_PROFILE_KINDS
KERNEL_KIND_NAMES = ("fft_groups_kernel", "spmm_kernel", "power_to_db_kernel", "spmm_db_fused_kernel",
                     "sdft_partial_kernel", "sdft_combine_kernel", "spmm_db_cluster_kernel", "reserved7")

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


class PvqtPeakParams(C.Structure):
    _fields_ = [("min_prominence", C.c_float), ("min_height", C.c_float)]


class PvqtAnalysisParams(C.Structure):
    _fields_ = [
        ("spectrogram_length", C.c_uint64),
        ("peak_config", PvqtPeakParams),
        ("bassline_peak_config", PvqtPeakParams),
        ("highest_bassnote", C.c_uint64),
        ("vqt_smoothing_duration_base_ns", C.c_uint64),
        ("vqt_smoothing_calmness_min", C.c_float),
        ("vqt_smoothing_calmness_max", C.c_float),
        ("note_calmness_smoothing_duration_ns", C.c_uint64),
        ("scene_calmness_smoothing_duration_ns", C.c_uint64),
        ("tuning_inaccuracy_smoothing_duration_ns", C.c_uint64),
        ("harmonic_threshold", C.c_float),
    ]


class PvqtRange(C.Structure):
    _fields_ = [("min_freq", C.c_float), ("octaves", C.c_uint32), ("buckets_per_octave", C.c_uint32)]


class PvqtAnalysisOutputs(C.Structure):
    _fields_ = [
        ("max_peaks", C.c_uint32),
        ("peak_count", C.c_void_p),
        ("peak_indices", C.c_void_p),
        ("peaks_continuous", C.c_void_p),
        ("x_vqt_smoothed", C.c_void_p),
        ("x_vqt_peakfiltered", C.c_void_p),
        ("x_vqt_afterglow", C.c_void_p),
        ("calmness", C.c_void_p),
        ("pitch_accuracy", C.c_void_p),
        ("pitch_deviation", C.c_void_p),
        ("smoothed_scene_calmness", C.c_void_p),
        ("smoothed_tuning_grid_inaccuracy", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/*.h declares
_FP = C.POINTER(C.c_float)
_VP = C.c_void_p
_SZ = C.c_size_t
SIGNATURES = {
    "pvqt_abi_version": (C.c_int, []),
    "pvqt_last_error_string": (C.c_char_p, []),
    "pvqt_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pvqt_device_attributes": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32)]),
    "pvqt_multi_pcie_probe": (C.c_int, [_VP, _SZ, _SZ, C.c_int, C.POINTER(C.c_double)]),
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
    "pvqt_kernel_filter_bandwidths": (C.c_int, [_VP, _FP, _FP, _SZ]),
    "pvqt_kernel_coverage_gaps": (C.c_int, [_VP, C.POINTER(C.c_uint32), _SZ, C.POINTER(_SZ)]),
    "pvqt_set_log_callback": (C.c_int, [_VP, _VP, C.c_int]),
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
    "pvqt_dev_flush_l2": (C.c_int, [_VP, _VP, _SZ]),
    "pvqt_synchronize": (C.c_int, [_VP]),
    "pvqt_event_create": (C.c_int, [_VP, C.POINTER(_VP)]),
    "pvqt_event_destroy": (C.c_int, [_VP, _VP]),
    "pvqt_event_record": (C.c_int, [_VP, _VP]),
    "pvqt_event_elapsed_ms": (C.c_int, [_VP, _VP, _VP, C.POINTER(C.c_float)]),
    "pvqt_launch_count": (C.c_uint64, [_VP]),
    "pvqt_set_profiling": (C.c_int, [_VP, C.c_int]),
    "pvqt_get_profile": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "pvqt_set_fused_epilogue": (C.c_int, [_VP, C.c_int]),
    "pvqt_set_sliding_dft": (C.c_int, [_VP, C.c_int]),
    "pvqt_plan_info": (C.c_int, [_VP, C.POINTER(C.c_int32), _SZ]),
    "pvqt_shard_range": (C.c_int, [_SZ, _SZ, _SZ, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "pvqt_frame_range_samples": (C.c_int, [_SZ, _SZ, _SZ, _SZ, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "pvqt_multi_create": (C.c_int, [C.POINTER(PvqtParams), C.c_int, C.POINTER(C.c_int), C.POINTER(_VP),
                                    C.POINTER(PvqtError)]),
    "pvqt_multi_destroy": (None, [_VP]),
    "pvqt_multi_num_devices": (C.c_int, [_VP]),
    "pvqt_multi_handle": (_VP, [_VP, C.c_int]),
    "pvqt_multi_calc_batch_db": (C.c_int, [_VP, _FP, _SZ, _SZ, _SZ, _FP]),
    "pvqt_multi_calc_streams_db": (C.c_int, [_VP, _FP, _SZ, _SZ, _SZ, _SZ, _SZ, _FP]),
    "pvqt_analysis_default_params": (C.c_int, [C.POINTER(PvqtAnalysisParams)]),
    "pvqt_analysis_create": (C.c_int, [C.POINTER(PvqtRange), C.POINTER(PvqtAnalysisParams), _SZ, C.c_int,
                                       C.POINTER(_VP)]),
    "pvqt_analysis_destroy": (None, [_VP]),
    "pvqt_analysis_n_buckets": (_SZ, [_VP]),
    "pvqt_analysis_n_streams": (_SZ, [_VP]),
    "pvqt_analysis_update_vqt_smoothing_duration": (C.c_int, [_VP, C.c_int, C.c_uint64]),
    "pvqt_analysis_preprocess_batch": (C.c_int, [_VP, _FP, _SZ, _SZ, C.c_uint64, C.POINTER(PvqtAnalysisOutputs)]),
    "pvqt_analysis_preprocess_device": (C.c_int, [_VP, _VP, _SZ, _SZ, C.c_uint64, C.POINTER(PvqtAnalysisOutputs),
                                                  _VP]),
    "pvqt_analysis_synchronize": (C.c_int, [_VP]),
    "pvqt_chroma": (C.c_int, [C.POINTER(PvqtRange), C.c_int, _FP, _SZ, _FP]),
    "pvqt_chroma_device": (C.c_int, [C.POINTER(PvqtRange), C.c_int, _VP, _SZ, _VP, _VP]),
    "pvqt_spectrogram_vqt": (C.c_int, [C.c_int, _FP, _SZ, _SZ, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), _SZ,
                                       C.POINTER(C.c_size_t)]),
    "pvqt_spectrogram_vqt_device": (C.c_int, [C.c_int, _VP, _SZ, _SZ, _VP, _VP, _SZ, C.POINTER(C.c_size_t), _VP]),
    "pvqt_spectrogram_peaks": (C.c_int, [C.c_int, C.POINTER(PvqtRange), _VP, _VP, _SZ, _SZ, C.POINTER(C.c_uint8), _SZ,
                                         C.POINTER(C.c_size_t)]),
    "pvqt_spectrogram_peaks_device": (C.c_int, [C.c_int, C.POINTER(PvqtRange), _VP, _VP, _SZ, _SZ, _VP, _SZ,
                                                C.POINTER(C.c_size_t), _VP]),
    "pvqt_calc_batch_analysis": (C.c_int, [_VP, _VP, _FP, _SZ, _SZ, _SZ, C.c_uint64, C.POINTER(PvqtAnalysisOutputs), _FP,
                                           C.POINTER(C.c_uint64)]),
    "pvqt_calc_streams_analysis": (C.c_int, [_VP, _VP, _FP, _SZ, _SZ, _SZ, _SZ, _SZ, C.c_uint64,
                                             C.POINTER(PvqtAnalysisOutputs), _FP, C.POINTER(C.c_uint64)]),
    # include/pvqt_agc.h
    "pvqt_agc_create": (C.c_int, [C.c_float, C.c_float, _SZ, C.c_int, C.POINTER(_VP)]),
    "pvqt_agc_destroy": (None, [_VP]),
    "pvqt_agc_n_streams": (_SZ, [_VP]),
    "pvqt_agc_gains": (C.c_int, [_VP, _FP]),
    "pvqt_agc_freeze_gain": (C.c_int, [_VP, C.c_int]),
    "pvqt_agc_process": (C.c_int, [_VP, _FP, _SZ, _SZ, _SZ, C.c_float]),
    "pvqt_agc_process_device": (C.c_int, [_VP, _VP, _SZ, _SZ, _SZ, C.c_float, _VP]),
    "pvqt_agc_synchronize": (C.c_int, [_VP]),
}

LOG_FN = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_void_p)   # pvqt_log_fn

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

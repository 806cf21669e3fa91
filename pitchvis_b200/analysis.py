"""Host-side mirror of `pitchvis_analysis::analysis` on top of the C ABI (include/pvqt_analysis.h).

`AnalysisParameters`, `PeakDetectionParameters`, `ContinuousPeak`, `AnalysisState.new/preprocess/
update_vqt_smoothing_duration` keep the reference's names (analysis.rs:36-410).  One `AnalysisState`
holds `n_streams` independent states; `preprocess_batch` advances all of them by T frames and returns
the reference's public fields after every frame.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import _ffi
from ._ffi import PvqtAnalysisOutputs, PvqtAnalysisParams, PvqtPeakParams, PvqtRange
from .vqt import PvqtRuntimeError, VqtRange

MS = 1_000_000


@dataclass
class PeakDetectionParameters:
    """peak_detection.rs:10-15"""
    min_prominence: float
    min_height: float


@dataclass
class AnalysisParameters:
    """analysis.rs:36-98 (defaults = `impl Default`); durations in nanoseconds"""
    spectrogram_length: int = 400
    peak_config: PeakDetectionParameters = field(default_factory=lambda: PeakDetectionParameters(10.0, 4.0))
    bassline_peak_config: PeakDetectionParameters = field(default_factory=lambda: PeakDetectionParameters(5.0, 3.5))
    highest_bassnote: int = 12 * 2 + 4
    vqt_smoothing_duration_base: int = 70 * MS
    vqt_smoothing_calmness_min: float = 0.6
    vqt_smoothing_calmness_max: float = 2.0
    note_calmness_smoothing_duration: int = 3_500 * MS
    scene_calmness_smoothing_duration: int = 800 * MS
    tuning_inaccuracy_smoothing_duration: int = 4_000 * MS
    harmonic_threshold: float = 0.3

    def to_c(self) -> PvqtAnalysisParams:
        return PvqtAnalysisParams(
            self.spectrogram_length,
            PvqtPeakParams(self.peak_config.min_prominence, self.peak_config.min_height),
            PvqtPeakParams(self.bassline_peak_config.min_prominence, self.bassline_peak_config.min_height),
            self.highest_bassnote, self.vqt_smoothing_duration_base, self.vqt_smoothing_calmness_min,
            self.vqt_smoothing_calmness_max, self.note_calmness_smoothing_duration,
            self.scene_calmness_smoothing_duration, self.tuning_inaccuracy_smoothing_duration,
            self.harmonic_threshold)


def _check(rc: int):
    if rc == _ffi.PVQT_OK:
        return
    if rc == _ffi.PVQT_BAD_LENGTH:
        raise ValueError(_ffi.last_error())  # the reference asserts (analysis.rs:289)
    raise PvqtRuntimeError(rc, _ffi.last_error())


VECTOR_FIELDS = ("x_vqt_smoothed", "x_vqt_peakfiltered", "x_vqt_afterglow", "calmness", "pitch_accuracy",
                 "pitch_deviation")


class AnalysisState:
    """`AnalysisState` for n_streams independent audio streams on one GPU."""

    def __init__(self, range: VqtRange, params: Optional[AnalysisParameters] = None, n_streams: int = 1,
                 device: int = 0):
        self._lib = _ffi.load()
        self.range = range
        self.params = params if params is not None else AnalysisParameters()
        self.n_streams = n_streams
        self.n_buckets = range.n_buckets()
        self._h = C.c_void_p()
        r = PvqtRange(range.min_freq, range.octaves, range.buckets_per_octave)
        p = self.params.to_c()
        _check(self._lib.pvqt_analysis_create(C.byref(r), C.byref(p), n_streams, device, C.byref(self._h)))

    @classmethod
    def new(cls, range: VqtRange, params: AnalysisParameters) -> "AnalysisState":
        return cls(range, params)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pvqt_analysis_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_vqt_smoothing_duration(self, new_duration_ns: Optional[int]):
        _check(self._lib.pvqt_analysis_update_vqt_smoothing_duration(
            self._h, 0 if new_duration_ns is None else 1, new_duration_ns or 0))

    def _result_buffers(self, S: int, T: int, max_peaks: int, vectors):
        NB = self.n_buckets
        res = {
            "peak_count": np.zeros((S, T), np.uint32),
            "peak_indices": np.zeros((S, T, max_peaks), np.uint32),
            "peaks_continuous": np.zeros((S, T, max_peaks, 2), np.float32),
            "smoothed_scene_calmness": np.zeros((S, T), np.float32),
            "smoothed_tuning_grid_inaccuracy": np.zeros((S, T), np.float32),
        }
        names = VECTOR_FIELDS if vectors is True else (tuple(vectors) if vectors else ())
        for name in names:
            if name not in VECTOR_FIELDS:
                raise ValueError(f"unknown result vector {name!r}")
            res[name] = np.zeros((S, T, NB), np.float32)
        out = PvqtAnalysisOutputs()
        out.max_peaks = max_peaks
        for name, arr in res.items():
            setattr(out, name, arr.ctypes.data_as(C.c_void_p))
        return res, out

    @staticmethod
    def _check_peak_capacity(res, max_peaks: int):
        most = int(res["peak_count"].max()) if res["peak_count"].size else 0
        if most > max_peaks:
            raise ValueError(f"a frame has {most} peaks but max_peaks is {max_peaks}: the stored peak lists are "
                             "truncated (the states have advanced); pass a larger max_peaks")

    def preprocess_batch(self, db: np.ndarray, frame_time_ns: int, max_peaks: int = 64,
                         vectors=True) -> Dict[str, np.ndarray]:
        """db: [n_streams][T][n_buckets] (or [T][n_buckets] for one stream).  Returns per-frame results.
        `vectors`: True = every per-bin result, False = none, or an iterable of names out of VECTOR_FIELDS."""
        db = np.ascontiguousarray(db, np.float32)
        if db.ndim == 2:
            db = db[None]
        if db.ndim != 3 or db.shape[0] != self.n_streams:
            raise ValueError("db must be [n_streams][n_frames][n_buckets]")
        S, T, NB = db.shape
        res, out = self._result_buffers(S, T, max_peaks, vectors)
        _check(self._lib.pvqt_analysis_preprocess_batch(
            self._h, db.ctypes.data_as(C.POINTER(C.c_float)), NB, T, frame_time_ns, C.byref(out)))
        self._check_peak_capacity(res, max_peaks)
        return res

    def calculate_and_preprocess(self, vqt, audio, hop: int, frame_time_ns: int, frames_per_stream: Optional[int] = None,
                                 max_peaks: int = 64, vectors=False, return_db: bool = False) -> Dict[str, np.ndarray]:
        """VQT + preprocess in one library call (pvqt_calc_streams_analysis, BASELINE configs[4]): host audio in
        ([n_samples] for a single-stream state, else [n_streams][n_samples]), the spectra stay in HBM between the
        transform and the analysis epilogue, only the results come back.  `res["d2h_bytes"]` is what was copied."""
        a = np.ascontiguousarray(audio, np.float32)
        if a.ndim == 1:
            a = a[None]
        if a.ndim != 2 or a.shape[0] != self.n_streams:
            raise ValueError("audio must be [n_streams][n_samples]")
        if int(hop) <= 0:
            raise ValueError("hop must be positive")
        S, n_samples = a.shape
        if frames_per_stream is None:
            frames_per_stream = vqt.frames_in(n_samples, hop)
        T = int(frames_per_stream)
        res, out = self._result_buffers(S, T, max_peaks, vectors)
        db = np.empty((S, T, self.n_buckets), np.float32) if return_db else None
        moved = C.c_uint64(0)
        _check(self._lib.pvqt_calc_streams_analysis(
            vqt.handle, self._h, a.ctypes.data_as(C.POINTER(C.c_float)), S, n_samples, n_samples, int(hop), T,
            frame_time_ns, C.byref(out), db.ctypes.data_as(C.POINTER(C.c_float)) if return_db else None, C.byref(moved)))
        self._check_peak_capacity(res, max_peaks)
        if return_db:
            res["db"] = db
        res["d2h_bytes"] = int(moved.value)
        return res

    def preprocess(self, x_vqt, frame_time_ns: int) -> Dict[str, np.ndarray]:
        """`AnalysisState::preprocess` for one frame of a single-stream state."""
        x = np.ascontiguousarray(x_vqt, np.float32)
        if x.ndim != 1:
            raise ValueError("x_vqt must be one frame")
        if self.n_streams != 1:
            raise ValueError("preprocess() is the single-stream entry; use preprocess_batch")
        if x.shape[0] != self.n_buckets:
            raise ValueError("x_vqt.len() must equal range.n_buckets()")
        r = self.preprocess_batch(x[None, None, :], frame_time_ns, max_peaks=min(256, self.n_buckets // 2 + 1))
        n = int(r["peak_count"][0, 0])
        out = {k: v[0, 0] for k, v in r.items()}
        out["peaks"] = set(int(i) for i in r["peak_indices"][0, 0, :n])
        out["peaks_continuous"] = r["peaks_continuous"][0, 0, :n]
        return out


def chroma(db: np.ndarray, range: Optional[VqtRange] = None, device: int = 0) -> np.ndarray:
    """Pitch-class energies of dB spectra (pitchvis_viewer/src/display_system/update.rs:1104-1131): [frames][12],
    the per-class sums of 10^(dB/10) divided by their maximum."""
    rng = range if range is not None else VqtRange()
    x = np.ascontiguousarray(np.atleast_2d(db), np.float32)
    if x.shape[1] != rng.n_buckets():
        raise ValueError("db must be [frames][n_buckets]")
    out = np.empty((x.shape[0], 12), np.float32)
    r = PvqtRange(rng.min_freq, rng.octaves, rng.buckets_per_octave)
    _check(_ffi.load().pvqt_chroma(C.byref(r), device, x.ctypes.data_as(C.POINTER(C.c_float)), x.shape[0],
                                   out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


def spectrogram_vqt(smoothed: np.ndarray, bin_rgb: np.ndarray, image: np.ndarray, write_index: int, device: int = 0) -> int:
    """The viewer's spectrogram ring in VQT mode (pitchvis_viewer/src/display_system/update.rs:930-1088) for all
    frames of `smoothed` [frames][n_buckets] at once.  `image` is the RGBA8 ring [height][n_buckets][4], updated in
    place; `bin_rgb` [n_buckets][3] are the per-bin colour bytes (pitchvis_colors::calculate_color, scaled as
    update.rs:986-988).  Returns the new write index."""
    x = np.ascontiguousarray(np.atleast_2d(smoothed), np.float32)
    rgb = np.ascontiguousarray(bin_rgb, np.uint8)
    if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 4 or not image.flags.c_contiguous:
        raise ValueError("image must be a C-contiguous uint8 array [height][n_buckets][4]")
    if image.shape[1] != x.shape[1] or rgb.shape != (x.shape[1], 3):
        raise ValueError("smoothed [frames][n], bin_rgb [n][3] and image [height][n][4] must agree on n")
    w = C.c_size_t(int(write_index))
    u8 = C.POINTER(C.c_uint8)
    _check(_ffi.load().pvqt_spectrogram_vqt(device, x.ctypes.data_as(C.POINTER(C.c_float)), x.shape[0], x.shape[1],
                                            rgb.ctypes.data_as(u8), image.ctypes.data_as(u8), image.shape[0], C.byref(w)))
    return int(w.value)


def spectrogram_peaks(peaks_continuous: np.ndarray, peak_count: np.ndarray, image: np.ndarray, write_index: int,
                      range: Optional[VqtRange] = None, device: int = 0) -> int:
    """The viewer's spectrogram ring in Peaks mode (pitchvis_viewer/src/display_system/update.rs:997-1062) for all frames
    at once.  `peaks_continuous` [frames][max_peaks][2] and `peak_count` [frames] are what `AnalysisState.preprocess_batch`
    returns for one stream; `image` is the RGBA8 ring [height][n_buckets][4], updated in place.  Returns the new index."""
    rng = range if range is not None else VqtRange()
    pk = np.ascontiguousarray(peaks_continuous, np.float32)
    cnt = np.ascontiguousarray(peak_count, np.uint32)
    if pk.ndim != 3 or pk.shape[2] != 2 or cnt.shape != (pk.shape[0],):
        raise ValueError("peaks_continuous must be [frames][max_peaks][2] and peak_count [frames]")
    if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 4 or not image.flags.c_contiguous or \
            image.shape[1] != rng.n_buckets():
        raise ValueError("image must be a C-contiguous uint8 array [height][n_buckets][4]")
    w = C.c_size_t(int(write_index))
    r = PvqtRange(rng.min_freq, rng.octaves, rng.buckets_per_octave)
    _check(_ffi.load().pvqt_spectrogram_peaks(device, C.byref(r), pk.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p),
                                              pk.shape[1], pk.shape[0], image.ctypes.data_as(C.POINTER(C.c_uint8)),
                                              image.shape[0], C.byref(w)))
    return int(w.value)


def ml_input_windows(history: np.ndarray, t: int) -> np.ndarray:
    """Model inputs as pitchvis_viewer/src/ml_system.rs:50-68 builds them: window i is the concatenation of frames
    i .. i + t - 1 of `history` [frames][n] (the reference takes the last t frames per call).  A zero-copy view
    [frames - t + 1][t * n] of the batched output: nothing to compute."""
    h = np.ascontiguousarray(history, np.float32)
    if h.ndim != 2 or t < 1 or h.shape[0] < t:
        raise ValueError("history must be [frames][n] with frames >= t >= 1")
    n = h.shape[1]
    return np.lib.stride_tricks.as_strided(h, shape=(h.shape[0] - t + 1, t * n), strides=(h.strides[0], h.strides[1]),
                                           writeable=False)

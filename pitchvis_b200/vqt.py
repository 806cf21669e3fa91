"""Host-side mirror of `pitchvis_analysis::vqt` on top of the C ABI (include/pvqt.h).

Same names, argument meaning and error behaviour as the reference
(pitchvis_analysis/src/vqt.rs): `VqtRange`, `VqtParameters` (+ defaults), `Vqt(params)`
(= `Vqt::new`, raising `VqtError` subclasses), `Vqt.calculate_vqt_instant_in_db`,
`Vqt.kernel()`, `Vqt.delay`, plus the new batched entry points.  Python is only the
test/bench host here; the Rust shim in rust/ and the C++ header in pitchvis_b200/cpp/
bind the same ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _ffi
from ._ffi import PvqtCsrView, PvqtError, PvqtParams

# vqt.rs:180-214
DEFAULT_SR = 22050
DEFAULT_N_FFT = 2 * 16384
DEFAULT_MIN_FREQ = 55.0
DEFAULT_UPSCALE_FACTOR = 1
DEFAULT_BUCKETS_PER_SEMITONE = 7 * DEFAULT_UPSCALE_FACTOR
DEFAULT_BUCKETS_PER_OCTAVE = 12 * DEFAULT_BUCKETS_PER_SEMITONE
DEFAULT_OCTAVES = 7
DEFAULT_SPARSITY_QUANTILE = 0.999
# f32 constants in the reference: evaluate the products in f32, not in Python's f64
DEFAULT_Q = float(np.float32(1.6) / np.float32(DEFAULT_UPSCALE_FACTOR))
DEFAULT_GAMMA = float(np.float32(4.8) * np.float32(DEFAULT_Q))


@dataclass
class VqtRange:
    """vqt.rs:239-262"""
    min_freq: float = DEFAULT_MIN_FREQ
    octaves: int = DEFAULT_OCTAVES
    buckets_per_octave: int = DEFAULT_BUCKETS_PER_OCTAVE

    def n_buckets(self) -> int:
        return self.buckets_per_octave * self.octaves


@dataclass
class VqtParameters:
    """vqt.rs:279-348 (field defaults = `impl Default`)"""
    sr: float = float(DEFAULT_SR)
    n_fft: int = DEFAULT_N_FFT
    range: VqtRange = field(default_factory=VqtRange)
    sparsity_quantile: float = DEFAULT_SPARSITY_QUANTILE
    quality: float = DEFAULT_Q
    gamma: float = DEFAULT_GAMMA

    @staticmethod
    def default() -> "VqtParameters":
        return VqtParameters()

    @staticmethod
    def hires() -> "VqtParameters":
        """The 'hi-res' configuration of SURVEY.md section 8 (the crate's UPSCALE_FACTOR=2 convention)."""
        return VqtParameters(sr=44100.0, n_fft=65536, range=VqtRange(55.0, 8, 168), sparsity_quantile=0.999,
                             quality=0.8, gamma=3.84)

    def to_c(self) -> PvqtParams:
        return PvqtParams(self.sr, self.n_fft, self.range.min_freq, self.range.octaves,
                          self.range.buckets_per_octave, self.sparsity_quantile, self.quality, self.gamma)


class VqtError(Exception):
    """vqt.rs:352-366"""


class AboveNyquist(VqtError):
    def __init__(self, highest_frequency: float, nyquist_frequency: float):
        self.highest_frequency = highest_frequency
        self.nyquist_frequency = nyquist_frequency
        super().__init__(
            f"the highest VQT bin frequency ({highest_frequency} Hz) exceeds the Nyquist frequency "
            f"({nyquist_frequency} Hz); reduce octaves or increase the sample rate")


class WindowExceedsNFft(VqtError):
    def __init__(self, window_length: float, n_fft: int):
        self.window_length = window_length
        self.n_fft = n_fft
        super().__init__(
            f"the longest filter window ({window_length} samples) exceeds n_fft ({n_fft} samples); "
            "increase n_fft or gamma, or decrease quality")


class PvqtRuntimeError(RuntimeError):
    """Any non-reference failure of the library (CUDA errors, unsupported sizes, bad arguments)."""

    def __init__(self, status: int, message: str):
        self.status = status
        super().__init__(f"pvqt status {status}: {message}")


@dataclass
class CsMat:
    """sprs::CsMat<Complex32> view (rows = filters, cols = half spectrum)."""
    rows: int
    cols: int
    indptr: np.ndarray
    indices: np.ndarray
    data: np.ndarray  # complex64

    def nnz(self) -> int:
        return int(self.indices.shape[0])


@dataclass
class WindowGroup:
    """vqt.rs:388-410"""
    window: Tuple[int, int]
    filter_bank: CsMat
    negative_filter_bank: Optional[CsMat]

    def window_size(self) -> int:
        return self.window[1] - self.window[0]


@dataclass
class VqtKernel:
    """vqt.rs:413-415"""
    window_groups: List[WindowGroup]


def _check(rc: int):
    if rc == _ffi.PVQT_OK:
        return
    msg = _ffi.last_error()
    if rc == _ffi.PVQT_BAD_LENGTH:
        # the reference panics here (vqt.rs:867-871)
        raise ValueError(msg)
    raise PvqtRuntimeError(rc, msg)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _as_f32(a, name: str) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 0:
        raise ValueError(f"{name} must be an array")
    return a


def _check_out(out: Optional[np.ndarray], shape) -> np.ndarray:
    """A caller-supplied result buffer goes to the library as a raw pointer: it must be exactly what gets written."""
    if out is None:
        return np.empty(shape, np.float32)
    if not isinstance(out, np.ndarray) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or \
            tuple(out.shape) != tuple(shape):
        raise ValueError(f"out must be a C-contiguous float32 array of shape {tuple(shape)}")
    return out


def _check_hop(hop) -> int:
    hop = int(hop)
    if hop <= 0:
        raise ValueError("hop must be positive")
    return hop


def _csmat(view: PvqtCsrView) -> CsMat:
    nnz = int(view.nnz)
    indptr = np.ctypeslib.as_array(view.indptr, shape=(view.rows + 1,)).copy()
    if nnz:
        indices = np.ctypeslib.as_array(view.indices, shape=(nnz,)).copy()
        data = np.ctypeslib.as_array(view.data, shape=(2 * nnz,)).copy().view(np.complex64)
    else:
        indices = np.zeros(0, np.int32)
        data = np.zeros(0, np.complex64)
    return CsMat(int(view.rows), int(view.cols), indptr, indices, data)


class Vqt:
    """`pitchvis_analysis::vqt::Vqt` backed by the sm_100a kernels (vqt.rs:440-917)."""

    def __init__(self, params: Optional[VqtParameters] = None, device: int = 0):
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        self._params = params if params is not None else VqtParameters.default()
        cp = self._params.to_c()
        err = PvqtError()
        rc = self._lib.pvqt_create(C.byref(cp), device, C.byref(self._h), C.byref(err))
        if rc == _ffi.PVQT_ABOVE_NYQUIST:
            raise AboveNyquist(err.highest_frequency, err.nyquist_frequency)
        if rc == _ffi.PVQT_WINDOW_EXCEEDS_NFFT:
            raise WindowExceedsNFft(err.window_length, int(err.n_fft))
        if rc != _ffi.PVQT_OK:
            raise PvqtRuntimeError(rc, _ffi.last_error())
        self.delay = float(self._lib.pvqt_delay_seconds(self._h))  # seconds (Duration in the reference)
        self.n_buckets = int(self._lib.pvqt_n_buckets(self._h))
        self.n_fft = int(self._lib.pvqt_n_fft(self._h))
        self.device = device

    # Vqt::new spelled as in the reference
    @classmethod
    def new(cls, params: VqtParameters, device: int = 0) -> "Vqt":
        return cls(params, device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pvqt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def params(self) -> VqtParameters:
        return self._params

    def kernel(self) -> VqtKernel:
        groups = []
        for g in range(int(self._lib.pvqt_num_window_groups(self._h))):
            b, e = C.c_uint64(), C.c_uint64()
            _check(self._lib.pvqt_group_window(self._h, g, C.byref(b), C.byref(e)))
            pos, neg = PvqtCsrView(), PvqtCsrView()
            _check(self._lib.pvqt_group_csr(self._h, g, 0, C.byref(pos)))
            _check(self._lib.pvqt_group_csr(self._h, g, 1, C.byref(neg)))
            negm = _csmat(neg) if neg.nnz > 0 else None  # vqt.rs:751
            groups.append(WindowGroup((int(b.value), int(e.value)), _csmat(pos), negm))
        return VqtKernel(groups)

    def group_columns(self, g: int) -> Tuple[int, int, int]:
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(self._lib.pvqt_group_columns(self._h, g, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    @property
    def spec_stride(self) -> int:
        return int(self._lib.pvqt_spec_stride(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.pvqt_launch_count(self._h))

    def set_fused_epilogue(self, mode: int) -> int:
        """Tuning / test switch: 0 = unfused K-spmm + K-db, 1 = K-spmm-db one CTA per tile, 2 = cluster form,
        3 = persistent warp-specialised pipeline (default where it fits).  Returns the mode in effect (a mode the
        kernel does not fit falls back: 3 -> 1 -> 0)."""
        return int(self._lib.pvqt_set_fused_epilogue(self._h, int(mode)))

    def plan_info(self) -> dict:
        out = (C.c_int32 * 10)()
        _check(self._lib.pvqt_plan_info(self._h, out, 10))
        keys = ("cluster_size", "clusters_resident", "cluster_coef_bytes", "cluster_max_rows", "fused_warps",
                "fft_block_threads", "spec_stride", "sdft_plans", "fused_walk_slots", "sdft_group_mask")
        return dict(zip(keys, [int(x) for x in out]))

    def set_sliding_dft(self, mode) -> int:
        """Tuning / test switch: 0 / False keeps every window group on the per-frame FFT path in batched calls,
        1 = K-sdft with the partial sums on the FP32 pipe, 2 / True (default) = on the tensor cores (mma.sync),
        3 = on the tensor cores through tcgen05 / TMEM (opt-in; slower than 2 in round 1)."""
        m = 2 if mode is True else (0 if mode is False else int(mode))
        return int(self._lib.pvqt_set_sliding_dft(self._h, m))

    # ---- per-frame entry point (vqt.rs:866) ----------------------------------------------
    def calculate_vqt_instant_in_db(self, x) -> np.ndarray:
        x = _as_f32(x, "x")
        if x.ndim != 1:
            raise ValueError("input must be exactly n_fft samples")
        out = np.empty(self.n_buckets, np.float32)
        _check(self._lib.pvqt_calc_instant_db(self._h, _fptr(x), x.shape[0], _fptr(out)))
        return out

    # ---- batched entry points (new) --------------------------------------------------------
    def frames_in(self, n_samples: int, hop: int) -> int:
        return int(self._lib.pvqt_frames_in(self._h, n_samples, hop))

    def calculate_vqt_batch_in_db(self, audio, hop: int, n_frames: Optional[int] = None,
                                  out: Optional[np.ndarray] = None) -> np.ndarray:
        """frame t = audio[t*hop : t*hop + n_fft]"""
        audio = _as_f32(audio, "audio")
        if audio.ndim != 1:
            raise ValueError("audio must be one recording [n_samples] (use calculate_vqt_streams_in_db for several)")
        hop = _check_hop(hop)
        if n_frames is None:
            n_frames = self.frames_in(audio.shape[0], hop)
        out = _check_out(out, (int(n_frames), self.n_buckets))
        _check(self._lib.pvqt_calc_batch_db(self._h, _fptr(audio), audio.shape[0], hop, n_frames, _fptr(out)))
        return out

    def calculate_vqt_frames_in_db(self, frames) -> np.ndarray:
        frames = _as_f32(frames, "frames")
        if frames.ndim != 2 or frames.shape[1] != self.n_fft:
            raise ValueError("input must be exactly n_fft samples")
        out = np.empty((frames.shape[0], self.n_buckets), np.float32)
        _check(self._lib.pvqt_calc_frames_db(self._h, _fptr(frames), frames.shape[0], _fptr(out)))
        return out

    def calculate_vqt_streams_in_db(self, audio, hop: int, frames_per_stream: Optional[int] = None,
                                    out: Optional[np.ndarray] = None) -> np.ndarray:
        """audio[n_streams][n_samples] -> out[n_streams][frames_per_stream][n_buckets]"""
        audio = _as_f32(audio, "audio")
        if audio.ndim != 2:
            raise ValueError("audio must be [n_streams][n_samples]")
        n_streams, n_samples = audio.shape
        hop = _check_hop(hop)
        if frames_per_stream is None:
            frames_per_stream = self.frames_in(n_samples, hop)
        out = _check_out(out, (n_streams, int(frames_per_stream), self.n_buckets))
        _check(self._lib.pvqt_calc_streams_db(self._h, _fptr(audio), n_streams, n_samples, n_samples, hop,
                                              frames_per_stream, _fptr(out)))
        return out


class DeviceBuffer:
    """A raw device allocation made through the C ABI (no torch)."""

    def __init__(self, vqt: Vqt, nbytes: int):
        self._lib = _ffi.load()
        self._vqt = vqt
        self.nbytes = int(nbytes)
        self.ptr = C.c_void_p()
        _check(self._lib.pvqt_dev_alloc(vqt.device, self.nbytes, C.byref(self.ptr)))

    def upload(self, host: np.ndarray, sync: bool = True):
        host = np.ascontiguousarray(host)
        assert host.nbytes <= self.nbytes
        _check(self._lib.pvqt_memcpy_h2d(self._vqt.handle, self.ptr, host.ctypes.data_as(C.c_void_p), host.nbytes,
                                         0 if sync else 1))

    def download(self, shape, dtype=np.float32) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        _check(self._lib.pvqt_memcpy_d2h(self._vqt.handle, out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes, 0))
        return out

    def free(self):
        if self.ptr and self.ptr.value:
            self._lib.pvqt_dev_free(self._vqt.device, self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def calc_db_device(vqt: Vqt, d_audio: DeviceBuffer, n_streams: int, stream_stride: int, hop: int,
                   frames_per_stream: int, d_out: DeviceBuffer, d_power: Optional[DeviceBuffer] = None):
    """Device-resident entry (pvqt_calc_db_device); asynchronous on the handle's stream."""
    lib = _ffi.load()
    _check(lib.pvqt_calc_db_device(vqt.handle, d_audio.ptr, n_streams, stream_stride, hop, frames_per_stream,
                                   d_out.ptr, d_power.ptr if d_power is not None else None, None))


def fft_device(vqt: Vqt, d_audio: DeviceBuffer, n_streams: int, stream_stride: int, hop: int,
               frames_per_stream: int, d_spec: DeviceBuffer):
    lib = _ffi.load()
    _check(lib.pvqt_fft_device(vqt.handle, d_audio.ptr, n_streams, stream_stride, hop, frames_per_stream,
                               d_spec.ptr, None))


def synchronize(vqt: Vqt):
    _check(_ffi.load().pvqt_synchronize(vqt.handle))


class MultiVqt:
    """One Vqt replicated over several GPUs of one box, single process (pvqt_multi_*)."""

    def __init__(self, params: Optional[VqtParameters] = None, devices: Optional[List[int]] = None):
        self._lib = _ffi.load()
        self._params = params if params is not None else VqtParameters.default()
        if devices is None:
            n = C.c_int()
            _check(self._lib.pvqt_device_count(C.byref(n)))
            devices = list(range(n.value))
        self.devices = list(devices)
        ids = (C.c_int * len(self.devices))(*self.devices)
        self._h = C.c_void_p()
        err = PvqtError()
        cp = self._params.to_c()
        rc = self._lib.pvqt_multi_create(C.byref(cp), len(self.devices), ids, C.byref(self._h), C.byref(err))
        if rc == _ffi.PVQT_ABOVE_NYQUIST:
            raise AboveNyquist(err.highest_frequency, err.nyquist_frequency)
        if rc == _ffi.PVQT_WINDOW_EXCEEDS_NFFT:
            raise WindowExceedsNFft(err.window_length, int(err.n_fft))
        if rc != _ffi.PVQT_OK:
            raise PvqtRuntimeError(rc, _ffi.last_error())
        self.n_buckets = self._params.range.n_buckets()
        self.n_fft = self._params.n_fft

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pvqt_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def device_handle(self, index: int) -> C.c_void_p:
        return C.c_void_p(self._lib.pvqt_multi_handle(self._h, index))

    def pcie_probe(self, h2d_bytes: int, d2h_bytes: int, reps: int = 5) -> float:
        """Seconds for `reps` x (h2d_bytes in + d2h_bytes out) on every device at once (pvqt_multi_pcie_probe)."""
        s = C.c_double()
        _check(self._lib.pvqt_multi_pcie_probe(self._h, h2d_bytes, d2h_bytes, reps, C.byref(s)))
        return float(s.value)

    def calculate_vqt_batch_in_db(self, audio, hop: int, n_frames: Optional[int] = None) -> np.ndarray:
        audio = _as_f32(audio, "audio")
        if audio.ndim != 1:
            raise ValueError("audio must be one recording [n_samples]")
        hop = _check_hop(hop)
        if n_frames is None:
            n_frames = (audio.shape[0] - self.n_fft) // hop + 1 if audio.shape[0] >= self.n_fft else 0
        out = np.empty((n_frames, self.n_buckets), np.float32)
        _check(self._lib.pvqt_multi_calc_batch_db(self._h, _fptr(audio), audio.shape[0], hop, n_frames, _fptr(out)))
        return out

    def calculate_vqt_streams_in_db(self, audio, hop: int, frames_per_stream: Optional[int] = None) -> np.ndarray:
        audio = _as_f32(audio, "audio")
        if audio.ndim != 2:
            raise ValueError("audio must be [n_streams][n_samples]")
        n_streams, n_samples = audio.shape
        hop = _check_hop(hop)
        if frames_per_stream is None:
            frames_per_stream = (n_samples - self.n_fft) // hop + 1 if n_samples >= self.n_fft else 0
        out = np.empty((n_streams, frames_per_stream, self.n_buckets), np.float32)
        _check(self._lib.pvqt_multi_calc_streams_db(self._h, _fptr(audio), n_streams, n_samples, n_samples, hop,
                                                    frames_per_stream, _fptr(out)))
        return out


def _raise_build_error(rc: int, err: PvqtError):
    if rc == _ffi.PVQT_ABOVE_NYQUIST:
        raise AboveNyquist(err.highest_frequency, err.nyquist_frequency)
    if rc == _ffi.PVQT_WINDOW_EXCEEDS_NFFT:
        raise WindowExceedsNFft(err.window_length, int(err.n_fft))
    raise PvqtRuntimeError(rc, _ffi.last_error())


def filter_bank_params(params: VqtParameters):
    """`Vqt::filter_bank_params` (vqt.rs:517-587): list of (freq, window_length, factor, min_window)."""
    lib = _ffi.load()
    n = params.range.n_buckets()
    arr = (_ffi.PvqtFilterParams * n)()
    err = PvqtError()
    cp = params.to_c()
    rc = lib.pvqt_filter_bank_params(C.byref(cp), arr, n, C.byref(err))
    if rc != _ffi.PVQT_OK:
        _raise_build_error(rc, err)
    return [(f.freq, f.window_length, int(f.sr_downscaling_factor), int(f.minimum_needed_window_size)) for f in arr]


class HostKernel:
    """The VqtKernel + delay built on the host only (pvqt_kernel_*): no GPU needed."""

    def __init__(self, params: Optional[VqtParameters] = None):
        self._lib = _ffi.load()
        self._params = params if params is not None else VqtParameters.default()
        self._h = C.c_void_p()
        err = PvqtError()
        cp = self._params.to_c()
        rc = self._lib.pvqt_kernel_create(C.byref(cp), C.byref(self._h), C.byref(err))
        if rc != _ffi.PVQT_OK:
            _raise_build_error(rc, err)
        self.delay = float(self._lib.pvqt_kernel_delay_seconds(self._h))
        self.n_buckets = int(self._lib.pvqt_kernel_n_buckets(self._h))

    def __del__(self):
        try:
            if self._h.value:
                self._lib.pvqt_kernel_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def filter_bandwidths(self):
        """(-3 dB band start, end) in Hz of every filter: the reference's calculate_bandwidth diagnostic (vqt.rs:962-989)."""
        lo, hi = np.empty(self.n_buckets, np.float32), np.empty(self.n_buckets, np.float32)
        _check(self._lib.pvqt_kernel_filter_bandwidths(self._h, _fptr(lo), _fptr(hi), self.n_buckets))
        return lo, hi

    def coverage_gaps(self) -> List[int]:
        """Filters below which the reference warns about a coverage gap (vqt.rs:695-710)."""
        n = C.c_size_t(0)
        _check(self._lib.pvqt_kernel_coverage_gaps(self._h, None, 0, C.byref(n)))
        out = (C.c_uint32 * max(1, n.value))()
        _check(self._lib.pvqt_kernel_coverage_gaps(self._h, out, n.value, C.byref(n)))
        return [int(out[i]) for i in range(n.value)]

    def kernel(self) -> VqtKernel:
        groups = []
        for g in range(int(self._lib.pvqt_kernel_num_window_groups(self._h))):
            b, e = C.c_uint64(), C.c_uint64()
            _check(self._lib.pvqt_kernel_group_window(self._h, g, C.byref(b), C.byref(e)))
            pos, neg = PvqtCsrView(), PvqtCsrView()
            _check(self._lib.pvqt_kernel_group_csr(self._h, g, 0, C.byref(pos)))
            _check(self._lib.pvqt_kernel_group_csr(self._h, g, 1, C.byref(neg)))
            groups.append(WindowGroup((int(b.value), int(e.value)), _csmat(pos), _csmat(neg) if neg.nnz > 0 else None))
        return VqtKernel(groups)


_log_keepalive = []


def set_log_callback(fn, max_level: int = 2):
    """Install `fn(level, message)` as the library's log sink (1 = warn, 2 = info, 3 = debug), or None to remove it.
    Mirrors the `log` crate lines of the reference: delay (info), coverage gaps (warn), kernel structure (debug)."""
    lib = _ffi.load()
    if fn is None:
        lib.pvqt_set_log_callback(None, None, 0)
        _log_keepalive.clear()
        return
    cb = _ffi.LOG_FN(lambda level, msg, _user: fn(int(level), msg.decode("utf-8", "replace")))
    _log_keepalive[:] = [cb]
    lib.pvqt_set_log_callback(C.cast(cb, C.c_void_p), None, int(max_level))

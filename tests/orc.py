"""ctypes binding of the CPU oracle (oracle/liboracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by pitchvis_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")


class OrcParams(C.Structure):
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


class OrcError(C.Structure):
    _fields_ = [("code", C.c_int), ("a", C.c_float), ("b", C.c_float), ("n", C.c_uint64)]


class OrcFilterParams(C.Structure):
    _fields_ = [
        ("freq", C.c_float),
        ("window_length", C.c_float),
        ("sr_downscaling_factor", C.c_uint64),
        ("minimum_needed_window_size", C.c_uint64),
    ]


class OrcCsr(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("nnz", C.c_int64),
        ("indptr", C.POINTER(C.c_int32)),
        ("indices", C.POINTER(C.c_int32)),
        ("data", C.POINTER(C.c_float)),
    ]


class OrcGroup(C.Structure):
    _fields_ = [
        ("window_begin", C.c_uint64),
        ("window_end", C.c_uint64),
        ("filter_bank", OrcCsr),
        ("negative_filter_bank", OrcCsr),
    ]


ORC_OK, ORC_ABOVE_NYQUIST, ORC_WINDOW_EXCEEDS_NFFT, ORC_ASSERT, ORC_BAD_LENGTH = range(5)

_lib = None


def _cpu_tag() -> str:
    """Identifies the host CPU the oracle was compiled for (-march=native): model name + feature flags."""
    try:
        with open("/proc/cpuinfo") as fh:
            txt = fh.read()
        model = next((ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith("model name")), "?")
        flags = next((ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith("flags")), "")
        return model + " | " + " ".join(sorted(flags.split()))
    except OSError:
        return "unknown"


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so if missing, stale, or built with -march=native on a different CPU (the built
    file travels to the GPU box, whose host CPU may differ).  gcc, a few seconds."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    tag_path = _LIB_PATH + ".cpu"
    stale = not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    try:
        with open(tag_path) as fh:
            same_cpu = fh.read() == _cpu_tag()
    except OSError:
        same_cpu = False
    if force or stale or not same_cpu:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
        with open(tag_path, "w") as fh:
            fh.write(_cpu_tag())
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    fp = C.POINTER(C.c_float)
    L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
    L.orc_default_params.restype = None
    L.orc_filter_bank_params.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcFilterParams), C.POINTER(OrcError)]
    L.orc_vqt_new.argtypes = [C.POINTER(OrcParams), C.POINTER(C.c_void_p), C.POINTER(OrcError)]
    L.orc_vqt_free.argtypes = [C.c_void_p]
    L.orc_vqt_free.restype = None
    L.orc_n_buckets.argtypes = [C.c_void_p]
    L.orc_n_buckets.restype = C.c_size_t
    L.orc_delay_seconds.argtypes = [C.c_void_p]
    L.orc_delay_seconds.restype = C.c_double
    L.orc_num_groups.argtypes = [C.c_void_p]
    L.orc_num_groups.restype = C.c_size_t
    L.orc_group_at.argtypes = [C.c_void_p, C.c_size_t]
    L.orc_group_at.restype = C.POINTER(OrcGroup)
    L.orc_vqt_set_group.argtypes = [
        C.c_void_p, C.c_size_t, C.c_int, C.c_int32, C.c_int32, C.c_int64,
        C.POINTER(C.c_int32), C.POINTER(C.c_int32), fp,
    ]
    L.orc_calc_instant_db.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_int, fp, fp]
    L.orc_calc_batch_db.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, fp]
    L.orc_power_to_db.argtypes = [fp, C.c_size_t, fp]
    L.orc_power_to_db.restype = None
    L.orc_test_create_sines.argtypes = [C.POINTER(OrcParams), fp, C.c_size_t, C.c_float, fp]
    L.orc_test_create_sines.restype = None
    L.orc_max_threads.restype = C.c_int
    _lib = L
    return L


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_params() -> OrcParams:
    p = OrcParams()
    lib().orc_default_params(C.byref(p))
    return p


def make_params(**kw) -> OrcParams:
    p = default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def hires_params() -> OrcParams:
    """SURVEY.md section 8 'hi-res' configuration (config 4)."""
    return make_params(sr=44100.0, n_fft=65536, min_freq=55.0, octaves=8, buckets_per_octave=168,
                       quality=0.8, gamma=3.84, sparsity_quantile=0.999)


class OracleError(Exception):
    def __init__(self, err: OrcError):
        self.code, self.a, self.b, self.n = err.code, err.a, err.b, err.n
        super().__init__(f"oracle error code={err.code} a={err.a} b={err.b} n={err.n}")


@dataclass
class CsrView:
    rows: int
    cols: int
    indptr: np.ndarray
    indices: np.ndarray
    data: np.ndarray  # complex64

    @property
    def nnz(self) -> int:
        return int(self.indices.shape[0])


def _csr_view(m: OrcCsr) -> CsrView:
    nnz = int(m.nnz)
    indptr = np.ctypeslib.as_array(m.indptr, shape=(m.rows + 1,)).copy()
    if nnz:
        indices = np.ctypeslib.as_array(m.indices, shape=(nnz,)).copy()
        data = np.ctypeslib.as_array(m.data, shape=(2 * nnz,)).copy().view(np.complex64)
    else:
        indices = np.zeros(0, np.int32)
        data = np.zeros(0, np.complex64)
    return CsrView(int(m.rows), int(m.cols), indptr, indices, data)


class OracleVqt:
    """Mirror of pitchvis_analysis::vqt::Vqt on the CPU oracle."""

    def __init__(self, params: OrcParams | None = None):
        self.params = params if params is not None else default_params()
        self._h = C.c_void_p()
        err = OrcError()
        rc = lib().orc_vqt_new(C.byref(self.params), C.byref(self._h), C.byref(err))
        if rc != ORC_OK:
            raise OracleError(err)
        self.n_buckets = int(lib().orc_n_buckets(self._h))
        self.n_fft = int(self.params.n_fft)
        self.delay = float(lib().orc_delay_seconds(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().orc_vqt_free(self._h)
            self._h = C.c_void_p()

    @property
    def num_groups(self) -> int:
        return int(lib().orc_num_groups(self._h))

    def group(self, g: int):
        grp = lib().orc_group_at(self._h, g).contents
        return (int(grp.window_begin), int(grp.window_end)), _csr_view(grp.filter_bank), _csr_view(
            grp.negative_filter_bank)

    def set_group(self, g: int, neg: bool, rows, cols, indptr, indices, data_c64):
        indptr = np.ascontiguousarray(indptr, np.int32)
        indices = np.ascontiguousarray(indices, np.int32)
        data = np.ascontiguousarray(data_c64, np.complex64).view(np.float32)
        rc = lib().orc_vqt_set_group(
            self._h, g, int(neg), rows, cols, indices.shape[0],
            indptr.ctypes.data_as(C.POINTER(C.c_int32)), indices.ctypes.data_as(C.POINTER(C.c_int32)),
            _fptr(data))
        assert rc == ORC_OK

    def calculate_vqt_instant_in_db(self, x: np.ndarray, mode: int = 0, return_power: bool = False):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(self.n_buckets, np.float32)
        pw = np.empty(self.n_buckets, np.float32)
        rc = lib().orc_calc_instant_db(self._h, _fptr(x), x.shape[0], mode, _fptr(out), _fptr(pw))
        if rc == ORC_BAD_LENGTH:
            raise ValueError("input must be exactly n_fft samples")
        assert rc == ORC_OK
        return (out, pw) if return_power else out

    def calculate_batch_db(self, audio: np.ndarray, hop: int, n_frames: int | None = None, mode: int = 0,
                           n_threads: int = 0) -> np.ndarray:
        audio = np.ascontiguousarray(audio, np.float32)
        if n_frames is None:
            n_frames = (audio.shape[0] - self.n_fft) // hop + 1 if audio.shape[0] >= self.n_fft else 0
        out = np.empty((n_frames, self.n_buckets), np.float32)
        rc = lib().orc_calc_batch_db(self._h, _fptr(audio), audio.shape[0], hop, n_frames, mode, n_threads,
                                     _fptr(out))
        if rc == ORC_BAD_LENGTH:
            raise ValueError("audio too short for the requested frames")
        assert rc == ORC_OK
        return out


def filter_bank_params(params: OrcParams):
    nb = params.octaves * params.buckets_per_octave
    arr = (OrcFilterParams * nb)()
    err = OrcError()
    rc = lib().orc_filter_bank_params(C.byref(params), arr, C.byref(err))
    if rc != ORC_OK:
        raise OracleError(err)
    return arr


def power_to_db(power: np.ndarray) -> np.ndarray:
    power = np.ascontiguousarray(power, np.float32)
    out = np.empty_like(power)
    lib().orc_power_to_db(_fptr(power), power.shape[0], _fptr(out))
    return out


def test_create_sines(params: OrcParams, freqs, t_diff: float = 0.0) -> np.ndarray:
    f = np.ascontiguousarray(freqs, np.float32)
    wave = np.empty(int(params.n_fft), np.float32)
    lib().orc_test_create_sines(C.byref(params), _fptr(f), f.shape[0], t_diff, _fptr(wave))
    return wave


test_create_sines.__test__ = False  # not a pytest test


# ---------------------------------------------------------------------------------------------------
# analysis oracle (oracle/analysis_oracle.c)
# ---------------------------------------------------------------------------------------------------
class OrcPeakParams(C.Structure):
    _fields_ = [("min_prominence", C.c_float), ("min_height", C.c_float)]


class OrcAnalysisParams(C.Structure):
    _fields_ = [
        ("spectrogram_length", C.c_uint64),
        ("peak_config", OrcPeakParams),
        ("bassline_peak_config", OrcPeakParams),
        ("highest_bassnote", C.c_uint64),
        ("vqt_smoothing_duration_base_ns", C.c_uint64),
        ("vqt_smoothing_calmness_min", C.c_float),
        ("vqt_smoothing_calmness_max", C.c_float),
        ("note_calmness_smoothing_duration_ns", C.c_uint64),
        ("scene_calmness_smoothing_duration_ns", C.c_uint64),
        ("tuning_inaccuracy_smoothing_duration_ns", C.c_uint64),
        ("harmonic_threshold", C.c_float),
    ]


class OrcContinuousPeak(C.Structure):
    _fields_ = [("center", C.c_float), ("size", C.c_float)]


_alib = None


def alib() -> C.CDLL:
    global _alib
    if _alib is not None:
        return _alib
    L = lib()
    fp = C.POINTER(C.c_float)
    u32p = C.POINTER(C.c_uint32)
    L.orc_analysis_default_params.argtypes = [C.POINTER(OrcAnalysisParams)]
    L.orc_analysis_default_params.restype = None
    L.orc_analysis_new.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.POINTER(OrcAnalysisParams)]
    L.orc_analysis_new.restype = C.c_void_p
    L.orc_analysis_free.argtypes = [C.c_void_p]
    L.orc_analysis_free.restype = None
    L.orc_analysis_update_vqt_smoothing_duration.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    L.orc_analysis_update_vqt_smoothing_duration.restype = None
    L.orc_analysis_preprocess.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_uint64]
    L.orc_analysis_n_buckets.argtypes = [C.c_void_p]
    L.orc_analysis_n_buckets.restype = C.c_size_t
    L.orc_analysis_peaks.argtypes = [C.c_void_p, u32p, C.c_size_t]
    L.orc_analysis_peaks.restype = C.c_size_t
    L.orc_analysis_peaks_continuous.argtypes = [C.c_void_p, C.POINTER(OrcContinuousPeak), C.c_size_t]
    L.orc_analysis_peaks_continuous.restype = C.c_size_t
    L.orc_analysis_vectors.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, fp]
    L.orc_analysis_vectors.restype = None
    L.orc_analysis_scene_calmness.argtypes = [C.c_void_p]
    L.orc_analysis_scene_calmness.restype = C.c_float
    L.orc_analysis_tuning_inaccuracy.argtypes = [C.c_void_p]
    L.orc_analysis_tuning_inaccuracy.restype = C.c_float
    L.orc_find_peaks.argtypes = [fp, C.c_size_t, C.c_float, C.c_float, C.c_uint32, C.c_int, u32p, C.c_size_t]
    L.orc_find_peaks.restype = C.c_size_t
    L.orc_ema_update.argtypes = [C.c_float, C.c_int, C.c_uint64, C.c_float, C.c_uint64]
    L.orc_ema_update.restype = C.c_float
    _alib = L
    return L


def analysis_default_params() -> OrcAnalysisParams:
    p = OrcAnalysisParams()
    alib().orc_analysis_default_params(C.byref(p))
    return p


def find_peaks(x: np.ndarray, min_prominence: float, min_height: float, buckets_per_octave: int,
               order: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.shape[0], np.uint32)
    n = alib().orc_find_peaks(_fptr(x), x.shape[0], min_prominence, min_height, buckets_per_octave, order,
                              out.ctypes.data_as(C.POINTER(C.c_uint32)), out.shape[0])
    return out[:n].copy()


def ema_update(y: float, horizon_ns, new_value: float, timestep_ns: int) -> float:
    return float(alib().orc_ema_update(y, 0 if horizon_ns is None else 1, horizon_ns or 0, new_value, timestep_ns))


class OracleAnalysisState:
    """Mirror of pitchvis_analysis::analysis::AnalysisState on the CPU oracle."""

    def __init__(self, min_freq=55.0, octaves=7, buckets_per_octave=84, params: OrcAnalysisParams | None = None):
        self.params = params if params is not None else analysis_default_params()
        self._h = C.c_void_p(alib().orc_analysis_new(min_freq, octaves, buckets_per_octave, C.byref(self.params)))
        self.n = int(alib().orc_analysis_n_buckets(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            alib().orc_analysis_free(self._h)
            self._h = C.c_void_p()

    def update_vqt_smoothing_duration(self, duration_ns):
        alib().orc_analysis_update_vqt_smoothing_duration(self._h, 0 if duration_ns is None else 1, duration_ns or 0)

    def preprocess(self, x_vqt: np.ndarray, frame_time_ns: int):
        x = np.ascontiguousarray(x_vqt, np.float32)
        rc = alib().orc_analysis_preprocess(self._h, _fptr(x), x.shape[0], frame_time_ns)
        if rc == ORC_BAD_LENGTH:
            raise ValueError("x_vqt.len() must equal range.n_buckets()")

    @property
    def peaks(self) -> np.ndarray:
        out = np.empty(self.n, np.uint32)
        m = alib().orc_analysis_peaks(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), self.n)
        return out[:m].copy()

    @property
    def peaks_continuous(self) -> np.ndarray:
        arr = (OrcContinuousPeak * self.n)()
        m = alib().orc_analysis_peaks_continuous(self._h, arr, self.n)
        return np.array([(arr[i].center, arr[i].size) for i in range(m)], np.float32).reshape(m, 2)

    def vectors(self):
        out = [np.empty(self.n, np.float32) for _ in range(6)]
        alib().orc_analysis_vectors(self._h, *[_fptr(o) for o in out])
        names = ["x_vqt_smoothed", "x_vqt_peakfiltered", "x_vqt_afterglow", "calmness", "pitch_accuracy",
                 "pitch_deviation"]
        return dict(zip(names, out))

    @property
    def smoothed_scene_calmness(self) -> float:
        return float(alib().orc_analysis_scene_calmness(self._h))

    @property
    def smoothed_tuning_grid_inaccuracy(self) -> float:
        return float(alib().orc_analysis_tuning_inaccuracy(self._h))


# ---------------------------------------------------------------------------------------------------
# AGC oracle (oracle/agc_oracle.c)
# ---------------------------------------------------------------------------------------------------
def agc_check(desired_output_rms: float, distortion_factor: float) -> int:
    L = lib()
    L.orc_agc_check.argtypes = [C.c_float, C.c_float]
    return int(L.orc_agc_check(desired_output_rms, distortion_factor))


def agc_process(samples: np.ndarray, desired_output_rms: float, distortion_factor: float, gain: float, frozen: bool):
    """MonoAgc::process on one chunk; returns (processed samples, new gain)."""
    L = lib()
    L.orc_agc_process.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_float, C.c_float, C.POINTER(C.c_float), C.c_int]
    L.orc_agc_process.restype = None
    x = np.array(samples, np.float32, copy=True)
    g = C.c_float(gain)
    L.orc_agc_process(_fptr(x), x.shape[0], desired_output_rms, distortion_factor, C.byref(g), 1 if frozen else 0)
    return x, float(g.value)


def agc_process_chunks(samples: np.ndarray, chunk: int, desired_output_rms: float, distortion_factor: float,
                       silence_threshold: float = 1e-6, gain: float = 1.0):
    L = lib()
    L.orc_agc_process_chunks.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_float,
                                         C.POINTER(C.c_float)]
    L.orc_agc_process_chunks.restype = None
    x = np.array(samples, np.float32, copy=True)
    g = C.c_float(gain)
    L.orc_agc_process_chunks(_fptr(x), x.shape[0], chunk, desired_output_rms, distortion_factor, silence_threshold,
                             C.byref(g))
    return x, float(g.value)


def chroma(db: np.ndarray, min_freq: float = 55.0, buckets_per_octave: int = 84) -> np.ndarray:
    """oracle/chroma_oracle.c: pitch-class energies of one dB frame (update.rs:1104-1131)."""
    L = lib()
    L.orc_chroma.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_float, C.c_uint32, C.POINTER(C.c_float)]
    L.orc_chroma.restype = None
    x = np.ascontiguousarray(db, np.float32)
    out = np.empty(12, np.float32)
    L.orc_chroma(_fptr(x), x.shape[0], min_freq, buckets_per_octave, _fptr(out))
    return out


def spectrogram_vqt(smoothed: np.ndarray, bin_rgb: np.ndarray, image: np.ndarray, write_index: int) -> int:
    """oracle/spectrogram_oracle.c: the spectrogram ring in VQT mode, one update_spectrogram_system call per frame
    (update.rs:930-1088).  `image` [height][n][4] uint8 is updated in place; returns the new write index."""
    L = lib()
    u8 = C.POINTER(C.c_uint8)
    L.orc_spectrogram_vqt_step.argtypes = [C.POINTER(C.c_float), C.c_size_t, u8, u8, C.c_size_t, C.POINTER(C.c_size_t)]
    L.orc_spectrogram_vqt_step.restype = None
    x = np.ascontiguousarray(np.atleast_2d(smoothed), np.float32)
    rgb = np.ascontiguousarray(bin_rgb, np.uint8)
    assert image.dtype == np.uint8 and image.flags.c_contiguous and image.shape[1:] == (x.shape[1], 4)
    w = C.c_size_t(write_index)
    for t in range(x.shape[0]):
        L.orc_spectrogram_vqt_step(_fptr(x[t]), x.shape[1], rgb.ctypes.data_as(u8), image.ctypes.data_as(u8), image.shape[0],
                                   C.byref(w))
    return int(w.value)


def calculate_color(buckets_per_octave: int, bucket: float) -> np.ndarray:
    """oracle/spectrogram_oracle.c: pitchvis_colors::calculate_color with the crate's COLORS / GRAY_LEVEL / EASING_POW."""
    L = lib()
    L.orc_calculate_color.argtypes = [C.c_uint32, C.c_float, C.POINTER(C.c_float)]
    L.orc_calculate_color.restype = None
    out = np.empty(3, np.float32)
    L.orc_calculate_color(buckets_per_octave, bucket, _fptr(out))
    return out


def spectrogram_peaks(peaks_continuous: np.ndarray, peak_count: np.ndarray, image: np.ndarray, write_index: int,
                      buckets_per_octave: int = 84) -> int:
    """oracle/spectrogram_oracle.c: the spectrogram ring in Peaks mode, one update_spectrogram_system call per frame
    (update.rs:997-1062).  peaks_continuous [frames][max_peaks][2], peak_count [frames]."""
    L = lib()
    u8 = C.POINTER(C.c_uint8)
    L.orc_spectrogram_peaks_step.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_uint32, C.c_size_t, u8, C.c_size_t,
                                             C.POINTER(C.c_size_t)]
    L.orc_spectrogram_peaks_step.restype = None
    pk = np.ascontiguousarray(peaks_continuous, np.float32)
    assert image.dtype == np.uint8 and image.flags.c_contiguous and image.shape[2] == 4
    w = C.c_size_t(write_index)
    for t in range(pk.shape[0]):
        n = min(int(peak_count[t]), pk.shape[1])
        row = np.ascontiguousarray(pk[t, :n])
        L.orc_spectrogram_peaks_step(_fptr(row) if n else None, n, buckets_per_octave, image.shape[1],
                                     image.ctypes.data_as(u8), image.shape[0], C.byref(w))
    return int(w.value)

"""Host logic of the product without a GPU: the C++ kernel builder against the oracle's, the
C-ABI symbol table against include/pvqt.h, error mapping, and the shard arithmetic."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import orc
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        text = open(os.path.join(inc, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(pvqt_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) > 40
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    exported = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (pvqt_[a-z0-9_]+)", exported))
    assert declared <= exported
    # and the ctypes table covers the same set
    assert declared == set(_ffi.SIGNATURES) | {n for n in declared if n not in _ffi.SIGNATURES and False} or \
        declared - set(_ffi.SIGNATURES) == set(), declared - set(_ffi.SIGNATURES)


def test_abi_version_and_defaults(built_lib):
    lib = _ffi.load()
    assert lib.pvqt_abi_version() == 1
    p = _ffi.PvqtParams()
    assert lib.pvqt_default_params(C.byref(p)) == 0
    o = orc.default_params()
    for f, _ in _ffi.PvqtParams._fields_:
        assert getattr(p, f) == getattr(o, f), f
    assert lib.pvqt_params_n_buckets(C.byref(p)) == 588
    d = pv.VqtParameters.default()
    assert (d.sr, d.n_fft, d.range.min_freq, d.range.octaves, d.range.buckets_per_octave) == (22050.0, 32768, 55.0, 7, 84)
    assert np.float32(d.quality) == np.float32(1.6) and np.float32(d.gamma) == np.float32(4.8) * np.float32(1.6)


@pytest.mark.parametrize("name", ["default", "hires", "train"])
def test_builder_matches_oracle(built_lib, name):
    if name == "default":
        pp, op = pv.VqtParameters.default(), orc.default_params()
    elif name == "hires":
        pp, op = pv.VqtParameters.hires(), orc.hires_params()
    else:
        # pitchvis_train/src/train.rs:30-41
        pp = pv.VqtParameters(range=pv.VqtRange(55.0, 7, 36), quality=10.0, gamma=53.0)
        op = orc.make_params(buckets_per_octave=36, quality=10.0, gamma=53.0)
    hk = pv.HostKernel(pp)
    ov = orc.OracleVqt(op)
    assert hk.n_buckets == ov.n_buckets
    assert hk.delay == ov.delay
    k = hk.kernel()
    assert len(k.window_groups) == ov.num_groups
    for g, wg in enumerate(k.window_groups):
        w, K, Kn = ov.group(g)
        assert wg.window == w
        for mine, ref in ((wg.filter_bank, K), (wg.negative_filter_bank, Kn)):
            if ref.nnz == 0:
                assert mine is None or mine.nnz() == 0
                continue
            assert (mine.rows, mine.cols) == (ref.rows, ref.cols)
            np.testing.assert_array_equal(mine.indptr, ref.indptr)
            np.testing.assert_array_equal(mine.indices, ref.indices)   # same sparsity pattern
            # two independent f64 FFTs rounded to f32: equal up to a last-bit rounding flip
            np.testing.assert_allclose(mine.data.view(np.float32), ref.data.view(np.float32), rtol=3e-7, atol=1e-12)


def test_filter_bank_params_match_oracle(built_lib):
    mine = pv.filter_bank_params(pv.VqtParameters.default())
    ref = orc.filter_bank_params(orc.default_params())
    assert len(mine) == len(ref) == 588
    for a, b in zip(mine, ref):
        assert a == (b.freq, b.window_length, b.sr_downscaling_factor, b.minimum_needed_window_size)
    # 8 rate groups (M = 128 .. 1) at the defaults, SURVEY.md section 0
    assert sorted({a[2] for a in mine}) == [1, 2, 4, 8, 16, 32, 64, 128]


def test_construction_errors(built_lib):
    with pytest.raises(pv.AboveNyquist) as e:
        pv.HostKernel(pv.VqtParameters(range=pv.VqtRange(55.0, 8, 84)))
    assert e.value.nyquist_frequency == 11025.0 and e.value.highest_frequency > 11025.0
    with pytest.raises(pv.WindowExceedsNFft) as e:
        pv.HostKernel(pv.VqtParameters(n_fft=2048))
    assert e.value.n_fft == 2048 and e.value.window_length > 2048


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device Vqt::new must fail loudly, never compute on the CPU."""
    lib = _ffi.load()
    n = C.c_int(0)
    rc = lib.pvqt_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pv.PvqtRuntimeError) as e:
        pv.Vqt()
    assert e.value.status == _ffi.PVQT_CUDA_ERROR


def test_shard_range_and_halo(built_lib):
    lib = _ffi.load()
    b, e = C.c_size_t(), C.c_size_t()
    for n_units in (0, 1, 7, 3507, 4096):
        for parts in (1, 2, 3, 8):
            cover = []
            for p in range(parts):
                assert lib.pvqt_shard_range(n_units, parts, p, C.byref(b), C.byref(e)) == 0
                cover.append((b.value, e.value))
            assert cover[0][0] == 0 and cover[-1][1] == n_units
            assert all(cover[i][1] == cover[i + 1][0] for i in range(parts - 1))
            sizes = [y - x for x, y in cover]
            assert max(sizes) - min(sizes) <= 1
    assert lib.pvqt_shard_range(10, 0, 0, C.byref(b), C.byref(e)) != 0
    s0, s1 = C.c_size_t(), C.c_size_t()
    assert lib.pvqt_frame_range_samples(32768, 368, 100, 200, C.byref(s0), C.byref(s1)) == 0
    assert (s0.value, s1.value) == (100 * 368, 199 * 368 + 32768)
    assert lib.pvqt_frame_range_samples(32768, 368, 5, 5, C.byref(s0), C.byref(s1)) == 0
    assert (s0.value, s1.value) == (0, 0)


def test_log_callback_and_coverage_gap_diagnostic(built_lib):
    # the reference's `log` lines (vqt.rs:468, :661-667, :688-710, :741-746) through pvqt_set_log_callback, and
    # calculate_bandwidth's -3 dB bands (vqt.rs:962-989)
    lines = []
    pv.set_log_callback(lambda level, msg: lines.append((level, msg)), 3)
    try:
        k = pv.HostKernel(pv.VqtParameters.default())
    finally:
        pv.set_log_callback(None)
    assert (2, "VQT analysis delay: 98 ms.") in lines
    dbg = [m for lv, m in lines if lv == 3]
    assert "window (24576, 32768) (8192 samples): 122 filters in 2 rate group(s)" in dbg
    assert "window (24576, 32768): kernel nnz 3742, conjugate-part nnz 369" in dbg
    assert "window (30087, 31111): kernel nnz 1578, conjugate-part nnz 0" in dbg
    assert sum(m.startswith("filter at ") for m in dbg) == 588
    assert not [m for lv, m in lines if lv == 1] and k.coverage_gaps() == []      # no gap at the defaults
    lo, hi = k.filter_bandwidths()
    fps = pv.filter_bank_params(pv.VqtParameters.default())
    f = np.array([p[0] for p in fps])
    assert np.all(lo < f) and np.all(f < hi) and np.all(hi - lo < 0.35 * f)
    # a sharper bank leaves gaps between neighbouring -3 dB bands: the reference warns, so does the sink
    warns = []
    pv.set_log_callback(lambda level, msg: warns.append((level, msg)), 1)
    try:
        k2 = pv.HostKernel(pv.VqtParameters(quality=3.2, n_fft=65536))
    finally:
        pv.set_log_callback(None)
    gaps = k2.coverage_gaps()
    assert gaps and len(warns) == len(gaps) and all(lv == 1 and m.startswith("coverage gap below the filter at") for lv, m in warns)
    n_before = len(warns)
    pv.HostKernel(pv.VqtParameters(quality=3.2, n_fft=65536))                        # sink removed: silent
    assert len(warns) == n_before


def test_wrapper_argument_validation(built_lib):
    # a caller-supplied result buffer goes to the library as a raw pointer: wrong dtype / shape / layout is refused
    # before any pointer is formed (no GPU needed: the checks run first)
    class _Stub(pv.Vqt):
        def __init__(self):
            self.n_buckets, self.n_fft = 588, 32768

        def frames_in(self, n, hop):
            return (n - self.n_fft) // hop + 1 if n >= self.n_fft else 0

        def close(self):
            pass

    v = _Stub()
    audio = np.zeros(32768 + 368, np.float32)
    for bad in (np.zeros((2, 588), np.float64), np.zeros((3, 588), np.float32), np.zeros((2, 1176), np.float32)[:, ::2]):
        with pytest.raises(ValueError):
            v.calculate_vqt_batch_in_db(audio, 368, out=bad)
    with pytest.raises(ValueError):
        v.calculate_vqt_batch_in_db(audio, 0)
    with pytest.raises(ValueError):
        v.calculate_vqt_batch_in_db(audio, -368)
    with pytest.raises(ValueError):
        v.calculate_vqt_batch_in_db(np.zeros((2, 40000), np.float32), 368)
    with pytest.raises(ValueError):
        v.calculate_vqt_streams_in_db(np.zeros((2, 40000), np.float32), 368, out=np.zeros((2, 20, 587), np.float32))

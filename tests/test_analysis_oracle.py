"""Pins the analysis oracle (oracle/analysis_oracle.c) against the reference's own tests for the
AnalysisState epilogue and against scipy.signal.find_peaks (the documented stand-in for the
find_peaks 0.1.5 crate).  CPU only."""
import numpy as np
import pytest
import scipy.signal

import orc
from pitchvis_b200 import synth

MS = 1_000_000


def test_ema_basic():
    # util.rs:143-186: 250 ms steps vs 125 ms steps agree within 0.05
    lo = hi = 0.0
    for target in (1.0, 2.0, 3.0, 4.0):
        for _ in range(2):
            lo = orc.ema_update(lo, 1000 * MS, target, 250 * MS)
        for _ in range(4):
            hi = orc.ema_update(hi, 1000 * MS, target, 125 * MS)
    assert abs(lo - hi) < 0.05


def test_ema_limit():
    # util.rs:188-225: n updates of timestep/n compose exactly
    def run(n):
        y = 0.0
        for _ in range(n):
            y = orc.ema_update(y, 1000 * MS, 1.0, (500 // n) * MS)
        return y
    high, medium, low = run(100), run(10), run(3)
    n_horizon = 1.0 / 0.005
    alpha = 2.0 / (n_horizon + 1.0)
    calc = 0.0 + (1.0 - np.exp(-alpha * 100)) * (1.0 - 0.0)
    assert abs(low - high) < 0.02 and abs(low - medium) < 0.02 and abs(low - calc) < 0.02
    # time_horizon None -> passthrough (util.rs:117-120)
    assert orc.ema_update(3.0, None, 7.5, 10 * MS) == 7.5


def test_analysis_does_something():
    # analysis.rs:415-428
    a = orc.OracleAnalysisState(55.0, 2, 24)
    a.preprocess(np.zeros(48, np.float32), 1_000_000_000)
    assert np.all(a.vectors()["x_vqt_smoothed"] == 0.0)
    with pytest.raises(ValueError):
        a.preprocess(np.zeros(47, np.float32), 1_000_000_000)   # analysis.rs:289


def test_doc_example_runs():
    # analysis.rs:110-118
    a = orc.OracleAnalysisState(55.0, 8, 24)
    a.preprocess(np.zeros(8 * 24, np.float32), 30 * MS)
    assert a.peaks.size == 0


def test_vqt_close_frequencies(oracle_default):
    # lib.rs:16-48: two tones a semitone apart, 333 Hz upward -> exactly two peaks (117 cases)
    v = oracle_default
    counts = []
    for i in range(int(2.6 * 30), 7 * 30 - 15):
        log_note = np.float32(i) / np.float32(30)
        f1 = np.float32(55.0) * np.float32(2.0) ** log_note
        f2 = np.float32(55.0) * np.float32(2.0) ** np.float32(log_note + np.float32(1.0 / 12.0))
        x_vqt = v.calculate_vqt_instant_in_db(orc.test_create_sines(v.params, [f1, f2]), mode=1)
        a = orc.OracleAnalysisState()
        a.preprocess(x_vqt, 1100 * MS)
        counts.append(a.peaks.size)
    assert len(counts) == 117
    assert all(c == 2 for c in counts), counts


def _scipy_peaks(x, prom, height, bpo):
    dist = int(round(bpo * 0.4 / 12.0))
    pk, _ = scipy.signal.find_peaks(x, height=height, prominence=prom, distance=dist if dist > 0 else None)
    return pk[pk >= -(-(bpo // 12) // 2)]


def test_find_peaks_matches_scipy_on_vqt_frames(oracle_default):
    v = oracle_default
    audio = synth.polyphonic_chords(6.0, 22050.0, seed=3)
    frames = v.calculate_batch_db(audio, 368, mode=1)
    assert frames.shape[0] > 200
    n_peaks = 0
    for x in frames[::3]:
        for prom, h in ((10.0, 4.0), (5.0, 3.5)):
            mine = orc.find_peaks(x, prom, h, 84, order=0)
            np.testing.assert_array_equal(mine, _scipy_peaks(x, prom, h, 84))
            # the alternative filter order (prominence before distance) gives the same set
            np.testing.assert_array_equal(mine, orc.find_peaks(x, prom, h, 84, order=1))
            n_peaks += mine.size
    assert n_peaks > 500


def test_find_peaks_matches_scipy_on_adversarial_input():
    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(3, 200))
        x = rng.normal(0, 6, n).astype(np.float32)
        x = np.abs(x)
        if trial % 3 == 0:
            # plateaus (runs of equal samples); distinct plateaus get distinct heights, because scipy
            # breaks exact height ties between *separate* peaks by an unstable argsort (platform-defined)
            x = np.round(x / 3) * 3
            run = np.cumsum(np.r_[0, np.diff(x) != 0])
            x = x + run * 1e-3
        x = x.astype(np.float32)
        for bpo in (12, 36, 84, 168):
            mine = orc.find_peaks(x, 2.0, 1.0, bpo, order=0)
            np.testing.assert_array_equal(mine, _scipy_peaks(x, 2.0, 1.0, bpo), err_msg=f"trial {trial} bpo {bpo}")


def test_preprocess_sequence_properties(oracle_default):
    """A run over chord frames: invariants of the epilogue state that do not depend on find_peaks details."""
    v = oracle_default
    audio = synth.polyphonic_chords(4.0, 22050.0, seed=5)
    frames = v.calculate_batch_db(audio, 368, mode=1)
    a = orc.OracleAnalysisState()
    ft = 16_689_342  # 368 / 22050 s
    prev_after = np.zeros(588, np.float32)
    for x in frames:
        a.preprocess(x, ft)
        vec = a.vectors()
        sm = vec["x_vqt_smoothed"]
        pk = a.peaks
        assert np.all(np.diff(pk) >= 3)                                   # min distance 3 bins
        assert np.all(pk >= 4)                                            # lowest half semitone dropped
        pf = vec["x_vqt_peakfiltered"]
        assert np.count_nonzero(pf) <= pk.size and np.all(pf[pk] == sm[pk])
        assert np.all(vec["x_vqt_afterglow"] >= sm)                        # afterglow.rs:17-19
        decay = np.float32(0.85) - np.float32(0.15) * (np.arange(588, dtype=np.float32) / np.float32(588))
        np.testing.assert_array_equal(vec["x_vqt_afterglow"], np.maximum(prev_after * decay, sm))
        prev_after = vec["x_vqt_afterglow"]
        pc = a.peaks_continuous
        assert pc.shape[0] == pk.size
        assert np.all(np.abs(pc[:, 0] - pk) <= 1.0 + 1e-4) and np.all(np.diff(pc[:, 0]) > 0)
        assert 0.0 <= a.smoothed_scene_calmness <= 1.0
        assert np.all((vec["calmness"] >= 0) & (vec["calmness"] <= 1))
    assert a.smoothed_scene_calmness > 0.0 and a.peaks.size > 0

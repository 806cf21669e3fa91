"""GPU parity of the AnalysisState epilogue (K-analysis) against the CPU oracle.

Layer (i) of SURVEY.md section 7: both sides are fed the *same* dB bits (the GPU VQT output), so any
difference comes from the epilogue itself.  Peak index sets must match exactly in every frame; the
float state may differ by libm last-ulp effects (device expf/powf/logf vs glibc), bounded below."""
import numpy as np
import pytest

import orc
import pitchvis_b200 as pv
from pitchvis_b200 import synth

pytestmark = pytest.mark.gpu

HOP = synth.HOP_DEFAULT
FRAME_NS = 16_689_342   # 368 / 22050 s


def _oracle_run(db, frame_ns, range_=(55.0, 7, 84), params=None, max_peaks=64):
    T, NB = db.shape
    a = orc.OracleAnalysisState(*range_, params=params)
    out = {"peaks": [], "cont": [], "scene": np.zeros(T, np.float32), "tuning": np.zeros(T, np.float32),
           "vec": {k: np.zeros((T, NB), np.float32) for k in pv.analysis.VECTOR_FIELDS}}
    for t in range(T):
        a.preprocess(db[t], frame_ns)
        out["peaks"].append(a.peaks)
        out["cont"].append(a.peaks_continuous)
        out["scene"][t] = a.smoothed_scene_calmness
        out["tuning"][t] = a.smoothed_tuning_grid_inaccuracy
        for k, v in a.vectors().items():
            out["vec"][k][t] = v
    return out


def _compare(res, s, ref, T):
    mismatches = []
    for t in range(T):
        n = int(res["peak_count"][s, t])
        got = res["peak_indices"][s, t, :n]
        if not np.array_equal(got, ref["peaks"][t]):
            mismatches.append((t, got.tolist(), ref["peaks"][t].tolist()))
    assert not mismatches, f"{len(mismatches)} frames with different peak sets, first: {mismatches[:3]}"
    for t in range(T):
        n = int(res["peak_count"][s, t])
        np.testing.assert_allclose(res["peaks_continuous"][s, t, :n, 0], ref["cont"][t][:, 0], atol=1e-3)  # bins
        np.testing.assert_allclose(res["peaks_continuous"][s, t, :n, 1], ref["cont"][t][:, 1], atol=1e-3)  # dB
    np.testing.assert_allclose(res["smoothed_scene_calmness"][s], ref["scene"], atol=2e-5)
    np.testing.assert_allclose(res["smoothed_tuning_grid_inaccuracy"][s], ref["tuning"], atol=2e-3)
    for k in pv.analysis.VECTOR_FIELDS:
        np.testing.assert_allclose(res[k][s], ref["vec"][k], atol=2e-4, err_msg=k)


@pytest.fixture(scope="module")
def vqt(built_lib):
    v = pv.Vqt()
    yield v
    v.close()


def test_zeros_and_wrong_length(built_lib):
    # analysis.rs:415-428 and the assert at analysis.rs:289
    a = pv.AnalysisState(pv.VqtRange(55.0, 2, 24))
    r = a.preprocess(np.zeros(48, np.float32), 1_000_000_000)
    assert np.all(r["x_vqt_smoothed"] == 0.0) and r["peaks"] == set()
    with pytest.raises(ValueError):
        a.preprocess(np.zeros(47, np.float32), 1_000_000_000)
    a.close()


def test_two_close_tones_give_two_peaks(vqt):
    # lib.rs:16-48 through the GPU VQT and the GPU epilogue (every 6th of the 117 cases)
    op = orc.default_params()
    for i in range(int(2.6 * 30), 7 * 30 - 15, 6):
        f1 = np.float32(55.0) * np.float32(2.0) ** (np.float32(i) / np.float32(30))
        f2 = np.float32(55.0) * np.float32(2.0) ** np.float32(np.float32(i) / np.float32(30) + np.float32(1 / 12))
        x_vqt = vqt.calculate_vqt_instant_in_db(orc.test_create_sines(op, [f1, f2]))
        a = pv.AnalysisState(pv.VqtRange())
        assert len(a.preprocess(x_vqt, 1100 * 1_000_000)["peaks"]) == 2, i
        a.close()


def test_sequence_matches_oracle(vqt):
    # BASELINE configs[4] on a 12 s recording (654 frames): same dB bits into both epilogues
    audio = synth.polyphonic_chords(12.0, 22050.0, seed=0)
    db = vqt.calculate_vqt_batch_in_db(audio, HOP)
    T = db.shape[0]
    a = pv.AnalysisState(pv.VqtRange(), n_streams=1)
    res = a.preprocess_batch(db, FRAME_NS)
    ref = _oracle_run(db, FRAME_NS)
    assert sum(len(p) for p in ref["peaks"]) > 2000
    _compare(res, 0, ref, T)
    a.close()


def test_batch_split_equals_one_call_and_streams_are_independent(vqt):
    # state is carried across calls: two calls of T/2 frames == one call of T frames, bit for bit;
    # and a stream's results do not depend on its neighbours in the batch
    chords = [synth.polyphonic_chords(4.0, 22050.0, seed=s) for s in (1, 2, 3)]
    db = vqt.calculate_vqt_streams_in_db(np.stack(chords), HOP)          # [3][T][588]
    S, T, NB = db.shape
    one = pv.AnalysisState(pv.VqtRange(), n_streams=S)
    full = one.preprocess_batch(db, FRAME_NS)
    two = pv.AnalysisState(pv.VqtRange(), n_streams=S)
    h = T // 2
    first, second = two.preprocess_batch(db[:, :h], FRAME_NS), two.preprocess_batch(db[:, h:], FRAME_NS)
    for k in full:
        np.testing.assert_array_equal(full[k], np.concatenate([first[k], second[k]], axis=1), err_msg=k)
    solo = pv.AnalysisState(pv.VqtRange(), n_streams=1)
    alone = solo.preprocess_batch(db[1], FRAME_NS)
    for k in full:
        np.testing.assert_array_equal(full[k][1], alone[k][0], err_msg=k)
    ref = _oracle_run(db[2], FRAME_NS)
    _compare(full, 2, ref, T)
    for s in (one, two, solo):
        s.close()


def test_smoothing_disabled_and_custom_parameters(vqt):
    # update_vqt_smoothing_duration(None) -> passthrough (analysis.rs:251-270); non-default thresholds
    audio = synth.polyphonic_chords(3.0, 22050.0, seed=9)
    db = vqt.calculate_vqt_batch_in_db(audio, HOP)
    T = db.shape[0]
    prm = pv.AnalysisParameters(peak_config=pv.PeakDetectionParameters(6.0, 2.0), highest_bassnote=40,
                                vqt_smoothing_duration_base=30 * 1_000_000)
    a = pv.AnalysisState(pv.VqtRange(), prm)
    a.update_vqt_smoothing_duration(None)
    res = a.preprocess_batch(db, 33_000_000)
    np.testing.assert_array_equal(res["x_vqt_smoothed"][0], db)
    op = orc.analysis_default_params()
    op.peak_config.min_prominence, op.peak_config.min_height = 6.0, 2.0
    op.highest_bassnote = 40
    op.vqt_smoothing_duration_base_ns = 30 * 1_000_000
    o = orc.OracleAnalysisState(params=op)
    o.update_vqt_smoothing_duration(None)
    for t in range(T):
        o.preprocess(db[t], 33_000_000)
        n = int(res["peak_count"][0, t])
        np.testing.assert_array_equal(res["peak_indices"][0, t, :n], o.peaks)
    a.close()


def test_config5_full_recording(vqt, oracle_default):
    """BASELINE configs[4]: the full 60 s recording (3507 frames) through VQT + epilogue.

    Layer (i): identical dB bits into both epilogues -> peak sets must be identical in every frame.
    Layer (ii): end to end against the all-CPU chain (exact-oracle VQT -> oracle epilogue).  The dB
    inputs then differ by up to 1e-3 dB, and the thresholds (prominence >= 10, height >= 4) are hard
    comparisons on EMA-smoothed values, so a near-tie may flip; every flipped peak must be such a
    near-tie (its oracle prominence or height within 5e-3 dB of the threshold), and there may be at
    most a handful."""
    import scipy.signal
    audio = synth.polyphonic_chords(60.0, 22050.0, seed=0)
    db = vqt.calculate_vqt_batch_in_db(audio, HOP)
    T = db.shape[0]
    assert T == 3507
    a = pv.AnalysisState(pv.VqtRange())
    res = a.preprocess_batch(db, FRAME_NS, vectors=False)
    ref = _oracle_run(db, FRAME_NS)
    bad = [t for t in range(T)
           if not np.array_equal(res["peak_indices"][0, t, :int(res["peak_count"][0, t])], ref["peaks"][t])]
    assert not bad, f"layer (i): {len(bad)} of {T} frames differ, first {bad[:5]}"
    np.testing.assert_allclose(res["smoothed_scene_calmness"][0], ref["scene"], atol=2e-5)
    a.close()

    db_cpu = oracle_default.calculate_batch_db(audio, HOP, mode=0)
    assert np.abs(db_cpu - db).max() <= 1e-3
    ref2 = _oracle_run(db_cpu, FRAME_NS)
    flipped = []
    for t in range(T):
        got = set(res["peak_indices"][0, t, :int(res["peak_count"][0, t])].tolist())
        want = set(ref2["peaks"][t].tolist())
        for p in got ^ want:
            flipped.append((t, p))
    assert len(flipped) <= 10, f"layer (ii): {len(flipped)} flipped peaks"
    for t, p in flipped:
        sm = ref2["vec"]["x_vqt_smoothed"][t]
        cfg = (5.0, 3.5) if p <= 28 else (10.0, 4.0)
        lo, hi = max(p - 2, 1), min(p + 3, 587)
        cands = [q for q in range(lo, hi) if sm[q - 1] < sm[q] >= sm[q + 1]]
        margins = []
        for q in cands:
            prom = scipy.signal.peak_prominences(sm, [q])[0][0]
            margins += [abs(prom - cfg[0]), abs(sm[q] - cfg[1])]
        assert margins and min(margins) < 5e-3, f"frame {t} bin {p}: not a near-tie (margins {margins})"


def test_single_call_pipeline_matches_two_calls(vqt):
    """pvqt_calc_batch_analysis (BASELINE configs[4] as one library call): the spectra stay in HBM between the transform
    and the epilogue.  Same bits as calculate_vqt_batch_in_db followed by preprocess_batch, exact peak sets against the
    oracle epilogue, and only the requested results cross PCIe."""
    audio = synth.polyphonic_chords(12.0, 22050.0, seed=4)
    db = vqt.calculate_vqt_batch_in_db(audio, HOP)
    T = db.shape[0]
    two = pv.AnalysisState(pv.VqtRange())
    want = two.preprocess_batch(db, FRAME_NS)
    one = pv.AnalysisState(pv.VqtRange())
    got = one.calculate_and_preprocess(vqt, audio, HOP, FRAME_NS, vectors=True, return_db=True)
    np.testing.assert_array_equal(got["db"][0], db)
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    ref = _oracle_run(db, FRAME_NS)
    _compare(got, 0, ref, T)
    # peaks only: a small fraction of the 4 * 588 bytes per frame the spectra would cost
    lean = pv.AnalysisState(pv.VqtRange())
    res = lean.calculate_and_preprocess(vqt, audio, HOP, FRAME_NS, max_peaks=32, vectors=False)
    np.testing.assert_array_equal(res["peak_count"], want["peak_count"])
    np.testing.assert_array_equal(res["peak_indices"], want["peak_indices"][:, :, :32])
    assert res["d2h_bytes"] == T * (4 + 32 * 4 + 32 * 8 + 4 + 4) < T * 588 * 4 // 5
    # a second call continues the recurrence: two halves == one call
    h = T // 2
    split = pv.AnalysisState(pv.VqtRange())
    n_half = (h - 1) * HOP + vqt.n_fft
    a1 = split.calculate_and_preprocess(vqt, audio[:n_half], HOP, FRAME_NS, frames_per_stream=h, vectors=True)
    a2 = split.calculate_and_preprocess(vqt, audio[h * HOP:], HOP, FRAME_NS, frames_per_stream=T - h, vectors=True)
    for k in want:
        np.testing.assert_array_equal(np.concatenate([a1[k], a2[k]], axis=1), want[k], err_msg=k)
    for s in (two, one, lean, split):
        s.close()


def test_single_call_pipeline_streams(vqt):
    chords = np.stack([synth.polyphonic_chords(3.0, 22050.0, seed=s) for s in (11, 12, 13, 14, 15)])
    db = vqt.calculate_vqt_streams_in_db(chords, HOP)
    S, T, _ = db.shape
    two = pv.AnalysisState(pv.VqtRange(), n_streams=S)
    want = two.preprocess_batch(db, FRAME_NS, vectors=("x_vqt_smoothed",))
    one = pv.AnalysisState(pv.VqtRange(), n_streams=S)
    got = one.calculate_and_preprocess(vqt, chords, HOP, FRAME_NS, vectors=("x_vqt_smoothed",))
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    with pytest.raises(ValueError):
        one.calculate_and_preprocess(vqt, chords[:3], HOP, FRAME_NS)        # another number of streams
    with pytest.raises(ValueError):
        one.calculate_and_preprocess(vqt, chords[:, :vqt.n_fft - 1], HOP, FRAME_NS, frames_per_stream=1)
    two.close()
    one.close()


def test_all_117_close_tone_cases_through_the_gpu(vqt):
    # the reference's test_vqt_close_frequencies (lib.rs:16-48), every case: GPU VQT (one batched call over independent
    # frames) + one GPU epilogue with 117 independent states, frame_time 1100 ms
    op = orc.default_params()
    cases = list(range(int(2.6 * 30), 7 * 30 - 15))
    assert len(cases) == 117
    frames = []
    for i in cases:
        f1 = np.float32(55.0) * np.float32(2.0) ** (np.float32(i) / np.float32(30))
        f2 = np.float32(55.0) * np.float32(2.0) ** np.float32(np.float32(i) / np.float32(30) + np.float32(1 / 12))
        frames.append(orc.test_create_sines(op, [f1, f2]))
    db = vqt.calculate_vqt_frames_in_db(np.stack(frames))
    a = pv.AnalysisState(pv.VqtRange(), n_streams=len(cases))
    res = a.preprocess_batch(db[:, None, :], 1100 * 1_000_000, vectors=False)
    assert res["peak_count"][:, 0].tolist() == [2] * len(cases)
    a.close()


@pytest.mark.parametrize("range_", [(55.0, 7, 84), (55.0, 8, 168), (55.0, 2, 24)])
def test_plateaus_ties_and_bin_counts(built_lib, range_):
    # Synthetic dB frames quantised to 0.5 dB: plateaus of equal maxima, equal-height neighbours inside the distance
    # window, long flat floors (the prominence walks run far) -- at 588, 1344 (more than 8 bins per search thread) and
    # 48 bins.  Peak sets must equal the oracle's in every frame.
    nb = range_[1] * range_[2]
    rng = np.random.default_rng(7)
    T = 48
    db = np.zeros((T, nb), np.float32)
    x = np.arange(nb)
    for t in range(T):
        frame = np.zeros(nb)
        for _ in range(int(rng.integers(5, 40))):
            c, w, h = rng.uniform(0, nb), rng.uniform(1.0, 12.0), rng.uniform(2.0, 45.0)
            frame += h * np.exp(-0.5 * ((x - c) / w) ** 2)
        frame += rng.uniform(0.0, 3.0, nb) * (rng.random(nb) < 0.3)
        db[t] = np.round(np.minimum(frame, 60.0) * 2.0) / 2.0          # ties and plateaus
    a = pv.AnalysisState(pv.VqtRange(*range_), n_streams=1)
    res = a.preprocess_batch(db, FRAME_NS)
    ref = _oracle_run(db, FRAME_NS, range_=range_)
    assert sum(len(p) for p in ref["peaks"]) > (50 if nb > 100 else 10)
    _compare(res, 0, ref, T)
    a.close()


def test_small_cta_form_gives_the_same_results(vqt, monkeypatch):
    # more than two streams per SM run in 128-thread CTAs (three searches one after the other): same results, bit for bit
    audio = synth.polyphonic_chords(6.0, 22050.0, seed=3)
    db = vqt.calculate_vqt_batch_in_db(audio, HOP)
    S = 3
    many = np.stack([np.roll(db, 7 * s, axis=0) for s in range(S)])
    big = pv.AnalysisState(pv.VqtRange(), n_streams=S)
    ref = big.preprocess_batch(many, FRAME_NS)
    big.close()
    monkeypatch.setenv("PVQT_ANALYSIS_SMALL_CTA", "1")
    small = pv.AnalysisState(pv.VqtRange(), n_streams=S)
    got = small.preprocess_batch(many, FRAME_NS)
    small.close()
    for k in ref:
        assert np.array_equal(ref[k], got[k]), k


def test_frame_by_frame_pipeline_matches_the_batch(vqt):
    # the viewer's loop: one frame of audio per call through VQT + AnalysisState (captured graph after the second call)
    # against the same recording in one batch call; then a changed smoothing duration must reach the captured launch
    audio = synth.polyphonic_chords(3.0, 22050.0, seed=5)
    n_fft = 32768
    T = 60
    vqt.set_sliding_dft(0)   # one frame per call transforms every window group with the FFT: the batch must too for equal bits
    try:
        batch = pv.AnalysisState(pv.VqtRange())
        ref = batch.calculate_and_preprocess(vqt, audio[:(T - 1) * HOP + n_fft], HOP, FRAME_NS, max_peaks=48)
        live = pv.AnalysisState(pv.VqtRange())
        for t in range(T):
            r = live.calculate_and_preprocess(vqt, audio[t * HOP:t * HOP + n_fft], HOP, FRAME_NS, frames_per_stream=1, max_peaks=48)
            for k in ("peak_count", "peak_indices", "peaks_continuous", "smoothed_scene_calmness", "smoothed_tuning_grid_inaccuracy"):
                assert np.array_equal(r[k][0, 0], ref[k][0, t]), (k, t)
        batch.update_vqt_smoothing_duration(20 * 1_000_000)
        live.update_vqt_smoothing_duration(20 * 1_000_000)
        ref2 = batch.calculate_and_preprocess(vqt, audio[T * HOP:(T + 9) * HOP + n_fft], HOP, FRAME_NS, max_peaks=48)
        for t in range(10):
            r = live.calculate_and_preprocess(vqt, audio[(T + t) * HOP:(T + t) * HOP + n_fft], HOP, FRAME_NS, frames_per_stream=1,
                                              max_peaks=48)
            for k in ("peak_count", "peak_indices", "smoothed_scene_calmness"):
                assert np.array_equal(r[k][0, 0], ref2[k][0, t]), (k, t)
    finally:
        vqt.set_sliding_dft(2)
    batch.close()
    live.close()


def test_frame_pipeline_with_no_results_requested(vqt):
    # one frame per call with an outputs struct that names no array: the states advance, nothing is copied
    import ctypes as C
    from pitchvis_b200 import _ffi
    lib = _ffi.load()
    audio = synth.polyphonic_chords(2.0, 22050.0, seed=1)
    a = pv.AnalysisState(pv.VqtRange())
    outs = _ffi.PvqtAnalysisOutputs()
    outs.max_peaks = 8
    moved = C.c_uint64(123)
    for t in range(3):
        x = np.ascontiguousarray(audio[t * HOP:t * HOP + 32768])
        rc = lib.pvqt_calc_batch_analysis(vqt.handle, a._h, x.ctypes.data_as(C.POINTER(C.c_float)), 32768, HOP, 1, FRAME_NS,
                                          C.byref(outs), None, C.byref(moved))
        assert rc == 0 and moved.value == 0
    # the state did advance: the next frame's results equal those of a state that saw the same four frames with results
    b = pv.AnalysisState(pv.VqtRange())
    for t in range(3):
        b.calculate_and_preprocess(vqt, audio[t * HOP:t * HOP + 32768], HOP, FRAME_NS, frames_per_stream=1, max_peaks=32)
    ra = a.calculate_and_preprocess(vqt, audio[3 * HOP:3 * HOP + 32768], HOP, FRAME_NS, frames_per_stream=1, max_peaks=32)
    rb = b.calculate_and_preprocess(vqt, audio[3 * HOP:3 * HOP + 32768], HOP, FRAME_NS, frames_per_stream=1, max_peaks=32)
    for k in ("peak_count", "peak_indices", "smoothed_scene_calmness"):
        assert np.array_equal(ra[k], rb[k]), k
    a.close()
    b.close()

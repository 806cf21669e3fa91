"""The reference's own property tests, run through the CUDA path and the C ABI:

  test_vqt_bandwidths                   vqt.rs:996-1027    all 11,740 tones in ONE pvqt_calc_frames_db call
  test_vqt_group_boundary_continuity    vqt.rs:1032-1076
  test_vqt_high_frequencies             lib.rs:50-72
  (test_vqt_close_frequencies, lib.rs:16-48: tests/test_gpu_analysis.py::test_all_117_close_tone_cases_through_the_gpu)

plus the contract that a frame's result depends on its n_fft samples only: every entry point, every hop (odd hops
included), every shard boundary gives the same bits."""
import ctypes as C

import numpy as np
import pytest

import orc
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vqt(built_lib):
    v = pv.Vqt()
    yield v
    v.close()


def _tones(freqs):
    """test_create_sines (util.rs:62-79) for one tone per frame; only the samples the windows read are generated
    (the reference fills all n_fft, the transform reads the last 8192)."""
    op = orc.default_params()
    return np.stack([orc.test_create_sines(op, [f]) for f in freqs])


def test_vqt_bandwidths_all_11740_tones_in_one_call(vqt):
    # vqt.rs:996-1027: sweep 20 tones per bucket; the strongest single-bin response may exceed the weakest summed
    # response by less than 3 dB
    idx = np.arange(10, 588 * 20 - 10)
    assert idx.size == 11740
    freqs = np.float32(55.0) * np.float32(2.0) ** (idx.astype(np.float32) / np.float32(84.0 * 20.0))
    db = vqt.calculate_vqt_frames_in_db(_tones(freqs))
    assert db.shape == (11740, 588)
    max_single, min_sum = float(db.max()), float(db.sum(axis=1).min())
    assert max_single - min_sum < 3.0
    # and the same frames against the oracle, a spread-out sample of them
    o = orc.OracleVqt()
    pick = idx[::587] - 10
    ref = np.stack([o.calculate_vqt_instant_in_db(orc.test_create_sines(o.params, [freqs[i]]), mode=0) for i in pick])
    assert np.abs(db[pick] - ref).max() <= 1e-3


def test_vqt_group_boundary_continuity(vqt):
    # vqt.rs:1032-1076 (filter_bank_params is private in the reference; exported here as pvqt_filter_bank_params)
    fps = pv.filter_bank_params(pv.VqtParameters.default())
    boundaries = [fps[i + 1][0] for i in range(len(fps) - 1) if fps[i][2] != fps[i + 1][2]]
    assert len(boundaries) == 7
    freqs = [np.float32(b) * np.float32(2.0) ** np.float32(i / (20 * 4.0 * 12.0)) for b in boundaries for i in range(-20, 21)]
    resp = vqt.calculate_vqt_frames_in_db(_tones(freqs)).max(axis=1).reshape(len(boundaries), 41)
    spread = resp.max(axis=1) - resp.min(axis=1)
    assert np.all(spread < 3.0), spread


def test_vqt_high_frequencies(vqt):
    # lib.rs:50-72: on-grid tones of every octave respond within 6 dB of each other
    freqs = [np.float32(55.0) * np.float32(2.0) ** np.float32(i + j / (12.0 * 30)) for i in range(7) for j in range(30)]
    top = vqt.calculate_vqt_frames_in_db(_tones(freqs)).max(axis=1)
    assert top.min() > top.max() - 6.0


def test_delay_and_structure_known_answers(vqt):
    # vqt.rs:1078-1085 test_vqt_delay; VQT_REVIEW.md:363-369
    assert int(vqt.delay * 1000) == 98 < 100
    k = vqt.kernel()
    assert [g.window_size() for g in k.window_groups] == [8192, 4096, 2048, 1024]
    assert sum(g.negative_filter_bank.nnz() for g in k.window_groups if g.negative_filter_bank is not None) == 379


def test_results_do_not_depend_on_addresses_hops_or_shards():
    """hi-res parameters, hop 735 (odd: consecutive frames alternate between even and odd sample offsets).  With every
    window group on the per-frame FFT path a frame's bits depend on its samples only: batch == independent frames ==
    per-frame calls == every frame-range shard, whatever the parity of the shard's first frame.  With the sliding
    partial-DFT path on (the default for overlapping frames) the sharded and the unsharded batch are bit-identical
    too (chunk sums do not depend on where a shard starts)."""
    p = pv.VqtParameters.hires()
    hop = synth.HOP_HIRES
    v = pv.Vqt(p)
    n = C.c_int()
    _ffi.load().pvqt_device_count(C.byref(n))
    devs = list(range(n.value)) if n.value > 1 else [0, 0]
    m3 = pv.MultiVqt(p, (devs * 3)[:3])      # 3 shards of 41 frames: f0 = 0, 14, 28 -> sample offsets 0, 10290, 20580 ...
    m2 = pv.MultiVqt(p, devs[:2])
    try:
        n_frames = 41
        audio = synth.polyphonic_chords(3.0, 44100.0, seed=21)[:p.n_fft + (n_frames - 1) * hop]
        frames = np.stack([audio[t * hop:t * hop + p.n_fft] for t in range(n_frames)])
        for sliding in (False, True):
            v.set_sliding_dft(sliding)
            for m in (m2, m3):
                for i in range(len(m.devices)):
                    _ffi.load().pvqt_set_sliding_dft(_ffi.load().pvqt_multi_handle(m._h, i), 2 if sliding else 0)
            batch = v.calculate_vqt_batch_in_db(audio, hop)
            # frame-range shards: 21 + 20 frames (the second shard starts at sample 21 * 735 = 15435: odd), 14 + 14 + 13
            np.testing.assert_array_equal(m2.calculate_vqt_batch_in_db(audio, hop), batch)
            np.testing.assert_array_equal(m3.calculate_vqt_batch_in_db(audio, hop), batch)
            # a batch that starts one frame later holds the same frames at the other address parity
            np.testing.assert_array_equal(v.calculate_vqt_batch_in_db(audio[hop:], hop), batch[1:])
            if not sliding:
                np.testing.assert_array_equal(v.calculate_vqt_frames_in_db(frames), batch)
                for t in (0, 1, 2, 39, 40):
                    np.testing.assert_array_equal(v.calculate_vqt_instant_in_db(frames[t]), batch[t])
    finally:
        m2.close()
        m3.close()
        v.close()


def test_hires_60s_against_the_oracle():
    """The benchmarked hi-res workload (`bench.py --workload hires60`: 60 s at 44.1 kHz, 3511 frames of 1344 bins) against
    the oracle's exact mode (f64 FFTs, all host threads: seconds), every frame."""
    p = pv.VqtParameters.hires()
    v = pv.Vqt(p)
    try:
        audio = synth.polyphonic_chords(60.0, 44100.0, seed=0)
        got = v.calculate_vqt_batch_in_db(audio, synth.HOP_HIRES)
        assert got.shape == (3511, 1344)
        o = orc.OracleVqt(orc.hires_params())
        for g, wg in enumerate(v.kernel().window_groups):
            fb = wg.filter_bank
            o.set_group(g, False, fb.rows, fb.cols, fb.indptr, fb.indices, fb.data)
            nb = wg.negative_filter_bank
            if nb is not None:
                o.set_group(g, True, nb.rows, nb.cols, nb.indptr, nb.indices, nb.data)
        ref = o.calculate_batch_db(audio, synth.HOP_HIRES, mode=0)
        assert np.abs(got - ref).max() <= 1e-3
    finally:
        v.close()


def test_multi_device_requests_fail_loudly():
    # asking for a device the box does not have is an error, never a silent fallback to device 0
    n = C.c_int()
    _ffi.load().pvqt_device_count(C.byref(n))
    with pytest.raises(pv.PvqtRuntimeError):
        pv.MultiVqt(pv.VqtParameters.default(), [0, n.value])
    with pytest.raises(pv.PvqtRuntimeError):
        pv.Vqt(device=n.value)


def test_multi_gpu_on_distinct_devices():
    """pvqt_multi_* over DISTINCT devices, bit-identical to one device.  Needs a box with at least 2 GPUs (the driver's
    scaling step, `gpurun --gpus 2`); on a 1-GPU box the test is skipped, not faked with [0, 0]."""
    n = C.c_int()
    _ffi.load().pvqt_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("one CUDA device on this box")
    v = pv.Vqt()
    m = pv.MultiVqt(pv.VqtParameters.default(), list(range(n.value)))
    try:
        chords = synth.polyphonic_chords(8.0, 22050.0, seed=0)
        np.testing.assert_array_equal(m.calculate_vqt_batch_in_db(chords, synth.HOP_DEFAULT),
                                      v.calculate_vqt_batch_in_db(chords, synth.HOP_DEFAULT))
        nsm = v.n_fft + 4 * synth.HOP_DEFAULT
        streams = np.stack([chords[o:o + nsm] for o in range(0, 9000, 1000)])
        np.testing.assert_array_equal(m.calculate_vqt_streams_in_db(streams, synth.HOP_DEFAULT),
                                      v.calculate_vqt_streams_in_db(streams, synth.HOP_DEFAULT))
    finally:
        m.close()
        v.close()

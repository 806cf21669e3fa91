"""Spectrogram ring, VQT mode (pitchvis_viewer/src/display_system/update.rs:930-1088) and the
model-input windows of ml_system.rs:50-68."""
import numpy as np
import pytest

import orc
import pitchvis_b200 as pv


def _numpy_ring(smoothed, rgb, image, w):
    """Independent restatement of update.rs:959-1080 (numpy f32 arithmetic, one frame at a time)."""
    h, n = image.shape[0], image.shape[1]
    f32 = np.float32
    for x in smoothed:
        mx = f32(max(f32(0.0), x.max()))
        if mx > 0:
            d = f32(1.0) - x / (mx + f32(0.001))
            b = np.clip((f32(1.0) - d * d) * f32(1.5), f32(0.0), f32(1.0))
        else:
            b = np.zeros(n, f32)
        a = np.clip(b * f32(255.0) * f32(1.2), 0.0, 255.0).astype(np.uint8)     # truncating
        row = h - 1 - w
        image[row, :, :3] = rgb
        image[row, :, 3] = a
        w = (w + 1) % h
        image[h - 1 - w] = 0
    return w


def _inputs(frames, n, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 45.0, (frames, n)).astype(np.float32)
    x[rng.uniform(size=x.shape) < 0.3] = 0.0
    if frames > 2:
        x[1] = 0.0                                   # silence: brightness 0 (update.rs:973-975)
    return x, rng.integers(0, 256, (n, 3), dtype=np.uint8)


def test_oracle_ring_matches_numpy_restatement():
    for frames, n, h, w0 in ((7, 588, 16, 0), (40, 588, 16, 5), (3, 48, 2, 1), (5, 12, 1, 0), (16, 100, 17, 16)):
        x, rgb = _inputs(frames, n, frames)
        a = np.random.default_rng(9).integers(0, 256, (h, n, 4), dtype=np.uint8)
        b = a.copy()
        wa = orc.spectrogram_vqt(x, rgb, a, w0)
        wb = _numpy_ring(x, rgb, b, w0)
        assert wa == wb == (w0 + frames) % h
        np.testing.assert_array_equal(a, b)
    # known answers: the loudest bin saturates (1 - 0.001/(m + 0.001) squared away, times 1.5 -> clamp 1 -> 255 * 1.2 -> 255);
    # half the maximum gives (1 - 0.25) * 1.5 > 1 as well; a tenth gives (1 - 0.81) * 1.5 * 306 = 87
    img = np.zeros((4, 3, 4), np.uint8)
    orc.spectrogram_vqt(np.array([[40.0, 20.0, 4.0]], np.float32), np.zeros((3, 3), np.uint8), img, 0)
    assert img[3, :, 3].tolist() == [255, 255, 87] and not img[:3].any()


def test_ml_input_windows_are_views_of_the_history():
    h = np.arange(7 * 5, dtype=np.float32).reshape(7, 5)
    w = pv.ml_input_windows(h, 3)
    assert w.shape == (5, 15)
    for i in range(5):
        np.testing.assert_array_equal(w[i], h[i:i + 3].ravel())     # ml_system.rs:56-60: frames len-T+i, i < T
    with pytest.raises(ValueError):
        pv.ml_input_windows(h, 8)


@pytest.mark.gpu
def test_gpu_ring_is_bit_identical_to_the_oracle(built_lib):
    for frames, n, h, w0 in ((1, 588, 256, 0), (100, 588, 256, 250), (700, 588, 256, 3), (9, 48, 2, 1), (4, 12, 1, 0),
                             (255, 1344, 256, 17)):
        x, rgb = _inputs(frames, n, frames + 1)
        ref = np.random.default_rng(3).integers(0, 256, (h, n, 4), dtype=np.uint8)
        got = ref.copy()
        w_ref = orc.spectrogram_vqt(x, rgb, ref, w0)
        w_got = pv.spectrogram_vqt(x, rgb, got, w0)
        assert w_got == w_ref
        np.testing.assert_array_equal(got, ref)
    # two batched calls = one
    x, rgb = _inputs(300, 588, 77)
    a = np.zeros((64, 588, 4), np.uint8)
    b = a.copy()
    w = pv.spectrogram_vqt(x[:123], rgb, a, 0)
    w = pv.spectrogram_vqt(x[123:], rgb, a, w)
    assert w == pv.spectrogram_vqt(x, rgb, b, 0)
    np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        pv.spectrogram_vqt(x, rgb, np.zeros((64, 100, 4), np.uint8), 0)
    with pytest.raises(pv.PvqtRuntimeError):
        pv.spectrogram_vqt(x, rgb, a, 64)            # write index outside the ring


# ---- SpectrogramMode::Peaks (update.rs:997-1062) --------------------------------------------------------------------
def _peak_lists(frames, max_peaks, n, seed):
    rng = np.random.default_rng(seed)
    pk = np.zeros((frames, max_peaks, 2), np.float32)
    cnt = rng.integers(0, max_peaks + 1, frames).astype(np.uint32)
    for t in range(frames):
        c = np.sort(rng.uniform(0.0, n - 1.0, int(cnt[t]))).astype(np.float32)
        pk[t, :cnt[t], 0] = c
        pk[t, :cnt[t], 1] = rng.uniform(0.0, 40.0, int(cnt[t])).astype(np.float32)
    if frames > 2:
        cnt[1] = 0                                   # a frame without peaks leaves its (cleared) row alone
        pk[2, :cnt[2], 1] = 0.0                      # max_size == 0: nothing is drawn (update.rs:1008)
    if frames > 3 and cnt[3] >= 2:
        pk[3, 1, 0] = pk[3, 0, 0] + np.float32(0.7)  # overlapping discs: the later peak owns the shared bins
    return pk, cnt


def test_oracle_peaks_mode_known_answers():
    # a peak exactly on a semitone takes the crate's colour of that pitch class (times 1.2, clamped): bin 0 of the
    # default range is A (55 Hz): semitone_offset = 84 - 21 = 63 -> 12 * 63 / 84 = 9 -> COLORS[9] = (1.00, 0.96, 0.03)
    img = np.zeros((4, 588, 4), np.uint8)
    pk = np.array([[[84.0, 30.0], [200.5, 15.0]]], np.float32)
    w = orc.spectrogram_peaks(pk, np.array([2], np.uint32), img, 0)
    assert w == 1 and not img[:3].any()
    row = img[3]
    assert row[84].tolist() == [255, 255, int(np.float32(7 / 255) * np.float32(255) * np.float32(1.2)), 255]
    lit = np.nonzero(row[:, 3])[0].tolist()
    # bins floor(c - 2) .. ceil(c + 2) - 1 with |bin - c| <= 2: the half-open range drops bin 86 of the integer centre
    # (update.rs:1031-1034), an asymmetry of the reference that is kept; alpha at distance 2 is exp(-2) * 306 = 41
    assert lit == [82, 83, 84, 85, 199, 200, 201, 202]
    assert row[82, 3] == 41 and row[83, 3] == row[85, 3] == 185
    # the weaker peak: brightness (1 - 0.25) * 1.5 clamps to 1, falloff exp(-0.25 / 2) -> 270 -> clamp 255
    assert row[200, 3] == 255 and row[199, 3] == int(np.float32(np.exp(np.float32(-2.25 / 2.0))) * np.float32(306.0))
    # colours are the LCh round trip of the crate: on a semitone exactly the table colour, halfway the L = 60 gray
    np.testing.assert_array_equal((orc.calculate_color(84, 7.0) * 255).round(), [2, 132, 181])
    np.testing.assert_array_equal((orc.calculate_color(84, 3.5) * 255).round(), [145, 145, 145])


@pytest.mark.gpu
def test_gpu_peaks_ring_is_bit_identical_to_the_oracle(built_lib):
    for frames, mp, h, w0, rng_ in ((1, 64, 256, 0, pv.VqtRange()), (90, 64, 256, 200, pv.VqtRange()),
                                    (300, 32, 64, 5, pv.VqtRange()), (7, 16, 2, 1, pv.VqtRange(55.0, 2, 24)),
                                    (5, 8, 1, 0, pv.VqtRange(55.0, 2, 24)), (40, 64, 50, 49, pv.VqtRange(55.0, 8, 168))):
        n = rng_.n_buckets()
        pk, cnt = _peak_lists(frames, mp, n, frames + mp)
        ref = np.random.default_rng(3).integers(0, 256, (h, n, 4), dtype=np.uint8)
        got = ref.copy()
        w_ref = orc.spectrogram_peaks(pk, cnt, ref, w0, rng_.buckets_per_octave)
        w_got = pv.spectrogram_peaks(pk, cnt, got, w0, rng_)
        assert w_got == w_ref
        np.testing.assert_array_equal(got, ref)
    # fed from K-analysis: the peaks of a real recording, two batched calls = one
    from pitchvis_b200 import synth
    v = pv.Vqt()
    a = pv.AnalysisState(pv.VqtRange())
    res = a.calculate_and_preprocess(v, synth.polyphonic_chords(6.0, 22050.0, seed=5), synth.HOP_DEFAULT, 16_689_342)
    pk, cnt = res["peaks_continuous"][0], res["peak_count"][0]
    assert cnt.max() > 3
    ref = np.zeros((128, 588, 4), np.uint8)
    got, two = ref.copy(), ref.copy()
    w_ref = orc.spectrogram_peaks(pk, cnt, ref, 0)
    assert pv.spectrogram_peaks(pk, cnt, got, 0) == w_ref
    np.testing.assert_array_equal(got, ref)
    w = pv.spectrogram_peaks(pk[:100], cnt[:100], two, 0)
    assert pv.spectrogram_peaks(pk[100:], cnt[100:], two, w) == w_ref
    np.testing.assert_array_equal(two, ref)
    a.close()
    v.close()

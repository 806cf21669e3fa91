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

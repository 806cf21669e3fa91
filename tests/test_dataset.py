"""Dataset caller: the batched call reproduces the reference's chunk loop
(pitchvis_train/src/train.rs:252-351), and the .npy layout matches train.rs:156-208."""
import os

import numpy as np
import pytest

import orc
import pitchvis_b200 as pv
from pitchvis_b200 import dataset, synth


def test_npy_layout(tmp_path):
    x = np.arange(3 * 252, dtype=np.float32).reshape(3, 252)
    t = np.ones((3, 128), np.float32)
    p = os.path.join(tmp_path, "data.npy")
    n = dataset.write_dataset_npy(p, x, t)
    d = np.load(p)
    assert n == 3 * 380 and d.shape == (3 * 380,) and d.dtype == np.dtype("<f4")
    np.testing.assert_array_equal(d.reshape(3, 380)[:, :252], x)
    np.testing.assert_array_equal(d.reshape(3, 380)[:, 252:], t)
    with pytest.raises(ValueError):
        dataset.write_dataset_npy(p, x, np.ones((3, 100), np.float32))


def test_train_parameters_on_the_oracle():
    # train.rs:30-41 -> 252 bins; the kernel builds (no VqtError) and its delay gives the reference's chunk length
    op = orc.make_params(n_fft=32768, min_freq=55.0, octaves=7, buckets_per_octave=36, quality=10.0, gamma=53.0)
    v = orc.OracleVqt(op)
    assert v.n_buckets == 252
    delay_ms = int(v.delay * 1000.0)
    chunk = (delay_ms * 22050 // 1000) // 64 * 64
    assert chunk > 0 and chunk % 64 == 0


@pytest.mark.gpu
def test_batched_call_equals_the_reference_chunk_loop(built_lib):
    vqt = pv.Vqt(dataset.TRAIN_PARAMS)
    try:
        assert vqt.n_buckets == 252
        chunk = dataset.train_chunk_samples(vqt)
        audio = (synth.polyphonic_chords(3.0, 22050.0, seed=21) * np.float32(0.3)).astype(np.float32)
        got = dataset.annotated_vqt(vqt, audio, agc=pv.MonoAgc(0.07, 0.001))
        # the reference loop, restated with the oracles: ring buffer of 2 * SR zeros, AGC per chunk, VQT of the last
        # N_FFT samples every third chunk (train.rs:276-341)
        op = orc.make_params(n_fft=32768, min_freq=55.0, octaves=7, buckets_per_octave=36, quality=10.0, gamma=53.0)
        o = orc.OracleVqt(op)
        ring = np.zeros(2 * 22050, np.float32)
        gain = 1.0
        ref = []
        for c in range(audio.shape[0] // chunk):
            x = audio[c * chunk:(c + 1) * chunk]
            y, gain = orc.agc_process_chunks(x, chunk, 0.07, 0.001, 1e-6, gain)
            ring = np.concatenate([ring[chunk:], y])
            if (c + 1) % 3 == 0:
                ref.append(o.calculate_vqt_instant_in_db(ring[-32768:], 0))
        ref = np.stack(ref)
        assert got.shape == ref.shape and ref.shape[0] >= 5
        assert np.abs(got - ref).max() <= 1e-3
    finally:
        vqt.close()

"""AGC pre-stage: dagc_fork::MonoAgc.

CPU: the oracle restates the reference's own test (dagc_fork/src/lib.rs:93-108) and the constructor checks.
GPU: K-agc through the C ABI is bit-identical to the oracle (same f32 operations in the same order)."""
import numpy as np
import pytest

import orc
import pitchvis_b200 as pv
from pitchvis_b200 import synth


def test_oracle_it_works():
    # lib.rs:93-108: frozen -> gain stays 1.0 and samples unchanged; unfrozen -> gain moves
    x = np.array([0.5, 1.0, -0.2], np.float32)
    y, g = orc.agc_process(x, 0.001, 0.0001, 1.0, frozen=True)
    assert g == 1.0 and np.array_equal(y, x)
    y, g = orc.agc_process(x, 0.001, 0.0001, 1.0, frozen=False)
    assert g != 1.0
    # first sample is scaled by the initial gain 1.0, the update uses the scaled sample (lib.rs:77-84)
    g1 = np.float32(1.0) + np.float32(0.0001) * (np.float32(1.0) - np.float32(0.25) / np.float32(0.001))
    g1 = max(g1, np.float32(0.0001))
    assert y[0] == x[0] and y[1] == np.float32(x[1] * g1)


def test_oracle_constructor_checks():
    # lib.rs:35-47
    assert orc.agc_check(0.07, 0.0001) == 0
    assert orc.agc_check(0.0, 0.0001) == 1 and orc.agc_check(float("inf"), 0.1) == 1 and orc.agc_check(float("nan"), 0.1) == 1
    assert orc.agc_check(0.07, -0.1) == 2 and orc.agc_check(0.07, 1.5) == 2 and orc.agc_check(0.07, float("nan")) == 2
    assert orc.agc_check(0.07, 0.0) == 0 and orc.agc_check(0.07, 1.0) == 0


def test_oracle_converges_to_target_and_freezes_on_silence():
    # a steady tone is driven towards mean(x^2) == desired_output_rms (the recurrence's fixed point), and silent
    # chunks (sum x^2 < 1e-6, audio_desktop.rs:106-107) leave the gain untouched
    sr = 22050
    t = np.arange(4 * sr) / sr
    x = (0.01 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    y, g = orc.agc_process_chunks(x, 512, 0.07, 0.001)
    tail = y[-sr:]
    assert abs(float(np.mean(tail.astype(np.float64) ** 2)) - 0.07) < 0.01
    z = np.zeros(2048, np.float32)
    _, g2 = orc.agc_process_chunks(z, 512, 0.07, 0.001, gain=g)
    assert g2 == g
    _, g3 = orc.agc_process_chunks(z, 512, 0.07, 0.001, silence_threshold=-1.0, gain=g)
    assert g3 > g  # never frozen: silence pushes the gain up


@pytest.mark.gpu
def test_gpu_matches_oracle_bit_for_bit(built_lib):
    rng = np.random.default_rng(3)
    streams = np.stack([synth.polyphonic_chords(1.0, 22050.0, seed=s)[:20000] * np.float32(a)
                        for s, a in ((0, 1.0), (1, 0.05), (2, 3.0))])
    streams[1, 5000:9000] = 0.0  # a silent stretch: frozen chunks
    streams = streams.astype(np.float32)
    for chunk, thr in ((441, 1e-6), (1024, 1e-6), (0, -1.0), (300, 1e-6)):
        agc = pv.MonoAgc(0.07, 0.0001, n_streams=3)
        got = agc.process_chunks(streams, chunk, thr)
        gains = agc.gains
        for s in range(3):
            ref, g = orc.agc_process_chunks(streams[s], chunk, 0.07, 0.0001, thr)
            np.testing.assert_array_equal(got[s], ref)
            assert gains[s] == np.float32(g)
        # state carries over between calls: two halves == one call
        agc2 = pv.MonoAgc(0.07, 0.0001, n_streams=3)
        h = (streams.shape[1] // 2 // max(chunk, 1)) * max(chunk, 1) if chunk else streams.shape[1] // 2
        if chunk:
            a = agc2.process_chunks(streams[:, :h], chunk, thr)
            b = agc2.process_chunks(streams[:, h:], chunk, thr)
            np.testing.assert_array_equal(np.concatenate([a, b], axis=1), got)
        agc.close(); agc2.close()


@pytest.mark.gpu
def test_gpu_reference_api_and_errors(built_lib):
    # the reference's own test through the mirror (lib.rs:93-108)
    agc = pv.MonoAgc(0.001, 0.0001)
    assert agc.gain() == 1.0 and not agc.is_gain_frozen()
    agc.freeze_gain(True)
    assert agc.is_gain_frozen()
    x = np.array([0.5, 1.0, -0.2], np.float32)
    y = agc.process(x)
    assert agc.gain() == 1.0 and np.array_equal(y, x)
    agc.freeze_gain(False)
    agc.process(x)
    assert agc.gain() != 1.0
    agc.close()
    with pytest.raises(pv.AgcError):
        pv.MonoAgc(0.0, 0.0001)
    with pytest.raises(pv.AgcError):
        pv.MonoAgc(0.07, 1.5)


@pytest.mark.gpu
def test_gpu_agc_then_vqt_pipeline(built_lib, oracle_default):
    # AGC -> VQT, the order of the reference's audio callback and dataset loop, against the two oracles chained
    audio = (synth.polyphonic_chords(2.5, 22050.0, seed=11) * np.float32(0.2)).astype(np.float32)
    agc = pv.MonoAgc(0.07, 0.001)
    levelled = agc.process_chunks(audio, 2112)       # train.rs:128-129: chunk = delay rounded down to x64 = 2112
    ref_levelled, _ = orc.agc_process_chunks(audio, 2112, 0.07, 0.001)
    np.testing.assert_array_equal(levelled, ref_levelled)
    v = pv.Vqt(pv.VqtParameters.default())
    hop = 3 * 2112                                   # train.rs:43: every STEP_SIZE_IN_CHUNKS = 3 chunks
    got = v.calculate_vqt_batch_in_db(levelled, hop)
    ref = oracle_default.calculate_batch_db(ref_levelled, hop, mode=0)
    assert got.shape == ref.shape and got.shape[0] >= 3
    assert np.abs(got - ref).max() <= 1e-3
    v.close(); agc.close()

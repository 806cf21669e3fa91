"""Chroma reduction (pitchvis_viewer/src/display_system/update.rs:1104-1131)."""
import numpy as np
import pytest

import orc
import pitchvis_b200 as pv
from pitchvis_b200 import synth


def test_oracle_single_tone_lands_on_its_pitch_class(oracle_default):
    # 440 Hz = A: pitch class 9 relative to C (update.rs:1107-1112: C4 = 261.626 Hz is class 0)
    db = oracle_default.calculate_vqt_instant_in_db(orc.test_create_sines(oracle_default.params, [440.0]), 0)
    c = orc.chroma(db)
    assert c.shape == (12,) and int(c.argmax()) == 9 and c.max() == 1.0 and c.min() >= 0.0
    # silence: every power is 10^0 = 1, classes hold equal bin counts up to rounding -> all within [0.9, 1]
    z = orc.chroma(np.zeros(588, np.float32))
    assert z.max() == 1.0 and z.min() > 0.9


@pytest.mark.gpu
def test_gpu_chroma_matches_oracle(built_lib):
    v = pv.Vqt(pv.VqtParameters.default())
    try:
        audio = synth.polyphonic_chords(3.0, 22050.0, seed=5)
        db = v.calculate_vqt_batch_in_db(audio, synth.HOP_DEFAULT)
        got = pv.chroma(db)
        assert got.shape == (db.shape[0], 12)
        ref = np.stack([orc.chroma(db[t]) for t in range(db.shape[0])])
        assert np.abs(got - ref).max() <= 2e-5          # device powf vs glibc powf: an ulp or two
        assert np.all(got.max(axis=1) == 1.0)
        # other range
        rng = pv.VqtRange(110.0, 2, 24)
        x = np.random.default_rng(1).uniform(0, 40, (5, 48)).astype(np.float32)
        assert np.abs(pv.chroma(x, rng) - np.stack([orc.chroma(r, 110.0, 24) for r in x])).max() <= 2e-5
        with pytest.raises(ValueError):
            pv.chroma(np.zeros((2, 100), np.float32))
    finally:
        v.close()

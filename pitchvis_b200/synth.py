"""Synthetic audio of the shapes BASELINE.json names.

`test_create_sines` restates the reference's test-signal generator
(pitchvis_analysis/src/util.rs:62-79); `polyphonic_chords` is the "random chords"
recording used by configs 2-5.  Pure numpy; used by tests/ and bench.py.
"""
from __future__ import annotations

import numpy as np

HOP_DEFAULT = 368    # round(22050 / 60): the viewer's 60 FPS call rate
HOP_HIRES = 735      # 44100 / 60


def test_create_sines(sr: float, n_fft: int, freqs, t_diff: float = 0.0) -> np.ndarray:
    """util.rs:62-79: amplitude 1/12 per sine, phase evaluated in f32, left to right."""
    f32 = np.float32
    i = np.arange(n_fft, dtype=np.float32)
    wave = np.zeros(n_fft, np.float32)
    for f in freqs:
        a = i + f32(t_diff) * f32(sr)
        a = a * f32(2.0)
        a = a * f32(np.pi)
        a = a / f32(sr)
        a = a * f32(f)
        wave += (np.sin(a) / f32(12.0)).astype(np.float32)
    return wave


test_create_sines.__test__ = False  # not a pytest test


def polyphonic_chords(seconds: float, sr: float = 22050.0, seed: int = 0) -> np.ndarray:
    """Random chords: every 0.5 s draw 3-6 distinct MIDI notes in [33, 117);
    each note has 6 harmonics of amplitude a/h (below Nyquist), a ~ U[0.02, 0.1], random phase, 10 ms
    raised-cosine on/off ramps; plus white noise (sigma 1e-3); scaled to RMS 0.07; float32."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    seg = int(round(0.5 * sr))
    ramp = int(round(0.010 * sr))
    out = np.zeros(n, np.float64)
    t_seg = np.arange(seg, dtype=np.float64) / sr
    env = np.ones(seg)
    r = 0.5 - 0.5 * np.cos(np.pi * np.arange(ramp) / ramp)
    env[:ramp] = r
    env[seg - ramp:] = r[::-1]
    for s0 in range(0, n, seg):
        m = min(seg, n - s0)
        k = int(rng.integers(3, 7))
        notes = rng.choice(np.arange(33, 117), size=k, replace=False)
        chunk = np.zeros(seg)
        for note in notes:
            f0 = 440.0 * 2.0 ** ((note - 69) / 12.0)
            a = rng.uniform(0.02, 0.1)
            for h in range(1, 7):
                if h * f0 >= sr / 2:
                    break
                ph = rng.uniform(0.0, 2.0 * np.pi)
                chunk += (a / h) * np.sin(2.0 * np.pi * h * f0 * t_seg + ph)
        out[s0:s0 + m] = (chunk * env)[:m]
    out += rng.normal(0.0, 1e-3, n)
    rms = np.sqrt(np.mean(out * out))
    out *= 0.07 / rms
    return out.astype(np.float32)


def frames_in(n_samples: int, n_fft: int, hop: int) -> int:
    return 0 if n_samples < n_fft else (n_samples - n_fft) // hop + 1

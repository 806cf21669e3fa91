"""K-analysis on many streams (device time of the launch through the host entry): python scripts/analysis_many.py [streams] [frames]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitchvis_b200 as pv
from pitchvis_b200 import synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
v = pv.Vqt()
audio = synth.polyphonic_chords(12.0, 22050.0, seed=0)
db = v.calculate_vqt_batch_in_db(audio, synth.HOP_DEFAULT)[:T]
many = np.ascontiguousarray(np.stack([np.roll(db, s % T, axis=0) for s in range(S)]))
a = pv.AnalysisState(pv.VqtRange(), n_streams=S)
a.preprocess_batch(many[:, :8], 16_689_342, vectors=False)
t0 = time.perf_counter(); a.preprocess_batch(many, 16_689_342, vectors=False); dt = time.perf_counter() - t0
print(f"{S} streams x {T} frames: {dt * 1e3:.1f} ms per call (host dB in, {many.nbytes / 1e6:.0f} MB)")

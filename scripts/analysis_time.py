"""Wall time of K-analysis on one 60 s recording (3507 frames, scalars and peaks back, no per-bin vectors) and on 256 streams.
   python scripts/analysis_time.py   (any build; PVQT_LIB selects a variant)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitchvis_b200 as pv
from pitchvis_b200 import synth
v = pv.Vqt()
audio = synth.polyphonic_chords(60.0, 22050.0, seed=0)
db = v.calculate_vqt_batch_in_db(audio, synth.HOP_DEFAULT)
n = db.shape[0]
st = pv.AnalysisState(pv.VqtRange())
st.preprocess_batch(db[:64], 16_689_342, vectors=False)
best = 1e9
for _ in range(5):
    t0 = time.perf_counter(); st.preprocess_batch(db, 16_689_342, vectors=False); best = min(best, time.perf_counter() - t0)
print(f"one stream: {n} frames in {best * 1e3:.1f} ms = {best / n * 1e6:.2f} us per frame")
S = 256
many = np.ascontiguousarray(np.stack([np.roll(db[:511], s, axis=0) for s in range(S)]))
sm = pv.AnalysisState(pv.VqtRange(), n_streams=S)
sm.preprocess_batch(many[:, :16], 16_689_342, vectors=False)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); sm.preprocess_batch(many, 16_689_342, vectors=False); best = min(best, time.perf_counter() - t0)
print(f"{S} streams x 511 frames in {best * 1e3:.1f} ms = {S * 511 / best / 1e6:.2f} M frames/s (host dB in: {many.nbytes / 1e6:.0f} MB)")

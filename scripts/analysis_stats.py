"""Per-phase cycles of K-analysis (a -DPVQT_ANALYSIS_STATS build: scripts/build_variant.sh astats -DPVQT_ANALYSIS_STATS).
   PVQT_LIB=pitchvis_b200/lib/libpvqt_astats.so python scripts/analysis_stats.py"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt()
audio = synth.polyphonic_chords(60.0, 22050.0, seed=0)
db = v.calculate_vqt_batch_in_db(audio, synth.HOP_DEFAULT)
n = db.shape[0]
buf = np.zeros(16, np.uint64)
lib.pvqt_debug_analysis_stats.argtypes = [C.c_void_p, C.c_int]
st = pv.AnalysisState(pv.VqtRange())
st.preprocess_batch(db[:64], 16_689_342)
lib.pvqt_debug_analysis_stats(buf.ctypes.data, 1)
t0 = time.perf_counter()
res = st.preprocess_batch(db, 16_689_342)
dt = time.perf_counter() - t0
lib.pvqt_debug_analysis_stats(buf.ctypes.data, 1)
names = ["EMA of the spectrum", "three peak searches", "compaction", "enhance + bass promotion (per peak) | the two 588-term sums", "afterglow / calmness (per bin)",
         "sequential passes over the peaks", "outputs", "frame total", "  peak search: maxima", "  peak search: min distance", "  peak search: prominence (thread 0's warp)"]
print(f"{n} frames, {dt * 1e3:.1f} ms per call = {dt / n * 1e6:.1f} us per frame (host buffers in and out); cycles per frame, thread 0:")
for i, nm in enumerate(names):
    if not nm: continue
    if 8 <= i < 11: nm = nm + " [bass search group only]"
    print(f"  {nm:38s} {buf[i] / n:9.0f}")

"""Per-frame latency of VQT + AnalysisState through one library call (pvqt_calc_batch_analysis with one frame): what the
reference's viewer does 60 times a second (vqt_system.rs:40-68 + analysis_system.rs:10-20)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt()
a = pv.AnalysisState(pv.VqtRange())
audio = synth.polyphonic_chords(8.0, 22050.0, seed=0)
n_fft, hop = 32768, synth.HOP_DEFAULT
FT = 16_689_342
out = _ffi.PvqtAnalysisOutputs()
mp = 32
bufs = {"peak_count": np.zeros(1, np.uint32), "peak_indices": np.zeros(mp, np.uint32), "peaks_continuous": np.zeros((mp, 2), np.float32),
        "smoothed_scene_calmness": np.zeros(1, np.float32), "smoothed_tuning_grid_inaccuracy": np.zeros(1, np.float32)}
out.max_peaks = mp
for k, b in bufs.items():
    setattr(out, k, b.ctypes.data_as(C.c_void_p))
moved = C.c_uint64(0)
FP = C.POINTER(C.c_float)
ts = []
for t in range(300):
    x = np.ascontiguousarray(audio[t * hop:t * hop + n_fft])
    t0 = time.perf_counter()
    rc = lib.pvqt_calc_batch_analysis(v.handle, a._h, x.ctypes.data_as(FP), n_fft, hop, 1, FT, C.byref(out), None, C.byref(moved))
    ts.append(time.perf_counter() - t0)
    assert rc == 0, _ffi.last_error()
ts = np.array(ts[50:]) * 1e6
print(f"VQT + AnalysisState, one frame per call: p50 {np.percentile(ts, 50):.1f} us, p99 {np.percentile(ts, 99):.1f} us, min {ts.min():.1f} us "
      f"({int(bufs['peak_count'][0])} peaks in the last frame)")

"""Per-phase cycles of K-fft work items (a -DPVQT_FFT_STATS build: scripts/build_variant.sh fftstats -DPVQT_FFT_STATS).
   PVQT_LIB=pitchvis_b200/lib/libpvqt_fftstats.so python scripts/fft_stats.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt()
audio = synth.polyphonic_chords(60.0, 22050.0, seed=0)
n = v.frames_in(audio.shape[0], synth.HOP_DEFAULT)
d_a = pv.DeviceBuffer(v, audio.nbytes); d_a.upload(audio)
d_o = pv.DeviceBuffer(v, n * 588 * 4)
buf = np.zeros((16, 16), np.uint64)
lib.pvqt_debug_fft_stats.argtypes = [C.c_void_p, C.c_int]
for _ in range(3):
    pv.calc_db_device(v, d_a, 1, 0, synth.HOP_DEFAULT, n, d_o)
lib.pvqt_debug_fft_stats(buf.ctypes.data, 1)
reps = 10
for _ in range(reps):
    pv.calc_db_device(v, d_a, 1, 0, synth.HOP_DEFAULT, n, d_o)
lib.pvqt_debug_fft_stats(buf.ctypes.data, 1)
names = ["addressing+edge", "p0 loads issued", "p0 butterflies+stores", "p0 barrier", "p1 loads", "p1 bfly+stores", "p1 barrier",
         "p2 loads", "p2 bfly+stores", "p2 barrier", "p3 loads", "p3 bfly+stores", "p3 barrier", "split step", "last barrier"]
print("mean cycles of thread 0 per work item and phase (a stamp after a load / barrier is its issue time: the wait shows in the next phase)")
for l2 in range(16):
    ctas = int(buf[l2, 15])
    if ctas == 0:
        continue
    per = buf[l2, :15].astype(np.float64) / ctas
    print(f"N_c = {1 << l2}: {ctas // reps} items per launch, {per.sum():.0f} cycles per item = {per.sum() / 1965:.2f} us")
    for i, nm in enumerate(names):
        if per[i] > 0:
            print(f"   {nm:24s} {per[i]:8.0f}")

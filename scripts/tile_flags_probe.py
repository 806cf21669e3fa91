import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt(); v.set_sliding_dft(0)
audio = synth.polyphonic_chords(60.0, 22050.0, seed=0); hop = 368
n = synth.frames_in(audio.shape[0], v.n_fft, hop)
d_audio = pv.DeviceBuffer(v, audio.nbytes); d_audio.upload(audio)
d_out = pv.DeviceBuffer(v, n * 588 * 4); d_flush = pv.DeviceBuffer(v, 512 << 20)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.pvqt_event_create(v.handle, C.byref(e0)); lib.pvqt_event_create(v.handle, C.byref(e1))
ts = []
for i in range(30):
    lib.pvqt_dev_flush_l2(v.handle, d_flush.ptr, 512 << 20)
    lib.pvqt_event_record(v.handle, e0)
    pv.calc_db_device(v, d_audio, 1, 0, hop, n, d_out)
    lib.pvqt_event_record(v.handle, e1)
    ms = C.c_float(); lib.pvqt_event_elapsed_ms(v.handle, e0, e1, C.byref(ms)); ts.append(ms.value * 1e3)
print("PVQT_TILE_FLAGS=%s, sliding DFT off: step median %.1f us" % (os.environ.get("PVQT_TILE_FLAGS"), float(np.median(ts[5:]))))

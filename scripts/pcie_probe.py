"""Diagnostic (GPU): pinned H2D / D2H bandwidth and per-call time of the host entry point."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt()
h = v.handle
n = 64 << 20
pin = C.c_void_p(); lib.pvqt_host_alloc_pinned(n, C.byref(pin))
dev = pv.DeviceBuffer(v, n)
for nbytes in (1 << 20, 5292000, 8248464, 64 << 20):
    for name, fn in (("h2d", lambda: lib.pvqt_memcpy_h2d(h, dev.ptr, pin, nbytes, 0)), ("d2h", lambda: lib.pvqt_memcpy_d2h(h, pin, dev.ptr, nbytes, 0))):
        fn(); t0 = time.perf_counter()
        for _ in range(10): fn()
        dt = (time.perf_counter() - t0) / 10
        print(f"{name} {nbytes/1e6:8.2f} MB  {dt*1e6:8.1f} us  {nbytes/dt/1e9:6.1f} GB/s")
audio = synth.polyphonic_chords(60.0, 22050.0, 0)
nf = synth.frames_in(audio.shape[0], v.n_fft, 368)
pin_in, pin_out = C.c_void_p(), C.c_void_p()
lib.pvqt_host_alloc_pinned(audio.nbytes, C.byref(pin_in)); lib.pvqt_host_alloc_pinned(nf * 588 * 4, C.byref(pin_out))
C.memmove(pin_in, audio.ctypes.data, audio.nbytes)
fp = C.POINTER(C.c_float)
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10): lib.pvqt_calc_batch_db(h, C.cast(pin_in, fp), audio.shape[0], 368, nf, C.cast(pin_out, fp))
    print("host entry, pinned, per call us:", (time.perf_counter() - t0) / 10 * 1e6, "launch_count", v.launch_count)
out = np.empty((nf, 588), np.float32)
t0 = time.perf_counter()
for _ in range(5): v.calculate_vqt_batch_in_db(audio, 368, out=out)
print("host entry, pageable numpy, per call us:", (time.perf_counter() - t0) / 5 * 1e6)

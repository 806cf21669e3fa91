"""Diagnostic (GPU): device-timed throughput when the launch is large enough to fill the GPU (BASELINE configs[2]
in small: S streams x 10 s, 511 frames each), beside the 3507-frame bench workload.  Not a bench line."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt()
h = v.handle
one = synth.polyphonic_chords(10.0, 22050.0, seed=0)
for S in (16, 128, 1024):
    audio = np.tile(one, (S, 1))
    fps = synth.frames_in(one.shape[0], v.n_fft, 368)
    d_audio = pv.DeviceBuffer(v, audio.nbytes); d_audio.upload(audio)
    d_out = pv.DeviceBuffer(v, S * fps * 588 * 4)
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.pvqt_event_create(h, C.byref(e0)); lib.pvqt_event_create(h, C.byref(e1))
    for _ in range(2): pv.calc_db_device(v, d_audio, S, one.shape[0], 368, fps, d_out)
    pv.synchronize(v)
    reps = 5
    lib.pvqt_event_record(h, e0)
    for _ in range(reps): pv.calc_db_device(v, d_audio, S, one.shape[0], 368, fps, d_out)
    lib.pvqt_event_record(h, e1)
    ms = C.c_float(); lib.pvqt_event_elapsed_ms(h, e0, e1, C.byref(ms))
    frames = S * fps * reps
    print(f"{S:5d} streams x {fps} frames: {frames / (ms.value * 1e-3) / 1e6:7.2f} M frames/s  ({ms.value / reps:8.3f} ms per pass, "
          f"{frames / reps * 35120 / (ms.value / reps * 1e-3) / 1e9:7.1f} GB/s algorithmic)")
    d_audio.free(); d_out.free()
# per-kernel split of one pass at 128 streams
S = 128
audio = np.tile(one, (S, 1)); fps = synth.frames_in(one.shape[0], v.n_fft, 368)
d_audio = pv.DeviceBuffer(v, audio.nbytes); d_audio.upload(audio); d_out = pv.DeviceBuffer(v, S * fps * 588 * 4)
lib.pvqt_set_profiling(h, 1)
pv.calc_db_device(v, d_audio, S, one.shape[0], 368, fps, d_out)
k_ms = (C.c_double * _ffi.PROFILE_KINDS)(); k_n = (C.c_uint64 * _ffi.PROFILE_KINDS)()
lib.pvqt_get_profile(h, 1, k_ms, k_n); lib.pvqt_set_profiling(h, 0)
print({_ffi.KERNEL_KIND_NAMES[i]: (round(k_ms[i], 3), int(k_n[i])) for i in range(_ffi.PROFILE_KINDS) if k_n[i]})

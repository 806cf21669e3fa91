"""Diagnostic (GPU): device-timed throughput of a large stream batch against PVQT_CHUNK_FRAMES (frames per launch)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for c in (8192, 16384, 32768, 65536, 131072, 262144):
        env = dict(os.environ, PVQT_CHUNK_FRAMES=str(c))
        subprocess.run([sys.executable, __file__, "run"], env=env)
    sys.exit(0)
sys.path.insert(0, ROOT)
import numpy as np
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt(); h = v.handle
one = synth.polyphonic_chords(10.0, 22050.0, seed=0)        # BASELINE configs[2] shape: 511 frames per stream
fps = synth.frames_in(one.shape[0], v.n_fft, 368)
S = 1024
audio = np.stack([np.roll(one, 1000 * s) for s in range(S)])
d_audio = pv.DeviceBuffer(v, audio.nbytes); d_audio.upload(audio)
d_out = pv.DeviceBuffer(v, S * fps * 588 * 4)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.pvqt_event_create(h, C.byref(e0)); lib.pvqt_event_create(h, C.byref(e1))
for _ in range(2): pv.calc_db_device(v, d_audio, S, one.shape[0], 368, fps, d_out)
pv.synchronize(v)
reps = 3
lib.pvqt_event_record(h, e0)
for _ in range(reps): pv.calc_db_device(v, d_audio, S, one.shape[0], 368, fps, d_out)
lib.pvqt_event_record(h, e1)
ms = C.c_float(); lib.pvqt_event_elapsed_ms(h, e0, e1, C.byref(ms))
print(f"chunk {os.environ.get('PVQT_CHUNK_FRAMES')}: {S} x {fps} frames: {S * fps * reps / (ms.value * 1e-3) / 1e6:7.2f} M frames/s")

"""Diagnostic (GPU, needs a -DPVQT_PHASE_TIMERS build: scripts/build_variant.sh pt -DPVQT_PHASE_TIMERS):
where one step of the bench workload spends its time, from %globaltimer stamps per CTA and phase."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt()
audio = synth.polyphonic_chords(60.0, 22050.0, seed=0)
hop = 368
n = synth.frames_in(audio.shape[0], v.n_fft, hop)
d_audio = pv.DeviceBuffer(v, audio.nbytes); d_audio.upload(audio)
d_out = pv.DeviceBuffer(v, n * 588 * 4)
d_flush = pv.DeviceBuffer(v, 512 << 20)
for _ in range(5):
    lib.pvqt_dev_flush_l2(v.handle, d_flush.ptr, 512 << 20)
    pv.calc_db_device(v, d_audio, 1, 0, hop, n, d_out)
pv.synchronize(v)
buf = np.zeros((2, 8192, 8), np.uint64)
lib.pvqt_debug_phase_stamps.argtypes = [C.c_void_p]
assert lib.pvqt_debug_phase_stamps(buf.ctypes.data) == 0
fft = buf[0][buf[0][:, 0] > 0]
sp = buf[1][buf[1][:, 0] > 0]
t0 = int(min(fft[:, 0].min(), sp[:, 0].min()))
f = (fft[:, :2].astype(np.int64) - t0) / 1e3
s = (sp[:, :7].astype(np.int64) - t0) / 1e3
print(f"K-fft: {len(fft)} CTAs, first start {f[:,0].min():.1f} us, last start {f[:,0].max():.1f}, last end {f[:,1].max():.1f}; "
      f"CTA duration median {np.median(f[:,1]-f[:,0]):.1f} us (p10 {np.percentile(f[:,1]-f[:,0],10):.1f}, p90 {np.percentile(f[:,1]-f[:,0],90):.1f})")
names = ["start", "prologue done", "after pdl_wait", "staged+combined", "walk done", "ls written", "end"]
print(f"K-spmm-db: {len(sp)} CTAs; per phase stamp: min / median / max over CTAs (us since the first K-fft CTA)")
for i, nm in enumerate(names):
    print(f"  {nm:18s} {s[:,i].min():7.1f} {np.median(s[:,i]):7.1f} {s[:,i].max():7.1f}")
d = np.diff(s, axis=1)
print("  phase durations (median / p90 us):", {names[i + 1]: (round(float(np.median(d[:, i])), 1), round(float(np.percentile(d[:, i], 90)), 1)) for i in range(6)})
sm = sp[:, 7].astype(int)
print("  CTAs per SM: min", np.bincount(sm, minlength=148).min(), "max", np.bincount(sm).max(), "SMs used", len(np.unique(sm)))

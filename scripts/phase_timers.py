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
if os.environ.get('PT_SDFT') is not None:
    print('sliding dft mode ->', v.set_sliding_dft(int(os.environ['PT_SDFT'])))
audio = synth.polyphonic_chords(60.0, 22050.0, seed=0)
hop = 368
n = synth.frames_in(audio.shape[0], v.n_fft, hop)
d_audio = pv.DeviceBuffer(v, audio.nbytes); d_audio.upload(audio)
d_out = pv.DeviceBuffer(v, n * 588 * 4)
d_flush = pv.DeviceBuffer(v, 512 << 20)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.pvqt_event_create(v.handle, C.byref(e0)); lib.pvqt_event_create(v.handle, C.byref(e1))
for _ in range(8):
    lib.pvqt_dev_flush_l2(v.handle, d_flush.ptr, 512 << 20)
    lib.pvqt_event_record(v.handle, e0)
    pv.calc_db_device(v, d_audio, 1, 0, hop, n, d_out)
    lib.pvqt_event_record(v.handle, e1)
    ms = C.c_float(); lib.pvqt_event_elapsed_ms(v.handle, e0, e1, C.byref(ms))
    print(f"step (CUDA events): {ms.value * 1e3:.1f} us")
pv.synchronize(v)
buf = np.zeros((2, 8192, 8), np.uint64)
lib.pvqt_debug_phase_stamps.argtypes = [C.c_void_p]
assert lib.pvqt_debug_phase_stamps(buf.ctypes.data) == 0
sd = np.zeros((1024, 2), np.uint64)
lib.pvqt_debug_sdft_stamps.argtypes = [C.c_void_p]
assert lib.pvqt_debug_sdft_stamps(sd.ctypes.data) == 0
sd = sd[sd[:, 0] > 0]
fft = buf[0][buf[0][:, 0] > 0]
sp = buf[1][buf[1][:, 0] > 0]
t0 = int(min(fft[:, 0].min(), sp[:, 0].min()))
f = (fft[:, :2].astype(np.int64) - t0) / 1e3
s = (sp[:, [0, 1, 7, 2, 3, 4, 5, 6]].astype(np.int64) - t0) / 1e3
if len(sd):
    q = (sd.astype(np.int64) - t0) / 1e3
    print(f"K-sdft: {len(sd)} CTAs, start {q[:,0].min():.1f}..{q[:,0].max():.1f} us, end {q[:,1].min():.1f}..{q[:,1].max():.1f} us (relative to the first K-fft CTA)")
print(f"K-fft: {len(fft)} CTAs, first start {f[:,0].min():.1f} us, last start {f[:,0].max():.1f}, last end {f[:,1].max():.1f}; "
      f"CTA duration median {np.median(f[:,1]-f[:,0]):.1f} us (p10 {np.percentile(f[:,1]-f[:,0],10):.1f}, p90 {np.percentile(f[:,1]-f[:,0],90):.1f})")
for g in range(4):
    if not (fft[:, 3] == g).any(): continue
    m = fft[:, 3] == g
    d = f[m, 1] - f[m, 0]
    print(f"  FFT group index {g}: {m.sum()} CTAs, start {f[m,0].min():.1f}..{f[m,0].max():.1f} us, duration median {np.median(d):.2f} (p10 {np.percentile(d,10):.2f}, p90 {np.percentile(d,90):.2f}), last end {f[m,1].max():.1f}")
for t in (2, 5, 10, 15, 20, 25, 30, 35, 40, 45, 50):
    act = (f[:, 0] <= t) & (f[:, 1] > t)
    sms = fft[act, 2].astype(int)
    print(f"  t = {t:2d} us: {act.sum()} K-fft CTAs active on {len(np.unique(sms))} SMs (max {np.bincount(sms).max() if act.any() else 0} per SM)")
names = ["start", "prologue done", "early combine", "after pdl_wait", "staged+combined", "walk done", "ls written", "end"]
print(f"K-spmm-db: {len(sp)} CTAs; per phase stamp: min / median / max over CTAs (us since the first K-fft CTA)")
for i, nm in enumerate(names):
    print(f"  {nm:18s} {s[:,i].min():7.1f} {np.median(s[:,i]):7.1f} {s[:,i].max():7.1f}")
d = np.diff(s, axis=1)
print("  phase durations (median / p90 us):", {names[i + 1]: (round(float(np.median(d[:, i])), 1), round(float(np.percentile(d[:, i], 90)), 1)) for i in range(7)})

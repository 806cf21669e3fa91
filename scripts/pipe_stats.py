"""Per-phase cycles of the K-spmm-db pipeline form (a -DPVQT_PIPE_STATS build: scripts/build_variant.sh stats -DPVQT_PIPE_STATS).
   PVQT_LIB=pitchvis_b200/lib/libpvqt_stats.so python scripts/pipe_stats.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pitchvis_b200 as pv
from pitchvis_b200 import _ffi, synth
lib = _ffi.load()
v = pv.Vqt()
# default: chords60 (one stream, 3507 frames); "python scripts/pipe_stats.py 256" = 256 streams of 10 s (one 130,816-frame launch)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if S == 1:
    audio = synth.polyphonic_chords(60.0, 22050.0, seed=0)
    n_per = v.frames_in(audio.shape[0], synth.HOP_DEFAULT)
    stride = 0
else:
    one = synth.polyphonic_chords(10.0, 22050.0, seed=0)
    n_per = v.frames_in(one.shape[0], synth.HOP_DEFAULT)
    audio = np.ascontiguousarray(np.stack([np.roll(one, 977 * s) for s in range(S)]))
    stride = audio.shape[1]
n = n_per * S
d_a = pv.DeviceBuffer(v, audio.nbytes); d_a.upload(audio)
d_o = pv.DeviceBuffer(v, n * 588 * 4)
buf = (C.c_longlong * (256 * 16))()
for _ in range(3):
    pv.calc_db_device(v, d_a, S, stride, synth.HOP_DEFAULT, n_per, d_o)
lib.pvqt_debug_pipe_stats.argtypes = [C.c_void_p, C.c_int]
lib.pvqt_debug_pipe_stats(buf, 1)
reps = 10
for _ in range(reps):
    pv.calc_db_device(v, d_a, S, stride, synth.HOP_DEFAULT, n_per, d_o)
lib.pvqt_debug_pipe_stats(buf, 1)
s = np.array(buf, dtype=np.int64).reshape(256, 16)[:148] / reps
names = ["W wait FULL", "W walk", "W wait LSEMPTY", "W ls+arrive", "H pdl_wait", "H wait EMPTY", "H stage issue", "H combine",
         "H cp.async wait", "H wait LSFULL", "H epilogue"]
print(f"cycles per launch ({(n + 7) // 8 / 148:.1f} tiles per CTA), median / max over CTAs; 1965 cycles = 1 us")
for i, nm in enumerate(names):
    print(f"{nm:18s} {np.median(s[:, i]):9.0f} {s[:, i].max():9.0f}")

# one flushed step with globaltimer stamps of K-fft (needs -DPVQT_PHASE_TIMERS too) and of the pipeline CTAs
if hasattr(lib, "pvqt_debug_phase_stamps"):
    d_flush = pv.DeviceBuffer(v, 512 << 20)
    for _ in range(3):
        lib.pvqt_dev_flush_l2(v.handle, d_flush.ptr, 512 << 20)
        pv.calc_db_device(v, d_a, S, stride, synth.HOP_DEFAULT, n_per, d_o)
    pv.synchronize(v)
    ph = np.zeros((2, 8192, 8), np.uint64)
    lib.pvqt_debug_phase_stamps.argtypes = [C.c_void_p]
    lib.pvqt_debug_phase_stamps(ph.ctypes.data)
    st = np.zeros((256, 8), np.uint64)
    lib.pvqt_debug_pipe_stamps.argtypes = [C.c_void_p]
    lib.pvqt_debug_pipe_stamps(st.ctypes.data)
    fft = ph[0][ph[0][:, 0] > 0]
    st = st[:148]
    t0 = int(fft[:, 0].min())
    f = (fft[:, :2].astype(np.int64) - t0) / 1e3
    print(f"K-fft: {len(fft)} CTAs, starts {f[:,0].min():.1f}..{f[:,0].max():.1f} us, ends {f[:,1].min():.1f}..{f[:,1].max():.1f} us")
    for t in (30, 34, 38, 40, 42, 44, 46, 48):
        act = (f[:, 0] <= t) & (f[:, 1] > t)
        print(f"  t = {t} us: {act.sum()} K-fft CTAs active on {len(np.unique(fft[act, 2].astype(int)))} SMs")
    q = (st[:, :5].astype(np.int64) - t0) / 1e3
    for i, nm in enumerate(["CTA start", "after pdl_wait", "first tile ready", "last walk done", "CTA end"]):
        print(f"  pipe {nm:18s} min {q[:,i].min():6.1f}  median {np.median(q[:,i]):6.1f}  max {q[:,i].max():6.1f} us")

"""Diagnostic (GPU): where does the GPU-vs-exact-oracle error come from?  FFT stage vs SpMM stage."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
import pitchvis_b200 as pv
from pitchvis_b200 import synth

v = pv.Vqt(); o = orc.OracleVqt()
HOP = 368; n_frames = 64
chords = synth.polyphonic_chords(8.0, 22050.0, seed=0)
audio = chords[:v.n_fft + (n_frames - 1) * HOP]
d_audio = pv.DeviceBuffer(v, audio.nbytes); d_audio.upload(audio)
stride = v.spec_stride
d_spec = pv.DeviceBuffer(v, n_frames * stride * 8)
pv.fft_device(v, d_audio, 1, 0, HOP, n_frames, d_spec)
spec = d_spec.download((n_frames, stride), np.complex64)
d_out = pv.DeviceBuffer(v, n_frames * 588 * 4); d_pow = pv.DeviceBuffer(v, n_frames * 588 * 4)
pv.calc_db_device(v, d_audio, 1, 0, HOP, n_frames, d_out, d_pow)
p_gpu = d_pow.download((n_frames, 588)).astype(np.float64)

k = v.kernel()
p_exact = np.zeros((n_frames, 588)); p_gpuspec64 = np.zeros((n_frames, 588))
row0 = 0
for g, wg in enumerate(k.window_groups):
    first, n_cols, off = v.group_columns(g)
    wb, we = wg.window
    import scipy.sparse as sp
    K = sp.csr_matrix((wg.filter_bank.data.astype(np.complex128), wg.filter_bank.indices, wg.filter_bank.indptr), shape=(wg.filter_bank.rows, wg.filter_bank.cols))
    Kn = None
    if wg.negative_filter_bank is not None:
        nb = wg.negative_filter_bank
        Kn = sp.csr_matrix((nb.data.astype(np.complex128), nb.indices, nb.indptr), shape=(nb.rows, nb.cols))
    ferr = 0; fmax = 0
    for t in range(n_frames):
        X = np.fft.rfft(audio[t * HOP + wb:t * HOP + we].astype(np.float64))
        Xg = np.zeros_like(X); Xg[first:first + n_cols] = spec[t, off:off + n_cols]
        ferr = max(ferr, np.abs(Xg[first:first + n_cols] - X[first:first + n_cols]).max()); fmax = max(fmax, np.abs(X).max())
        y = K @ X; yg = K @ Xg
        if Kn is not None:
            y = y + np.conj(Kn @ X); yg = yg + np.conj(Kn @ Xg)
        p_exact[t, row0:row0 + K.shape[0]] = np.abs(y) ** 2
        p_gpuspec64[t, row0:row0 + K.shape[0]] = np.abs(yg) ** 2
    print(f"group {g}: N={we-wb} max |X_gpu - X_exact| / max|X| = {ferr / fmax:.3e}")
    row0 += K.shape[0]

rel = 10 * np.log10(p_exact / p_exact.max(axis=1, keepdims=True))
def report(name, a, b):
    err = np.abs(10 * np.log10(a / b))
    for lo, hi in [(-40, 0), (-60, -40), (-80, -60)]:
        m = (rel >= lo) & (rel < hi)
        print(f"  {name:34s} [{lo},{hi}) dB: max {err[m].max():.2e} dB  rms {np.sqrt((err[m]**2).mean()):.2e}")
report("GPU total vs exact", p_gpu, p_exact)
report("FFT error only (f64 SpMM on GPU X)", p_gpuspec64, p_exact)
report("SpMM error only (GPU vs f64 SpMM)", p_gpu, p_gpuspec64)

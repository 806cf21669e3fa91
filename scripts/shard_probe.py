import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import pitchvis_b200 as pv
from pitchvis_b200 import synth, _ffi
p = pv.VqtParameters.hires(); hop = synth.HOP_HIRES
v = pv.Vqt(p); v.set_sliding_dft(False)
n_frames = 41
audio = synth.polyphonic_chords(3.0, 44100.0, seed=21)[:p.n_fft + (n_frames - 1) * hop]
batch = v.calculate_vqt_batch_in_db(audio, hop)
for parts in (2, 3, 4, 5):
    bad = 0
    for i in range(parts):
        base, rem = divmod(n_frames, parts)
        f0 = i * base + min(i, rem); f1 = f0 + base + (1 if i < rem else 0)
        s0, s1 = f0 * hop, (f1 - 1) * hop + p.n_fft
        got = v.calculate_vqt_batch_in_db(audio[s0:s1], hop)
        d = (got != batch[f0:f1])
        bad += int(d.sum())
        if d.any():
            fr, bn = np.nonzero(d)
            print(parts, i, f0, f1, 'mismatch', d.sum(), 'frames', sorted(set(fr.tolist()))[:10], 'bins', bn.min(), bn.max())
    print('parts', parts, 'bad', bad)
# repeatability of the same call
a = v.calculate_vqt_batch_in_db(audio[:p.n_fft + 13 * hop], hop)
b = v.calculate_vqt_batch_in_db(audio[:p.n_fft + 13 * hop], hop)
print('repeat equal', np.array_equal(a, b), 'vs batch', np.array_equal(a, batch[:14]))
for nf in range(1, 20):
    g = v.calculate_vqt_batch_in_db(audio[:p.n_fft + (nf - 1) * hop], hop)
    print(nf, np.array_equal(g, batch[:nf]), end='; ')
print()

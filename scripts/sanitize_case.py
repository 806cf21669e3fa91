"""Small end-to-end case for compute-sanitizer (memcheck): every kernel of the library once, small sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pitchvis_b200 as pv
from pitchvis_b200 import synth
v = pv.Vqt(pv.VqtParameters.default())
audio = synth.polyphonic_chords(2.2, 22050.0, seed=3)
for mode in (3, 1, 0, 2):                    # K-spmm-db pipeline form, one-CTA form, unfused pair, cluster form
    v.set_fused_epilogue(mode)
    for sd in (2, 1, 0):                     # K-sdft on tensor cores, on the FP32 pipe, off
        v.set_sliding_dft(sd)
        out = v.calculate_vqt_batch_in_db(audio, 368)
        assert np.all(np.isfinite(out)), (mode, sd)
v.set_fused_epilogue(3); v.set_sliding_dft(2)
one = v.calculate_vqt_instant_in_db(audio[:v.n_fft])
streams = np.stack([audio[:v.n_fft + 30 * 333], audio[100:100 + v.n_fft + 30 * 333]])
v.calculate_vqt_streams_in_db(streams, 333)
st = pv.AnalysisState(pv.VqtRange()); st.preprocess_batch(out[:8], 16_689_342)
st.calculate_and_preprocess(v, audio, 368, 16_689_342); st.close()   # VQT + AnalysisState in one call (spectra stay in HBM)
agc = pv.MonoAgc(0.07, 0.001, n_streams=2); agc.process_chunks(streams, 441); agc.close()
pv.chroma(out[:5])
v.close()
h = pv.Vqt(pv.VqtParameters.hires())
ah = synth.polyphonic_chords(2.0, 44100.0, seed=4)
h.calculate_vqt_batch_in_db(ah, 735); h.close()
print("sanitize case ok")

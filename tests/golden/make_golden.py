"""Generate the golden fixtures under tests/golden/ (committed; the GPU box has no /root/reference).

    python tests/golden/make_golden.py

Provenance.  The reference (heinzelotto/pitchvis, Rust) ships NO golden arrays and cannot be built in this
image (no rustc/cargo), so these vectors come from the CPU oracle (oracle/vqt_oracle.c, analysis_oracle.c:
a line-by-line C restatement, pinned to every known answer and property test the reference holds --
tests/test_oracle_known_answers.py, tests/test_analysis_oracle.py) in its exact mode (f64 FFT, mode 0).
They freeze that oracle: a change of the oracle or of the kernel builder that moves any value shows up
here, on CPU, before the GPU parity tests consume the same files.

Files
  vqt_default_sines.npz   config 1: test_create_sines(440 Hz + 4 harmonics) (util.rs:62-79) -> 588 dB values
  vqt_default_chords.npz  8 frames of the seeded chord generator (config 2's signal) -> [8][588] dB + power
  vqt_hires_chords.npz    config 4 parameters, 3 frames -> [3][1344] dB
  kernel_summary.json     per parameter set: windows, rows, nnz, conj-part nnz, delay, CRC of the CSR arrays
  analysis_chords.npz     config 5: AnalysisState over 48 frames of the golden dB input: peak sets, counts,
                          continuous peaks, smoothed scene calmness
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import orc  # noqa: E402
from pitchvis_b200 import synth  # noqa: E402

HOP, HOP_HIRES = synth.HOP_DEFAULT, synth.HOP_HIRES
FRAME_TIME_NS = 16_689_342  # 368 / 22050 s


def kernel_summary(v: orc.OracleVqt):
    groups = []
    for g in range(v.num_groups):
        (wb, we), K, Kn = v.group(g)
        groups.append({
            "window": [int(wb), int(we)], "rows": int(K.rows), "cols": int(K.cols), "nnz": int(K.nnz),
            "neg_nnz": int(Kn.nnz),
            "indices_crc32": zlib.crc32(np.ascontiguousarray(K.indices, np.int32).tobytes()),
            "indptr_crc32": zlib.crc32(np.ascontiguousarray(K.indptr, np.int32).tobytes()),
            # values are f32 results of libm calls (cosf/sinf/hypotf): keep a tolerance-friendly digest
            "abs_sum": float(np.abs(K.data.astype(np.complex128)).sum()),
            "neg_abs_sum": float(np.abs(Kn.data.astype(np.complex128)).sum()) if Kn.nnz else 0.0,
        })
    return {"n_buckets": int(v.n_buckets), "delay_ms": float(v.delay * 1e3), "groups": groups}


def main():
    d = orc.OracleVqt()
    h = orc.OracleVqt(orc.hires_params())
    t = orc.OracleVqt(orc.make_params(n_fft=16384, min_freq=110.0, octaves=6, buckets_per_octave=48, quality=1.0,
                                      gamma=20.0))
    with open(os.path.join(HERE, "kernel_summary.json"), "w") as fh:
        json.dump({"default": kernel_summary(d), "hires": kernel_summary(h), "six_octaves_48": kernel_summary(t)},
                  fh, indent=1)

    x = orc.test_create_sines(d.params, [440, 880, 1320, 1760, 2200])
    db, power = d.calculate_vqt_instant_in_db(x, 0, True)
    np.savez_compressed(os.path.join(HERE, "vqt_default_sines.npz"), x=x, db=db, power=power)

    chords = synth.polyphonic_chords(3.0, 22050.0, seed=0)
    n = 8
    audio = chords[:d.n_fft + (n - 1) * HOP * 3]
    hop = HOP * 3  # three viewer hops apart: covers a chord change
    dbs, pws = [], []
    for i in range(n):
        a, b = d.calculate_vqt_instant_in_db(audio[i * hop:i * hop + d.n_fft], 0, True)
        dbs.append(a)
        pws.append(b)
    np.savez_compressed(os.path.join(HERE, "vqt_default_chords.npz"), audio=audio, hop=np.int64(hop),
                        db=np.stack(dbs), power=np.stack(pws))

    ch = synth.polyphonic_chords(2.0, 44100.0, seed=4)
    ah = ch[:h.n_fft + 2 * HOP_HIRES]
    np.savez_compressed(os.path.join(HERE, "vqt_hires_chords.npz"), audio=ah, hop=np.int64(HOP_HIRES),
                        db=np.stack([h.calculate_vqt_instant_in_db(ah[i * HOP_HIRES:i * HOP_HIRES + h.n_fft], 0)
                                     for i in range(3)]))

    # analysis epilogue over 48 consecutive frames (hop 368) of the chord signal
    T = 48
    a5 = chords[:d.n_fft + (T - 1) * HOP]
    db5 = d.calculate_batch_db(a5, HOP, mode=0)
    st = orc.OracleAnalysisState()
    counts, idx, cont, calm = [], np.full((T, 64), -1, np.int32), np.zeros((T, 64, 2), np.float32), []
    for i in range(T):
        st.preprocess(db5[i], FRAME_TIME_NS)
        p = np.sort(st.peaks)
        counts.append(len(p))
        idx[i, :len(p)] = p
        pc = st.peaks_continuous
        cont[i, :len(pc)] = pc
        calm.append(st.smoothed_scene_calmness)
    np.savez_compressed(os.path.join(HERE, "analysis_chords.npz"), db=db5, frame_time_ns=np.int64(FRAME_TIME_NS),
                        peak_count=np.array(counts, np.int32), peak_indices=idx, peaks_continuous=cont,
                        scene_calmness=np.array(calm, np.float32))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

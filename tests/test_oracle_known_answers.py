"""Pins the CPU oracle (oracle/vqt_oracle.c) against every known answer and property test the
reference holds for the VQT path (SURVEY.md section 4 / BASELINE.md section 2).  CPU only."""
import numpy as np
import pytest

import orc


def test_n_buckets_default(oracle_default):
    # vqt.rs:236 doc-test: n_buckets() == 7 * 84
    assert oracle_default.n_buckets == 7 * 84 == 588


def test_window_groups_and_fft_sizes(oracle_default):
    # vqt.rs:133-134: 4 real FFTs of 8192, 4096, 2048, 1024 points
    v = oracle_default
    assert v.num_groups == 4
    sizes = [v.group(g)[0][1] - v.group(g)[0][0] for g in range(4)]
    assert sizes == [8192, 4096, 2048, 1024]
    windows = [v.group(g)[0] for g in range(4)]
    assert windows == [(24576, 32768), (28551, 32647), (29575, 31623), (30087, 31111)]
    assert [v.group(g)[1].rows for g in range(4)] == [122, 252, 168, 46]


def test_kernel_nnz(oracle_default):
    # VQT_REVIEW.md:369: ~18k non-zeros, 379 of them in the conjugate-part matrices
    v = oracle_default
    pos = [v.group(g)[1].nnz for g in range(4)]
    neg = [v.group(g)[2].nnz for g in range(4)]
    assert sum(neg) == 379
    assert neg == [369, 10, 0, 0]
    assert pos == [3742, 6644, 5743, 1578]
    assert 17000 < sum(pos) + sum(neg) < 19000


def test_delay(oracle_default):
    # vqt.rs:1078-1085 test_vqt_delay: delay.as_millis() < 100; VQT_REVIEW.md:363: 98 ms
    assert int(oracle_default.delay * 1000) == 98


def test_csr_rows_sorted(oracle_default):
    v = oracle_default
    for g in range(v.num_groups):
        for m in v.group(g)[1:]:
            for r in range(m.rows):
                idx = m.indices[m.indptr[r]:m.indptr[r + 1]]
                assert np.all(np.diff(idx) > 0)
            if m.nnz:
                assert m.indices.max() < m.cols


def test_fft_convention_unnormalised():
    # vqt.rs:1087-1128: forward FFT is unnormalised, half spectrum = lower half of the complex FFT.
    # A kernel of one unit coefficient at column c turns the oracle into "read |X[c]|^2".
    p = orc.make_params()
    v = orc.OracleVqt(p)
    (wb, we), K, _ = v.group(3)
    n = we - wb
    rng = np.random.default_rng(1)
    x = np.zeros(v.n_fft, np.float32)
    x[wb:we] = rng.standard_normal(n).astype(np.float32)
    ref = np.fft.rfft(x[wb:we].astype(np.float64))
    cols = [0, 1, 7, n // 4, n // 2]
    rows = K.rows
    for mode in (0, 1):
        indptr = np.zeros(rows + 1, np.int32)
        indptr[1:len(cols) + 1] = np.arange(1, len(cols) + 1)
        indptr[len(cols) + 1:] = len(cols)
        v2 = orc.OracleVqt(p)
        for g in range(3):
            r = v2.group(g)[1].rows
            v2.set_group(g, False, r, v2.group(g)[1].cols, np.zeros(r + 1, np.int32), np.zeros(0, np.int32),
                         np.zeros(0, np.complex64))
            v2.set_group(g, True, r, v2.group(g)[1].cols, np.zeros(r + 1, np.int32), np.zeros(0, np.int32),
                         np.zeros(0, np.complex64))
        v2.set_group(3, False, rows, K.cols, indptr, np.array(cols, np.int32), np.ones(len(cols), np.complex64))
        _, pw = v2.calculate_vqt_instant_in_db(x, mode, return_power=True)
        off = 122 + 252 + 168
        got = pw[off:off + len(cols)]
        np.testing.assert_allclose(got, np.abs(ref[cols]) ** 2, rtol=2e-5)


def test_single_tone_440(oracle_default):
    # amplitude 1/12 at 440 Hz peaks at bin 252 with
    # 10*log10((sqrt(sr)/24)^2 / 0.09) = 26.29 dB
    v = oracle_default
    x = orc.test_create_sines(v.params, [440.0])
    db = v.calculate_vqt_instant_in_db(x)
    assert int(db.argmax()) == 252
    assert abs(float(db.max()) - 10 * np.log10((np.sqrt(22050.0) / 24) ** 2 / 0.09)) < 0.01


def test_zero_input_gives_zero_output(oracle_default):
    v = oracle_default
    db = v.calculate_vqt_instant_in_db(np.zeros(v.n_fft, np.float32))
    assert np.all(db == 0.0)


def test_wrong_length_panics(oracle_default):
    # vqt.rs:867-871
    with pytest.raises(ValueError):
        oracle_default.calculate_vqt_instant_in_db(np.zeros(100, np.float32))


def test_errors():
    # vqt.rs:518-528 / 567-573
    with pytest.raises(orc.OracleError) as e:
        orc.OracleVqt(orc.make_params(octaves=8))
    assert e.value.code == orc.ORC_ABOVE_NYQUIST and e.value.b == 11025.0
    with pytest.raises(orc.OracleError) as e:
        orc.OracleVqt(orc.make_params(n_fft=2048))
    assert e.value.code == orc.ORC_WINDOW_EXCEEDS_NFFT and e.value.n == 2048


def test_vqt_high_frequencies(oracle_default):
    # lib.rs:50-72: per-octave on-grid tones respond within 6 dB of each other
    v = oracle_default
    inf, sup = np.inf, 0.0
    for i in range(7):
        for j in range(30):
            freq = np.float32(55.0) * np.float32(2.0) ** np.float32(i + j / (12.0 * 30))
            db = v.calculate_vqt_instant_in_db(orc.test_create_sines(v.params, [freq]), mode=1)
            inf, sup = min(inf, db.max()), max(sup, db.max())
    assert inf > sup - 6.0


def test_vqt_group_boundary_continuity(oracle_default):
    # vqt.rs:1032-1076: +-quarter-semitone sweep across each rate-group boundary, spread < 3 dB
    v = oracle_default
    fps = orc.filter_bank_params(v.params)
    boundaries = [fps[i + 1].freq for i in range(len(fps) - 1)
                  if fps[i].sr_downscaling_factor != fps[i + 1].sr_downscaling_factor]
    assert boundaries
    for b in boundaries:
        resp = []
        for i in range(-20, 21):
            freq = np.float32(b) * np.float32(2.0) ** np.float32(i / (20 * 4.0 * 12.0))
            resp.append(v.calculate_vqt_instant_in_db(orc.test_create_sines(v.params, [freq]), mode=1).max())
        assert max(resp) - min(resp) < 3.0, f"boundary {b}"


def test_vqt_bandwidths_subsampled(oracle_default):
    # vqt.rs:996-1027 sweeps 11,740 tones; here every 7th (1,677 tones) to keep the CPU suite short
    v = oracle_default
    max_single, min_sum = 0.0, np.inf
    for i in range(10, 588 * 20 - 10, 7):
        freq = np.float32(55.0) * np.float32(2.0) ** np.float32(i / (84.0 * 20.0))
        db = v.calculate_vqt_instant_in_db(orc.test_create_sines(v.params, [freq]), mode=1)
        max_single, min_sum = max(max_single, db.max()), min(min_sum, db.sum())
    assert max_single - min_sum < 3.0


def test_f32_mode_tracks_exact_mode(oracle_default):
    # the reference-faithful f32 path and the f64 "exact" path agree far inside the 1e-3 dB budget
    # on strong bins; this bounds the reference's own rounding noise
    v = oracle_default
    x = orc.test_create_sines(v.params, [440, 880, 1320, 1760, 2200])
    a = v.calculate_vqt_instant_in_db(x, 0)
    b = v.calculate_vqt_instant_in_db(x, 1)
    assert np.abs(a - b).max() < 1e-4


def test_hires_structure():
    # hi-res -> 1344 bins, 5 FFTs 16384..1024, 41,870 + 742 nnz
    v = orc.OracleVqt(orc.hires_params())
    assert v.n_buckets == 1344 and v.num_groups == 5
    assert [v.group(g)[0][1] - v.group(g)[0][0] for g in range(5)] == [16384, 8192, 4096, 2048, 1024]
    assert sum(v.group(g)[1].nnz for g in range(5)) == 41870
    assert sum(v.group(g)[2].nnz for g in range(5)) == 742


def test_batch_equals_instant(oracle_default):
    v = oracle_default
    rng = np.random.default_rng(3)
    audio = (0.05 * rng.standard_normal(v.n_fft + 5 * 368)).astype(np.float32)
    out = v.calculate_batch_db(audio, 368, mode=1, n_threads=2)
    assert out.shape == (6, 588)
    for t in range(6):
        np.testing.assert_array_equal(out[t], v.calculate_vqt_instant_in_db(audio[t * 368:t * 368 + v.n_fft], 1))

/*
 * vqt_oracle.h -- CPU restatement of pitchvis_analysis's VQT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pitchvis_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, as the checker or the timed CPU baseline.
 *
 * Every function cites the reference lines it follows (paths relative to the
 * upstream repo heinzelotto/pitchvis).
 *
 * Pinning status: the Rust reference cannot be compiled in this image (no
 * rustc/cargo) and ships no golden arrays.  This restatement is pinned against
 * every known answer and property test the reference holds for the path
 * (tests/test_oracle_known_answers.py): 588 buckets, 4 window groups with FFT
 * sizes 8192/4096/2048/1024, 379 conjugate-part non-zeros of ~18k, delay 98 ms,
 * unnormalised forward FFT convention, the 3 dB / 6 dB flatness properties, and
 * the "two tones -> two peaks" test.  Array-valued parity with rustfft/sprs
 * output is therefore "known-answer pinned", not "golden-vector pinned"; the
 * peak *index* semantics of the third-party find_peaks 0.1.5 crate are
 * unpinned (see analysis_oracle.c).
 */
#ifndef VQT_ORACLE_H
#define VQT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* vqt.rs:239-262 (VqtRange) + vqt.rs:279-331 (VqtParameters), flattened. */
typedef struct orc_params {
    float    sr;
    uint64_t n_fft;
    float    min_freq;
    uint32_t octaves;
    uint32_t buckets_per_octave;
    float    sparsity_quantile;
    float    quality;
    float    gamma;
} orc_params;

/* vqt.rs:352-366 (VqtError) */
enum {
    ORC_OK = 0,
    ORC_ABOVE_NYQUIST = 1,        /* a = highest_frequency, b = nyquist_frequency */
    ORC_WINDOW_EXCEEDS_NFFT = 2,  /* a = window_length,     n = n_fft             */
    ORC_ASSERT = 3,               /* a panic (assert!/expect) inside Vqt::new      */
    ORC_BAD_LENGTH = 4            /* vqt.rs:867-871 assert_eq!(x.len(), n_fft)     */
};

typedef struct orc_error {
    int      code;
    float    a, b;
    uint64_t n;
} orc_error;

/* vqt.rs:370-384 (FilterParams) */
typedef struct orc_filter_params {
    float    freq;
    float    window_length;
    uint64_t sr_downscaling_factor;
    uint64_t minimum_needed_window_size;
} orc_filter_params;

/* One CSR matrix of Complex32 (sprs::CsMat<Complex32>, vqt.rs:396,403). */
typedef struct orc_csr {
    int32_t  rows, cols;
    int64_t  nnz;
    int32_t *indptr;   /* rows + 1           */
    int32_t *indices;  /* nnz, ascending per row */
    float   *data;     /* 2 * nnz, interleaved (re, im) */
} orc_csr;

/* vqt.rs:388-410 (WindowGroup) */
typedef struct orc_group {
    uint64_t window_begin, window_end;
    orc_csr  filter_bank;
    orc_csr  negative_filter_bank; /* nnz == 0 <=> None (vqt.rs:751) */
} orc_group;

typedef struct orc_vqt orc_vqt;

void orc_default_params(orc_params *p);                       /* vqt.rs:180-214, 333-348 */
int  orc_filter_bank_params(const orc_params *p, orc_filter_params *out /* n_buckets */,
                            orc_error *err);                  /* vqt.rs:517-587 */
int  orc_vqt_new(const orc_params *p, orc_vqt **out, orc_error *err); /* vqt.rs:465-505 */
void orc_vqt_free(orc_vqt *v);
size_t orc_n_buckets(const orc_vqt *v);                       /* vqt.rs:259-261 */
double orc_delay_seconds(const orc_vqt *v);                   /* vqt.rs:756, Duration::from_secs_f32 */
size_t orc_num_groups(const orc_vqt *v);
const orc_group *orc_group_at(const orc_vqt *v, size_t g);

/* Replace the kernel of `v` by caller-supplied CSR arrays (deep copy).  Lets the
 * parity tests run the oracle's runtime on the product's host-built kernel
 * (what Vqt::kernel() exposes, vqt.rs:511-513). */
int  orc_vqt_set_group(orc_vqt *v, size_t g, int neg, int32_t rows, int32_t cols, int64_t nnz,
                       const int32_t *indptr, const int32_t *indices, const float *data);

/* vqt.rs:866-916 + 922-954.  mode 0: "exact" (f64 FFT, f64 accumulation of the f32
 * kernel, power rounded to f32, then power_to_db in f32 exactly as the reference);
 * mode 1: "faithful f32" (f32 FFT, f32 sequential CSR sums in the reference's
 * order).  out_db: n_buckets.  out_power (optional): |z|^2 before dB, n_buckets. */
int  orc_calc_instant_db(orc_vqt *v, const float *x, size_t n, int mode, float *out_db,
                         float *out_power);

/* Sliding batched variant, frame t = audio[t*hop .. t*hop+n_fft) (template:
 * pitchvis_train/src/train.rs:276-341).  n_threads <= 0: all cores (OpenMP), one
 * scratch per thread as train.rs:146-154 does with one Vqt per worker. */
int  orc_calc_batch_db(orc_vqt *v, const float *audio, size_t n_samples, size_t hop,
                       size_t n_frames, int mode, int n_threads, float *out_db);

void orc_power_to_db(const float *power, size_t n, float *out_db); /* vqt.rs:922-954 */

/* util.rs:62-79 (test_create_sines): adds to a zeroed buffer of n_fft samples. */
void orc_test_create_sines(const orc_params *p, const float *freqs, size_t n_freqs, float t_diff,
                           float *wave /* n_fft */);

int  orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

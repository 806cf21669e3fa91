/*
 * vqt_oracle.c -- CPU restatement of pitchvis_analysis::vqt (plain C).
 *
 * TEST INFRASTRUCTURE ONLY -- see vqt_oracle.h for who may use this and for the
 * pinning status.  Not product code; the product has its own kernel builder and
 * CUDA runtime under pitchvis_b200/csrc/ and never calls into this file.
 *
 * All f32 arithmetic below mirrors the op order of the Rust source; compile with
 * -ffp-contract=off so gcc does not fuse a*b+c where Rust would not.
 *
 * Third-party arithmetic that is not in the reference tree (Cargo.lock pins):
 *   rustfft 6.4.1 / realfft 3.5.0  -> any unnormalised forward DFT
 *                                     X[k] = sum x[n] e^{-2 pi i k n / N}
 *                                     (convention pinned by vqt.rs:1087-1128);
 *                                     restated here in f64 (radix-2) and f32 (Stockham).
 *   sprs 0.11.4                    -> TriMat::to_csr (column-sorted rows) and
 *                                     prod::mul_acc_mat_vec_csr (row-wise sequential sum).
 *   apodize 1.0.0                  -> hanning_iter: symmetric Hann in f64,
 *                                     0.5 - 0.5 cos(2 pi n / (len - 1)).
 *   num-complex 0.4.6              -> norm() = hypot, norm_sqr, exp() = from_polar.
 */
#include "vqt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI_F 3.14159274101257324219f /* std::f32::consts::PI */
#define ORC_PI_D 3.14159265358979323846

typedef struct { float re, im; } c32;
typedef struct { double re, im; } c64;

struct orc_vqt {
    orc_params params;
    size_t     n_buckets;
    size_t     n_groups;
    orc_group *groups;
    double     delay_s;
};

/* ------------------------------------------------------------------------- */
/* small helpers                                                             */
/* ------------------------------------------------------------------------- */

static void csr_free(orc_csr *m)
{
    free(m->indptr);
    free(m->indices);
    free(m->data);
    memset(m, 0, sizeof(*m));
}

/* Rust `x as usize` / `x as u32` for f32: truncate toward zero, saturate, NaN -> 0. */
static uint64_t f32_as_u64(float x)
{
    if (!(x > 0.0f)) return 0;
    if (x >= 18446744073709551615.0f) return UINT64_MAX;
    return (uint64_t)x;
}

/* ------------------------------------------------------------------------- */
/* f64 FFT (radix-2, in place) -- used for kernel construction and mode 0     */
/* ------------------------------------------------------------------------- */

static void fft_f64(c64 *a, size_t n)
{
    /* n must be a power of two; otherwise fall back to the O(n^2) DFT below. */
    size_t j = 0;
    for (size_t i = 1; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { c64 t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len >> 1;
        for (size_t k = 0; k < half; ++k) {
            double ang = -2.0 * ORC_PI_D * (double)k / (double)len;
            double wr = cos(ang), wi = sin(ang);
            for (size_t i = k; i < n; i += len) {
                c64 u = a[i], v = a[i + half];
                double tr = v.re * wr - v.im * wi, ti = v.re * wi + v.im * wr;
                a[i].re = u.re + tr; a[i].im = u.im + ti;
                a[i + half].re = u.re - tr; a[i + half].im = u.im - ti;
            }
        }
    }
}

static void dft_f64_any(c64 *a, size_t n)
{
    if (n && (n & (n - 1)) == 0) { fft_f64(a, n); return; }
    c64 *o = (c64 *)calloc(n, sizeof(c64));
    for (size_t k = 0; k < n; ++k) {
        double sr = 0, si = 0;
        for (size_t m = 0; m < n; ++m) {
            double ang = -2.0 * ORC_PI_D * (double)((k * m) % n) / (double)n;
            double wr = cos(ang), wi = sin(ang);
            sr += a[m].re * wr - a[m].im * wi;
            si += a[m].re * wi + a[m].im * wr;
        }
        o[k].re = sr; o[k].im = si;
    }
    memcpy(a, o, n * sizeof(c64));
    free(o);
}

/* ------------------------------------------------------------------------- */
/* f32 FFT (Stockham autosort, radix 4 + 2) -- mode 1 and the CPU baseline    */
/* ------------------------------------------------------------------------- */

typedef struct {
    size_t n;       /* complex length (real length / 2) */
    c32   *tw;      /* exp(-2 pi i k / n), k < n          */
    c32   *rtw;     /* exp(-2 pi i k / (2n)), k <= n/2: real-FFT split twiddles */
} f32_plan;

static void f32_plan_init(f32_plan *pl, size_t n_complex)
{
    pl->n = n_complex;
    pl->tw = (c32 *)malloc(sizeof(c32) * n_complex);
    pl->rtw = (c32 *)malloc(sizeof(c32) * (n_complex / 2 + 1));
    for (size_t k = 0; k < n_complex; ++k) {
        double a = -2.0 * ORC_PI_D * (double)k / (double)n_complex;
        pl->tw[k].re = (float)cos(a); pl->tw[k].im = (float)sin(a);
    }
    for (size_t k = 0; k <= n_complex / 2; ++k) {
        double a = -2.0 * ORC_PI_D * (double)k / (double)(2 * n_complex);
        pl->rtw[k].re = (float)cos(a); pl->rtw[k].im = (float)sin(a);
    }
}

static void f32_plan_free(f32_plan *pl) { free(pl->tw); free(pl->rtw); }

/* x -> result returned in whichever buffer holds it (pointer returned). */
static c32 *fft_f32_stockham(const f32_plan *pl, c32 *x, c32 *y)
{
    size_t N = pl->n, n = N, s = 1;
    const c32 *tw = pl->tw;
    while (n >= 4) {
        size_t n1 = n / 4, tstep = N / n;
        for (size_t p = 0; p < n1; ++p) {
            c32 w1 = tw[p * tstep], w2 = tw[2 * p * tstep], w3 = tw[3 * p * tstep];
            const c32 *xa = x + s * p, *xb = x + s * (p + n1), *xc = x + s * (p + 2 * n1),
                      *xd = x + s * (p + 3 * n1);
            c32 *y0 = y + s * (4 * p), *y1 = y0 + s, *y2 = y1 + s, *y3 = y2 + s;
            for (size_t q = 0; q < s; ++q) {
                float ar = xa[q].re, ai = xa[q].im, br = xb[q].re, bi = xb[q].im;
                float cr = xc[q].re, ci = xc[q].im, dr = xd[q].re, di = xd[q].im;
                float apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci;
                float bpdr = br + dr, bpdi = bi + di;
                /* j*(b-d) with the forward sign convention: -i*(b-d) */
                float jr = (bi - di), ji = -(br - dr);
                float t1r = amcr + jr, t1i = amci + ji;
                float t2r = apcr - bpdr, t2i = apci - bpdi;
                float t3r = amcr - jr, t3i = amci - ji;
                y0[q].re = apcr + bpdr; y0[q].im = apci + bpdi;
                y1[q].re = t1r * w1.re - t1i * w1.im; y1[q].im = t1r * w1.im + t1i * w1.re;
                y2[q].re = t2r * w2.re - t2i * w2.im; y2[q].im = t2r * w2.im + t2i * w2.re;
                y3[q].re = t3r * w3.re - t3i * w3.im; y3[q].im = t3r * w3.im + t3i * w3.re;
            }
        }
        n = n1; s *= 4;
        c32 *t = x; x = y; y = t;
    }
    if (n == 2) {
        const c32 *xa = x, *xb = x + s;
        for (size_t q = 0; q < s; ++q) {
            c32 a = xa[q], b = xb[q];
            y[q].re = a.re + b.re; y[q].im = a.im + b.im;
            y[q + s].re = a.re - b.re; y[q + s].im = a.im - b.im;
        }
        c32 *t = x; x = y; y = t;
    }
    return x;
}

/* Real forward FFT of length 2*pl->n: in (2n floats) -> spec (n + 1 complex).
 * work: 2n complex of scratch. */
static void rfft_f32(const f32_plan *pl, const float *in, c32 *spec, c32 *work)
{
    size_t n = pl->n;
    c32 *a = work, *b = work + n;
    memcpy(a, in, sizeof(float) * 2 * n); /* z[m] = x[2m] + i x[2m+1] */
    c32 *Z = fft_f32_stockham(pl, a, b);
    const c32 *rt = pl->rtw;
    spec[0].re = Z[0].re + Z[0].im; spec[0].im = 0.0f;
    spec[n].re = Z[0].re - Z[0].im; spec[n].im = 0.0f;
    for (size_t k = 1; k <= n / 2; ++k) {
        c32 zk = Z[k], zn = Z[n - k];
        float er = 0.5f * (zk.re + zn.re), ei = 0.5f * (zk.im - zn.im); /* even part */
        float dr = 0.5f * (zk.re - zn.re), di = 0.5f * (zk.im + zn.im);
        /* odd part O = -i * (dr + i di) = (di, -dr); X[k] = E + W^k O */
        float or_ = di, oi = -dr;
        c32 w = rt[k];
        float tr = or_ * w.re - oi * w.im, ti = or_ * w.im + oi * w.re;
        spec[k].re = er + tr; spec[k].im = ei + ti;
        /* X[n-k] = conj(E) - conj(W^k O) ... using W^{n-k} = -conj(W^k) */
        spec[n - k].re = er - tr; spec[n - k].im = -(ei - ti);
    }
}

/* ------------------------------------------------------------------------- */
/* parameters                                                                */
/* ------------------------------------------------------------------------- */

void orc_default_params(orc_params *p)
{
    /* vqt.rs:180-214 and Default impl vqt.rs:333-348 */
    const float q = 1.6f / 1.0f;          /* DEFAULT_Q = 1.6 / UPSCALE_FACTOR   */
    p->sr = 22050.0f;                     /* DEFAULT_SR                          */
    p->n_fft = 2 * 16384;                 /* DEFAULT_N_FFT                       */
    p->min_freq = 55.0f;                  /* DEFAULT_MIN_FREQ                    */
    p->octaves = 7;                       /* DEFAULT_OCTAVES                     */
    p->buckets_per_octave = 12 * 7 * 1;   /* 12 * DEFAULT_BUCKETS_PER_SEMITONE   */
    p->sparsity_quantile = 0.999f;        /* DEFAULT_SPARSITY_QUANTILE           */
    p->quality = q;
    p->gamma = 4.8f * q;                  /* DEFAULT_GAMMA = 4.8 * DEFAULT_Q     */
}

static size_t params_n_buckets(const orc_params *p)
{
    return (size_t)p->buckets_per_octave * (size_t)p->octaves; /* vqt.rs:259-261 */
}

/* vqt.rs:517-587 */
int orc_filter_bank_params(const orc_params *p, orc_filter_params *out, orc_error *err)
{
    size_t nb = params_n_buckets(p);
    float bpo = (float)p->buckets_per_octave;
    float highest = p->min_freq * powf(2.0f, (float)(nb - 1) / bpo);    /* :518-521 */
    float nyquist = p->sr / 2.0f;                                        /* :522 */
    if (highest > nyquist) {                                             /* :523-528 */
        if (err) { err->code = ORC_ABOVE_NYQUIST; err->a = highest; err->b = nyquist; err->n = 0; }
        return ORC_ABOVE_NYQUIST;
    }
    float r = powf(2.0f, 1.0f / bpo);                                    /* :532 */
    float alpha = (r * r - 1.0f) / (r * r + 1.0f);                       /* :533 */
    for (size_t k = 0; k < nb; ++k) {
        float freq = p->min_freq * powf(2.0f, (float)k / bpo);           /* :537-538 */
        float window_length = p->quality * p->sr / (alpha * freq + p->gamma); /* :539 */
        const float GRACE_FACTOR = 1.15f;                                /* :545 */
        float minimum_scaled_sr = ceilf(freq * 2.0f * GRACE_FACTOR);     /* :546 */
        uint64_t kk = f32_as_u64(floorf(log2f(p->sr / minimum_scaled_sr))) & 0xffffffffu; /* :548 */
        uint64_t factor = (uint64_t)1 << kk;                             /* :549 */
        uint64_t k2 = f32_as_u64(floorf(log2f((float)p->n_fft / window_length))) & 0xffffffffu; /* :554 */
        uint64_t min_ws = p->n_fft >> k2;                                /* :555 */
        out[k].freq = freq;
        out[k].window_length = window_length;
        out[k].sr_downscaling_factor = factor;
        out[k].minimum_needed_window_size = min_ws;
    }
    float longest = out[0].window_length;                                /* :567 */
    if (longest > (float)p->n_fft) {                                     /* :568-573 */
        if (err) { err->code = ORC_WINDOW_EXCEEDS_NFFT; err->a = longest; err->b = 0; err->n = p->n_fft; }
        return ORC_WINDOW_EXCEEDS_NFFT;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* kernel construction                                                       */
/* ------------------------------------------------------------------------- */

static int cmp_f32_total(const void *a, const void *b)
{
    /* f32::total_cmp on non-negative finite values == numeric order */
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* vqt.rs:769-852.  Returns 0 or ORC_ASSERT.  v: scaled_n_fft complex (f32). */
static int calculate_filter(float sr, float sparsity_quantile, uint64_t sr_scaling,
                            const orc_filter_params *fp, uint64_t win_begin, uint64_t win_end,
                            float window_center, c32 *v, size_t scaled_n_fft)
{
    float m = (float)sr_scaling;
    float scaled_freq = fp->freq * m;                                      /* :778 */
    float scaled_window_length = fp->window_length / m;                    /* :779 */
    uint64_t len = f32_as_u64(roundf(scaled_window_length));               /* :780 */
    float scaled_window_center = (window_center - (float)win_begin) / m;   /* :781 */
    uint64_t center = f32_as_u64(floorf(scaled_window_center));            /* :782 */
    if ((win_end - win_begin) / sr_scaling != scaled_n_fft) return ORC_ASSERT; /* :783 */
    if (len > scaled_n_fft) return ORC_ASSERT;                             /* :785 */
    if (center < len / 2) return ORC_ASSERT;                               /* :786-788 */
    uint64_t filter_begin = center - len / 2;
    if (filter_begin + len > scaled_n_fft) return ORC_ASSERT;              /* :789-792 */

    memset(v, 0, sizeof(c32) * scaled_n_fft);                              /* :796 */
    for (uint64_t i = 0; i < len; ++i) {                                   /* :797-800 */
        /* apodize::hanning_iter: f64, symmetric */
        double w = (len > 1) ? 0.5 - 0.5 * cos(2.0 * ORC_PI_D * (double)i / (double)(len - 1)) : 1.0;
        /* Complex32::i() * 2.0 * PI * (i as f32) * scaled_freq / sr, left to right in f32 */
        float th = 1.0f * 2.0f;
        th = th * ORC_PI_F;
        th = th * (float)i;
        th = th * scaled_freq;
        th = th / sr;
        float wf = (float)w;
        v[filter_begin + i].re = wf * cosf(th);  /* exp(0 + i th) = (cos th, sin th) */
        v[filter_begin + i].im = wf * sinf(th);
    }

    float norm_1 = 0.0f;                                                   /* :804 */
    for (size_t i = 0; i < scaled_n_fft; ++i) norm_1 += hypotf(v[i].re, v[i].im);
    for (size_t i = 0; i < scaled_n_fft; ++i) { v[i].re /= norm_1; v[i].im /= norm_1; } /* :805 */

    /* :808 rustfft forward FFT (f32 there; f64 here, rounded to f32 -- equivalent
     * up to rustfft's own rounding noise, see DESIGN.md "kernel values") */
    c64 *t = (c64 *)malloc(sizeof(c64) * scaled_n_fft);
    for (size_t i = 0; i < scaled_n_fft; ++i) { t[i].re = v[i].re; t[i].im = v[i].im; }
    dft_f64_any(t, scaled_n_fft);
    for (size_t i = 0; i < scaled_n_fft; ++i) {
        v[i].re = (float)t[i].re;
        v[i].im = -(float)t[i].im;                                          /* :811 conj */
    }
    free(t);

    /* :813-842 sparsify */
    float *resp = (float *)malloc(sizeof(float) * scaled_n_fft);
    for (size_t i = 0; i < scaled_n_fft; ++i) resp[i] = hypotf(v[i].re, v[i].im);
    qsort(resp, scaled_n_fft, sizeof(float), cmp_f32_total);               /* :823 */
    float v_abs_sum = 0.0f;                                                /* :824 */
    for (size_t i = 0; i < scaled_n_fft; ++i) v_abs_sum += resp[i];
    float accum = 0.0f;
    size_t cutoff_idx = 0;
    float limit = (1.0f - sparsity_quantile) * v_abs_sum;                  /* :827 */
    while (accum < limit) {
        if (cutoff_idx >= scaled_n_fft) { free(resp); return ORC_ASSERT; } /* index panic */
        accum += resp[cutoff_idx];
        cutoff_idx += 1;
    }
    float cutoff_value = cutoff_idx == 0 ? 0.0f : resp[cutoff_idx - 1];    /* :831-835 */
    for (size_t i = 0; i < scaled_n_fft; ++i) {                            /* :837-842 */
        if (hypotf(v[i].re, v[i].im) < cutoff_value) { v[i].re = 0.0f; v[i].im = 0.0f; }
    }
    free(resp);
    return ORC_OK;
}

typedef struct { int32_t row, col; float re, im; } triplet;

static int cmp_triplet(const void *a, const void *b)
{
    const triplet *x = (const triplet *)a, *y = (const triplet *)b;
    if (x->row != y->row) return (x->row > y->row) - (x->row < y->row);
    return (x->col > y->col) - (x->col < y->col);
}

/* sprs TriMat::to_csr: rows with ascending column indices (no duplicates occur here). */
static void triplets_to_csr(triplet *t, size_t nnz, int32_t rows, int32_t cols, orc_csr *m)
{
    qsort(t, nnz, sizeof(triplet), cmp_triplet);
    m->rows = rows; m->cols = cols; m->nnz = (int64_t)nnz;
    m->indptr = (int32_t *)calloc((size_t)rows + 1, sizeof(int32_t));
    m->indices = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
    m->data = (float *)malloc(sizeof(float) * 2 * (nnz ? nnz : 1));
    for (size_t i = 0; i < nnz; ++i) {
        m->indptr[t[i].row + 1] += 1;
        m->indices[i] = t[i].col;
        m->data[2 * i] = t[i].re; m->data[2 * i + 1] = t[i].im;
    }
    for (int32_t r = 0; r < rows; ++r) m->indptr[r + 1] += m->indptr[r];
}

/* vqt.rs:599-759 */
int orc_vqt_new(const orc_params *p, orc_vqt **out, orc_error *err)
{
    orc_error local; if (!err) err = &local;
    memset(err, 0, sizeof(*err));
    *out = NULL;
    size_t nb = params_n_buckets(p);
    orc_filter_params *filters = (orc_filter_params *)malloc(sizeof(orc_filter_params) * (nb ? nb : 1));
    int rc = orc_filter_bank_params(p, filters, err);
    if (rc) { free(filters); return rc; }

    float max_window_length = filters[0].window_length;                    /* :604 */
    float window_center = (float)p->n_fft - max_window_length / 2.0f;      /* :605 */

    /* rate groups: runs of equal sr_downscaling_factor (:616-642) */
    typedef struct { uint64_t m, wb, we; size_t first, count; } rate_group;
    rate_group *rgs = (rate_group *)malloc(sizeof(rate_group) * nb);
    size_t n_rg = 0;
    for (size_t i = 0; i < nb;) {
        size_t j = i;
        uint64_t ws = 0;
        while (j < nb && filters[j].sr_downscaling_factor == filters[i].sr_downscaling_factor) {
            if (filters[j].minimum_needed_window_size > ws) ws = filters[j].minimum_needed_window_size;
            ++j;
        }
        rate_group g; g.m = filters[i].sr_downscaling_factor; g.first = i; g.count = j - i;
        float half = (float)ws / 2.0f;
        if ((window_center + half) < (float)p->n_fft) {                    /* :627 */
            g.wb = f32_as_u64(window_center - half);                       /* :630 */
            g.we = f32_as_u64(window_center + half);                       /* :631 */
        } else {
            g.wb = p->n_fft - ws; g.we = p->n_fft;                         /* :634 */
        }
        rgs[n_rg++] = g;
        i = j;
    }

    float kernel_gain = sqrtf(p->sr);                                      /* :646 */

    orc_vqt *v = (orc_vqt *)calloc(1, sizeof(orc_vqt));
    v->params = *p; v->n_buckets = nb;
    v->groups = (orc_group *)calloc(n_rg, sizeof(orc_group));

    /* window groups: runs of rate groups with equal window (:653-754) */
    for (size_t a = 0; a < n_rg;) {
        size_t b = a;
        size_t n_filters = 0;
        while (b < n_rg && rgs[b].wb == rgs[a].wb && rgs[b].we == rgs[a].we) { n_filters += rgs[b].count; ++b; }
        uint64_t wb = rgs[a].wb, we = rgs[a].we;
        uint64_t window_size = we - wb;                                    /* :657 */
        size_t n_spectrum = window_size / 2 + 1;                           /* :658 */
        size_t cap = 0;
        for (size_t g = a; g < b; ++g) cap += rgs[g].count * (size_t)(window_size / rgs[g].m);
        triplet *pos = (triplet *)malloc(sizeof(triplet) * (cap ? cap : 1));
        triplet *neg = (triplet *)malloc(sizeof(triplet) * (cap ? cap : 1));
        size_t npos = 0, nneg = 0;
        int32_t row = 0;
        for (size_t g = a; g < b && !rc; ++g) {
            uint64_t m = rgs[g].m;
            size_t scaled_n_fft = (size_t)(window_size / m);               /* :674 */
            c32 *fv = (c32 *)malloc(sizeof(c32) * (scaled_n_fft ? scaled_n_fft : 1));
            for (size_t f = 0; f < rgs[g].count; ++f) {
                rc = calculate_filter(p->sr, p->sparsity_quantile, m, &filters[rgs[g].first + f], wb, we,
                                      window_center, fv, scaled_n_fft);
                if (rc) break;
                for (size_t j = 0; j < scaled_n_fft; ++j) {                /* :725-735 */
                    if (fv[j].re == 0.0f && fv[j].im == 0.0f) continue;
                    float vr = fv[j].re * kernel_gain, vi = fv[j].im * kernel_gain; /* z * gain   */
                    vr = vr / (float)window_size; vi = vi / (float)window_size;     /* / window_size */
                    if (j <= scaled_n_fft / 2) {
                        triplet t = { row, (int32_t)j, vr, vi }; pos[npos++] = t;
                    } else {
                        triplet t = { row, (int32_t)(scaled_n_fft - j), vr, -vi }; neg[nneg++] = t;
                    }
                }
                row += 1;
            }
            free(fv);
        }
        if (rc) {
            free(pos); free(neg); free(rgs); free(filters);
            orc_vqt_free(v);
            err->code = rc;
            return rc;
        }
        orc_group *grp = &v->groups[v->n_groups++];
        grp->window_begin = wb; grp->window_end = we;
        triplets_to_csr(pos, npos, (int32_t)n_filters, (int32_t)n_spectrum, &grp->filter_bank);
        triplets_to_csr(neg, nneg, (int32_t)n_filters, (int32_t)n_spectrum, &grp->negative_filter_bank);
        free(pos); free(neg);
        a = b;
    }

    /* :756 Duration::from_secs_f32((n_fft as f32 - window_center) / sr) */
    v->delay_s = (double)(((float)p->n_fft - window_center) / p->sr);
    free(rgs); free(filters);
    *out = v;
    return ORC_OK;
}

void orc_vqt_free(orc_vqt *v)
{
    if (!v) return;
    for (size_t g = 0; g < v->n_groups; ++g) {
        csr_free(&v->groups[g].filter_bank);
        csr_free(&v->groups[g].negative_filter_bank);
    }
    free(v->groups);
    free(v);
}

size_t orc_n_buckets(const orc_vqt *v) { return v->n_buckets; }
double orc_delay_seconds(const orc_vqt *v) { return v->delay_s; }
size_t orc_num_groups(const orc_vqt *v) { return v->n_groups; }
const orc_group *orc_group_at(const orc_vqt *v, size_t g) { return g < v->n_groups ? &v->groups[g] : NULL; }

int orc_vqt_set_group(orc_vqt *v, size_t g, int neg, int32_t rows, int32_t cols, int64_t nnz,
                      const int32_t *indptr, const int32_t *indices, const float *data)
{
    if (g >= v->n_groups) return ORC_ASSERT;
    orc_csr *m = neg ? &v->groups[g].negative_filter_bank : &v->groups[g].filter_bank;
    csr_free(m);
    m->rows = rows; m->cols = cols; m->nnz = nnz;
    m->indptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)rows + 1));
    m->indices = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
    m->data = (float *)malloc(sizeof(float) * 2 * (size_t)(nnz ? nnz : 1));
    memcpy(m->indptr, indptr, sizeof(int32_t) * ((size_t)rows + 1));
    if (nnz) {
        memcpy(m->indices, indices, sizeof(int32_t) * (size_t)nnz);
        memcpy(m->data, data, sizeof(float) * 2 * (size_t)nnz);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* runtime                                                                   */
/* ------------------------------------------------------------------------- */

/* vqt.rs:922-954, in f32 and in the reference's op order */
void orc_power_to_db(const float *power, size_t n, float *out)
{
    const float REF_POWER = 0.3f * 0.3f;
    const float A_MIN = 1e-6f * 1e-6f;
    const float TOP_DB = 60.0f;
    float ref_db = 10.0f * log10f(REF_POWER);                              /* :927 */
    float mx = -3.40282347e+38f, mn = 3.40282347e+38f;                     /* f32::MIN / MAX */
    for (size_t i = 0; i < n; ++i) {
        float pw = power[i] > A_MIN ? power[i] : A_MIN;                    /* .max(A_MIN) */
        float ls = 10.0f * log10f(pw) - ref_db;                            /* :930 */
        out[i] = ls;
        mx = ls > mx ? ls : mx;
        mn = ls < mn ? ls : mn;
    }
    float floor_ = mx - TOP_DB;                                            /* :939 */
    float log_spec_min = mn > floor_ ? mn : floor_;                        /* :940 */
    for (size_t i = 0; i < n; ++i) {                                       /* :944-951 */
        float clamped = out[i] > floor_ ? out[i] : floor_;
        out[i] = (log_spec_min > 0.0f) ? (clamped - log_spec_min) : (clamped > 0.0f ? clamped : 0.0f);
    }
}

typedef struct {
    size_t    n_groups;
    f32_plan *plans;   /* per group (mode 1) */
    c32      *spec32;  /* max n_spectrum      */
    c32      *work32;  /* 2 * max n/2 complex  */
    c64      *buf64;   /* max window size      */
    float    *power;   /* n_buckets            */
} scratch;

static void scratch_init(scratch *s, const orc_vqt *v, int mode)
{
    memset(s, 0, sizeof(*s));
    size_t maxw = 0;
    for (size_t g = 0; g < v->n_groups; ++g) {
        size_t w = v->groups[g].window_end - v->groups[g].window_begin;
        if (w > maxw) maxw = w;
    }
    s->n_groups = v->n_groups;
    s->power = (float *)malloc(sizeof(float) * (v->n_buckets ? v->n_buckets : 1));
    if (mode == 1) {
        s->plans = (f32_plan *)calloc(v->n_groups, sizeof(f32_plan));
        for (size_t g = 0; g < v->n_groups; ++g)
            f32_plan_init(&s->plans[g], (v->groups[g].window_end - v->groups[g].window_begin) / 2);
        s->spec32 = (c32 *)malloc(sizeof(c32) * (maxw / 2 + 1));
        s->work32 = (c32 *)malloc(sizeof(c32) * (maxw + 2));
    } else {
        s->buf64 = (c64 *)malloc(sizeof(c64) * (maxw ? maxw : 1));
    }
}

static void scratch_free(scratch *s)
{
    if (s->plans) { for (size_t g = 0; g < s->n_groups; ++g) f32_plan_free(&s->plans[g]); free(s->plans); }
    free(s->spec32); free(s->work32); free(s->buf64); free(s->power);
}

/* vqt.rs:873-913 for one frame; writes |z|^2 (f32) per bucket into s->power */
static void frame_power(const orc_vqt *v, const float *x, int mode, scratch *s)
{
    size_t offset = 0;
    for (size_t g = 0; g < v->n_groups; ++g) {
        const orc_group *grp = &v->groups[g];
        size_t wb = grp->window_begin, ws = grp->window_end - grp->window_begin;
        const orc_csr *K = &grp->filter_bank, *Kn = &grp->negative_filter_bank;
        size_t rows = (size_t)K->rows;
        if (mode == 1) {
            rfft_f32(&s->plans[g], x + wb, s->spec32, s->work32);          /* :881-887 */
            const c32 *X = s->spec32;
            for (size_t r = 0; r < rows; ++r) {
                float ar = 0.0f, ai = 0.0f;                                /* x_vqt.fill(0) :873 */
                for (int32_t e = K->indptr[r]; e < K->indptr[r + 1]; ++e) { /* :890-894 */
                    float kr = K->data[2 * e], ki = K->data[2 * e + 1];
                    c32 xv = X[K->indices[e]];
                    ar += kr * xv.re - ki * xv.im;
                    ai += kr * xv.im + ki * xv.re;
                }
                if (Kn->nnz > 0) {                                         /* :896-910 */
                    float nr = 0.0f, ni = 0.0f;
                    for (int32_t e = Kn->indptr[r]; e < Kn->indptr[r + 1]; ++e) {
                        float kr = Kn->data[2 * e], ki = Kn->data[2 * e + 1];
                        c32 xv = X[Kn->indices[e]];
                        nr += kr * xv.re - ki * xv.im;
                        ni += kr * xv.im + ki * xv.re;
                    }
                    ar += nr; ai += -ni;                                   /* acc += neg.conj() */
                }
                s->power[offset + r] = ar * ar + ai * ai;                  /* norm_sqr :930 */
            }
        } else {
            c64 *b = s->buf64;
            for (size_t i = 0; i < ws; ++i) { b[i].re = x[wb + i]; b[i].im = 0.0; }
            dft_f64_any(b, ws);
            for (size_t r = 0; r < rows; ++r) {
                double ar = 0.0, ai = 0.0;
                for (int32_t e = K->indptr[r]; e < K->indptr[r + 1]; ++e) {
                    double kr = K->data[2 * e], ki = K->data[2 * e + 1];
                    c64 xv = b[K->indices[e]];
                    ar += kr * xv.re - ki * xv.im;
                    ai += kr * xv.im + ki * xv.re;
                }
                if (Kn->nnz > 0) {
                    double nr = 0.0, ni = 0.0;
                    for (int32_t e = Kn->indptr[r]; e < Kn->indptr[r + 1]; ++e) {
                        double kr = Kn->data[2 * e], ki = Kn->data[2 * e + 1];
                        c64 xv = b[Kn->indices[e]];
                        nr += kr * xv.re - ki * xv.im;
                        ni += kr * xv.im + ki * xv.re;
                    }
                    ar += nr; ai -= ni;
                }
                s->power[offset + r] = (float)(ar * ar + ai * ai);
            }
        }
        offset += rows;                                                    /* :912 */
    }
}

int orc_calc_instant_db(orc_vqt *v, const float *x, size_t n, int mode, float *out_db, float *out_power)
{
    if (n != v->params.n_fft) return ORC_BAD_LENGTH;                       /* :867-871 */
    scratch s; scratch_init(&s, v, mode);
    frame_power(v, x, mode, &s);
    if (out_power) memcpy(out_power, s.power, sizeof(float) * v->n_buckets);
    orc_power_to_db(s.power, v->n_buckets, out_db);                        /* :915 */
    scratch_free(&s);
    return ORC_OK;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_calc_batch_db(orc_vqt *v, const float *audio, size_t n_samples, size_t hop, size_t n_frames,
                      int mode, int n_threads, float *out_db)
{
    size_t n_fft = v->params.n_fft;
    if (n_frames == 0) return ORC_OK;
    if (n_samples < n_fft || (n_frames - 1) * hop + n_fft > n_samples) return ORC_BAD_LENGTH;
    size_t nb = v->n_buckets;
#ifdef _OPENMP
    int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel num_threads(nt)
    {
        scratch s; scratch_init(&s, v, mode);
#pragma omp for schedule(static)
        for (long long t = 0; t < (long long)n_frames; ++t) {
            frame_power(v, audio + (size_t)t * hop, mode, &s);
            orc_power_to_db(s.power, nb, out_db + (size_t)t * nb);
        }
        scratch_free(&s);
    }
#else
    (void)n_threads;
    scratch s; scratch_init(&s, v, mode);
    for (size_t t = 0; t < n_frames; ++t) {
        frame_power(v, audio + t * hop, mode, &s);
        orc_power_to_db(s.power, nb, out_db + t * nb);
    }
    scratch_free(&s);
#endif
    return ORC_OK;
}

/* util.rs:62-79 */
void orc_test_create_sines(const orc_params *p, const float *freqs, size_t n_freqs, float t_diff, float *wave)
{
    for (size_t i = 0; i < p->n_fft; ++i) wave[i] = 0.0f;
    for (size_t f = 0; f < n_freqs; ++f) {
        for (size_t i = 0; i < p->n_fft; ++i) {
            /* (((i as f32 + t_diff * sr) * 2.0 * PI / sr) * f).sin() / 12.0 */
            float a = (float)i + t_diff * p->sr;
            a = a * 2.0f;
            a = a * ORC_PI_F;
            a = a / p->sr;
            a = a * freqs[f];
            wave[i] += sinf(a) / 12.0f;
        }
    }
}

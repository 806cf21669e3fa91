/*
 * analysis_oracle.c -- CPU restatement of AnalysisState::preprocess and its modules (plain C).
 *
 * TEST INFRASTRUCTURE ONLY -- see analysis_oracle.h for the pinning status and the assumption made
 * about the third-party find_peaks 0.1.5 crate (Cargo.lock:2949; call site peak_detection.rs:31-42).
 * f32 arithmetic follows the op order of the Rust sources (build with -ffp-contract=off).
 */
#include "analysis_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int      has_horizon;   /* Option<Duration>::is_some() */
    uint64_t horizon_ns;
    float    y;
} ema_t;

struct orc_analysis {
    orc_analysis_params params;
    float    min_freq;
    uint32_t octaves, bpo;
    size_t   n;
    ema_t   *x_vqt_smoothed;            /* analysis.rs:128 */
    float   *x_vqt_peakfiltered;        /* :132 */
    float   *x_vqt_afterglow;           /* :136 */
    uint32_t *peaks; size_t n_peaks;    /* :139 (HashSet<usize>, kept ascending here) */
    orc_continuous_peak *peaks_continuous; size_t n_cont;  /* :142 */
    ema_t   *calmness;                  /* :149 */
    ema_t   *released_note_calmness;    /* :153 */
    float   *pitch_accuracy, *pitch_deviation;  /* :156,159 */
    ema_t    smoothed_scene_calmness;   /* :162 */
    ema_t    smoothed_tuning_grid_inaccuracy;  /* :176 */
    float   *values;                    /* scratch: x_vqt_smoothed_values */
};

/* ---- std::time::Duration helpers ------------------------------------------------------------ */
static float dur_as_secs_f32(uint64_t ns)
{
    /* Duration::as_secs_f32: (secs as f32) + (nanos as f32) / 1e9 */
    uint64_t secs = ns / 1000000000ull;
    uint32_t nanos = (uint32_t)(ns % 1000000000ull);
    return (float)secs + (float)nanos / 1000000000.0f;
}
static float dur_as_millis_f32(uint64_t ns) { return (float)(ns / 1000000ull); } /* as_millis() as f32 */
static uint64_t f32_as_u64(float x)
{
    if (!(x > 0.0f)) return 0;
    if (x >= 18446744073709551615.0f) return UINT64_MAX;
    return (uint64_t)x;
}
static uint64_t dur_from_millis_f32(float ms) { return f32_as_u64(ms) * 1000000ull; } /* from_millis(x as u64) */

/* ---- EmaMeasurement, util.rs:91-137 ---------------------------------------------------------- */
static void ema_update(ema_t *e, float new_value, uint64_t timestep_ns)
{
    if (e->has_horizon) {
        float alpha = 1.0f - expf(-2.0f * dur_as_secs_f32(timestep_ns) / dur_as_secs_f32(e->horizon_ns)); /* :108 */
        e->y = e->y + alpha * (new_value - e->y);                                                           /* :124 */
    } else {
        e->y = new_value;                                                                                   /* :119 */
    }
}

float orc_ema_update(float y, int has_horizon, uint64_t horizon_ns, float new_value, uint64_t timestep_ns)
{
    ema_t e = { has_horizon, horizon_ns, y };
    ema_update(&e, new_value, timestep_ns);
    return e.y;
}

/* ---- find_peaks 0.1.5, assumed == scipy.signal.find_peaks(height, distance, prominence) ------- */
typedef struct { float pr; uint32_t idx; } prio_t;
static int cmp_prio(const void *a, const void *b)
{
    const prio_t *x = (const prio_t *)a, *y = (const prio_t *)b;
    if (x->pr != y->pr) return (x->pr > y->pr) - (x->pr < y->pr);
    return (x->idx > y->idx) - (x->idx < y->idx);   /* stable: equal heights stay in index order */
}

static size_t select_by_distance(const float *x, uint32_t *pk, size_t m, size_t distance)
{
    if (m == 0 || distance <= 1) return m;
    unsigned char *keep = (unsigned char *)malloc(m);
    prio_t *order = (prio_t *)malloc(sizeof(prio_t) * m);
    memset(keep, 1, m);
    for (size_t i = 0; i < m; ++i) { order[i].pr = x[pk[i]]; order[i].idx = (uint32_t)i; }
    qsort(order, m, sizeof(prio_t), cmp_prio);
    for (size_t t = m; t-- > 0;) {          /* highest priority first */
        size_t j = order[t].idx;
        if (!keep[j]) continue;
        for (size_t k = j; k-- > 0 && pk[j] - pk[k] < distance;) keep[k] = 0;
        for (size_t k = j + 1; k < m && pk[k] - pk[j] < distance; ++k) keep[k] = 0;
    }
    size_t w = 0;
    for (size_t i = 0; i < m; ++i) if (keep[i]) pk[w++] = pk[i];
    free(keep); free(order);
    return w;
}

static float prominence_of(const float *x, size_t n, size_t peak)
{
    float h = x[peak], left_min = h, right_min = h;
    for (size_t i = peak; ; --i) {
        if (x[i] > h) break;
        if (x[i] < left_min) left_min = x[i];
        if (i == 0) break;
    }
    for (size_t i = peak; i < n; ++i) {
        if (x[i] > h) break;
        if (x[i] < right_min) right_min = x[i];
    }
    return h - (left_min > right_min ? left_min : right_min);
}

size_t orc_find_peaks(const float *x, size_t n, float min_prominence, float min_height, uint32_t bpo, int order,
                      uint32_t *out, size_t cap)
{
    if (n < 3) return 0;
    uint32_t *pk = (uint32_t *)malloc(sizeof(uint32_t) * n);
    size_t m = 0;
    /* strict local maxima, plateaus reported at their middle (PeakFinder / Peak::middle_position) */
    for (size_t i = 1, imax = n - 1; i < imax; ++i) {
        if (x[i - 1] < x[i]) {
            size_t ahead = i + 1;
            while (ahead < imax && x[ahead] == x[i]) ++ahead;
            if (x[ahead] < x[i]) {
                pk[m++] = (uint32_t)((i + ahead - 1) / 2);
                i = ahead;
            }
        }
    }
    /* with_min_height (peak_detection.rs:33) */
    size_t w = 0;
    for (size_t i = 0; i < m; ++i) if (x[pk[i]] >= min_height) pk[w++] = pk[i];
    m = w;
    /* with_min_distance: round(bpo * 0.4 / 12) bins (peak_detection.rs:37-40) */
    size_t min_sep = (size_t)f32_as_u64(roundf((float)bpo * 0.4f / 12.0f));
    if (order == 0 && min_sep > 0) m = select_by_distance(x, pk, m, min_sep);
    /* with_min_prominence (peak_detection.rs:32) */
    w = 0;
    for (size_t i = 0; i < m; ++i) if (prominence_of(x, n, pk[i]) >= min_prominence) pk[w++] = pk[i];
    m = w;
    if (order == 1 && min_sep > 0) m = select_by_distance(x, pk, m, min_sep);
    /* drop the lowest half semitone (peak_detection.rs:45-50) */
    size_t min_bin = ((size_t)bpo / 12 + 1) / 2;   /* (bpo / 12).div_ceil(2) */
    w = 0;
    for (size_t i = 0; i < m; ++i) if (pk[i] >= min_bin) { if (w < cap) out[w] = pk[i]; ++w; }
    free(pk);
    return w;
}

/* ---- enhance_peaks_continuous, peak_detection.rs:61-148 -------------------------------------- */
static orc_continuous_peak enhance_one(const orc_analysis *a, const float *vqt, size_t p)
{
    orc_continuous_peak r;
    size_t n = a->n;
    if (p < 1 || p > n - 2) { r.center = (float)p; r.size = vqt[p]; return r; }      /* :71-77 */
    float bins = (float)a->bpo;
    float f_prev = a->min_freq * powf(2.0f, (float)(p - 1) / bins);                  /* :81-83 */
    float f_curr = a->min_freq * powf(2.0f, (float)p / bins);
    float f_next = a->min_freq * powf(2.0f, (float)(p + 1) / bins);
    float l0 = logf(f_prev), l1 = logf(f_curr), l2 = logf(f_next);                   /* :86 */
    float a0 = vqt[p - 1], a1 = vqt[p], a2 = vqt[p + 1];                             /* :87 */
    float denom = (l0 - l1) * (l0 - l2) * (l1 - l2);                                 /* :91 */
    if (fabsf(denom) < 1.1920929e-7f) { r.center = (float)p; r.size = vqt[p]; return r; } /* :93-99 */
    float qa = (l2 * (a1 - a0) + l0 * (a2 - a1) + l1 * (a0 - a2)) / denom;           /* :102-105 */
    float qb = ((l2 * l2) * (a0 - a1) + (l0 * l0) * (a1 - a2) + (l1 * l1) * (a2 - a0)) / denom; /* :107-110 */
    float log_f_peak;
    if (fabsf(qa) < 1.1920929e-7f) {                                                 /* :113-118 */
        log_f_peak = l1;
    } else {
        float v = -qb / (2.0f * qa);
        log_f_peak = v < l0 ? l0 : (v > l2 ? l2 : v);                                /* clamp(l0, l2) */
    }
    float f_peak = expf(log_f_peak);                                                 /* :124 */
    float center = bins * log2f(f_peak / a->min_freq);                               /* :125 */
    float hi = (float)n - 1.0f;
    float cc = center < 0.0f ? 0.0f : (center > hi ? hi : center);                   /* :130-131 */
    size_t lower = (size_t)f32_as_u64(floorf(cc));                                   /* :133 */
    size_t upper = lower + 1 < n - 1 ? lower + 1 : n - 1;                            /* :134 */
    float fract = cc - truncf(cc);                                                   /* :135 f32::fract */
    float size = vqt[lower] * (1.0f - fract) + vqt[upper] * fract;                   /* :137 */
    r.center = cc;
    r.size = size > 0.0f ? size : 0.0f;                                              /* :141 */
    return r;
}

static int cmp_center(const void *a, const void *b)
{
    float x = ((const orc_continuous_peak *)a)->center, y = ((const orc_continuous_peak *)b)->center;
    return (x > y) - (x < y);
}

/* ---- promote_bass_peaks_with_harmonics, peak_detection.rs:172-241 ----------------------------- */
static void promote_bass(const orc_analysis *a, orc_continuous_peak *pc, size_t m, const float *vqt)
{
    const float weights[4] = { 0.5f, 0.3f, 0.15f, 0.05f };                           /* :194 */
    size_t n = a->n;
    for (size_t i = 0; i < m; ++i) {
        if (pc[i].center > (float)a->params.highest_bassnote) continue;              /* :181 */
        float fundamental_freq = a->min_freq * powf(2.0f, pc[i].center / (float)a->bpo);   /* :186-187 */
        float fundamental_power = powf(10.0f, pc[i].size / 10.0f);                   /* :190 */
        float harmonic_score = 0.0f;
        for (int h = 2; h <= 5; ++h) {
            float harmonic_freq = fundamental_freq * (float)h;                       /* :197 */
            if (!(harmonic_freq >= a->min_freq)) continue;                           /* :200-204 */
            float harmonic_bin = (log2f(harmonic_freq) - log2f(a->min_freq)) * (float)a->bpo;
            if (harmonic_bin >= 0.0f && harmonic_bin < (float)n) {                   /* :207 */
                size_t lo = (size_t)f32_as_u64(floorf(harmonic_bin));                /* :209 */
                size_t hi = (size_t)f32_as_u64(ceilf(harmonic_bin));
                if (hi > n - 1) hi = n - 1;                                          /* :210 */
                float frac = harmonic_bin - truncf(harmonic_bin);                    /* :211 */
                float amp = lo == hi ? vqt[lo] : vqt[lo] * (1.0f - frac) + vqt[hi] * frac;  /* :213-217 */
                float harmonic_power = powf(10.0f, amp / 10.0f);                     /* :220 */
                float threshold_power = fundamental_power * a->params.harmonic_threshold;   /* :223 */
                if (harmonic_power > threshold_power) harmonic_score += harmonic_power * weights[h - 2]; /* :224-227 */
            }
        }
        if (harmonic_score > 0.0f) {                                                 /* :232-239 */
            float fp = fundamental_power > 1e-6f ? fundamental_power : 1e-6f;
            float boost = 1.0f + 0.5f * (harmonic_score / fp);
            float capped = boost < 1.5f ? boost : 1.5f;
            pc[i].size += 10.0f * log10f(capped);
        }
    }
}

/* ---- update_calmness, calmness.rs:23-95 -------------------------------------------------------- */
static void update_calmness(orc_analysis *a, const float *x_vqt, const float *smoothed, uint64_t frame_time_ns)
{
    size_t n = a->n;
    unsigned char *around = (unsigned char *)calloc(n, 1);
    uint32_t *pk = (uint32_t *)malloc(sizeof(uint32_t) * n);
    int radius = (int)(a->bpo / 12 / 3);                                             /* :36 */
    size_t m = orc_find_peaks(x_vqt, n, a->params.peak_config.min_prominence, a->params.peak_config.min_height,
                              a->bpo, 0, pk, n);                                     /* :39 unsmoothed */
    for (size_t i = 0; i < m; ++i) {
        int lo = (int)pk[i] - radius; if (lo < 0) lo = 0;                            /* :41-43 half-open */
        int hi = (int)pk[i] + radius; if (hi > (int)n) hi = (int)n;
        for (int j = lo; j < hi; ++j) around[j] = 1;
    }
    float weighted_calmness_sum = 0.0f, weight_sum = 0.0f;                           /* :49-50 */
    for (size_t b = 0; b < n; ++b) {
        if (around[b]) {
            ema_update(&a->calmness[b], 1.0f, frame_time_ns);                        /* :60 */
            a->released_note_calmness[b] = a->calmness[b];                           /* :63 */
            float amplitude_power = powf(10.0f, smoothed[b] / 10.0f);                /* :66-67 */
            weighted_calmness_sum += a->calmness[b].y * amplitude_power;             /* :69 */
            weight_sum += amplitude_power;                                           /* :70 */
        } else {
            ema_update(&a->calmness[b], 0.0f, frame_time_ns);                        /* :73 */
            ema_update(&a->released_note_calmness[b], 0.0f, frame_time_ns);          /* :74 */
            float rc = a->released_note_calmness[b].y;
            if (rc > 0.01f) {                                                        /* :78 */
                float rw = rc * 0.3f;                                                /* :80 */
                weighted_calmness_sum += rc * rw;
                weight_sum += rw;
            }
        }
    }
    if (weight_sum > 0.0f)                                                           /* :87-90 */
        ema_update(&a->smoothed_scene_calmness, weighted_calmness_sum / weight_sum, frame_time_ns);
    free(around); free(pk);
}

/* ---- public ---------------------------------------------------------------------------------- */
void orc_analysis_default_params(orc_analysis_params *p)
{
    /* analysis.rs:72-98 */
    p->spectrogram_length = 400;
    p->peak_config.min_prominence = 10.0f; p->peak_config.min_height = 4.0f;
    p->bassline_peak_config.min_prominence = 5.0f; p->bassline_peak_config.min_height = 3.5f;
    p->highest_bassnote = 12 * 2 + 4;
    p->vqt_smoothing_duration_base_ns = 70ull * 1000000ull;
    p->vqt_smoothing_calmness_min = 0.6f;
    p->vqt_smoothing_calmness_max = 2.0f;
    p->note_calmness_smoothing_duration_ns = 3500ull * 1000000ull;
    p->scene_calmness_smoothing_duration_ns = 800ull * 1000000ull;
    p->tuning_inaccuracy_smoothing_duration_ns = 4000ull * 1000000ull;
    p->harmonic_threshold = 0.3f;
}

static float frequency_multiplier(const orc_analysis *a, size_t bin)
{
    float octave_fraction = (float)bin / (float)a->bpo / (float)a->octaves;          /* analysis.rs:201-202 */
    return 1.5f - 0.5f * octave_fraction;                                            /* :203 */
}

orc_analysis *orc_analysis_new(float min_freq, uint32_t octaves, uint32_t bpo, const orc_analysis_params *p)
{
    orc_analysis *a = (orc_analysis *)calloc(1, sizeof(*a));
    a->params = *p; a->min_freq = min_freq; a->octaves = octaves; a->bpo = bpo;
    size_t n = a->n = (size_t)octaves * bpo;
    a->x_vqt_smoothed = (ema_t *)calloc(n ? n : 1, sizeof(ema_t));
    a->calmness = (ema_t *)calloc(n ? n : 1, sizeof(ema_t));
    a->released_note_calmness = (ema_t *)calloc(n ? n : 1, sizeof(ema_t));
    a->x_vqt_peakfiltered = (float *)calloc(n ? n : 1, sizeof(float));
    a->x_vqt_afterglow = (float *)calloc(n ? n : 1, sizeof(float));
    a->pitch_accuracy = (float *)calloc(n ? n : 1, sizeof(float));
    a->pitch_deviation = (float *)calloc(n ? n : 1, sizeof(float));
    a->values = (float *)calloc(n ? n : 1, sizeof(float));
    a->peaks = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    a->peaks_continuous = (orc_continuous_peak *)calloc(n ? n : 1, sizeof(orc_continuous_peak));
    for (size_t b = 0; b < n; ++b) {
        float duration_ms = dur_as_millis_f32(p->vqt_smoothing_duration_base_ns) * frequency_multiplier(a, b); /* :204-205 */
        a->x_vqt_smoothed[b].has_horizon = 1;
        a->x_vqt_smoothed[b].horizon_ns = dur_from_millis_f32(duration_ms);          /* :206 */
        a->calmness[b].has_horizon = 1; a->calmness[b].horizon_ns = p->note_calmness_smoothing_duration_ns;
        a->released_note_calmness[b] = a->calmness[b];
    }
    a->smoothed_scene_calmness.has_horizon = 1;
    a->smoothed_scene_calmness.horizon_ns = p->scene_calmness_smoothing_duration_ns;
    a->smoothed_tuning_grid_inaccuracy.has_horizon = 1;
    a->smoothed_tuning_grid_inaccuracy.horizon_ns = p->tuning_inaccuracy_smoothing_duration_ns;
    return a;
}

void orc_analysis_free(orc_analysis *a)
{
    if (!a) return;
    free(a->x_vqt_smoothed); free(a->calmness); free(a->released_note_calmness); free(a->x_vqt_peakfiltered);
    free(a->x_vqt_afterglow); free(a->pitch_accuracy); free(a->pitch_deviation); free(a->values); free(a->peaks);
    free(a->peaks_continuous); free(a);
}

void orc_analysis_update_vqt_smoothing_duration(orc_analysis *a, int has_duration, uint64_t duration_ns)
{
    a->params.vqt_smoothing_duration_base_ns = has_duration ? duration_ns : 0;       /* analysis.rs:253 */
    for (size_t b = 0; b < a->n; ++b) {
        if (has_duration) {                                                          /* :257-264 */
            float duration_ms = dur_as_millis_f32(duration_ns) * frequency_multiplier(a, b);
            a->x_vqt_smoothed[b].has_horizon = 1;
            a->x_vqt_smoothed[b].horizon_ns = dur_from_millis_f32(duration_ms);
        } else {
            a->x_vqt_smoothed[b].has_horizon = 0;                                    /* :267 */
        }
    }
}

int orc_analysis_preprocess(orc_analysis *a, const float *x_vqt, size_t n_in, uint64_t frame_time_ns)
{
    size_t n = a->n;
    if (n_in != n) return 4;                                                         /* analysis.rs:289 */
    float calmness = a->smoothed_scene_calmness.y;                                   /* :295 */
    float calmness_multiplier = a->params.vqt_smoothing_calmness_min +
        (a->params.vqt_smoothing_calmness_max - a->params.vqt_smoothing_calmness_min) * calmness; /* :296-298 */
    uint64_t base_ms_int = a->params.vqt_smoothing_duration_base_ns / 1000000ull;
    for (size_t b = 0; b < n; ++b) {                                                 /* :301-323 */
        if (base_ms_int > 0) {
            float duration_ms = (float)base_ms_int * frequency_multiplier(a, b) * calmness_multiplier; /* :315-317 */
            a->x_vqt_smoothed[b].has_horizon = 1;
            a->x_vqt_smoothed[b].horizon_ns = dur_from_millis_f32(duration_ms);      /* :319 */
        }
        ema_update(&a->x_vqt_smoothed[b], x_vqt[b], frame_time_ns);                  /* :322 */
        a->values[b] = a->x_vqt_smoothed[b].y;                                       /* :325-329 */
    }
    const float *sm = a->values;

    /* peaks: bass config up to highest_bassnote, general config above (analysis.rs:332-349) */
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    size_t m = orc_find_peaks(sm, n, a->params.bassline_peak_config.min_prominence,
                              a->params.bassline_peak_config.min_height, a->bpo, 0, tmp, n);
    a->n_peaks = 0;
    for (size_t i = 0; i < m; ++i) if (tmp[i] <= a->params.highest_bassnote) a->peaks[a->n_peaks++] = tmp[i];
    m = orc_find_peaks(sm, n, a->params.peak_config.min_prominence, a->params.peak_config.min_height, a->bpo, 0, tmp, n);
    for (size_t i = 0; i < m; ++i) if (tmp[i] > a->params.highest_bassnote) a->peaks[a->n_peaks++] = tmp[i];
    free(tmp);

    /* continuous peaks (analysis.rs:351-361) */
    a->n_cont = a->n_peaks;
    for (size_t i = 0; i < a->n_peaks; ++i) a->peaks_continuous[i] = enhance_one(a, sm, a->peaks[i]);
    qsort(a->peaks_continuous, a->n_cont, sizeof(orc_continuous_peak), cmp_center); /* peak_detection.rs:145 */
    promote_bass(a, a->peaks_continuous, a->n_cont, sm);

    /* apply_peak_filter (afterglow.rs:27-36) and update_afterglow (afterglow.rs:10-21) */
    for (size_t b = 0; b < n; ++b) a->x_vqt_peakfiltered[b] = 0.0f;
    for (size_t i = 0; i < a->n_peaks; ++i) a->x_vqt_peakfiltered[a->peaks[i]] = sm[a->peaks[i]];
    for (size_t b = 0; b < n; ++b) {
        float x = a->x_vqt_afterglow[b];
        x *= 0.85f - 0.15f * ((float)b / (float)n);
        if (x < sm[b]) x = sm[b];
        a->x_vqt_afterglow[b] = x;
    }

    update_calmness(a, x_vqt, sm, frame_time_ns);                                    /* analysis.rs:378-387 */

    /* update_tuning_inaccuracy, pitch_analysis.rs:48-75 */
    float inaccuracy_sum = 0.0f, power_sum = 0.0f;
    for (size_t i = 0; i < a->n_cont; ++i) {
        float power = powf(10.0f, a->peaks_continuous[i].size / 10.0f);
        power_sum += power;
        float cs = a->peaks_continuous[i].center * 12.0f / (float)a->bpo;
        inaccuracy_sum += fabsf(cs - roundf(cs)) * power;
    }
    float avg = power_sum > 0.0f ? inaccuracy_sum / power_sum : 0.0f;
    ema_update(&a->smoothed_tuning_grid_inaccuracy, 100.0f * avg, frame_time_ns);

    /* update_pitch_accuracy_and_deviation, pitch_analysis.rs:12-42 */
    for (size_t b = 0; b < n; ++b) { a->pitch_accuracy[b] = 0.0f; a->pitch_deviation[b] = 0.0f; }
    for (size_t i = 0; i < a->n_cont; ++i) {
        float cs = a->peaks_continuous[i].center * 12.0f / (float)a->bpo;
        float deviation = cs - roundf(cs);
        float drift = fabsf(deviation);
        float accuracy = 1.0f - 2.0f * drift; if (accuracy < 0.0f) accuracy = 0.0f;
        size_t bin = (size_t)f32_as_u64(roundf(a->peaks_continuous[i].center));
        if (bin < n) { a->pitch_accuracy[bin] = accuracy; a->pitch_deviation[bin] = deviation; }
    }
    return 0;
}

size_t orc_analysis_n_buckets(const orc_analysis *a) { return a->n; }

size_t orc_analysis_peaks(const orc_analysis *a, uint32_t *out, size_t cap)
{
    for (size_t i = 0; i < a->n_peaks && i < cap; ++i) out[i] = a->peaks[i];
    return a->n_peaks;
}

size_t orc_analysis_peaks_continuous(const orc_analysis *a, orc_continuous_peak *out, size_t cap)
{
    for (size_t i = 0; i < a->n_cont && i < cap; ++i) out[i] = a->peaks_continuous[i];
    return a->n_cont;
}

void orc_analysis_vectors(const orc_analysis *a, float *smoothed, float *peakfiltered, float *afterglow, float *calmness,
                          float *pitch_accuracy, float *pitch_deviation)
{
    for (size_t b = 0; b < a->n; ++b) {
        if (smoothed) smoothed[b] = a->x_vqt_smoothed[b].y;
        if (peakfiltered) peakfiltered[b] = a->x_vqt_peakfiltered[b];
        if (afterglow) afterglow[b] = a->x_vqt_afterglow[b];
        if (calmness) calmness[b] = a->calmness[b].y;
        if (pitch_accuracy) pitch_accuracy[b] = a->pitch_accuracy[b];
        if (pitch_deviation) pitch_deviation[b] = a->pitch_deviation[b];
    }
}

float orc_analysis_scene_calmness(const orc_analysis *a) { return a->smoothed_scene_calmness.y; }
float orc_analysis_tuning_inaccuracy(const orc_analysis *a) { return a->smoothed_tuning_grid_inaccuracy.y; }

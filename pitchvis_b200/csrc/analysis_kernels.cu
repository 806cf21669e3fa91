// analysis_kernels.cu -- K-analysis: the AnalysisState epilogue on sm_100a, and its C ABI
// (include/pvqt_analysis.h).
//
// Replaces AnalysisState::preprocess (analysis.rs:288-404) and the modules it calls.  The state of a
// stream is a recurrence in time (EMA, calmness feedback into the EMA horizon, afterglow), so the
// kernel runs one CTA per stream and walks the frames in order; all per-bin work of a frame is
// spread over the CTA, the three peak searches of a frame run cooperatively in shared memory.
//
// Built with -fmad=false: the reference's f32 expressions are not FMA-contracted, and a horizon is
// truncated to whole milliseconds (analysis.rs:319), so rounding differences are kept to libm's
// the rare last-bit differences of correctly rounded transcendentals (cr_* below).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <math_constants.h>

#include "last_error.hpp"
#include "pvqt_analysis.h"
#include "pvqt_internal.hpp"

namespace {

int afail(pvqt_status st, const std::string &m) { pvqt_detail::set_last_error(m); return st; }
int acuda(cudaError_t e, const char *what)
{
    pvqt_detail::set_last_error(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
    return PVQT_CUDA_ERROR;
}
#define ACUDA(call)                                           \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return acuda(_e, #call);       \
    } while (0)

constexpr int kThreads = 256;
constexpr int kMaxPeaksSmem = 256;  // peaks of one frame kept in shared memory (588 bins, distance 3 -> <= 196)

struct AnalysisKernelParams {
    pvqt_analysis_params prm;
    float    min_freq;
    int32_t  octaves, bpo, nb;
    int32_t  has_horizon;        // x_vqt_smoothed time horizon is Some(..)
    float   *st_smoothed, *st_calm, *st_released, *st_afterglow, *st_scalar;  // per-stream state
    const float *db;             // [S][T][NB]
    uint32_t n_frames;
    uint64_t frame_time_ns;
    pvqt_analysis_outputs out;   // device pointers
};

// ---- std::time::Duration / EmaMeasurement -------------------------------------------------------
__host__ __device__ inline float dur_as_secs_f32(uint64_t ns)
{
    return (float)(ns / 1000000000ull) + (float)(uint32_t)(ns % 1000000000ull) / 1000000000.0f;
}
__device__ __forceinline__ uint64_t f32_as_u64(float x)
{
    if (!(x > 0.0f)) return 0;
    if (x >= 18446744073709551615.0f) return 0xffffffffffffffffull;
    return (uint64_t)x;
}
// Correctly rounded f32 transcendentals.  The reference calls the platform libm (glibc's expf / logf /
// log2f / log10f / powf are correctly rounded in all but vanishingly rare cases); CUDA's f32 versions
// may be 1-2 ulp off, and the log-parabola refinement (peak_detection.rs:86-118) amplifies one ulp of
// logf into ~0.1 bin.  Evaluating in f64 and rounding once gives the same bits as the CPU path; the
// calls are per bin / per peak, a few hundred per frame, so FP64 throughput is irrelevant here.
__device__ __forceinline__ float cr_expf(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float cr_logf(float x) { return (float)log((double)x); }
__device__ __forceinline__ float cr_log2f(float x) { return (float)log2((double)x); }
__device__ __forceinline__ float cr_log10f(float x) { return (float)log10((double)x); }
__device__ __forceinline__ float cr_powf(float a, float b) { return (float)pow((double)a, (double)b); }

// alpha = 1 - exp(-2 dt / tau), util.rs:108
__device__ __forceinline__ float ema_alpha(float timestep_s, uint64_t horizon_ns)
{
    return 1.0f - cr_expf(-2.0f * timestep_s / dur_as_secs_f32(horizon_ns));
}
__device__ __forceinline__ float ema_step(float y, float alpha, float x) { return y + alpha * (x - y); }  // util.rs:124

// ---- cooperative peak search (find_peaks wrapper, peak_detection.rs:26-51) -------------------------
enum : unsigned char { kNone = 0, kUndecided = 1, kKept = 2, kRemoved = 3 };

// st[b] = 1 for every peak of x[0..n), else 0.  All threads of the CTA must call.
__device__ void find_peaks_block(const float *x, int n, float min_prominence, float min_height, int distance,
                                 int min_bin, unsigned char *st)
{
    const int tid = threadIdx.x;
    for (int b = tid; b < n; b += kThreads) st[b] = kNone;
    __syncthreads();
    // strict local maxima, plateaus at their middle, then min_height
    for (int b = tid + 1; b < n - 1; b += kThreads) {
        if (x[b - 1] < x[b]) {
            int ahead = b + 1;
            while (ahead < n - 1 && x[ahead] == x[b]) ++ahead;
            if (x[ahead] < x[b] && x[b] >= min_height) st[(b + ahead - 1) >> 1] = kUndecided;
        }
    }
    __syncthreads();
    // min_distance: taller peaks win (greedy by height == repeated "local champion" rounds)
    if (distance > 1) {
        int pending;
        do {
            for (int b = tid; b < n; b += kThreads) {
                if (st[b] != kUndecided) continue;
                bool has_kept = false, top = true;
                const int lo = max(b - distance + 1, 0), hi = min(b + distance - 1, n - 1);
                for (int j = lo; j <= hi; ++j) {
                    if (j == b) continue;
                    const unsigned char s = st[j];
                    if (s == kKept) has_kept = true;
                    else if (s == kUndecided && (x[j] > x[b] || (x[j] == x[b] && j > b))) top = false;
                }
                if (!has_kept && top) st[b] = kKept;
            }
            __syncthreads();
            pending = 0;
            for (int b = tid; b < n; b += kThreads) {
                if (st[b] != kUndecided) continue;
                bool has_kept = false;
                const int lo = max(b - distance + 1, 0), hi = min(b + distance - 1, n - 1);
                for (int j = lo; j <= hi; ++j) has_kept |= (j != b && st[j] == kKept);
                if (has_kept) st[b] = kRemoved;
                else pending = 1;
            }
            pending = __syncthreads_or(pending);
        } while (pending);
    } else {
        for (int b = tid; b < n; b += kThreads) if (st[b] == kUndecided) st[b] = kKept;
        __syncthreads();
    }
    // min_prominence, then drop the lowest half semitone.  The bases of a peak are the minima of the stretches to its
    // left and right over which nothing is higher (the walk stops at the first x[i] > h): for the tallest peaks those
    // stretches span the whole spectrum, and a thread walking them alone (one dependent shared-memory load per step) cost
    // 9 us per side and call.  A warp walks a peak's stretch 32 bins at a time instead: ballot for the first higher
    // bin, min over the lanes before it.  min is exact whatever the order, so the result is the sequential walk's.
    const int lane = tid & 31;
    for (int b0 = (tid & ~31); b0 < n; b0 += kThreads) {       // the warp's 32 consecutive bins of this round
        const int b = b0 + lane;
        const bool kept = b < n && st[b] == kKept;
        unsigned todo = __ballot_sync(0xffffffffu, kept);
        unsigned char r = 0;
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int p = b0 + src;
            const float h = x[p];
            float mins[2];
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                float m = h;
                for (int base = p;; base += side ? 32 : -32) {
                    const int i = side ? base + lane : base - lane;
                    const bool in = side ? i < n : i >= 0;
                    const float v = in ? x[i] : h;
                    const unsigned stop = __ballot_sync(0xffffffffu, in && !(v <= h));   // first bin that is higher (or NaN)
                    const int first = stop ? __ffs(stop) - 1 : 32;
                    if (in && lane < first) m = fminf(m, v);
                    const bool more = side ? base + 32 < n : base - 32 >= 0;
                    if (stop || !more) break;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
                mins[side] = m;
            }
            if (lane == src) r = (h - fmaxf(mins[0], mins[1]) >= min_prominence && p >= min_bin) ? 1 : 0;
        }
        if (b < n) st[b] = r;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads) analysis_kernel(const __grid_constant__ AnalysisKernelParams P)
{
    extern __shared__ __align__(16) unsigned char a_smem[];
    const int n = P.nb, tid = threadIdx.x, stream = blockIdx.x;
    float *xraw = reinterpret_cast<float *>(a_smem);
    float *sm = xraw + n, *calm = sm + n, *released = calm + n, *aglow = released + n, *termc = aglow + n,
          *termw = termc + n, *pacc = termw + n, *pdev = pacc + n;
    float2 *pk_cont = reinterpret_cast<float2 *>(pdev + n);
    float *pk_power = reinterpret_cast<float *>(pk_cont + kMaxPeaksSmem);
    int *pk_idx = reinterpret_cast<int *>(pk_power + kMaxPeaksSmem);
    int *scan = pk_idx + kMaxPeaksSmem;                       // kThreads + 1
    unsigned char *st_bass = reinterpret_cast<unsigned char *>(scan + kThreads + 2);
    unsigned char *st_gen = st_bass + n, *st_raw = st_gen + n;
    __shared__ float s_scene, s_tuning;
    __shared__ int s_npeaks;

    const size_t so = (size_t)stream * n;
    for (int b = tid; b < n; b += kThreads) {
        sm[b] = P.st_smoothed[so + b];
        calm[b] = P.st_calm[so + b];
        released[b] = P.st_released[so + b];
        aglow[b] = P.st_afterglow[so + b];
    }
    if (tid == 0) { s_scene = P.st_scalar[2 * stream]; s_tuning = P.st_scalar[2 * stream + 1]; }
    __syncthreads();

    const pvqt_analysis_params &prm = P.prm;
    const float ft = dur_as_secs_f32(P.frame_time_ns);
    const float alpha_calm = ema_alpha(ft, prm.note_calmness_smoothing_duration_ns);
    const float alpha_scene = ema_alpha(ft, prm.scene_calmness_smoothing_duration_ns);
    const float alpha_tuning = ema_alpha(ft, prm.tuning_inaccuracy_smoothing_duration_ns);
    const uint64_t base_ms = prm.vqt_smoothing_duration_base_ns / 1000000ull;   // as_millis()
    const float bpo_f = (float)P.bpo;
    const int distance = (int)f32_as_u64(roundf(bpo_f * 0.4f / 12.0f));         // peak_detection.rs:37
    const int min_bin = (P.bpo / 12 + 1) / 2;                                   // peak_detection.rs:45
    const int radius = P.bpo / 12 / 3;                                          // calmness.rs:36
    const int hb = prm.highest_bassnote > 0x7fffffffull ? 0x7fffffff : (int)prm.highest_bassnote;
    const int chunk = (n + kThreads - 1) / kThreads;

    for (uint32_t t = 0; t < P.n_frames; ++t) {
        const size_t fo = ((size_t)stream * P.n_frames + t);
        const float *x = P.db + fo * n;
        // ---- calmness-adaptive EMA of the dB spectrum (analysis.rs:295-329) -----------------------
        const float calmness_multiplier =
            prm.vqt_smoothing_calmness_min + (prm.vqt_smoothing_calmness_max - prm.vqt_smoothing_calmness_min) * s_scene;
        for (int b = tid; b < n; b += kThreads) {
            const float xv = x[b];
            xraw[b] = xv;
            float y = sm[b];
            if (!P.has_horizon) {
                y = xv;                                                           // util.rs:117-120
            } else {
                uint64_t horizon_ns = 0;                                          // base 0 ms: horizon stays 0 ms
                if (base_ms > 0) {
                    const float octave_fraction = (float)b / bpo_f / (float)P.octaves;
                    const float frequency_multiplier = 1.5f - 0.5f * octave_fraction;
                    const float duration_ms = (float)base_ms * frequency_multiplier * calmness_multiplier;
                    horizon_ns = f32_as_u64(duration_ms) * 1000000ull;            // from_millis(x as u64)
                }
                y = ema_step(y, ema_alpha(ft, horizon_ns), xv);
            }
            sm[b] = y;
        }
        __syncthreads();

        // ---- peaks: bass config up to highest_bassnote, general config above (analysis.rs:332-349) --
        find_peaks_block(sm, n, prm.bassline_peak_config.min_prominence, prm.bassline_peak_config.min_height, distance,
                         min_bin, st_bass);
        find_peaks_block(sm, n, prm.peak_config.min_prominence, prm.peak_config.min_height, distance, min_bin, st_gen);
        // unsmoothed peaks for the calmness update (calmness.rs:39)
        find_peaks_block(xraw, n, prm.peak_config.min_prominence, prm.peak_config.min_height, distance, min_bin, st_raw);

        // ordered compaction of the peak set
        {
            int cnt = 0;
            const int b0 = tid * chunk, b1 = min(b0 + chunk, n);
            for (int b = b0; b < b1; ++b) cnt += (b <= hb ? st_bass[b] : st_gen[b]);
            // exclusive scan of the per-thread counts: warp shuffles, then the 8 warp totals (a single thread adding 256
            // shared-memory entries one after the other cost 4.6 us per frame)
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if ((tid & 31) >= o) incl += up;
            }
            if ((tid & 31) == 31) scan[tid >> 5] = incl;
            __syncthreads();
            int before = 0;
            for (int wi = 0; wi < (tid >> 5); ++wi) before += scan[wi];
            if (tid == kThreads - 1) s_npeaks = before + incl;
            int w = before + incl - cnt;
            __syncthreads();
            for (int b = b0; b < b1; ++b)
                if (b <= hb ? st_bass[b] : st_gen[b]) { if (w < kMaxPeaksSmem) pk_idx[w] = b; ++w; }
        }
        __syncthreads();
        const int n_peaks = s_npeaks, n_stored = min(n_peaks, kMaxPeaksSmem);

        // ---- enhance_peaks_continuous + promote_bass_peaks_with_harmonics (per peak) -----------------
        for (int i = tid; i < n_stored; i += kThreads) {
            const int p = pk_idx[i];
            float center = (float)p, size = sm[p];
            if (p >= 1 && p <= n - 2) {                                           // peak_detection.rs:71-77
                const float f_prev = P.min_freq * cr_powf(2.0f, (float)(p - 1) / bpo_f);
                const float f_curr = P.min_freq * cr_powf(2.0f, (float)p / bpo_f);
                const float f_next = P.min_freq * cr_powf(2.0f, (float)(p + 1) / bpo_f);
                const float l0 = cr_logf(f_prev), l1 = cr_logf(f_curr), l2 = cr_logf(f_next);
                const float a0 = sm[p - 1], a1 = sm[p], a2 = sm[p + 1];
                const float denom = (l0 - l1) * (l0 - l2) * (l1 - l2);
                if (!(fabsf(denom) < 1.1920929e-7f)) {
                    const float qa = (l2 * (a1 - a0) + l0 * (a2 - a1) + l1 * (a0 - a2)) / denom;
                    const float qb = ((l2 * l2) * (a0 - a1) + (l0 * l0) * (a1 - a2) + (l1 * l1) * (a2 - a0)) / denom;
                    float log_f_peak = l1;
                    if (!(fabsf(qa) < 1.1920929e-7f)) {
                        const float v = -qb / (2.0f * qa);
                        log_f_peak = v < l0 ? l0 : (v > l2 ? l2 : v);
                    }
                    const float f_peak = cr_expf(log_f_peak);
                    const float c = bpo_f * cr_log2f(f_peak / P.min_freq);
                    const float hi = (float)n - 1.0f;
                    const float cc = c < 0.0f ? 0.0f : (c > hi ? hi : c);
                    const int lower = (int)f32_as_u64(floorf(cc));
                    const int upper = min(lower + 1, n - 1);
                    const float fract = cc - truncf(cc);
                    const float s = sm[lower] * (1.0f - fract) + sm[upper] * fract;
                    center = cc;
                    size = s > 0.0f ? s : 0.0f;
                }
            }
            if (!(center > (float)prm.highest_bassnote)) {                        // peak_detection.rs:181
                const float fundamental_freq = P.min_freq * cr_powf(2.0f, center / bpo_f);
                const float fundamental_power = cr_powf(10.0f, size / 10.0f);
                const float weights[4] = {0.5f, 0.3f, 0.15f, 0.05f};
                float harmonic_score = 0.0f;
#pragma unroll
                for (int h = 2; h <= 5; ++h) {
                    const float harmonic_freq = fundamental_freq * (float)h;
                    if (!(harmonic_freq >= P.min_freq)) continue;
                    const float harmonic_bin = (cr_log2f(harmonic_freq) - cr_log2f(P.min_freq)) * bpo_f;
                    if (harmonic_bin >= 0.0f && harmonic_bin < (float)n) {
                        const int lo = (int)f32_as_u64(floorf(harmonic_bin));
                        const int hi2 = min((int)f32_as_u64(ceilf(harmonic_bin)), n - 1);
                        const float frac = harmonic_bin - truncf(harmonic_bin);
                        const float amp = lo == hi2 ? sm[lo] : sm[lo] * (1.0f - frac) + sm[hi2] * frac;
                        const float harmonic_power = cr_powf(10.0f, amp / 10.0f);
                        if (harmonic_power > fundamental_power * prm.harmonic_threshold)
                            harmonic_score += harmonic_power * weights[h - 2];
                    }
                }
                if (harmonic_score > 0.0f) {
                    const float boost = 1.0f + 0.5f * (harmonic_score / fmaxf(fundamental_power, 1e-6f));
                    size += 10.0f * cr_log10f(fminf(boost, 1.5f));
                }
            }
            pk_cont[i] = make_float2(center, size);
            pk_power[i] = cr_powf(10.0f, size / 10.0f);                              // pitch_analysis.rs:58
        }

        // ---- peak filter, afterglow, per-bin calmness (afterglow.rs:10-36, calmness.rs:52-85) ---------
        for (int b = tid; b < n; b += kThreads) {
            const float s = sm[b];
            float g = aglow[b];
            g *= 0.85f - 0.15f * ((float)b / (float)n);
            if (g < s) g = s;
            aglow[b] = g;
            bool has_peak = false;                                                // [p - r, p + r) around unsmoothed peaks
            for (int p = max(b - radius + 1, 0); p <= min(b + radius, n - 1); ++p) has_peak |= st_raw[p] != 0;
            float c = calm[b], tc = 0.0f, tw = 0.0f;
            if (has_peak) {
                c = ema_step(c, alpha_calm, 1.0f);
                released[b] = c;
                const float amplitude_power = cr_powf(10.0f, s / 10.0f);
                tc = c * amplitude_power;
                tw = amplitude_power;
            } else {
                c = ema_step(c, alpha_calm, 0.0f);
                const float rc = ema_step(released[b], alpha_calm, 0.0f);
                released[b] = rc;
                if (rc > 0.01f) { tw = rc * 0.3f; tc = rc * tw; }
            }
            calm[b] = c;
            termc[b] = tc;
            termw[b] = tw;
            pacc[b] = 0.0f;
            pdev[b] = 0.0f;
        }
        __syncthreads();

        // ---- the reference's sequential f32 sums, in its order (calmness.rs:49-90, pitch_analysis.rs:54-74)
        if (tid == 0) {
            float wc = 0.0f, ws = 0.0f;
            for (int b = 0; b < n; ++b) { wc += termc[b]; ws += termw[b]; }
            if (ws > 0.0f) s_scene = ema_step(s_scene, alpha_scene, wc / ws);
        } else if (tid == 32) {
            float inaccuracy_sum = 0.0f, power_sum = 0.0f;
            for (int i = 0; i < n_stored; ++i) {
                power_sum += pk_power[i];
                const float cs = pk_cont[i].x * 12.0f / bpo_f;
                inaccuracy_sum += fabsf(cs - roundf(cs)) * pk_power[i];
            }
            const float avg = power_sum > 0.0f ? inaccuracy_sum / power_sum : 0.0f;
            s_tuning = ema_step(s_tuning, alpha_tuning, 100.0f * avg);
        } else if (tid == 64) {
            for (int i = 0; i < n_stored; ++i) {                                  // pitch_analysis.rs:24-41
                const float cs = pk_cont[i].x * 12.0f / bpo_f;
                const float deviation = cs - roundf(cs);
                const int bin = (int)f32_as_u64(roundf(pk_cont[i].x));
                if (bin < n) { pacc[bin] = fmaxf(1.0f - 2.0f * fabsf(deviation), 0.0f); pdev[bin] = deviation; }
            }
        }
        __syncthreads();

        // ---- per-frame outputs ------------------------------------------------------------------------
        const pvqt_analysis_outputs &O = P.out;
        if (tid == 0) {
            if (O.peak_count) O.peak_count[fo] = (uint32_t)n_peaks;
            if (O.smoothed_scene_calmness) O.smoothed_scene_calmness[fo] = s_scene;
            if (O.smoothed_tuning_grid_inaccuracy) O.smoothed_tuning_grid_inaccuracy[fo] = s_tuning;
        }
        const int n_out = min(n_stored, (int)O.max_peaks);
        for (int i = tid; i < n_out; i += kThreads) {
            if (O.peak_indices) O.peak_indices[fo * O.max_peaks + i] = (uint32_t)pk_idx[i];
            if (O.peaks_continuous) O.peaks_continuous[fo * O.max_peaks + i] = pvqt_continuous_peak{pk_cont[i].x, pk_cont[i].y};
        }
        for (int b = tid; b < n; b += kThreads) {
            const size_t o = fo * n + b;
            if (O.x_vqt_smoothed) O.x_vqt_smoothed[o] = sm[b];
            if (O.x_vqt_peakfiltered) O.x_vqt_peakfiltered[o] = (b <= hb ? st_bass[b] : st_gen[b]) ? sm[b] : 0.0f;
            if (O.x_vqt_afterglow) O.x_vqt_afterglow[o] = aglow[b];
            if (O.calmness) O.calmness[o] = calm[b];
            if (O.pitch_accuracy) O.pitch_accuracy[o] = pacc[b];
            if (O.pitch_deviation) O.pitch_deviation[o] = pdev[b];
        }
        __syncthreads();
    }

    for (int b = tid; b < n; b += kThreads) {
        P.st_smoothed[so + b] = sm[b];
        P.st_calm[so + b] = calm[b];
        P.st_released[so + b] = released[b];
        P.st_afterglow[so + b] = aglow[b];
    }
    if (tid == 0) { P.st_scalar[2 * stream] = s_scene; P.st_scalar[2 * stream + 1] = s_tuning; }
}

size_t analysis_smem_bytes(int nb)
{
    return sizeof(float) * 9 * (size_t)nb + sizeof(float2) * kMaxPeaksSmem + sizeof(float) * kMaxPeaksSmem +
           sizeof(int) * kMaxPeaksSmem + sizeof(int) * (kThreads + 2) + 3 * (size_t)nb + 16;
}

}  // namespace

struct pvqt_analysis {
    pvqt_analysis_params params{};
    pvqt_range range{};
    size_t n_streams = 0, nb = 0;
    int device = 0;
    int has_horizon = 1;
    cudaStream_t stream = nullptr;
    float *st_smoothed = nullptr, *st_calm = nullptr, *st_released = nullptr, *st_afterglow = nullptr, *st_scalar = nullptr;
    // device mirrors of the result buffers of the single-call pipeline (pvqt_calc_*_analysis), reused across calls
    static constexpr int kOutputs = 11;
    void *mirror[kOutputs] = {};
    size_t mirror_bytes[kOutputs] = {};
};

namespace {

// the members of pvqt_analysis_outputs in declaration order, with their bytes per frame
inline void *&out_member(pvqt_analysis_outputs &o, int i)
{
    void **m[pvqt_analysis::kOutputs] = {
        reinterpret_cast<void **>(&o.peak_count), reinterpret_cast<void **>(&o.peak_indices),
        reinterpret_cast<void **>(&o.peaks_continuous), reinterpret_cast<void **>(&o.x_vqt_smoothed),
        reinterpret_cast<void **>(&o.x_vqt_peakfiltered), reinterpret_cast<void **>(&o.x_vqt_afterglow),
        reinterpret_cast<void **>(&o.calmness), reinterpret_cast<void **>(&o.pitch_accuracy),
        reinterpret_cast<void **>(&o.pitch_deviation), reinterpret_cast<void **>(&o.smoothed_scene_calmness),
        reinterpret_cast<void **>(&o.smoothed_tuning_grid_inaccuracy)};
    return *m[i];
}
inline size_t out_bytes_per_frame(int i, size_t nb, size_t max_peaks)
{
    switch (i) {
    case 0: return sizeof(uint32_t);
    case 1: return max_peaks * sizeof(uint32_t);
    case 2: return max_peaks * sizeof(pvqt_continuous_peak);
    case 9: case 10: return sizeof(float);
    default: return nb * sizeof(float);
    }
}

}  // namespace

extern "C" {

int pvqt_analysis_default_params(pvqt_analysis_params *p)
{
    if (!p) return afail(PVQT_INVALID_ARGUMENT, "null argument");
    // analysis.rs:72-98
    p->spectrogram_length = 400;
    p->peak_config = {10.0f, 4.0f};
    p->bassline_peak_config = {5.0f, 3.5f};
    p->highest_bassnote = 12 * 2 + 4;
    p->vqt_smoothing_duration_base_ns = 70ull * 1000000ull;
    p->vqt_smoothing_calmness_min = 0.6f;
    p->vqt_smoothing_calmness_max = 2.0f;
    p->note_calmness_smoothing_duration_ns = 3500ull * 1000000ull;
    p->scene_calmness_smoothing_duration_ns = 800ull * 1000000ull;
    p->tuning_inaccuracy_smoothing_duration_ns = 4000ull * 1000000ull;
    p->harmonic_threshold = 0.3f;
    return PVQT_OK;
}

int pvqt_analysis_create(const pvqt_range *range, const pvqt_analysis_params *params, size_t n_streams, int device,
                         pvqt_analysis **out)
{
    if (!range || !params || !out || n_streams == 0) return afail(PVQT_INVALID_ARGUMENT, "bad argument");
    *out = nullptr;
    const size_t nb = (size_t)range->octaves * range->buckets_per_octave;
    if (nb < 3 || nb > 4096) return afail(PVQT_UNSUPPORTED, "n_buckets must be in [3, 4096]");
    {   // peaks are strict local maxima at least max(2, min_distance) bins apart: the kernel keeps kMaxPeaksSmem of them
        const float d = std::round((float)range->buckets_per_octave * 0.4f / 12.0f);   // peak_detection.rs:37
        const size_t apart = std::max<size_t>(2, d > 0.0f ? (size_t)d : 0);
        if ((nb + apart - 1) / apart > (size_t)kMaxPeaksSmem)
            return afail(PVQT_UNSUPPORTED, "more than " + std::to_string(kMaxPeaksSmem) +
                                               " peaks per frame are possible for this range (n_buckets / min_distance)");
    }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess) return acuda(e, "cudaGetDeviceCount (no CPU fallback exists)");
    if (device < 0 || device >= n_dev) return afail(PVQT_INVALID_ARGUMENT, "device out of range");
    ACUDA(cudaSetDevice(device));
    std::unique_ptr<pvqt_analysis> a(new pvqt_analysis());
    a->params = *params;
    a->range = *range;
    a->n_streams = n_streams;
    a->nb = nb;
    a->device = device;
    ACUDA(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
    const size_t bytes = n_streams * nb * sizeof(float);
    // AnalysisState::new: every EMA starts at 0, afterglow at 0 (analysis.rs:199-239)
    for (float **p : {&a->st_smoothed, &a->st_calm, &a->st_released, &a->st_afterglow}) {
        ACUDA(cudaMalloc(p, bytes));
        ACUDA(cudaMemset(*p, 0, bytes));
    }
    ACUDA(cudaMalloc(&a->st_scalar, n_streams * 2 * sizeof(float)));
    ACUDA(cudaMemset(a->st_scalar, 0, n_streams * 2 * sizeof(float)));
    ACUDA(cudaFuncSetAttribute(analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)analysis_smem_bytes((int)nb)));
    *out = a.release();
    return PVQT_OK;
}

void pvqt_analysis_destroy(pvqt_analysis *a)
{
    if (!a) return;
    cudaSetDevice(a->device);
    if (a->stream) { cudaStreamSynchronize(a->stream); cudaStreamDestroy(a->stream); }
    for (float *p : {a->st_smoothed, a->st_calm, a->st_released, a->st_afterglow, a->st_scalar}) cudaFree(p);
    for (void *p : a->mirror) if (p) cudaFree(p);
    delete a;
}

size_t pvqt_analysis_n_buckets(const pvqt_analysis *a) { return a ? a->nb : 0; }
size_t pvqt_analysis_n_streams(const pvqt_analysis *a) { return a ? a->n_streams : 0; }

int pvqt_analysis_update_vqt_smoothing_duration(pvqt_analysis *a, int has_duration, uint64_t duration_ns)
{
    if (!a) return afail(PVQT_INVALID_ARGUMENT, "null handle");
    a->params.vqt_smoothing_duration_base_ns = has_duration ? duration_ns : 0;  // analysis.rs:253
    a->has_horizon = has_duration ? 1 : 0;                                       // analysis.rs:257-268
    return PVQT_OK;
}

int pvqt_analysis_preprocess_device(pvqt_analysis *a, const float *d_db, size_t n_buckets, size_t n_frames,
                                    uint64_t frame_time_ns, const pvqt_analysis_outputs *d_out, void *cuda_stream)
{
    if (!a || !d_db) return afail(PVQT_INVALID_ARGUMENT, "null argument");
    if (n_buckets != a->nb) return afail(PVQT_BAD_LENGTH, "x_vqt.len() must equal range.n_buckets()");  // analysis.rs:289
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : a->stream;
    return pvqt_detail::analysis_run_device(a, d_db, 0, a->n_streams, n_frames, frame_time_ns, d_out, 0, st);
}

}  // extern "C"

int pvqt_detail::analysis_device(const pvqt_analysis *a) { return a ? a->device : -1; }

int pvqt_detail::analysis_run_device(pvqt_analysis *a, const float *d_db, size_t first_stream, size_t n_streams,
                                     size_t n_frames, uint64_t frame_time_ns, const pvqt_analysis_outputs *d_out,
                                     size_t out_frame_offset, cudaStream_t st)
{
    if (!a || !d_db) return afail(PVQT_INVALID_ARGUMENT, "null argument");
    if (first_stream + n_streams > a->n_streams) return afail(PVQT_INVALID_ARGUMENT, "stream range out of bounds");
    if (n_frames == 0 || n_streams == 0) return PVQT_OK;
    if (n_frames > 0xffffffffull) return afail(PVQT_INVALID_ARGUMENT, "too many frames");
    ACUDA(cudaSetDevice(a->device));
    AnalysisKernelParams P{};
    P.prm = a->params;
    P.min_freq = a->range.min_freq;
    P.octaves = (int32_t)a->range.octaves;
    P.bpo = (int32_t)a->range.buckets_per_octave;
    P.nb = (int32_t)a->nb;
    P.has_horizon = a->has_horizon;
    const size_t so = first_stream * a->nb;
    P.st_smoothed = a->st_smoothed + so; P.st_calm = a->st_calm + so; P.st_released = a->st_released + so;
    P.st_afterglow = a->st_afterglow + so; P.st_scalar = a->st_scalar + 2 * first_stream;
    P.db = d_db;
    P.n_frames = (uint32_t)n_frames;
    P.frame_time_ns = frame_time_ns;
    if (d_out) {
        P.out = *d_out;
        for (int i = 0; i < pvqt_analysis::kOutputs && out_frame_offset != 0; ++i)
            if (out_member(P.out, i))
                out_member(P.out, i) = static_cast<char *>(out_member(P.out, i)) +
                                       out_frame_offset * out_bytes_per_frame(i, a->nb, d_out->max_peaks);
    }
    analysis_kernel<<<(unsigned)n_streams, kThreads, analysis_smem_bytes((int)a->nb), st>>>(P);
    ACUDA(cudaGetLastError());
    return PVQT_OK;
}

int pvqt_detail::analysis_outputs_reserve(pvqt_analysis *a, const pvqt_analysis_outputs *host, size_t frames,
                                          pvqt_analysis_outputs *dev, cudaStream_t stream)
{
    *dev = pvqt_analysis_outputs{};
    if (!host) return PVQT_OK;
    ACUDA(cudaSetDevice(a->device));
    dev->max_peaks = host->max_peaks;
    pvqt_analysis_outputs h = *host;
    for (int i = 0; i < pvqt_analysis::kOutputs; ++i) {
        if (!out_member(h, i)) continue;
        const size_t bytes = std::max<size_t>(frames * out_bytes_per_frame(i, a->nb, host->max_peaks), 16);
        if (bytes > a->mirror_bytes[i]) {
            if (a->mirror[i]) cudaFree(a->mirror[i]);
            a->mirror[i] = nullptr;
            a->mirror_bytes[i] = 0;
            ACUDA(cudaMalloc(&a->mirror[i], bytes));
            a->mirror_bytes[i] = bytes;
        }
        // slots past a frame's peak_count stay 0
        if (i == 1 || i == 2) ACUDA(cudaMemsetAsync(a->mirror[i], 0, bytes, stream));
        out_member(*dev, i) = a->mirror[i];
    }
    return PVQT_OK;
}

int pvqt_detail::analysis_outputs_download(pvqt_analysis *a, const pvqt_analysis_outputs *host,
                                           const pvqt_analysis_outputs *dev, size_t frames, cudaStream_t stream,
                                           size_t *bytes_out)
{
    size_t total = 0;
    if (host) {
        pvqt_analysis_outputs h = *host, d = *dev;
        for (int i = 0; i < pvqt_analysis::kOutputs; ++i) {
            if (!out_member(h, i) || !out_member(d, i)) continue;
            const size_t bytes = frames * out_bytes_per_frame(i, a->nb, host->max_peaks);
            ACUDA(cudaMemcpyAsync(out_member(h, i), out_member(d, i), bytes, cudaMemcpyDeviceToHost, stream));
            total += bytes;
        }
    }
    if (bytes_out) *bytes_out = total;
    return PVQT_OK;
}

extern "C" {

int pvqt_analysis_synchronize(pvqt_analysis *a)
{
    if (!a) return afail(PVQT_INVALID_ARGUMENT, "null handle");
    ACUDA(cudaSetDevice(a->device));
    ACUDA(cudaStreamSynchronize(a->stream));
    return PVQT_OK;
}

int pvqt_analysis_preprocess_batch(pvqt_analysis *a, const float *db, size_t n_buckets, size_t n_frames,
                                   uint64_t frame_time_ns, const pvqt_analysis_outputs *out)
{
    if (!a || !db) return afail(PVQT_INVALID_ARGUMENT, "null argument");
    if (n_buckets != a->nb) return afail(PVQT_BAD_LENGTH, "x_vqt.len() must equal range.n_buckets()");
    if (n_frames == 0) return PVQT_OK;
    ACUDA(cudaSetDevice(a->device));
    const size_t S = a->n_streams, T = n_frames, NB = a->nb, P = out ? out->max_peaks : 0;
    std::vector<void *> owned;
    auto dalloc = [&](size_t bytes, void **p) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
        if (e == cudaSuccess) owned.push_back(*p);
        return e;
    };
    auto cleanup = [&]() { for (void *p : owned) cudaFree(p); };
    float *d_db = nullptr;
    cudaError_t e = dalloc(S * T * NB * sizeof(float), (void **)&d_db);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_db, db, S * T * NB * sizeof(float), cudaMemcpyHostToDevice, a->stream);
    pvqt_analysis_outputs d{};
    struct Copy { void *host; void *dev; size_t bytes; };
    std::vector<Copy> copies;
    auto want = [&](void *host, size_t bytes, void **dev) {
        if (!host || e != cudaSuccess) return;
        e = dalloc(bytes, dev);
        if (e == cudaSuccess) e = cudaMemsetAsync(*dev, 0, bytes, a->stream);  // slots past peak_count stay 0
        if (e == cudaSuccess) copies.push_back({host, *dev, bytes});
    };
    if (out) {
        d.max_peaks = out->max_peaks;
        want(out->peak_count, S * T * sizeof(uint32_t), (void **)&d.peak_count);
        want(out->peak_indices, S * T * P * sizeof(uint32_t), (void **)&d.peak_indices);
        want(out->peaks_continuous, S * T * P * sizeof(pvqt_continuous_peak), (void **)&d.peaks_continuous);
        want(out->x_vqt_smoothed, S * T * NB * sizeof(float), (void **)&d.x_vqt_smoothed);
        want(out->x_vqt_peakfiltered, S * T * NB * sizeof(float), (void **)&d.x_vqt_peakfiltered);
        want(out->x_vqt_afterglow, S * T * NB * sizeof(float), (void **)&d.x_vqt_afterglow);
        want(out->calmness, S * T * NB * sizeof(float), (void **)&d.calmness);
        want(out->pitch_accuracy, S * T * NB * sizeof(float), (void **)&d.pitch_accuracy);
        want(out->pitch_deviation, S * T * NB * sizeof(float), (void **)&d.pitch_deviation);
        want(out->smoothed_scene_calmness, S * T * sizeof(float), (void **)&d.smoothed_scene_calmness);
        want(out->smoothed_tuning_grid_inaccuracy, S * T * sizeof(float), (void **)&d.smoothed_tuning_grid_inaccuracy);
    }
    if (e != cudaSuccess) { cleanup(); return acuda(e, "allocate analysis staging"); }
    int rc = pvqt_analysis_preprocess_device(a, d_db, n_buckets, n_frames, frame_time_ns, &d, nullptr);
    if (rc != PVQT_OK) { cleanup(); return rc; }
    for (const Copy &c : copies)
        if ((e = cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, a->stream)) != cudaSuccess) break;
    if (e == cudaSuccess) e = cudaStreamSynchronize(a->stream);
    cleanup();
    if (e != cudaSuccess) return acuda(e, "analysis batch");
    return PVQT_OK;
}

}  // extern "C"

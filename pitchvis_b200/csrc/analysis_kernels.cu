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
#include <atomic>
#include <cmath>
#include <cstdlib>
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

constexpr int kThreads = 384;   // 12 warps: three search groups of four
constexpr int kSmallThreads = 128;   // the many-stream form (see the launch)
constexpr int kMaxPeaksSmem = 256;  // peaks of one frame kept in shared memory (588 bins, distance 3 -> <= 196)

struct AnalysisKernelParams {
    pvqt_analysis_params prm;
    float    min_freq;
    int32_t  octaves, bpo, nb;
    int32_t  has_horizon;        // x_vqt_smoothed time horizon is Some(..)
    float   *st_smoothed, *st_calm, *st_released, *st_afterglow, *st_scalar;  // per-stream state
    uint64_t *st_ema_ms;         // per stream and bin: the millisecond horizon st_ema_alpha was evaluated for (~0: none)
    float   *st_ema_alpha;
    const float *logf_table;     // ln of every bin's centre frequency (computed once per handle)
    const float *db;             // [S][T][NB]
    uint32_t n_frames;
    uint64_t frame_time_ns;
    pvqt_analysis_outputs out;   // device pointers
};

#ifdef PVQT_ANALYSIS_STATS
// Diagnostic build only (scripts/analysis_stats.py): cycles thread 0 of CTA 0 spends per phase of a frame, summed over the frames.
__device__ unsigned long long g_analysis_stats[16];
#define A_T(var) const long long var = clock64()
#define A_ACC(slot, t0, t1) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_analysis_stats[slot] += (unsigned long long)((t1) - (t0)); } while (0)
#else
#define A_T(var) do { } while (0)
#define A_ACC(slot, t0, t1) do { } while (0)
#endif

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
// the same for a horizon of whole milliseconds (Duration::from_millis(ms).as_secs_f32()): ms / 1000 s + (ms % 1000) * 1e6 ns,
// without the 64-bit division by 1e9 (the EMA of the spectrum re-evaluates alpha whenever a bin's horizon changes)
__device__ __forceinline__ float ema_alpha_ms(float timestep_s, uint64_t horizon_ms)
{
    if (horizon_ms > 0xffffffffull) return ema_alpha(timestep_s, horizon_ms * 1000000ull);
    const uint32_t ms = (uint32_t)horizon_ms, secs = ms / 1000u, sub_ns = (ms - secs * 1000u) * 1000000u;
    return 1.0f - cr_expf(-2.0f * timestep_s / ((float)secs + (float)sub_ns / 1000000000.0f));
}
__device__ __forceinline__ float ema_step(float y, float alpha, float x) { return y + alpha * (x - y); }  // util.rs:124

// ---- cooperative peak search (find_peaks wrapper, peak_detection.rs:26-51) -------------------------
enum : unsigned char { kNone = 0, kUndecided = 1, kKept = 2, kRemoved = 3 };

// Barrier among the `count` threads (a multiple of 32) of one search group; id 1..15.
__device__ __forceinline__ void group_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ bool group_sync_or(int id, int count, bool pred)
{
    unsigned r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %3, 0;\n\tbar.red.or.pred q, %1, %2, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(r) : "r"(id), "r"(count), "r"((unsigned)pred) : "memory");
    return r != 0;
}

// st[b] = 1 for every peak of x[0..n), else 0.  Called by the gsz threads (whole warps, gtid = 0..gsz-1) of one search
// group, which synchronise on named barrier `bar`: the three searches of a frame run side by side on three groups of
// warps -- each search is a chain of short, latency-bound phases, so a third of the CTA finishes one nearly as fast as
// the whole CTA did (10 k cycles each, one after the other, before: profiles/r02_f_analysis_phase_cycles.txt).
__device__ __forceinline__ void find_peaks_group(const float *x, int n, float min_prominence, float min_height, int distance,
                                 int min_bin, unsigned char *st, int gtid, int gsz, int bar, float *warp_min)
{
    // A thread owns bins gtid, gtid + gsz, ... (at most 32: n <= 4096, gsz >= 128) and keeps their state in two
    // register bit masks besides st[], which only the neighbours read.
    A_T(p0);
    // strict local maxima, plateaus at their middle, then min_height (peak_detection.rs / find_peaks: the scan marks the
    // middle (start + end) / 2 of a plateau whose two neighbours are lower; here every bin asks whether it is that
    // middle, so a thread writes its own bins only and no clearing pass is needed).  Also the minimum of x: a peak lower
    // than min_prominence above it cannot pass, whatever its surroundings.
    unsigned und = 0;
    float lo_x = CUDART_INF_F;
    for (int k0 = 0; gtid + k0 * gsz < n; k0 += 8) {       // eight of the thread's bins at a time: their loads issue together
        float vm[8], v0[8], vp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = gtid + (k0 + u) * gsz;
            const bool in = c < n;
            v0[u] = in ? x[c] : 0.0f;
            vm[u] = in && c >= 1 ? x[c - 1] : 0.0f;
            vp[u] = in && c <= n - 2 ? x[c + 1] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = gtid + (k0 + u) * gsz;
            if (c < n) {
                const float v = v0[u];
                lo_x = fminf(lo_x, v);
                bool peak = false;
                if (c >= 1 && c <= n - 2 && v >= min_height) {
                    if (vm[u] < v && vp[u] < v) {
                        peak = true;                            // the usual case: no plateau
                    } else if (vm[u] == v || vp[u] == v) {
                        int l = c, r = c;
                        while (l >= 1 && x[l - 1] == v) --l;
                        while (r <= n - 2 && x[r + 1] == v) ++r;
                        peak = l >= 1 && r <= n - 2 && x[l - 1] < v && x[r + 1] < v && ((l + r) >> 1) == c;
                    }
                }
                st[c] = peak ? kUndecided : kNone;
                und |= (peak ? 1u : 0u) << (k0 + u);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
    if ((gtid & 31) == 0) warp_min[gtid >> 5] = lo_x;
    group_sync(bar, gsz);
    float x_min = warp_min[0];
    for (int w = 1; w < (gsz >> 5); ++w) x_min = fminf(x_min, warp_min[w]);
    A_T(p1);
    A_ACC(8, p0, p1);
    // min_distance: taller peaks win (greedy by height == repeated "local champion" rounds)
    unsigned kept = 0;
    if (distance > 1) {
        bool pending;
        do {
            for (unsigned m = und; m; m &= m - 1) {
                const int k = __ffs(m) - 1, b = gtid + k * gsz;
                bool has_kept = false, top = true;
                const int lo = max(b - distance + 1, 0), hi = min(b + distance - 1, n - 1);
                const float xb = x[b];
                for (int j = lo; j <= hi; ++j) {
                    if (j == b) continue;
                    const unsigned char s = st[j];
                    if (s == kKept) has_kept = true;
                    else if (s == kUndecided) { const float xj = x[j]; if (xj > xb || (xj == xb && j > b)) top = false; }
                }
                if (!has_kept && top) { st[b] = kKept; kept |= 1u << k; und &= ~(1u << k); }
            }
            group_sync(bar, gsz);
            for (unsigned m = und; m; m &= m - 1) {
                const int k = __ffs(m) - 1, b = gtid + k * gsz;
                bool has_kept = false;
                const int lo = max(b - distance + 1, 0), hi = min(b + distance - 1, n - 1);
                for (int j = lo; j <= hi; ++j) has_kept |= (j != b && st[j] == kKept);
                if (has_kept) { st[b] = kRemoved; und &= ~(1u << k); }
            }
            pending = group_sync_or(bar, gsz, und != 0);
        } while (pending);
    } else {
        kept = und;
        group_sync(bar, gsz);
    }
    A_T(p2);
    A_ACC(9, p1, p2);
    // min_prominence, then drop the lowest half semitone.  A peak's prominence is its height over the higher of its two
    // bases, a base being the minimum of the stretch to that side over which nothing is higher than the peak.  Only the
    // comparison with min_prominence is needed, and f32 subtraction is monotonic, so
    //     h - max(base_l, base_r) >= P   <=>   on EACH side some bin with h - x[i] >= P comes before the first x[i] > h
    // and the walk of a side ends at the first bin that decides it -- a few bins for a narrow spectral peak, however
    // tall (walking the whole stretch of the tallest peaks, 32 bins at a time by a warp, cost 3.6-8.5 k cycles per
    // search).  One thread per kept bin, its own two short walks.
    unsigned pass = 0;
    for (unsigned m = kept; m; m &= m - 1) {
        const int k = __ffs(m) - 1, b = gtid + k * gsz;
        const float h = x[b];
        // every bin is >= x_min and f32 subtraction is monotonic: h - x[i] <= h - x_min < P on both sides
        bool ok = b >= min_bin && h - x_min >= min_prominence;
#pragma unroll
        for (int side = 0; side < 2 && ok; ++side) {
            bool deep = false;
            for (int i = side ? b + 1 : b - 1; side ? i < n : i >= 0; side ? ++i : --i) {
                const float v = x[i];
                if (!(v <= h)) break;                       // a higher bin (or NaN) ends the stretch
                if (h - v >= min_prominence) { deep = true; break; }
            }
            // (the peak itself belongs to the stretch: h - h >= P only for P <= 0)
            ok = deep || (h - h >= min_prominence);
        }
        pass |= (ok ? 1u : 0u) << k;
    }
    {
        int k = 0;
        for (int b = gtid; b < n; b += gsz, ++k) st[b] = (pass >> k) & 1u;
    }
    A_T(p3);
    A_ACC(10, p2, p3);
}

// Two register budgets: MIN_CTAS = 1 (167 registers, nothing spilled) for a few streams, where a frame's latency is the
// job's time; MIN_CTAS = 2 (80 registers, ~400 bytes spilled, two CTAs per SM) when there are more streams than SMs and the
// second resident CTA is worth more than the spills cost (one 60 s stream 38 against 42 ms; 1024 streams 5.1 against 5.4 M frames/s).
template <int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) analysis_kernel(const __grid_constant__ AnalysisKernelParams P)
{
    constexpr int kThreads = THREADS;   // (shadows the default CTA size below)
    extern __shared__ __align__(16) unsigned char a_smem[];
    const int n = P.nb, tid = threadIdx.x, stream = blockIdx.x;
    float *xraw = reinterpret_cast<float *>(a_smem);
    float *sm = xraw + n, *calm = sm + n, *released = calm + n, *aglow = released + n, *termc = aglow + n,
          *termw = termc + n, *pacc = termw + n, *pdev = pacc + n, *logf_bin = pdev + n;
    float2 *pk_cont = reinterpret_cast<float2 *>(logf_bin + n);
    float *pk_power = reinterpret_cast<float *>(pk_cont + kMaxPeaksSmem);
    float *pk_dev = pk_power + kMaxPeaksSmem;                 // deviation from the semitone grid, its |.| * power, the rounded bin
    float *pk_inacc = pk_dev + kMaxPeaksSmem;
    int *pk_bin = reinterpret_cast<int *>(pk_inacc + kMaxPeaksSmem);
    int *pk_idx = pk_bin + kMaxPeaksSmem;
    int *scan = pk_idx + kMaxPeaksSmem;                       // kThreads + 1
    unsigned char *st_bass = reinterpret_cast<unsigned char *>(scan + kThreads + 2);
    unsigned char *st_gen = st_bass + n, *st_raw = st_gen + n;
    __shared__ float s_scene, s_tuning, s_wc, s_ws;
    __shared__ float s_warp_min[kThreads / 32];
    __shared__ int s_npeaks;

    const size_t so = (size_t)stream * n;
    for (int b = tid; b < n; b += kThreads) {
        sm[b] = P.st_smoothed[so + b];
        calm[b] = P.st_calm[so + b];
        released[b] = P.st_released[so + b];
        aglow[b] = P.st_afterglow[so + b];
    }
    // ln of every bin's centre frequency (peak_detection.rs:79-85 evaluates it per peak and frame; it depends on the bin only)
    for (int b = tid; b < n; b += kThreads) logf_bin[b] = P.logf_table[b];
    if (tid == 0) { s_scene = P.st_scalar[2 * stream]; s_tuning = P.st_scalar[2 * stream + 1]; }
    __syncthreads();

    const pvqt_analysis_params &prm = P.prm;
    const float ft = dur_as_secs_f32(P.frame_time_ns);
    const float alpha_calm = ema_alpha(ft, prm.note_calmness_smoothing_duration_ns);
    const float alpha_scene = ema_alpha(ft, prm.scene_calmness_smoothing_duration_ns);
    const float alpha_tuning = ema_alpha(ft, prm.tuning_inaccuracy_smoothing_duration_ns);
    const uint64_t base_ms = prm.vqt_smoothing_duration_base_ns / 1000000ull;   // as_millis()
    const float bpo_f = (float)P.bpo;
    const float log2_min_freq = cr_log2f(P.min_freq);
    const int distance = (int)f32_as_u64(roundf(bpo_f * 0.4f / 12.0f));         // peak_detection.rs:37
    const int min_bin = (P.bpo / 12 + 1) / 2;                                   // peak_detection.rs:45
    const int radius = P.bpo / 12 / 3;                                          // calmness.rs:36
    const int hb = prm.highest_bassnote > 0x7fffffffull ? 0x7fffffff : (int)prm.highest_bassnote;
    const int chunk = (n + kThreads - 1) / kThreads;

    // alpha of a bin's EMA changes only when its horizon, truncated to whole milliseconds, does: a thread keeps the last
    // (milliseconds, alpha) of its bins in registers and evaluates the f64 exp only on a change (scene calmness drifts
    // slowly: a few bins per frame instead of all of them)
    constexpr int kBinsPerThread = 8;                     // n <= kBinsPerThread * kThreads uses the cache
    uint64_t ema_ms[kBinsPerThread];
    float ema_a[kBinsPerThread];
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {   // the cache outlives the launch (one frame per call: 588 f64 exps per call otherwise)
        const int b = tid + k * kThreads;
        ema_ms[k] = b < n ? P.st_ema_ms[so + b] : ~0ull;
        ema_a[k] = b < n ? P.st_ema_alpha[so + b] : 0.0f;
    }
    const bool ema_cached = n <= kBinsPerThread * kThreads;
    float base_fm[kBinsPerThread];   // base horizon x the bin's frequency multiplier (analysis.rs:309-318), frame-independent
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
        const float octave_fraction = (float)(tid + k * kThreads) / bpo_f / (float)P.octaves;
        const float frequency_multiplier = 1.5f - 0.5f * octave_fraction;
        base_fm[k] = (float)base_ms * frequency_multiplier;
    }
    float x_next[kBinsPerThread];
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
        const int b = tid + k * kThreads;
        x_next[k] = (b < n && P.n_frames > 0) ? P.db[(size_t)stream * P.n_frames * n + b] : 0.0f;
    }

    for (uint32_t t = 0; t < P.n_frames; ++t) {
        const size_t fo = ((size_t)stream * P.n_frames + t);
        const float *x = P.db + fo * n;
        A_T(a0);
        // ---- calmness-adaptive EMA of the dB spectrum (analysis.rs:295-329) -----------------------
        const float calmness_multiplier =
            prm.vqt_smoothing_calmness_min + (prm.vqt_smoothing_calmness_max - prm.vqt_smoothing_calmness_min) * s_scene;
#pragma unroll
        for (int k = 0; k < kBinsPerThread; ++k) {
            const int b = tid + k * kThreads;
            if (b >= n) break;
            const float xv = x_next[k];          // loaded one frame ahead (below): an L2 round trip off the frame's critical path
            if (t + 1 < P.n_frames) x_next[k] = x[(size_t)n + b];
            xraw[b] = xv;
            float y = sm[b];
            if (!P.has_horizon) {
                y = xv;                                                           // util.rs:117-120
            } else {
                uint64_t horizon_ms = 0;                                          // base 0 ms: horizon stays 0 ms
                if (base_ms > 0) horizon_ms = f32_as_u64(base_fm[k] * calmness_multiplier);   // from_millis(x as u64)
                if (horizon_ms != ema_ms[k]) {
                    ema_ms[k] = horizon_ms;
                    ema_a[k] = ema_alpha_ms(ft, horizon_ms);
                }
                y = ema_step(y, ema_a[k], xv);
            }
            sm[b] = y;
        }
        if (!ema_cached)   // more bins than the cache covers: the rest without it
            for (int b = tid + kBinsPerThread * kThreads; b < n; b += kThreads) {
                const float xv = x[b];
                xraw[b] = xv;
                float y = xv;
                if (P.has_horizon) {
                    uint64_t horizon_ms = 0;
                    if (base_ms > 0) {
                        const float octave_fraction = (float)b / bpo_f / (float)P.octaves;
                        const float frequency_multiplier = 1.5f - 0.5f * octave_fraction;
                        horizon_ms = f32_as_u64((float)base_ms * frequency_multiplier * calmness_multiplier);
                    }
                    y = ema_step(sm[b], ema_alpha(ft, horizon_ms * 1000000ull), xv);
                }
                sm[b] = y;
            }
        __syncthreads();
        A_T(a1);
        A_ACC(0, a0, a1);

        // ---- peaks: bass config up to highest_bassnote, general config above (analysis.rs:332-349) --
        // three searches side by side, a third of the CTA's warps each: the bass configuration, the general
        // configuration, and the unsmoothed spectrum for the calmness update (calmness.rs:39)
        if constexpr (kThreads >= 384) {
            constexpr int kGroup = kThreads / 3;
            static_assert(kThreads % 96 == 0, "three search groups of whole warps");
            const int group = tid / kGroup, gtid = tid - group * kGroup;
            if (group == 0)
                find_peaks_group(sm, n, prm.bassline_peak_config.min_prominence, prm.bassline_peak_config.min_height, distance,
                                 min_bin, st_bass, gtid, kGroup, 1, s_warp_min);
            else if (group == 1)
                find_peaks_group(sm, n, prm.peak_config.min_prominence, prm.peak_config.min_height, distance, min_bin, st_gen,
                                 gtid, kGroup, 2, s_warp_min + kGroup / 32);
            else
                find_peaks_group(xraw, n, prm.peak_config.min_prominence, prm.peak_config.min_height, distance, min_bin, st_raw,
                                 gtid, kGroup, 3, s_warp_min + 2 * (kGroup / 32));
        } else {
            // the small CTA of the many-stream form: one group, the three searches one after the other -- through ONE
            // copy of the search code (five CTAs per SM in different phases: 31 % of the stall samples were instruction
            // fetches with three inlined copies)
#pragma unroll 1
            for (int q = 0; q < 3; ++q) {
                const float *xs = q == 2 ? xraw : sm;
                unsigned char *sts = q == 0 ? st_bass : (q == 1 ? st_gen : st_raw);
                const float prom = q == 0 ? prm.bassline_peak_config.min_prominence : prm.peak_config.min_prominence;
                const float height = q == 0 ? prm.bassline_peak_config.min_height : prm.peak_config.min_height;
                find_peaks_group(xs, n, prom, height, distance, min_bin, sts, tid, kThreads, 1, s_warp_min);
            }
        }
        __syncthreads();

        A_T(a2);
        A_ACC(1, a1, a2);
        // (this per-bin pass needs the unsmoothed peaks only; it runs before the compaction so that the two long
        // sequential sums over its terms can run beside the per-peak work below)
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
        A_T(a2b);
        A_ACC(4, a2, a2b);
        // ordered compaction of the peak set (its barriers also publish the terms above)
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
        A_T(a3);
        A_ACC(2, a2b, a3);

        // ---- the reference's sequential f32 sums, in its order (calmness.rs:49-90, pitch_analysis.rs:54-74)
        // Four single threads on four warps: a lone warp issues an instruction every other cycle at best, so the two
        // 588-term sums of calmness.rs:49-90 (13 cycles per bin when one thread carried both chains, their loads and the
        // loop) go to a thread each, with nothing but a 16-byte load per four dependent adds in the loop.
        auto seq_sum = [&](const float *terms) {
            float acc = 0.0f;
            int b = 0;
            if ((n & 3) == 0) {
                const float4 *p4 = reinterpret_cast<const float4 *>(terms);
                const int n4 = n >> 2;
                float4 cur = p4[0];
                for (int q = 0; q < n4; ++q) {
                    const float4 nxt = p4[min(q + 1, n4 - 1)];
                    acc += cur.x;
                    acc += cur.y;
                    acc += cur.z;
                    acc += cur.w;
                    cur = nxt;
                }
                b = n;
            }
            for (; b < n; ++b) acc += terms[b];
            return acc;
        };
        // two threads on two warps carry the two sums beside the per-peak work: 256 and 288 in the large CTA, which own no
        // peak (at most 256 are stored); the last two warps' first threads in the small one (they refine their peaks after)
        constexpr int kSumA = kThreads >= 384 ? 256 : kThreads - 64, kSumB = kThreads >= 384 ? 288 : kThreads - 32;
        static_assert(kMaxPeaksSmem <= 256 && kSumA >= 64, "see above");
        if (tid == kSumA) s_wc = seq_sum(termc);
        else if (tid == kSumB) s_ws = seq_sum(termw);

        // ---- enhance_peaks_continuous + promote_bass_peaks_with_harmonics (per peak) -----------------
        for (int i = tid; i < n_stored; i += kThreads) {
            const int p = pk_idx[i];
            float center = (float)p, size = sm[p];
            if (p >= 1 && p <= n - 2) {                                           // peak_detection.rs:71-77
                const float l0 = logf_bin[p - 1], l1 = logf_bin[p], l2 = logf_bin[p + 1];
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
                    const float harmonic_bin = (cr_log2f(harmonic_freq) - log2_min_freq) * bpo_f;
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
            const float power = cr_powf(10.0f, size / 10.0f);                        // pitch_analysis.rs:58
            pk_power[i] = power;
            // the per-peak factors of the two sequential passes below (pitch_analysis.rs:24-41, :54-74), computed here in parallel
            const float cs = center * 12.0f / bpo_f;
            const float deviation = cs - roundf(cs);
            pk_dev[i] = deviation;
            pk_inacc[i] = fabsf(deviation) * power;
            pk_bin[i] = (int)f32_as_u64(roundf(center));
        }

        __syncthreads();
        A_T(a4);
        A_ACC(3, a3, a4);

        if (tid == 32) {
            float inaccuracy_sum = 0.0f, power_sum = 0.0f;
            for (int i = 0; i < n_stored; ++i) {
                power_sum += pk_power[i];
                inaccuracy_sum += pk_inacc[i];
            }
            const float avg = power_sum > 0.0f ? inaccuracy_sum / power_sum : 0.0f;
            s_tuning = ema_step(s_tuning, alpha_tuning, 100.0f * avg);
        } else if (tid == 64) {
            for (int i = 0; i < n_stored; ++i) {                                  // pitch_analysis.rs:24-41, in peak order
                const float deviation = pk_dev[i];
                const int bin = pk_bin[i];
                if (bin < n) { pacc[bin] = fmaxf(1.0f - 2.0f * fabsf(deviation), 0.0f); pdev[bin] = deviation; }
            }
        }
        __syncthreads();
        A_T(a6);
        A_ACC(5, a4, a6);

        // ---- per-frame outputs ------------------------------------------------------------------------
        const pvqt_analysis_outputs &O = P.out;
        if (tid == 0) {
            if (s_ws > 0.0f) s_scene = ema_step(s_scene, alpha_scene, s_wc / s_ws);    // calmness.rs:88-91
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
        A_T(a7);
        A_ACC(6, a6, a7);
        A_ACC(7, a0, a7);
    }

    for (int b = tid; b < n; b += kThreads) {
        P.st_smoothed[so + b] = sm[b];
        P.st_calm[so + b] = calm[b];
        P.st_released[so + b] = released[b];
        P.st_afterglow[so + b] = aglow[b];
    }
    if (tid == 0) { P.st_scalar[2 * stream] = s_scene; P.st_scalar[2 * stream + 1] = s_tuning; }
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
        const int b = tid + k * kThreads;
        if (b < n) { P.st_ema_ms[so + b] = ema_ms[k]; P.st_ema_alpha[so + b] = ema_a[k]; }
    }
}

// ln(min_freq * 2^(b / bpo)) for every bin, with the kernel's own correctly rounded functions (peak_detection.rs:79-85)
__global__ void analysis_tables_kernel(float min_freq, int bpo, int nb, float *logf_table)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) logf_table[b] = cr_logf(min_freq * cr_powf(2.0f, (float)b / (float)bpo));
}

size_t analysis_smem_bytes(int nb)
{
    return sizeof(float) * 10 * (size_t)nb + sizeof(float2) * kMaxPeaksSmem + sizeof(float) * 4 * kMaxPeaksSmem +
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
    uint64_t *st_ema_ms = nullptr;      // alpha cache of the spectrum's EMA, valid for frame time `ema_frame_time_ns`
    float *st_ema_alpha = nullptr, *logf_table = nullptr;
    uint64_t ema_frame_time_ns = 0;
    // device mirrors of the result buffers of the single-call pipeline (pvqt_calc_*_analysis), reused across calls
    static constexpr int kOutputs = 11;
    void *mirror[kOutputs] = {};
    size_t mirror_bytes[kOutputs] = {};
    uint64_t generation = next_generation_base();   // unique per handle (upper bits) and bumped when the parameters or a
                                                    // mirror change: captured per-frame launches key on it
    static uint64_t next_generation_base()
    {
        static std::atomic<uint64_t> handles{0};
        return (handles.fetch_add(1) + 1) << 32;
    }
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
    ACUDA(cudaMalloc(&a->st_ema_ms, n_streams * nb * sizeof(uint64_t)));
    ACUDA(cudaMemset(a->st_ema_ms, 0xff, n_streams * nb * sizeof(uint64_t)));
    ACUDA(cudaMalloc(&a->st_ema_alpha, bytes));
    ACUDA(cudaMemset(a->st_ema_alpha, 0, bytes));
    ACUDA(cudaMalloc(&a->logf_table, nb * sizeof(float)));
    analysis_tables_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, a->stream>>>(range->min_freq, (int)range->buckets_per_octave, (int)nb,
                                                                              a->logf_table);
    ACUDA(cudaGetLastError());
    ACUDA(cudaStreamSynchronize(a->stream));
    for (auto k : {(const void *)analysis_kernel<kThreads, 1>, (const void *)analysis_kernel<kThreads, 2>,
                   (const void *)analysis_kernel<kSmallThreads, 5>})
        ACUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)analysis_smem_bytes((int)nb)));
    *out = a.release();
    return PVQT_OK;
}

void pvqt_analysis_destroy(pvqt_analysis *a)
{
    if (!a) return;
    cudaSetDevice(a->device);
    if (a->stream) { cudaStreamSynchronize(a->stream); cudaStreamDestroy(a->stream); }
    for (float *p : {a->st_smoothed, a->st_calm, a->st_released, a->st_afterglow, a->st_scalar}) cudaFree(p);
    cudaFree(a->st_ema_ms);
    cudaFree(a->st_ema_alpha);
    cudaFree(a->logf_table);
    for (void *p : a->mirror) if (p) cudaFree(p);
    delete a;
}

size_t pvqt_analysis_n_buckets(const pvqt_analysis *a) { return a ? a->nb : 0; }
size_t pvqt_analysis_n_streams(const pvqt_analysis *a) { return a ? a->n_streams : 0; }

int pvqt_analysis_update_vqt_smoothing_duration(pvqt_analysis *a, int has_duration, uint64_t duration_ns)
{
    if (!a) return afail(PVQT_INVALID_ARGUMENT, "null handle");
    ++a->generation;
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
uint64_t pvqt_detail::analysis_generation(const pvqt_analysis *a) { return a ? a->generation : 0; }
void *&pvqt_detail::analysis_output_member(pvqt_analysis_outputs &o, int i) { return out_member(o, i); }
size_t pvqt_detail::analysis_output_bytes_per_frame(const pvqt_analysis *a, int i, size_t max_peaks)
{
    return out_bytes_per_frame(i, a->nb, max_peaks);
}

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
    // alpha depends on the frame time too: another frame time invalidates the whole cache (every stream's)
    if (a->ema_frame_time_ns != frame_time_ns) {
        ACUDA(cudaMemsetAsync(a->st_ema_ms, 0xff, a->n_streams * a->nb * sizeof(uint64_t), st));
        a->ema_frame_time_ns = frame_time_ns;
    }
    P.st_ema_ms = a->st_ema_ms + so;
    P.st_ema_alpha = a->st_ema_alpha + so;
    P.logf_table = a->logf_table;
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
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, a->device) != cudaSuccess) sms = 148;
    const size_t smem = analysis_smem_bytes((int)a->nb);
    // up to one stream per SM: the large CTA with all its registers (a frame's latency is the job's time); up to two:
    // the large CTA twice per SM; beyond: small CTAs, five per SM (1024 streams x 511 frames: 36.7 against 40.7 ms -- with
    // many streams the SM is no longer waiting but issuing, ~10 us of f64 transcendentals and searches per frame and SM)
    if (std::getenv("PVQT_ANALYSIS_SMALL_CTA")) analysis_kernel<kSmallThreads, 5><<<(unsigned)n_streams, kSmallThreads, smem, st>>>(P);   // (tests)
    else if ((size_t)n_streams <= (size_t)sms) analysis_kernel<kThreads, 1><<<(unsigned)n_streams, kThreads, smem, st>>>(P);
    else if ((size_t)n_streams <= 2 * (size_t)sms || std::getenv("PVQT_ANALYSIS_LARGE_CTA"))
        analysis_kernel<kThreads, 2><<<(unsigned)n_streams, kThreads, smem, st>>>(P);
    else analysis_kernel<kSmallThreads, 5><<<(unsigned)n_streams, kSmallThreads, smem, st>>>(P);
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
            ++a->generation;
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

#ifdef PVQT_ANALYSIS_STATS
extern "C" int pvqt_debug_analysis_stats(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_analysis_stats, sizeof(g_analysis_stats)) != cudaSuccess) return 7;
    if (reset) {
        static unsigned long long zeros[16] = {};
        cudaMemcpyToSymbol(g_analysis_stats, zeros, sizeof(zeros));
    }
    return 0;
}
#endif

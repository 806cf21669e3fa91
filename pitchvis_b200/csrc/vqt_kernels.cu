// vqt_kernels.cu -- sm_100a kernels of the VQT hot path.
//
//   K-fft   fft_groups_kernel    batched shared-memory Stockham real-to-complex FFT, one launch for
//                                all window groups of all frames; replaces the realfft
//                                `process_with_scratch` calls at vqt.rs:884-887.  Only the FFT bins
//                                the sparse kernel consumes are produced (output-pruned last pass).
//   K-spmm  spmm_kernel          batched complex banded SpMM, kernel coefficients stationary per
//                                64-row block, spectra staged per 8-frame tile with cp.async;
//                                includes the conjugate-part product (vqt.rs:889-910) and |z|^2.
//   K-db    power_to_db_kernel   power_to_db (vqt.rs:922-954): one warp per frame, shuffle reductions.
//
// All complex arithmetic is issued as packed f32x2 instructions (FADD2 / FMUL2 / FFMA2, new on
// sm_100): the FFT is issue-bound, and a complex add is one instruction instead of two, a complex
// multiply two instead of four (profiles/r01_a_baseline_instruction_mix.txt -> r01_b).
//
// FFT convention (pinned by vqt.rs:1087-1128): unnormalised forward transform,
// X[k] = sum_n x[n] exp(-2 pi i k n / N), half spectrum k = 0..N/2.
#include "vqt_device.cuh"
#include "device_helpers.cuh"
#include "sdft_combine.cuh"

#include <algorithm>

#include <math_constants.h>

#ifndef PVQT_FFT_MIN_BLOCKS
#define PVQT_FFT_MIN_BLOCKS 4   // resident CTAs of 256 threads the register allocation of K-fft is held to
#endif

namespace pvqt_dev {
#ifdef PVQT_PHASE_TIMERS
// Diagnostic build only (scripts/phase_timers.py): globaltimer stamps per CTA and phase.
__device__ unsigned long long g_phase_stamps[2][8192][8];
#define PVQT_STAMP(kernel, slot)                                                                          \
    do {                                                                                                  \
        if (threadIdx.x == 0 && blockIdx.x < 8192) {                                                      \
            unsigned long long t_;                                                                        \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                        \
            g_phase_stamps[kernel][blockIdx.x][slot] = t_;                                                \
        }                                                                                                 \
    } while (0)
cudaError_t read_phase_stamps(unsigned long long *out) { return cudaMemcpyFromSymbol(out, g_phase_stamps, sizeof(g_phase_stamps)); }
#else
#define PVQT_STAMP(kernel, slot) do { } while (0)
#endif
#ifdef PVQT_FFT_STATS
// Diagnostic build only (scripts/fft_stats.py): cycles thread 0 of every CTA spends per phase of its work item, summed
// per FFT size.  A stamp taken right after a barrier or a load is its ISSUE time: the wait shows up in the next phase.
__device__ unsigned long long g_fft_stats[16][16];   // [log2 N_c][phase]; phase 15 = CTAs
#define FFT_T(slot) do { if (threadIdx.x == 0) fft_t[slot] = clock64(); } while (0)
#else
#define FFT_T(slot) do { } while (0)
#endif
namespace {

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kCosPi8 = 0.92387953251128675613f;
constexpr float kSinPi8 = 0.38268343236508977173f;

// L1 policy of K-fft's two kinds of read-only loads: the audio of a work item is read once per CTA (the overlap of neighbouring
// frames is served by L2: they run on other SMs), the twiddle tables by every CTA of the SM.
__device__ __forceinline__ float2 ld_audio2(const float2 *p)
{
#if defined(PVQT_FFT_L1_HINTS)
    float2 v;
    asm volatile("ld.global.nc.L1::evict_first.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
#elif defined(PVQT_FFT_AUDIO_CG)
    return __ldcg(p);
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ float2 ld_twiddle(const float2 *p)
{
#if defined(PVQT_FFT_L1_HINTS)
    float2 v;
    asm volatile("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// multiply by W16^M = exp(-2 pi i M / 16), M compile-time
template <int M>
__device__ __forceinline__ float2 mul_w16(float2 a)
{
    if constexpr (M == 0) return a;
    else if constexpr (M == 1) return cmul(a, make_float2(kCosPi8, -kSinPi8));
    else if constexpr (M == 2)  // ((x + y) c, (y - x) c)
        return __fmul2_rn(__fadd2_rn(a, make_float2(a.y, -a.x)), make_float2(kSqrtHalf, kSqrtHalf));
    else if constexpr (M == 3) return cmul(a, make_float2(kSinPi8, -kCosPi8));
    else if constexpr (M == 4) return cmul_mi(a);
    else if constexpr (M == 6)  // ((y - x) c, -(x + y) c)
        return __fmul2_rn(__fadd2_rn(make_float2(a.y, -a.x), make_float2(-a.x, -a.y)),
                          make_float2(kSqrtHalf, kSqrtHalf));
    else if constexpr (M == 9) return cmul(a, make_float2(-kCosPi8, kSinPi8));
    else { static_assert(M < 0, "unsupported W16 power"); return a; }
}

__device__ __forceinline__ void fft2(float2 &a0, float2 &a1)
{
    const float2 t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}

// 4-point forward DFT, natural order in and out: 8 packed adds
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

// R-point forward DFT on registers v[0..R).  The result is left permuted:
// register j holds frequency out_index<R>(j).
template <int R>
__device__ __forceinline__ constexpr int out_index(int j)
{
    if constexpr (R == 16) return (j >> 2) + 4 * (j & 3);
    else if constexpr (R == 8) return (j >> 1) + 4 * (j & 1);
    else return j;
}

template <int R>
__device__ __forceinline__ void butterfly(float2 *v)
{
    if constexpr (R == 2) {
        fft2(v[0], v[1]);
    } else if constexpr (R == 4) {
        fft4(v[0], v[1], v[2], v[3]);
    } else if constexpr (R == 8) {
        // n = 2 n1 + n2, k = k1 + 4 k2
        fft4(v[0], v[2], v[4], v[6]);
        fft4(v[1], v[3], v[5], v[7]);
        v[3] = mul_w16<2>(v[3]);
        v[5] = mul_w16<4>(v[5]);
        v[7] = mul_w16<6>(v[7]);
        fft2(v[0], v[1]);
        fft2(v[2], v[3]);
        fft2(v[4], v[5]);
        fft2(v[6], v[7]);
    } else {
        static_assert(R == 16, "radix");
        // n = 4 n1 + n2, k = k1 + 4 k2
        fft4(v[0], v[4], v[8], v[12]);
        fft4(v[1], v[5], v[9], v[13]);
        fft4(v[2], v[6], v[10], v[14]);
        fft4(v[3], v[7], v[11], v[15]);
        // v[4 k1 + n2] *= W16^(n2 k1)
        v[5] = mul_w16<1>(v[5]);
        v[6] = mul_w16<2>(v[6]);
        v[7] = mul_w16<3>(v[7]);
        v[9] = mul_w16<2>(v[9]);
        v[10] = mul_w16<4>(v[10]);
        v[11] = mul_w16<6>(v[11]);
        v[13] = mul_w16<3>(v[13]);
        v[14] = mul_w16<6>(v[14]);
        v[15] = mul_w16<9>(v[15]);
        fft4(v[0], v[1], v[2], v[3]);
        fft4(v[4], v[5], v[6], v[7]);
        fft4(v[8], v[9], v[10], v[11]);
        fft4(v[12], v[13], v[14], v[15]);
    }
}

// radix of pass `pass` of the plan for N_c complex points (1 = no such pass)
__host__ __device__ constexpr int plan_radix(int nc, int pass)
{
    switch (nc) {
    case 32: return pass == 0 ? 16 : pass == 1 ? 2 : 1;
    case 64: return pass == 0 ? 16 : pass == 1 ? 4 : 1;
    case 128: return pass == 0 ? 16 : pass == 1 ? 8 : 1;
    case 256: return pass < 2 ? 16 : 1;
    case 512: return pass == 0 ? 16 : pass == 1 ? 8 : pass == 2 ? 4 : 1;
    case 1024: return pass < 2 ? 16 : pass == 2 ? 4 : 1;
    case 2048: return pass < 2 ? 16 : pass == 2 ? 8 : 1;
    case 4096: return pass < 3 ? 16 : 1;
    case 8192: return pass < 2 ? 16 : pass == 2 ? 8 : pass == 3 ? 4 : 1;
    case 16384: return pass < 3 ? 16 : pass == 3 ? 4 : 1;
    default: return 1;
    }
}

// shared-memory index padding: one float2 of padding per 16 keeps every pass conflict-free.
// pad_index(i + m) == pad_index(i) + m + m / 16 whenever m is a multiple of 16, which turns the
// per-point index arithmetic of a pass into compile-time offsets from one base pointer.
__host__ __device__ constexpr int pad_index(int i) { return i + (i >> 4); }
constexpr int kEdgeSlot = 16;   // a padding slot of every FFT buffer (N_c >= 32)

// Barrier among the T threads that share one FFT (several FFTs share a CTA when T < BLOCK).
// The barrier id must be an immediate: with `bar.sync %r, T` ptxas reserves all 16 named barriers for the CTA, and
// named barriers are an SM-wide pool -- K-fft then ran 2 CTAs per SM instead of the 4 its registers allow
// (globaltimer stamps per CTA, scripts/phase_timers.py).  With immediates a 256-thread CTA uses at most 5.
template <int T, int ID, int N>
__device__ __forceinline__ void fft_bar_case(int fid)
{
    if constexpr (ID + 1 < N) {
        if (fid == ID) asm volatile("bar.sync %0, %1;" ::"n"((ID + 1) & 15), "n"(T) : "memory");
        else fft_bar_case<T, ID + 1, N>(fid);
    } else {
        asm volatile("bar.sync %0, %1;" ::"n"((ID + 1) & 15), "n"(T) : "memory");
    }
}
template <int T, int BLOCK>
__device__ __forceinline__ void fft_sync(int fid)
{
    if constexpr (T >= BLOCK || T < 32) __syncthreads();
    else if constexpr (T == 32) __syncwarp();
#ifdef PVQT_FFT_OLD_BAR   // A/B only
    else asm volatile("bar.sync %0, %1;" ::"r"(fid + 1), "n"(T) : "memory");
#else
    else fft_bar_case<T, 0, BLOCK / T>(fid);
#endif
}

// One Stockham pass (decimation in time, autosort):
//   butterfly b (< NC/R), k = b mod NS:
//     v[r]  = in[b + r NC/R] * exp(-2 pi i r k / (NS R))
//     out[(b - k) R + k + r NS] = DFT_R(v)[r]
template <int NC, int PASS, int NS, int BLOCK>
__device__ __forceinline__ void fft_passes(float2 (&v)[kPointsPerThread], float2 *s, int t, int fid, const float *x,
                                           bool valid, const FftGroup &g, long long *fft_t, bool shifted, float2 &aux)
{
    (void)fft_t;
    constexpr int R = plan_radix(NC, PASS);
    constexpr int T = NC / kPointsPerThread;
    constexpr int NB = kPointsPerThread / R;
    constexpr int LD = NC / R;  // distance between the R inputs of a butterfly
    constexpr bool kFirst = PASS == 0;
    constexpr bool kLast = NS * R == NC;
    constexpr bool kLdConst = (LD % 16) == 0;
    constexpr bool kStConst = (NS % 16) == 0 || (NS == 1 && R == 16);

    // Last pass: butterfly b produces bins b + q NS.  The split step reads Z[c] and Z[NC - c] for
    // c in [col_lo, col_hi] only; butterflies none of whose outputs are read are skipped entirely.
    bool need[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + i * T;
        need[i] = !kLast || b <= g.col_hi || b + (R - 1) * NS >= NC - g.col_hi;
    }

#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + i * T;
        if constexpr (kFirst) {
            // z[m] = x[2m] + i x[2m+1]: the real window packed as N/2 complex points
            const float *xb = x + 2 * b;
            if ((reinterpret_cast<uintptr_t>(x) & 7) == 0) {  // even window start: one 8-byte load per point
#pragma unroll
                for (int r = 0; r < R; ++r)
                    v[i * R + r] = valid ? ld_audio2(reinterpret_cast<const float2 *>(xb) + r * LD) : make_float2(0.f, 0.f);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    v[i * R + r] = valid ? make_float2(__ldg(xb + 2 * r * LD), __ldg(xb + 2 * r * LD + 1))
                                         : make_float2(0.f, 0.f);
            }
        } else if (need[i]) {
            if constexpr (kLdConst) {
                const float2 *p = s + pad_index(b);
#pragma unroll
                for (int r = 0; r < R; ++r) v[i * R + r] = p[r * (LD + LD / 16)];
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) v[i * R + r] = s[pad_index(b + r * LD)];
            }
        }
    }
    FFT_T(1 + 3 * PASS);   // loads issued
    if constexpr (!kFirst) fft_sync<T, BLOCK>(fid);  // all reads done before the in-place overwrite

#pragma unroll
    for (int i = 0; i < NB; ++i) {
        if (!need[i]) continue;
        const int b = t + i * T;
        if constexpr (NS > 1) {
            const int k = b & (NS - 1);
            const float2 *tw = g.twiddle[PASS] + k;
#pragma unroll
            for (int r = 1; r < R; ++r) v[i * R + r] = cmul(v[i * R + r], ld_twiddle(tw + (r - 1) * NS));
        }
        butterfly<R>(&v[i * R]);

        const int k = b & (NS - 1);
        const int j0 = (b - k) * R + k;
        if constexpr (kStConst) {
            float2 *p = s + pad_index(j0);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                constexpr int dummy = 0;
                (void)dummy;
                const int q = out_index<R>(j);
                const int off = NS == 1 ? q : q * (NS + NS / 16);
                // last pass: every output of a butterfly that is needed at all is stored (two compares and a predicate
                // per output cost more than the store; butterflies with no consumed output were skipped above)
                p[off] = v[i * R + j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int o = j0 + out_index<R>(j) * NS;
                s[pad_index(o)] = v[i * R + j];
            }
        }
    }
    if constexpr (kFirst) {
        // the frame's edge term (loaded by the caller before the window, used only here: its latency hides behind the
        // first pass), parked in a padding slot of the frame's buffer for the split step
        if (shifted && t == 0) s[kEdgeSlot] = make_float2(aux.x - aux.y, 0.f);
    }
    if constexpr (kLast) {   // the split step's first twiddle, in flight across the barrier
        aux = (int)threadIdx.x <= g.col_hi - g.col_lo ? __ldg(g.split_twiddle + threadIdx.x) : make_float2(0.f, 0.f);
    }
    FFT_T(2 + 3 * PASS);   // butterflies and stores issued
    if constexpr (kLast) __syncthreads();   // the split step reads the buffers of every frame of the CTA
    else fft_sync<T, BLOCK>(fid);
    FFT_T(3 + 3 * PASS);   // barrier issued

    if constexpr (!kLast) fft_passes<NC, PASS + 1, NS * R, BLOCK>(v, s, t, fid, x, valid, g, fft_t, shifted, aux);
}

template <int NC, int BLOCK>
__device__ __forceinline__ void fft_group_body(const FftParams &P, const FftGroup &g, float2 *smem)
{
    constexpr int T = NC / kPointsPerThread;   // threads per FFT
    constexpr int FPC = BLOCK / T;             // frames per CTA
    static_assert(T >= 1 && FPC >= 1, "block too small for this FFT size");
    const int tid = threadIdx.x;
    const int fid = tid / T;
    const int t = tid - fid * T;
    float2 *s = smem + fid * pad_index(NC);
    // One wave of CTAs: CTA c of the group transforms the frames of work items c, c + n_ctas, ... (an item = FPC frames),
    // so that every SM finishes at about the same time instead of a last, mostly empty wave of one-item CTAs.
    const uint32_t n_items = (P.frames.n_frames + FPC - 1) / FPC;
#ifdef PVQT_FFT_STATS
    long long fft_t[14];
    for (auto &q : fft_t) q = 0;
    const long long t_cta = clock64();
#else
    long long *fft_t = nullptr;
#endif
#pragma unroll 1
    for (uint32_t item = blockIdx.x - g.cta_begin; item < n_items; item += (uint32_t)g.n_ctas) {
    const uint32_t local_frame = item * FPC + fid;
    const bool valid = local_frame < P.frames.n_frames;

    // frame -> (stream, frame in stream), 32-bit: frames_per_stream < 2^31 and n_frames <= 2^20 per launch
    const uint32_t ft = P.frames.first_t + (valid ? local_frame : 0);
    const uint32_t ds = ft / P.frames.frames_per_stream;
    const uint64_t stream = (uint64_t)P.frames.first_stream + ds;
    const uint64_t in_stream = ft - ds * P.frames.frames_per_stream;
    const float *x = P.frames.audio + stream * P.frames.stream_stride + in_stream * P.frames.hop + g.window_begin;
    // The nested windows are centred, so at the defaults every window starts on an odd sample and the packed loads
    // z[m] = (x[2m], x[2m+1]) would be 4-byte loads.  Transform the window moved one sample down instead (8-byte
    // aligned) and undo the shift on the consumed bins in the split step:
    //   X[k] = W^-k (X'[k] + x[w + N - 1] - x[w - 1]),   W = exp(-2 pi i / N)
    // (exact algebra: the two windows differ by one sample at either end, and W^(N k) = 1).
    // Which form a frame takes is a property of the window group alone (the parity of window_begin inside the frame),
    // never of the address: the same n_fft samples give the same bits through every entry point, at every hop, in
    // every shard.  The address only selects the load width of the first pass (8-byte loads when x - shifted is
    // 8-byte aligned -- every frame at even hops and strides -- else two 4-byte loads of the same values).
#ifndef PVQT_FFT_NO_SHIFT
    const bool shifted = (g.window_begin & 1) != 0;
#else
    const bool shifted = false;
#endif
    float2 aux = make_float2(0.f, 0.f);   // in: the two samples of the edge term; out: the split step's first twiddle
    if (shifted) {
        x -= 1;
        if (t == 0 && valid) aux = make_float2(__ldg(x + 2 * NC), __ldg(x));
    }

    float2 v[kPointsPerThread];
    FFT_T(0);
    fft_passes<NC, 0, 1, BLOCK>(v, s, t, fid, x, valid, g, fft_t, shifted, aux);   // ends with a CTA-wide barrier

    // Real-FFT split: with Z = FFT_{NC}(z), E/O the spectra of the even/odd samples,
    //   X[c] = E[c] + W_N^c O[c],  E = (Z[c] + conj Z[NC-c]) / 2,  O = -i (Z[c] - conj Z[NC-c]) / 2
    // written straight into the spectrum layout K-spmm-db stages.  The CTA's FPC frames are consecutive frames of one
    // tile, so the step runs over (column, group of VEC frames): the twiddle and the index arithmetic are shared by the
    // VEC frames and each thread stores whole 4-frame chunks (16 bytes; 8 when a CTA holds two frames) -- a quarter
    // of the instructions of one thread per (column, frame), which at the defaults were a quarter of the kernel's
    // (profiles/r02_d_fft_sass_segments.txt).  Frames past the end of the launch transformed zeros: their chunks
    // are zeros, inside the last tile of the scratch.
    {
        constexpr int VEC = FPC >= 4 ? 4 : FPC;    // frames per store
        constexpr int NV = FPC / VEC;              // frame groups per column
        constexpr int PAD = pad_index(NC);
        const uint32_t lf0 = item * FPC;
        const int n_cols = g.col_hi - g.col_lo + 1;
        const bool planes = P.plane_stride > 0;
        for (int idx = tid; idx < n_cols * NV; idx += BLOCK) {
            int vg = 0, i = idx;
            if constexpr (NV > 1) { vg = idx / n_cols; i = idx - vg * n_cols; }
            const int c = g.col_lo + i, col = g.spec_offset + i;
            const float2 *zk_p = smem + vg * (VEC * PAD) + pad_index(c & (NC - 1));
            const float2 *zn_p = smem + vg * (VEC * PAD) + pad_index((NC - c) & (NC - 1));
            const float2 *edge_p = smem + vg * (VEC * PAD) + kEdgeSlot;
            const float2 w = idx == tid && tid < n_cols ? aux : __ldg(g.split_twiddle + i);
            float xre[VEC], xim[VEC];
#pragma unroll
            for (int f = 0; f < VEC; ++f) {
                const float2 zk = zk_p[f * PAD], zn = zn_p[f * PAD];
                const float2 e = __fmul2_rn(__fadd2_rn(zk, make_float2(zn.x, -zn.y)), make_float2(0.5f, 0.5f));
                const float2 d = __fmul2_rn(__fadd2_rn(zk, make_float2(-zn.x, zn.y)), make_float2(0.5f, 0.5f));
                const float2 o = cmul_mi(d);
                float2 xc = cadd(e, cmul(o, w));
                if (shifted) xc = cmul(make_float2(xc.x + edge_p[f * PAD].x, xc.y), make_float2(w.x, -w.y));   // W^-c = conj(w)
                xre[f] = xc.x;
                xim[f] = xc.y;
            }
            // first frame of the group: a multiple of VEC, so the group lies in one half of one tile
            const uint32_t lf = lf0 + (uint32_t)(vg * VEC), tile = lf / kTileFrames;
            const int f0 = (int)(lf % kTileFrames), half = f0 >> 2;
            float *re_p, *im_p;
            if (planes) {   // [tile][chunk][column][4]: chunk = half (Re), 2 + half (Im)
                re_p = P.spec + (((size_t)tile * 4 + half) * P.plane_stride + col) * 4 + (f0 & 3);
                im_p = re_p + (size_t)2 * P.plane_stride * 4;
            } else {        // [tile][column][4 swizzled chunks][4], spec_index_re / _im
                const int sw = (col >> 1) & 3;
                float *rec = P.spec + ((size_t)tile * P.spec_stride + col) * (2 * kTileFrames) + (f0 & 3);
                re_p = rec + ((half ^ sw) << 2);
                im_p = rec + (((2 + half) ^ sw) << 2);
            }
            if constexpr (VEC == 4) {
                *reinterpret_cast<float4 *>(re_p) = make_float4(xre[0], xre[1], xre[2], xre[3]);
                *reinterpret_cast<float4 *>(im_p) = make_float4(xim[0], xim[1], xim[2], xim[3]);
            } else if constexpr (VEC == 2) {
                *reinterpret_cast<float2 *>(re_p) = make_float2(xre[0], xre[1]);
                *reinterpret_cast<float2 *>(im_p) = make_float2(xim[0], xim[1]);
            } else {
                if (valid) {   // one frame per CTA: nothing to pad
                    *re_p = xre[0];
                    *im_p = xim[0];
                }
            }
        }
    }
    FFT_T(13);   // split step issued
    __syncthreads();   // the next item's first pass overwrites the buffers the split step has just read
#ifdef PVQT_FFT_STATS
    if (threadIdx.x == 0) {
        constexpr int L2 = NC == 32 ? 5 : NC == 64 ? 6 : NC == 128 ? 7 : NC == 256 ? 8 : NC == 512 ? 9 : NC == 1024 ? 10 : NC == 2048 ? 11 : NC == 4096 ? 12 : NC == 8192 ? 13 : 14;
        const long long t_end = clock64();
        long long prev = t_cta;
        atomicAdd(&g_fft_stats[L2][0], (unsigned long long)(fft_t[0] - t_cta));   // item addressing, edge
        prev = fft_t[0];
        for (int q = 1; q <= 13; ++q) {
            if (fft_t[q] == 0) continue;
            atomicAdd(&g_fft_stats[L2][q], (unsigned long long)(fft_t[q] - prev));
            prev = fft_t[q];
        }
        atomicAdd(&g_fft_stats[L2][14], (unsigned long long)(t_end - prev));      // last barrier
        atomicAdd(&g_fft_stats[L2][15], 1ull);
    }
#endif
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK == 256 ? PVQT_FFT_MIN_BLOCKS : (BLOCK == 512 ? 2 : 1)) fft_groups_kernel(const __grid_constant__ FftParams P)
{
    extern __shared__ __align__(16) float2 fft_smem[];
    PVQT_STAMP(0, 0);
    pdl_launch_dependents();
    int gi = 0;
#pragma unroll 1
    while (gi + 1 < P.n_groups && (int)blockIdx.x >= P.group[gi + 1].cta_begin) ++gi;
    const FftGroup &g = P.group[gi];
    switch (g.log2_nc) {
#define PVQT_FFT_CASE(LOG2)                                                        \
    case LOG2:                                                                     \
        if constexpr ((1 << LOG2) / kPointsPerThread <= BLOCK)                     \
            fft_group_body<(1 << LOG2), BLOCK>(P, g, fft_smem);                    \
        break;
        PVQT_FFT_CASE(5)
        PVQT_FFT_CASE(6)
        PVQT_FFT_CASE(7)
        PVQT_FFT_CASE(8)
        PVQT_FFT_CASE(9)
        PVQT_FFT_CASE(10)
        PVQT_FFT_CASE(11)
        PVQT_FFT_CASE(12)
        PVQT_FFT_CASE(13)
        PVQT_FFT_CASE(14)
#undef PVQT_FFT_CASE
    default: break;
    }
    PVQT_STAMP(0, 1);
#ifdef PVQT_PHASE_TIMERS
    if (threadIdx.x == 0 && blockIdx.x < 8192) {
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        g_phase_stamps[0][blockIdx.x][2] = sm;
        g_phase_stamps[0][blockIdx.x][3] = gi;
    }
#endif
    // Tell K-spmm-db which tiles are complete: it then starts on a tile while other CTAs of this grid still run,
    // instead of after the grid and its flush (3 us after the last CTA, and all its CTAs in lock-step).
    if (P.tile_ready != nullptr) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t n_items = (P.frames.n_frames + g.frames_per_cta - 1) / g.frames_per_cta;
            for (uint32_t item = blockIdx.x - g.cta_begin; item < n_items; item += (uint32_t)g.n_ctas) {
                const uint32_t lf0 = item * g.frames_per_cta;
                const uint32_t lf1 = min(lf0 + (uint32_t)g.frames_per_cta, P.frames.n_frames);
                for (uint32_t tl = lf0 / kTileFrames; tl * kTileFrames < lf1; ++tl) atomicAdd(P.tile_ready + tl, 1u);
            }
        }
    }
    // Launched programmatically behind K-sdft (which runs beside this kernel): the grid must not complete before
    // that one has, so that the kernels after this one see the partial sums too.  One CTA -- the last one, which
    // starts when K-sdft is long complete -- waits for the whole grid: with the wait in every CTA the first wave
    // (2 CTAs per SM beside K-sdft's) finished its FFTs after 4 us and then held its SM slots until K-sdft ended
    // at 16 us, with no K-fft CTA running in between (globaltimer stamps, scripts/phase_timers.py).
#ifndef PVQT_FFT_WAIT_ALL
    if (P.wait_prior && blockIdx.x == gridDim.x - 1) pdl_wait();
#else
    if (P.wait_prior) pdl_wait();
#endif
    // The CTAs of one FFT group also run the combine step of the K-sdft groups for the frames they own,
    // reusing the FFT's shared memory.  The host picks the last group: its CTAs are scheduled when the
    // partial sums are long complete, so the wait below never holds SM slots.
    if (P.n_sdft > 0 && gi == P.combine_group) {
        pdl_wait();   // the partial sums must be complete (a second wait in the last CTA is harmless)
        __syncthreads();
        const uint32_t n_items = (P.frames.n_frames + g.frames_per_cta - 1) / g.frames_per_cta;
        for (uint32_t item = blockIdx.x - g.cta_begin; item < n_items; item += (uint32_t)g.n_ctas)
            for (int i = 0; i < P.n_sdft; ++i)
                sdft_combine_frames(P.sdft[i], item * g.frames_per_cta, g.frames_per_cta, fft_smem,
                                    sizeof(float2) * (size_t)pad_index(BLOCK * kPointsPerThread));
    }
}

// ------------------------------------------------------------------------------------------
// K-spmm: banded complex SpMM, kernel-stationary
// ------------------------------------------------------------------------------------------
// Grid: (row block, group of kSpmmWarps tiles), widest row blocks first.  Each warp owns one 8-frame
// tile: it stages the columns its block reads (contiguous 64-byte records of the tiled spectrum) with
// cp.async, then walks the block's band; lane l accumulates rows first_row + 2l and + 2l + 1 for the
// 8 frames.  No block-level barrier: staging buffers are warp-private.
__global__ void __launch_bounds__(kSpmmWarps * 32) spmm_kernel(const __grid_constant__ SpmmParams P)
{
    extern __shared__ __align__(16) float4 spmm_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned tile_groups = (P.n_tiles + kSpmmWarps - 1) / kSpmmWarps;
    const int rb = __ldg(P.block_order + blockIdx.x / tile_groups);
    const uint32_t tile = (blockIdx.x % tile_groups) * kSpmmWarps + warp;
    if (tile >= P.n_tiles) return;
    const SpmmRowBlock B = P.blocks[rb];

    float4 *buf = spmm_smem + (size_t)warp * P.max_cols * 4;
    const int4 meta = __ldg(P.lane_meta + rb * 32 + lane);
    const float4 *kv = P.values + (size_t)B.val_base * 32 + lane;
    const float4 *nv = P.values + (size_t)B.nval_base * 32 + lane;
    // Band widths are padded to multiples of kSpmmUnroll with zero coefficients and `values` has
    // kSpmmUnroll spare slots at its end, so the next group's coefficients are always fetched
    // unconditionally one full group ahead (they come from L2: the staging buffers leave little L1).
    float4 kq[kSpmmUnroll];
#pragma unroll
    for (int u = 0; u < kSpmmUnroll; ++u) kq[u] = __ldg(kv + u * 32);

    pdl_launch_dependents();
    pdl_wait();  // everything above is plan data; the spectra below come from K-fft
    {
        const float4 *src = reinterpret_cast<const float4 *>(P.spec) + ((size_t)tile * P.spec_stride + B.col_lo) * 4;
        const int n16 = B.n_cols * 4;
        for (int i = lane; i < n16; i += 32) cp_async16(buf + i, src + i);
    }
    cp_async_wait_all();
    __syncwarp();

    float2 re0[4], im0[4], re1[4], im1[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) re0[p] = im0[p] = re1[p] = im1[p] = make_float2(0.f, 0.f);

    // Lanes whose pair band is shorter than the block's skip the spectrum loads (x = 0).
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < B.width; j0 += kSpmmUnroll) {
        float4 kc[kSpmmUnroll];
#pragma unroll
        for (int u = 0; u < kSpmmUnroll; ++u) {
            kc[u] = kq[u];
            kq[u] = __ldg(kv + (j0 + kSpmmUnroll + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < kSpmmUnroll; ++u) {
            const int j = j0 + u;
            const int c = meta.x + j;
            const float4 *rec = buf + c * 4;
            const int sw = (c >> 1) & 3;  // col_lo is a multiple of 8: local and global swizzle agree
            float4 xr03 = zero4, xr47 = zero4, xi03 = zero4, xi47 = zero4;
            if (j < meta.y) {
                xr03 = rec[sw];
                xr47 = rec[1 ^ sw];
                xi03 = rec[2 ^ sw];
                xi47 = rec[3 ^ sw];
            }
            mac8<false>(re0, im0, kc[u].x, kc[u].y, xr03, xr47, xi03, xi47);
            mac8<false>(re1, im1, kc[u].z, kc[u].w, xr03, xr47, xi03, xi47);
        }
    }
    for (int j = 0; j < B.nwidth; ++j) {
        const float4 k = __ldg(nv + j * 32);  // conj(Kneg), see device plan
        const int c = meta.z + j;
        const float4 *rec = buf + c * 4;
        const int sw = (c >> 1) & 3;
        float4 xr03 = zero4, xr47 = zero4, xi03 = zero4, xi47 = zero4;
        if (j < meta.w) {
            xr03 = rec[sw];
            xr47 = rec[1 ^ sw];
            xi03 = rec[2 ^ sw];
            xi47 = rec[3 ^ sw];
        }
        mac8<true>(re0, im0, k.x, k.y, xr03, xr47, xi03, xi47);
        mac8<true>(re1, im1, k.z, k.w, xr03, xr47, xi03, xi47);
    }

    // |z|^2 (norm_sqr, vqt.rs:930); lanes write adjacent row pairs -> 256-byte runs per frame
    const int r_local = 2 * lane;
    const uint32_t frame0 = tile * kTileFrames;
    float *out = P.power + (size_t)frame0 * P.n_buckets + B.first_row + r_local;
    const float pr0[kTileFrames] = {re0[0].x, re0[0].y, re0[1].x, re0[1].y, re0[2].x, re0[2].y, re0[3].x, re0[3].y};
    const float pi0[kTileFrames] = {im0[0].x, im0[0].y, im0[1].x, im0[1].y, im0[2].x, im0[2].y, im0[3].x, im0[3].y};
    const float pr1[kTileFrames] = {re1[0].x, re1[0].y, re1[1].x, re1[1].y, re1[2].x, re1[2].y, re1[3].x, re1[3].y};
    const float pi1[kTileFrames] = {im1[0].x, im1[0].y, im1[1].x, im1[1].y, im1[2].x, im1[2].y, im1[3].x, im1[3].y};
#pragma unroll
    for (int f = 0; f < kTileFrames; ++f) {
        if (frame0 + f < P.n_frames) {
            if (r_local < B.n_rows) out[(size_t)f * P.n_buckets] = pr0[f] * pr0[f] + pi0[f] * pi0[f];
            if (r_local + 1 < B.n_rows) out[(size_t)f * P.n_buckets + 1] = pr1[f] * pr1[f] + pi1[f] * pi1[f];
        }
    }
}

// ------------------------------------------------------------------------------------------
// K-db: power_to_db, one warp per frame
// ------------------------------------------------------------------------------------------
constexpr int kDbWarps = 8;
constexpr int kDbMaxPerLane = 8;        // register-resident path: n_buckets <= 32 * 4 * 8 = 1024
// kVec: n_buckets % 4 == 0 and <= 1024 -> every lane keeps its float4s in registers (one pass over memory)
template <bool kVec>
__global__ void __launch_bounds__(kDbWarps * 32) power_to_db_kernel(const __grid_constant__ DbParams P)
{
    const int lane = threadIdx.x & 31;
    const uint32_t frame = blockIdx.x * kDbWarps + (threadIdx.x >> 5);
    pdl_launch_dependents();
    pdl_wait();  // P.power is written by K-spmm
    if (frame >= P.n_frames) return;
    const float *p = P.power + (size_t)frame * P.n_buckets;
    float *out = P.out_db + (size_t)frame * P.n_buckets;
    float mx = -CUDART_INF_F, mn = CUDART_INF_F;

    if constexpr (kVec) {
        const int n4 = P.n_buckets >> 2;
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        float4 l[kDbMaxPerLane];
#pragma unroll
        for (int i = 0; i < kDbMaxPerLane; ++i)
            l[i] = (lane + 32 * i < n4) ? p4[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kDbMaxPerLane; ++i) {
            if (lane + 32 * i < n4) {
                l[i] = make_float4(log_spec(l[i].x, P.ref_db), log_spec(l[i].y, P.ref_db), log_spec(l[i].z, P.ref_db),
                                   log_spec(l[i].w, P.ref_db));
                mx = fmaxf(mx, fmaxf(fmaxf(l[i].x, l[i].y), fmaxf(l[i].z, l[i].w)));
                mn = fminf(mn, fminf(fminf(l[i].x, l[i].y), fminf(l[i].z, l[i].w)));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // frame-wise max / min (vqt.rs:933-938)
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        const float floor_db = mx - kTopDb, log_spec_min = fmaxf(mn, floor_db);  // vqt.rs:939-940
        float4 *o4 = reinterpret_cast<float4 *>(out);
#pragma unroll
        for (int i = 0; i < kDbMaxPerLane; ++i)
            if (lane + 32 * i < n4)
                o4[lane + 32 * i] = make_float4(db_out(l[i].x, floor_db, log_spec_min), db_out(l[i].y, floor_db, log_spec_min),
                                                db_out(l[i].z, floor_db, log_spec_min), db_out(l[i].w, floor_db, log_spec_min));
    } else {
        for (int r = lane; r < P.n_buckets; r += 32) {
            const float l = log_spec(p[r], P.ref_db);
            mx = fmaxf(mx, l);
            mn = fminf(mn, l);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        const float floor_db = mx - kTopDb, log_spec_min = fmaxf(mn, floor_db);
        for (int r = lane; r < P.n_buckets; r += 32) out[r] = db_out(log_spec(p[r], P.ref_db), floor_db, log_spec_min);
    }
}


// ------------------------------------------------------------------------------------------
// K-spmm-db: banded complex SpMM + |z|^2 + power_to_db, one CTA per 8-frame tile
// ------------------------------------------------------------------------------------------
// The CTA stages the tile's whole spectrum (n_cols 64-byte records) once, every lane accumulates its
// two kernel rows for the 8 frames (same arithmetic and summation order as spmm_kernel), and because
// the CTA owns all n_buckets rows of its frames the frame-wise max / min of power_to_db (vqt.rs:933-940)
// are reduced in shared memory: the dB values are written once, coalesced, and the |z|^2 round trip
// through HBM/L2 and the third launch disappear.
// The tile is staged de-swizzled into four planes [chunk][column] of 16-byte entries (PLANE columns each, a
// compile-time stride), so a lane's four spectrum loads are one pointer plus immediates; band slots past a
// row's own band carry zero coefficients and columns past the staged range are zero-filled, so the band
// walk has no predicates at all (ncu source page: the address arithmetic, predicates and register clears
// of the swizzled, predicated form were 58 of the 90 instructions per slot).
template <int MAX_THREADS, int MIN_BLOCKS, int PLANE, int ROWS>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS) spmm_db_fused_kernel(const __grid_constant__ FusedParams P)
{
    constexpr int H = ROWS / 2;      // float4 coefficient entries per lane and slot (two rows each)
    constexpr int RING = kFusedRing;
    extern __shared__ __align__(16) float4 fused_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;
    PVQT_STAMP(1, 0);
    const FusedWarp W = P.warp[warp];
    const int4 meta = __ldg(P.lane_meta + warp * 32 + lane);
    const int2 rows = __ldg(P.lane_rows + warp * 32 + lane);
    const float4 *values = P.values + (size_t)(tile % P.values_copies) * P.values_stride;
    const float4 *kv = values + (size_t)W.val_base * (H * 32) + lane;
    // The coefficient stream goes through a lane-private ring in shared memory filled with cp.async RING slots
    // ahead: a register prefetch deep enough to cover the L2 latency does not fit the register budget of three
    // resident CTAs.  Lane l copies and reads only its own 16-byte entries, so no warp-level sync is needed.
    float4 *ring = fused_smem + 4 * PLANE + (size_t)warp * (RING * H * 32) + lane;
#pragma unroll
    for (int s = 0; s < RING; ++s) {  // `values` ends with spare slots
#pragma unroll
        for (int h = 0; h < H; ++h) cp_async16(ring + (s * H + h) * 32, kv + (s * H + h) * 32);
        asm volatile("cp.async.commit_group;\n" ::);
    }
    {   // columns the band walk may touch past the staged range
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = threadIdx.x; i < 4 * (P.cols_touched - P.n_cols); i += blockDim.x)
            fused_smem[(i & 3) * PLANE + P.n_cols + (i >> 2)] = z;
    }

    pdl_launch_dependents();
    PVQT_STAMP(1, 1);

    // K-sdft combine for this tile, straight into the planes (the columns of those groups are not in `spec`):
    // X_t[k] = sum_i phase[i][k] C[row(t) + i][k] (+ remainder), every CTA its own 8 frames.  As an epilogue of
    // K-fft's last CTAs the same step cost 8 us of that kernel's tail.
    auto combine = [&]() {
        for (int gi = 0; gi < P.n_sdft; ++gi) {
            const SdftParams &D = P.sdft[gi];
            const SdftGroup &G = D.g;
            const uint32_t lf0 = tile * kTileFrames, total = D.n_streams * D.frames;
            const int items = kTileFrames * G.nk;
            for (int item = threadIdx.x; item < items; item += blockDim.x) {
                const int fi = item / G.nk, k = item - fi * G.nk;
                const uint32_t lf = lf0 + fi;
                float2 x = make_float2(0.f, 0.f);
                if (lf < total) {
                    const uint32_t st = lf / D.frames, t = lf - st * D.frames;
                    const size_t row = (size_t)st * D.rows_per_stream + t;
                    x = sdft_dot<true>(D.partial_c + row * G.nk + k, G.phase + k, G.q, G.nk);
                    if (G.rem != 0) x = __fadd2_rn(x, cmul(__ldcg(D.partial_r + (row + G.q) * G.nk + k), __ldg(G.phase + G.q * G.nk + k)));
                }
                float *re_plane = reinterpret_cast<float *>(fused_smem + (fi >> 2) * PLANE + G.spec_offset + k);
                float *im_plane = reinterpret_cast<float *>(fused_smem + (2 + (fi >> 2)) * PLANE + G.spec_offset + k);
                re_plane[fi & 3] = x.x;
                im_plane[fi & 3] = x.y;
            }
        }
    };
    // Dependencies.  By default the kernel waits for the grid before it (K-fft, which itself outlasts K-sdft).  With
    // the completion counters armed it instead waits for exactly what this tile needs -- K-sdft's partial sums for
    // the combine, then the K-fft CTAs that write this tile -- so the CTAs of a step no longer start in lock-step
    // 3 us after K-fft's last CTA.  Every poll is bounded and falls back to the grid wait.
    __shared__ int dep_ready;
    auto poll = [&](auto done, int tries) {   // thread 0 polls, the CTA learns the outcome
        if (threadIdx.x == 0) {
            int ok = 0;
            for (int i = 0; i < tries && !(ok = done()); ++i) __nanosleep(100);
            dep_ready = ok;
        }
        __syncthreads();
        const bool r = dep_ready != 0;
        __syncthreads();
        return r;
    };
    auto load_acquire = [](const unsigned *p) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    };
    auto stage = [&]() {
        const float4 *src = reinterpret_cast<const float4 *>(P.spec) + (size_t)tile * P.spec_stride * 4;
        const int n16 = P.n_cols * 4;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) {
            const int c = i >> 2, p = i & 3;     // physical chunk p of column c
            const int q = p ^ ((c >> 1) & 3);    // logical chunk: 0,1 = Re frames 0-3, 4-7; 2,3 = Im
            bool from_sdft = false;              // columns the combine produces are not copied
            for (int gi = 0; gi < P.n_sdft; ++gi)
                from_sdft |= c >= P.sdft[gi].g.spec_offset && c < P.sdft[gi].g.spec_offset + P.sdft[gi].g.nk;
            if (!from_sdft) cp_async16(fused_smem + q * PLANE + c, src + i);
        }
    };
    if (P.tile_ready == nullptr) {
        pdl_wait();  // everything above is plan data; the spectra and partial sums below come from K-fft / K-sdft
        PVQT_STAMP(1, 2);
        stage();
        combine();   // while the cp.async copies are in flight
    } else {
        bool waited = false;
        if (P.n_sdft > 0) {
            const bool ok = P.sdft_done != nullptr &&
                            poll([&] { return (int)(load_acquire(P.sdft_done) - P.sdft_expected) >= 0; }, 256);
            if (!ok) { pdl_wait(); waited = true; }
            combine();
        }
        PVQT_STAMP(1, 7);
        if (!waited) {
            const uint32_t nf = min((uint32_t)kTileFrames, P.n_frames - tile * kTileFrames);
            uint32_t expected = 0;
            for (int gi = 0; gi < P.n_ready_groups; ++gi)
                expected += P.ready_fpc[gi] >= kTileFrames ? 1u : (nf + P.ready_fpc[gi] - 1) / P.ready_fpc[gi];
            if (!poll([&] { return load_acquire(P.tile_ready + tile) == expected; }, 4096)) pdl_wait();
        }
        if (threadIdx.x == 0) P.tile_ready[tile] = 0;   // every writer of this tile is through: reset for the next launch
        PVQT_STAMP(1, 2);
        stage();
    }
    cp_async_wait_all();
    __syncthreads();
    PVQT_STAMP(1, 3);

    float2 re[ROWS][4], im[ROWS][4];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int p = 0; p < 4; ++p) re[r][p] = im[r][p] = make_float2(0.f, 0.f);

    // One ring pass over the warp's slots: the band (y += K x), then the conjugate-part band (y += conj(Kneg x)),
    // whose slots follow the band's in `values` -- so its coefficients ride the same prefetch instead of paying a
    // dependent L2 / DRAM round trip per slot at the end of every CTA's critical path.
    auto next_slot = [&](int j, float4 (&k)[H]) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(RING - 1));  // slot j has landed
        float4 *slot = ring + (j & (RING - 1)) * (H * 32);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            k[h] = slot[h * 32];
            cp_async16(slot + h * 32, kv + ((j + RING) * H + h) * 32);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    {
        const float4 *xp = fused_smem + meta.x;
#pragma unroll 1
        for (int j = 0; j < W.width; ++j) {
            float4 k[H];
            next_slot(j, k);
            const float4 xr03 = xp[0], xr47 = xp[PLANE], xi03 = xp[2 * PLANE], xi47 = xp[3 * PLANE];
            ++xp;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                mac8<false>(re[2 * h], im[2 * h], k[h].x, k[h].y, xr03, xr47, xi03, xi47);
                mac8<false>(re[2 * h + 1], im[2 * h + 1], k[h].z, k[h].w, xr03, xr47, xi03, xi47);
            }
        }
    }
    {
        const float4 *xp = fused_smem + meta.z;
#pragma unroll 1
        for (int j = W.width; j < W.width + W.nwidth; ++j) {
            float4 k[H];  // conj(Kneg), see device plan
            next_slot(j, k);
            const float4 xr03 = xp[0], xr47 = xp[PLANE], xi03 = xp[2 * PLANE], xi47 = xp[3 * PLANE];
            ++xp;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                mac8<true>(re[2 * h], im[2 * h], k[h].x, k[h].y, xr03, xr47, xi03, xi47);
                mac8<true>(re[2 * h + 1], im[2 * h + 1], k[h].z, k[h].w, xr03, xr47, xi03, xi47);
            }
        }
    }

    // |z|^2 (norm_sqr) and log_spec (vqt.rs:930) per lane; the staging buffer becomes ls[frame][row]
    const uint32_t frame0 = tile * kTileFrames;
    const int nb = P.n_buckets;
    __syncthreads();  // every warp is done reading the staged spectrum
    PVQT_STAMP(1, 4);
    float *ls = reinterpret_cast<float *>(fused_smem);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (rows.y > r) {
#pragma unroll
            for (int f = 0; f < kTileFrames; ++f) {
                const float zr = (f & 1) ? re[r][f >> 1].y : re[r][f >> 1].x;
                const float zi = (f & 1) ? im[r][f >> 1].y : im[r][f >> 1].x;
                const float p = zr * zr + zi * zi;
                ls[f * nb + rows.x + r] = log_spec(p, P.ref_db);
                if (P.power != nullptr && frame0 + f < P.n_frames) P.power[(size_t)(frame0 + f) * nb + rows.x + r] = p;
            }
        }
    }
    __syncthreads();
    PVQT_STAMP(1, 5);

    // power_to_db's frame-wise part (vqt.rs:933-950): one warp per frame, coalesced stores
    const int n_warps = blockDim.x >> 5;
    for (int f = warp; f < kTileFrames; f += n_warps) {
        if (frame0 + f >= P.n_frames) break;
        const float *l = ls + f * nb;
        float *out = P.out_db + (size_t)(frame0 + f) * nb;
        float mx = -CUDART_INF_F, mn = CUDART_INF_F;
        const bool vec = (nb & 3) == 0 && (reinterpret_cast<uintptr_t>(P.out_db) & 15) == 0;
        if (vec) {
            const float4 *l4 = reinterpret_cast<const float4 *>(l);
            const int n4 = nb >> 2;
            for (int i = lane; i < n4; i += 32) {
                const float4 v = l4[i];
                mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
                mn = fminf(mn, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
            }
        } else {
            for (int r = lane; r < nb; r += 32) {
                mx = fmaxf(mx, l[r]);
                mn = fminf(mn, l[r]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        const float floor_db = mx - kTopDb, log_spec_min = fmaxf(mn, floor_db);  // vqt.rs:939-940
        if (vec) {
            const float4 *l4 = reinterpret_cast<const float4 *>(l);
            float4 *o4 = reinterpret_cast<float4 *>(out);
            const int n4 = nb >> 2;
            for (int i = lane; i < n4; i += 32) {
                const float4 v = l4[i];
                o4[i] = make_float4(db_out(v.x, floor_db, log_spec_min), db_out(v.y, floor_db, log_spec_min),
                                    db_out(v.z, floor_db, log_spec_min), db_out(v.w, floor_db, log_spec_min));
            }
        } else {
            for (int r = lane; r < nb; r += 32) out[r] = db_out(l[r], floor_db, log_spec_min);
        }
    }
#ifdef PVQT_PHASE_TIMERS
    __syncthreads();
    PVQT_STAMP(1, 6);
#endif
}

}  // namespace

size_t fft_smem_bytes(int block_threads)
{
    // every CTA holds block_threads * 16 complex points, whatever the FFT size
    return sizeof(float2) * (size_t)pad_index(block_threads * kPointsPerThread);
}

size_t spmm_smem_bytes(int max_cols) { return (size_t)kSpmmWarps * max_cols * kTileFrames * sizeof(float2); }

cudaError_t configure_kernels(int max_cols)
{
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(fft_groups_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fft_smem_bytes(256))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(fft_groups_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fft_smem_bytes(512))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(fft_groups_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fft_smem_bytes(1024))) != cudaSuccess) return e;
    // K-fft runs beside K-sdft (programmatic dependent launch).  CTAs of two kernels share an SM only under the same
    // L1 / shared-memory split, and the driver derives each kernel's split from its own occupancy: with K-fft at
    // 4 CTAs per SM (143 KB) and K-sdft at 100 KB, K-fft's CTAs waited until K-sdft had left the SMs (13 us; globaltimer
    // stamps, scripts/phase_timers.py).  Both kernels therefore name the same carve-out (kStepCarveoutPct, 164 KB).
    if (kStepCarveoutPct >= 0)
        for (const void *k : {(const void *)fft_groups_kernel<256>, (const void *)fft_groups_kernel<512>,
                              (const void *)fft_groups_kernel<1024>})
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, kStepCarveoutPct)) != cudaSuccess)
                return e;
    if (spmm_smem_bytes(max_cols) > 227 * 1024) return cudaErrorInvalidConfiguration;
    return cudaFuncSetAttribute(spmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)spmm_smem_bytes(max_cols));
}

namespace {
template <typename Kernel, typename Params>
cudaError_t launch_dependent(Kernel kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                             const Params &p)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}
}  // namespace

cudaError_t launch_fft(const FftParams &p, int total_ctas, int block_threads, cudaStream_t stream)
{
    const size_t smem = fft_smem_bytes(block_threads);
    if (p.wait_prior) {
        switch (block_threads) {
        case 256: return launch_dependent(fft_groups_kernel<256>, total_ctas, 256, smem, stream, p);
        case 512: return launch_dependent(fft_groups_kernel<512>, total_ctas, 512, smem, stream, p);
        case 1024: return launch_dependent(fft_groups_kernel<1024>, total_ctas, 1024, smem, stream, p);
        default: return cudaErrorInvalidConfiguration;
        }
    }
    switch (block_threads) {
    case 256: fft_groups_kernel<256><<<total_ctas, 256, smem, stream>>>(p); break;
    case 512: fft_groups_kernel<512><<<total_ctas, 512, smem, stream>>>(p); break;
    case 1024: fft_groups_kernel<1024><<<total_ctas, 1024, smem, stream>>>(p); break;
    default: return cudaErrorInvalidConfiguration;
    }
    return cudaGetLastError();
}

cudaError_t launch_spmm(const SpmmParams &p, cudaStream_t stream)
{
    const unsigned tile_groups = (p.n_tiles + kSpmmWarps - 1) / kSpmmWarps;
    return launch_dependent(spmm_kernel, tile_groups * p.n_blocks, kSpmmWarps * 32, spmm_smem_bytes(p.max_cols), stream, p);
}

// ---- K-spmm-db launch -------------------------------------------------------------------------
// plane width (columns) of the staged tile: the smallest instantiated width that holds `cols_touched`
int fused_plane_cols(int cols_touched)
{
    for (int p : {832, 1664})
        if (cols_touched <= p) return p;
    return 0;
}

size_t fused_smem_bytes(int cols_touched, int n_buckets, int n_warps, int rows_per_lane)
{
    const size_t plane = (size_t)4 * fused_plane_cols(cols_touched) * sizeof(float4);
    const size_t ls = (size_t)kTileFrames * n_buckets * sizeof(float);
    return std::max(plane, (ls + 15) & ~(size_t)15) +
           (size_t)n_warps * kFusedRing * (rows_per_lane / 2) * 32 * sizeof(float4);
}

bool fused_supported(int n_warps, int cols_touched, int n_buckets, int rows_per_lane)
{
    const int max_warps = rows_per_lane == 4 ? 16 : 32;
    return n_warps >= 1 && n_warps <= max_warps && fused_plane_cols(cols_touched) > 0 &&
           (size_t)kTileFrames * n_buckets * sizeof(float) <= (size_t)4 * fused_plane_cols(cols_touched) * sizeof(float4) &&
           fused_smem_bytes(cols_touched, n_buckets, n_warps, rows_per_lane) <= 200 * 1024;
}

namespace {
template <typename Fn>
cudaError_t with_fused_kernel(int n_warps, int plane, int rows_per_lane, Fn fn)
{
    if (rows_per_lane == 4) {   // 128 registers: three resident CTAs of <= 5 warps, one of <= 16
        if (plane == 832) {
            if (n_warps <= 5) return fn(spmm_db_fused_kernel<160, 3, 832, 4>);
            return fn(spmm_db_fused_kernel<512, 1, 832, 4>);
        }
        return fn(spmm_db_fused_kernel<512, 1, 1664, 4>);
    }
    // two rows per lane, 64 registers: 3 resident CTAs of <= 10 warps, 2 of <= 16, 1 beyond
    if (plane == 832) {
        if (n_warps <= 10) return fn(spmm_db_fused_kernel<320, 3, 832, 2>);
        if (n_warps <= 16) return fn(spmm_db_fused_kernel<512, 2, 832, 2>);
        return fn(spmm_db_fused_kernel<1024, 1, 832, 2>);
    }
    if (n_warps <= 16) return fn(spmm_db_fused_kernel<512, 1, 1664, 2>);
    return fn(spmm_db_fused_kernel<1024, 1, 1664, 2>);
}
}  // namespace

cudaError_t configure_fused(int n_warps, int cols_touched, int n_buckets, int rows_per_lane)
{
    const int smem = (int)fused_smem_bytes(cols_touched, n_buckets, n_warps, rows_per_lane);
    return with_fused_kernel(n_warps, fused_plane_cols(cols_touched), rows_per_lane, [&](auto kernel) {
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    });
}

cudaError_t launch_spmm_db_fused(const FusedParams &p, cudaStream_t stream)
{
    const size_t smem = fused_smem_bytes(p.cols_touched, p.n_buckets, p.n_warps, p.rows_per_lane);
    return with_fused_kernel(p.n_warps, fused_plane_cols(p.cols_touched), p.rows_per_lane, [&](auto kernel) {
        return launch_dependent(kernel, p.n_tiles, (unsigned)p.n_warps * 32, smem, stream, p);
    });
}

cudaError_t launch_power_to_db(const DbParams &p, cudaStream_t stream)
{
    const unsigned grid = (p.n_frames + kDbWarps - 1) / kDbWarps;
    const bool vec = (p.n_buckets % 4) == 0 && p.n_buckets <= 32 * 4 * kDbMaxPerLane &&
                     (reinterpret_cast<uintptr_t>(p.power) % 16) == 0 && (reinterpret_cast<uintptr_t>(p.out_db) % 16) == 0;
    return vec ? launch_dependent(power_to_db_kernel<true>, grid, kDbWarps * 32, 0, stream, p)
               : launch_dependent(power_to_db_kernel<false>, grid, kDbWarps * 32, 0, stream, p);
}

}  // namespace pvqt_dev

#ifdef PVQT_FFT_STATS
extern "C" int pvqt_debug_fft_stats(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, pvqt_dev::g_fft_stats, sizeof(pvqt_dev::g_fft_stats)) != cudaSuccess) return 7;
    if (reset) {
        static unsigned long long zeros[16][16] = {};
        cudaMemcpyToSymbol(pvqt_dev::g_fft_stats, zeros, sizeof(zeros));
    }
    return 0;
}
#endif

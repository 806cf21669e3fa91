// vqt_kernels.cu -- sm_100a kernels of the VQT hot path.
//
//   K-fft   fft_groups_kernel   batched shared-memory Stockham real-to-complex FFT, one
//                               launch for all window groups of all frames; replaces the
//                               realfft `process_with_scratch` calls at vqt.rs:884-887.
//                               Only the FFT bins the sparse kernel consumes are produced.
//   K-spmm  spmm_db_kernel      batched complex banded SpMM over a tile of frames, with the
//                               conjugate-part product (vqt.rs:889-910) and power_to_db
//                               (vqt.rs:922-954) fused as the epilogue.
//
// FFT convention (pinned by vqt.rs:1087-1128): unnormalised forward transform,
// X[k] = sum_n x[n] exp(-2 pi i k n / N), half spectrum k = 0..N/2.
#include "vqt_device.cuh"

#include <math_constants.h>

namespace pvqt_dev {
namespace {

// ------------------------------------------------------------------------------------------
// complex helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * (-i)
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kCosPi8 = 0.92387953251128675613f;
constexpr float kSinPi8 = 0.38268343236508977173f;

// multiply by W16^M = exp(-2 pi i M / 16), M compile-time
template <int M>
__device__ __forceinline__ float2 mul_w16(float2 a)
{
    if constexpr (M == 0) return a;
    else if constexpr (M == 1) return cmul(a, make_float2(kCosPi8, -kSinPi8));
    else if constexpr (M == 2) return make_float2((a.x + a.y) * kSqrtHalf, (a.y - a.x) * kSqrtHalf);
    else if constexpr (M == 3) return cmul(a, make_float2(kSinPi8, -kCosPi8));
    else if constexpr (M == 4) return cmul_mi(a);
    else if constexpr (M == 6) return make_float2((a.y - a.x) * kSqrtHalf, -(a.x + a.y) * kSqrtHalf);
    else if constexpr (M == 9) return cmul(a, make_float2(-kCosPi8, kSinPi8));
    else { static_assert(M < 0, "unsupported W16 power"); return a; }
}

__device__ __forceinline__ void fft2(float2 &a0, float2 &a1)
{
    float2 t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}

// 4-point forward DFT, natural order in and out
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
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

// shared-memory index padding: one float2 of padding per 16 keeps every pass conflict-free
__host__ __device__ constexpr int pad_index(int i) { return i + (i >> 4); }

// One Stockham pass (decimation in time, autosort):
//   butterfly b (< NC/R), k = b mod NS:
//     v[r]  = in[b + r NC/R] * exp(-2 pi i r k / (NS R))
//     out[(b - k) R + k + r NS] = DFT_R(v)[r]
template <int NC, int PASS, int NS>
__device__ __forceinline__ void fft_passes(float2 (&v)[kPointsPerThread], float2 *s, int t, const float *x,
                                           bool valid, const FftGroup &g)
{
    constexpr int R = plan_radix(NC, PASS);
    constexpr int T = NC / kPointsPerThread;
    constexpr int NB = kPointsPerThread / R;
    constexpr bool kFirst = PASS == 0;
    constexpr bool kLast = NS * R == NC;

#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + i * T;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int m = b + r * (NC / R);
            if constexpr (kFirst) {
                // z[m] = x[2m] + i x[2m+1]: the real window packed as N/2 complex points
                v[i * R + r] = valid ? make_float2(__ldg(x + 2 * m), __ldg(x + 2 * m + 1)) : make_float2(0.f, 0.f);
            } else {
                v[i * R + r] = s[pad_index(m)];
            }
        }
    }
    if constexpr (!kFirst) __syncthreads();  // all reads done before the in-place overwrite

#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + i * T;
        if constexpr (NS > 1) {
            const int k = b & (NS - 1);
            const float2 *tw = g.twiddle[PASS] + k;
#pragma unroll
            for (int r = 1; r < R; ++r) v[i * R + r] = cmul(v[i * R + r], __ldg(tw + (r - 1) * NS));
        }
        butterfly<R>(&v[i * R]);
    }

#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int b = t + i * T;
        const int k = b & (NS - 1);
        const int j0 = (b - k) * R + k;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int o = j0 + out_index<R>(j) * NS;
            if constexpr (kLast) {
                // only the bins the split step reads: Z[c] and Z[NC - c], c in [col_lo, col_hi]
                if (o <= g.col_hi || o >= NC - g.col_hi) s[pad_index(o)] = v[i * R + j];
            } else {
                s[pad_index(o)] = v[i * R + j];
            }
        }
    }
    __syncthreads();

    if constexpr (!kLast) fft_passes<NC, PASS + 1, NS * R>(v, s, t, x, valid, g);
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
    const uint32_t local_frame = (blockIdx.x - g.cta_begin) * FPC + fid;
    const bool valid = local_frame < P.frames.n_frames;

    const uint64_t f = P.frames.first_frame + (valid ? local_frame : 0);
    const uint64_t stream = f / P.frames.frames_per_stream;
    const uint64_t in_stream = f - stream * P.frames.frames_per_stream;
    const float *x = P.frames.audio + stream * P.frames.stream_stride + in_stream * P.frames.hop + g.window_begin;

    float2 *s = smem + fid * pad_index(NC);
    float2 v[kPointsPerThread];
    fft_passes<NC, 0, 1>(v, s, t, x, valid, g);

    // Real-FFT split: with Z = FFT_{NC}(z), E/O the spectra of the even/odd samples,
    //   X[c] = E[c] + W_N^c O[c],  E = (Z[c] + conj Z[NC-c]) / 2,  O = -i (Z[c] - conj Z[NC-c]) / 2
    if (valid) {
        float2 *out = P.spec + (uint64_t)local_frame * P.spec_stride + g.spec_offset;
        const int n_cols = g.col_hi - g.col_lo + 1;
        for (int i = t; i < n_cols; i += T) {
            const int c = g.col_lo + i;
            const float2 zk = s[pad_index(c & (NC - 1))];
            const float2 zn = s[pad_index((NC - c) & (NC - 1))];
            const float er = 0.5f * (zk.x + zn.x), ei = 0.5f * (zk.y - zn.y);
            const float orr = 0.5f * (zk.y + zn.y), oi = -0.5f * (zk.x - zn.x);
            const float2 w = __ldg(g.split_twiddle + i);
            out[i] = make_float2(er + (orr * w.x - oi * w.y), ei + (orr * w.y + oi * w.x));
        }
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) fft_groups_kernel(const __grid_constant__ FftParams P)
{
    extern __shared__ __align__(16) float2 fft_smem[];
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
}

// ------------------------------------------------------------------------------------------
// K-spmm: banded complex SpMM over a tile of F frames + power_to_db epilogue
// ------------------------------------------------------------------------------------------
constexpr int kSpmmThreads = 256;
constexpr float kAMin = 1e-6f * 1e-6f;  // vqt.rs:924
constexpr float kTopDb = 60.0f;         // vqt.rs:925

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
}

template <int F>
__global__ void __launch_bounds__(kSpmmThreads) spmm_db_kernel(const __grid_constant__ SpmmParams P)
{
    extern __shared__ __align__(16) unsigned char spmm_smem_raw[];
    const int S = P.spec_stride;
    const int NB = P.n_buckets;
    float2 *tile = reinterpret_cast<float2 *>(spmm_smem_raw);            // [F][S]
    float *ls = reinterpret_cast<float *>(tile + (size_t)F * S);         // [F][NB]
    float *red = ls + (size_t)F * NB;                                    // [2][F][warps]
    constexpr int kWarps = kSpmmThreads / 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t frame0 = blockIdx.x * F;
    const int n_valid = min((uint32_t)F, P.n_frames - frame0);

    // stage the spectra of this tile's frames (contiguous in global memory)
    {
        const float4 *src = reinterpret_cast<const float4 *>(P.spec + (size_t)frame0 * S);
        float4 *dst = reinterpret_cast<float4 *>(tile);
        const int n16_valid = n_valid * S / 2, n16 = F * S / 2;
        for (int i = tid; i < n16_valid; i += kSpmmThreads) cp_async16(dst + i, src + i);
        for (int i = n16_valid + tid; i < n16; i += kSpmmThreads) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        cp_async_wait_all();
    }
    __syncthreads();

    float run_max[F], run_min[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { run_max[f] = -CUDART_INF_F; run_min[f] = CUDART_INF_F; }

    for (int blk = warp; blk < P.n_blocks; blk += kWarps) {
        const SpmmBlock B = P.blocks[blk];
        const int row = blk * kSpmmRowsPerBlock + lane;
        const int2 rc = __ldg(P.row_cols + row);
        float2 acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = make_float2(0.f, 0.f);

        // y[r] += sum_c K[r,c] X[c]                                       (vqt.rs:889-894)
        const float2 *kv = P.values + (size_t)B.val_base * kSpmmRowsPerBlock + lane;
        for (int j = 0; j < B.width; ++j) {
            const float2 k = __ldg(kv + j * kSpmmRowsPerBlock);
            const int c = min(rc.x + j, S - 1);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float2 x = tile[f * S + c];
                acc[f].x = fmaf(k.x, x.x, acc[f].x);
                acc[f].x = fmaf(-k.y, x.y, acc[f].x);
                acc[f].y = fmaf(k.x, x.y, acc[f].y);
                acc[f].y = fmaf(k.y, x.x, acc[f].y);
            }
        }
        // y[r] += conj(sum_c Kneg[r,c] X[c]) = sum_c conj(Kneg[r,c]) conj(X[c])   (vqt.rs:896-910)
        const float2 *nv = P.values + (size_t)B.nval_base * kSpmmRowsPerBlock + lane;
        for (int j = 0; j < B.nwidth; ++j) {
            const float2 k = __ldg(nv + j * kSpmmRowsPerBlock);  // stored already conjugated
            const int c = min(rc.y + j, S - 1);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float2 x = tile[f * S + c];
                acc[f].x = fmaf(k.x, x.x, acc[f].x);
                acc[f].x = fmaf(k.y, x.y, acc[f].x);
                acc[f].y = fmaf(k.y, x.x, acc[f].y);
                acc[f].y = fmaf(-k.x, x.y, acc[f].y);
            }
        }

        if (row < NB) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float p = acc[f].x * acc[f].x + acc[f].y * acc[f].y;       // norm_sqr, vqt.rs:930
                if (P.out_power != nullptr && f < n_valid)
                    P.out_power[(size_t)(frame0 + f) * NB + row] = p;
                const float l = 10.0f * log10f(fmaxf(p, kAMin)) - P.ref_db;      // vqt.rs:930
                ls[f * NB + row] = l;
                run_max[f] = fmaxf(run_max[f], l);
                run_min[f] = fminf(run_min[f], l);
            }
        }
    }

    // frame-wise max / min of the log spectrum (vqt.rs:933-938): lanes, then warps
#pragma unroll
    for (int f = 0; f < F; ++f) {
        float mx = run_max[f], mn = run_min[f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (lane == 0) {
            red[f * kWarps + warp] = mx;
            red[(F + f) * kWarps + warp] = mn;
        }
    }
    __syncthreads();

    // clamp to 60 dB below the frame maximum and shift (vqt.rs:939-951)
    for (int f = 0; f < n_valid; ++f) {
        float mx = -CUDART_INF_F, mn = CUDART_INF_F;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            mx = fmaxf(mx, red[f * kWarps + w]);
            mn = fminf(mn, red[(F + f) * kWarps + w]);
        }
        const float floor_db = mx - kTopDb;
        const float log_spec_min = fmaxf(mn, floor_db);
        float *out = P.out_db + (size_t)(frame0 + f) * NB;
        for (int r = tid; r < NB; r += kSpmmThreads) {
            const float clamped = fmaxf(ls[f * NB + r], floor_db);
            out[r] = log_spec_min > 0.0f ? clamped - log_spec_min : fmaxf(clamped, 0.0f);
        }
    }
}

template <int F>
cudaError_t launch_spmm_t(const SpmmParams &p, cudaStream_t stream)
{
    const size_t smem = spmm_smem_bytes(F, p.spec_stride, p.n_buckets);
    const unsigned grid = (p.n_frames + F - 1) / F;
    spmm_db_kernel<F><<<grid, kSpmmThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int F>
cudaError_t configure_spmm_t(size_t smem)
{
    return cudaFuncSetAttribute(spmm_db_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace

size_t fft_smem_bytes(int block_threads)
{
    // every CTA holds block_threads * 16 complex points, whatever the FFT size
    return sizeof(float2) * (size_t)pad_index(block_threads * kPointsPerThread);
}

size_t spmm_smem_bytes(int frames_per_cta, int spec_stride, int n_buckets)
{
    return (size_t)frames_per_cta * spec_stride * sizeof(float2) + (size_t)frames_per_cta * n_buckets * sizeof(float) +
           2u * frames_per_cta * (kSpmmThreads / 32) * sizeof(float);
}

cudaError_t configure_kernels(int spec_stride, int n_buckets, int *spmm_frames_per_cta)
{
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(fft_groups_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fft_smem_bytes(256))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(fft_groups_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fft_smem_bytes(512))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(fft_groups_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fft_smem_bytes(1024))) != cudaSuccess) return e;
    // largest frame tile whose shared memory still lets two CTAs share an SM (or fits at all)
    const size_t kTwoPerSm = 110 * 1024, kMax = 227 * 1024;
    int f = 0;
    for (int cand : {8, 4, 2, 1})
        if (spmm_smem_bytes(cand, spec_stride, n_buckets) <= kTwoPerSm) { f = cand; break; }
    if (f == 0)
        for (int cand : {8, 4, 2, 1})
            if (spmm_smem_bytes(cand, spec_stride, n_buckets) <= kMax) { f = cand; break; }
    if (f == 0) return cudaErrorInvalidConfiguration;
    const size_t smem = spmm_smem_bytes(f, spec_stride, n_buckets);
    switch (f) {
    case 8: e = configure_spmm_t<8>(smem); break;
    case 4: e = configure_spmm_t<4>(smem); break;
    case 2: e = configure_spmm_t<2>(smem); break;
    default: e = configure_spmm_t<1>(smem); break;
    }
    *spmm_frames_per_cta = f;
    return e;
}

cudaError_t launch_fft(const FftParams &p, int total_ctas, int block_threads, cudaStream_t stream)
{
    const size_t smem = fft_smem_bytes(block_threads);
    switch (block_threads) {
    case 256: fft_groups_kernel<256><<<total_ctas, 256, smem, stream>>>(p); break;
    case 512: fft_groups_kernel<512><<<total_ctas, 512, smem, stream>>>(p); break;
    case 1024: fft_groups_kernel<1024><<<total_ctas, 1024, smem, stream>>>(p); break;
    default: return cudaErrorInvalidConfiguration;
    }
    return cudaGetLastError();
}

cudaError_t launch_spmm_db(const SpmmParams &p, int frames_per_cta, cudaStream_t stream)
{
    switch (frames_per_cta) {
    case 8: return launch_spmm_t<8>(p, stream);
    case 4: return launch_spmm_t<4>(p, stream);
    case 2: return launch_spmm_t<2>(p, stream);
    case 1: return launch_spmm_t<1>(p, stream);
    default: return cudaErrorInvalidConfiguration;
    }
}

}  // namespace pvqt_dev

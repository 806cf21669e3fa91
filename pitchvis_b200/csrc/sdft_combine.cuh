// sdft_combine.cuh -- the combine step of K-sdft as a device function, shared by the stand-alone
// sdft_combine_kernel and by K-fft (whose last group's CTAs run it as an epilogue, so the step costs no
// launch of its own).
#pragma once

#include "device_helpers.cuh"
#include "vqt_device.cuh"

namespace pvqt_dev {
namespace {

// X_t[k] = sum_{i<q} phase[i][k] C[row(t) + i][k] + phase[q][k] R[row(t) + q][k] for the local frames
// [lf0, lf0 + nfr) of the launch, all consumed bins k, accumulated with Kahan compensation in f32 (the rounding of the
// plain running sum would decide the accuracy of the path) and written into the
// tiled spectrum layout.  The chunk rows of the frames' run are staged with cp.async when they fit
// `smem` (one L2 round trip instead of q dependent ones); frames of another stream than the first
// frame's (only at stream boundaries) read global memory directly.  Called by every thread of the CTA.
__device__ __forceinline__ void sdft_combine_frames(const SdftParams &P, uint32_t lf0, int nfr, float2 *smem,
                                                    size_t smem_bytes)
{
    const SdftGroup &G = P.g;
    const uint32_t n_frames = P.n_streams * P.frames;
    __syncthreads();  // `smem` may still be read by the caller's previous phase
    if (lf0 >= n_frames) return;
    nfr = (int)min((uint32_t)nfr, n_frames - lf0);
    const uint32_t s0 = lf0 / P.frames, t0 = lf0 - s0 * P.frames;
    const size_t run_row0 = (size_t)s0 * P.rows_per_stream + t0;
    const uint32_t run_rows = min((uint32_t)(G.q + nfr - 1), P.rows_per_stream - t0);
    const bool staged = (size_t)run_rows * G.nk * sizeof(float2) <= smem_bytes;
    if (staged) {
        // whole 16-byte pieces; the run starts on an even element when row * nk is even, else one scalar first
        const float2 *src = P.partial_c + run_row0 * G.nk;
        const int n = (int)run_rows * G.nk;
        const int head = (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1);  // elements before 16-byte alignment
        float2 *dst = smem + head;  // keep dst and src congruent modulo 16 bytes (smem base is 16-byte aligned)
        if (threadIdx.x == 0 && head) dst[0] = src[0];
        const int n2 = (n - head) >> 1;
        for (int i = threadIdx.x; i < n2; i += blockDim.x) cp_async16(dst + head + 2 * i, src + head + 2 * i);
        if (threadIdx.x == 0 && ((n - head) & 1)) dst[n - 1] = src[n - 1];
        cp_async_wait_all();
    }
    __syncthreads();
    const int head_off = staged ? (int)((reinterpret_cast<uintptr_t>(P.partial_c + run_row0 * G.nk) >> 3) & 1) : 0;
    const int items = nfr * G.nk;
    const size_t row_stride = (size_t)G.nk;  // elements between consecutive chunk rows
    for (int item = threadIdx.x; item < items; item += blockDim.x) {
        const int fi = item / G.nk, k = item - fi * G.nk;
        const uint32_t lf = lf0 + fi;
        const uint32_t s = lf / P.frames, t = lf - s * P.frames;
        const size_t row = (size_t)s * P.rows_per_stream + t;
        const bool in_run = staged && s == s0;
        const float2 *c = in_run ? smem + head_off + (size_t)(t - t0) * G.nk + k : P.partial_c + row * G.nk + k;
        const float2 *w = G.phase + k;
        // Kahan-compensated f32 sum of the q (+1) products: the rounding of the plain running sum would be
        // the largest error of the whole path (4x); FP64 is avoided on purpose (slow pipe).
        float2 acc = make_float2(0.f, 0.f), comp = make_float2(0.f, 0.f);
        const float2 neg1 = make_float2(-1.f, -1.f);
#pragma unroll 4
        for (int i = 0; i <= G.q; ++i) {
            if (i == G.q && G.rem == 0) break;
            const float2 wv = __ldg(w);
            const float2 v = i < G.q ? *c : P.partial_r[(row + G.q) * G.nk + k];
            w += row_stride;
            c += row_stride;
            const float2 p = cmul(v, wv);
            const float2 y = __ffma2_rn(comp, neg1, p);           // p - comp
            const float2 tsum = __fadd2_rn(acc, y);
            const float2 d = __ffma2_rn(acc, neg1, tsum);         // (acc + y) - acc
            comp = __ffma2_rn(y, neg1, d);                        // ((acc + y) - acc) - y
            acc = tsum;
        }
        const float xr = acc.x, xi = acc.y;
        P.spec[spec_index_re(lf, G.spec_offset + k, P.spec_stride)] = xr;
        P.spec[spec_index_im(lf, G.spec_offset + k, P.spec_stride)] = xi;
    }
}

}  // namespace
}  // namespace pvqt_dev

// sdft_combine.cuh -- the combine step of K-sdft as a device function, shared by the stand-alone
// sdft_combine_kernel and by K-fft (whose last group's CTAs run it as an epilogue, so the step costs no
// launch of its own).
#pragma once

#include "device_helpers.cuh"
#include "vqt_device.cuh"

namespace pvqt_dev {
namespace {

// sum_i w[i * stride] * c[i * stride], i < n, as four interleaved partial sums (terms i mod 4) added at the
// end: free, and half the rounding error of one running sum (DESIGN.md, K-sdft accuracy).
// kCg: read the partial sums with ld.global.cg (L2 only) -- for the caller that starts on K-sdft's completion counter
// instead of a grid dependency and must not meet a line an earlier kernel left in this SM's L1.
template <bool kCg = false>
__device__ __forceinline__ float2 sdft_dot(const float2 *c, const float2 *w, int n, int stride)
{
    auto ld = [](const float2 *p) { return kCg ? __ldcg(p) : *p; };
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        const float2 w0 = w[0], w1 = w[stride], w2 = w[2 * stride], w3 = w[3 * stride];
        const float2 v0 = ld(c), v1 = ld(c + stride), v2 = ld(c + 2 * stride), v3 = ld(c + 3 * stride);
        w += 4 * stride;
        c += 4 * stride;
        a0 = __ffma2_rn(make_float2(w0.x, w0.x), v0, a0); a0 = __ffma2_rn(make_float2(-w0.y, w0.y), make_float2(v0.y, v0.x), a0);
        a1 = __ffma2_rn(make_float2(w1.x, w1.x), v1, a1); a1 = __ffma2_rn(make_float2(-w1.y, w1.y), make_float2(v1.y, v1.x), a1);
        a2 = __ffma2_rn(make_float2(w2.x, w2.x), v2, a2); a2 = __ffma2_rn(make_float2(-w2.y, w2.y), make_float2(v2.y, v2.x), a2);
        a3 = __ffma2_rn(make_float2(w3.x, w3.x), v3, a3); a3 = __ffma2_rn(make_float2(-w3.y, w3.y), make_float2(v3.y, v3.x), a3);
    }
    for (; i < n; ++i) {
        const float2 w0 = w[0], v0 = ld(c);
        w += stride;
        c += stride;
        a0 = __ffma2_rn(make_float2(w0.x, w0.x), v0, a0); a0 = __ffma2_rn(make_float2(-w0.y, w0.y), make_float2(v0.y, v0.x), a0);
    }
    return __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
}

// X_t[k] = sum_{i<q} phase[i][k] C[row(t) + i][k] + phase[q][k] R[row(t) + q][k] for the local frames
// [lf0, lf0 + nfr) of the launch and all consumed bins k, written into the tiled spectrum layout.  The
// chunk rows of the frames' run, their R rows and the phase table are staged into `smem` with cp.async when
// they fit (one L2 round trip instead of q dependent ones); frames of another stream than the first
// frame's (only at stream boundaries) read global memory directly.  Called by every thread of the CTA.
__device__ __forceinline__ void sdft_combine_frames(const SdftParams &P, uint32_t lf0, int nfr, float2 *smem,
                                                    size_t smem_bytes)
{
    const SdftGroup &G = P.g;
    const uint32_t n_frames = P.n_streams * P.frames;
    __syncthreads();  // `smem` may still be read by the caller's previous phase
    if (lf0 >= n_frames) return;
    nfr = (int)min((uint32_t)nfr, n_frames - lf0);
    const int nk = G.nk;
    const uint32_t s0 = lf0 / P.frames, t0 = lf0 - s0 * P.frames;
    const size_t run_row0 = (size_t)s0 * P.rows_per_stream + t0;
    const int run_rows = (int)min((uint32_t)(G.q + nfr - 1), P.rows_per_stream - t0);
    const int r_rows = G.rem != 0 ? max(0, min(nfr, (int)(P.rows_per_stream - t0) - G.q)) : 0;
    // smem: [1 pad][run_rows x nk] C | [r_rows x nk] R | [(q + 1) x nk] phase   (float2 each)
    const int n_c = run_rows * nk, n_r = r_rows * nk, n_w = (G.q + 1) * nk;
    const bool staged = (size_t)(n_c + n_r + n_w + 8) * sizeof(float2) <= smem_bytes;
    float2 *cs = smem, *rs = smem + ((n_c + 3) & ~1), *ws = rs + ((n_r + 3) & ~1);
    if (staged) {
        auto stage = [&](float2 *&dst, const float2 *src, int n) {
            // keep dst and src congruent modulo 16 bytes; odd ends as scalars
            const int head = (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1);
            dst += head;  // dst[j] <-> src[j]; dst + head is 16-byte aligned when dst was
            if (threadIdx.x == 0 && head && n > 0) dst[0] = src[0];
            const int n2 = (n - head) >> 1;
            for (int i = threadIdx.x; i < n2; i += blockDim.x) cp_async16(dst + head + 2 * i, src + head + 2 * i);
            if (threadIdx.x == 0 && n > head && ((n - head) & 1)) dst[n - 1] = src[n - 1];
        };
        stage(cs, P.partial_c + run_row0 * nk, n_c);
        stage(rs, P.partial_r + (run_row0 + G.q) * nk, n_r);
        stage(ws, G.phase, n_w);
        cp_async_wait_all();
    }
    __syncthreads();
    const int items = nfr * nk;
    for (int item = threadIdx.x; item < items; item += blockDim.x) {
        const int fi = item / nk, k = item - fi * nk;
        const uint32_t lf = lf0 + fi;
        const uint32_t s = lf / P.frames, t = lf - s * P.frames;
        const size_t row = (size_t)s * P.rows_per_stream + t;
        float2 x;
        if (staged && s == s0) {
            x = sdft_dot(cs + (size_t)(t - t0) * nk + k, ws + k, G.q, nk);
            if (G.rem != 0) {
                const float2 w = ws[G.q * nk + k], v = rs[(size_t)(t - t0) * nk + k];
                x = __fadd2_rn(x, cmul(v, w));
            }
        } else {
            x = sdft_dot(P.partial_c + row * nk + k, G.phase + k, G.q, nk);
            if (G.rem != 0) x = __fadd2_rn(x, cmul(P.partial_r[(row + G.q) * nk + k], __ldg(G.phase + G.q * nk + k)));
        }
        P.spec[spec_index_re(lf, G.spec_offset + k, P.spec_stride)] = x.x;
        P.spec[spec_index_im(lf, G.spec_offset + k, P.spec_stride)] = x.y;
    }
}

}  // namespace
}  // namespace pvqt_dev

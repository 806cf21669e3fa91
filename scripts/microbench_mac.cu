// microbench_mac.cu -- the FMA body of K-spmm-db's band walk in isolation: one slot = 32 FFMA2 on 16 packed accumulators
// (two kernel rows x eight frames, re / im), each accumulator updated twice per slot.  Cycles per slot per warp at 1..4
// warps per SM sub-partition, for several orderings of the 32 instructions.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 2048;

template <int ORDER>
__global__ void __launch_bounds__(512, 1) k(float *out, const float *in, int warps)
{
    if ((int)(threadIdx.x >> 5) >= warps) return;
    float2 re[2][4], im[2][4], xr[4], xi[4];
    float kk[4];
    for (int r = 0; r < 2; ++r) for (int p = 0; p < 4; ++p) { re[r][p] = make_float2(in[r * 4 + p] + threadIdx.x, in[8 + p]); im[r][p] = make_float2(in[p] - threadIdx.x, in[3]); }
    for (int p = 0; p < 4; ++p) { xr[p] = make_float2(in[16 + p] + threadIdx.x, in[20 + p]); xi[p] = make_float2(in[24 + p], in[28 + p] + threadIdx.x); }
    for (int i = 0; i < 4; ++i) kk[i] = in[32 + i] + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float kre = kk[2 * r], kim = kk[2 * r + 1];
            const float2 a = make_float2(kre, kre), b = make_float2(-kim, -kim), c = make_float2(kre, kre), d = make_float2(kim, kim);
            if (ORDER == 0) {          // as mac_slot / mac8: per p, the four updates back to back
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    re[r][p] = __ffma2_rn(a, xr[p], re[r][p]);
                    re[r][p] = __ffma2_rn(b, xi[p], re[r][p]);
                    im[r][p] = __ffma2_rn(c, xi[p], im[r][p]);
                    im[r][p] = __ffma2_rn(d, xr[p], im[r][p]);
                }
            } else {                   // all first updates, then all second updates (dependent distance 8)
#pragma unroll
                for (int p = 0; p < 4; ++p) { re[r][p] = __ffma2_rn(a, xr[p], re[r][p]); im[r][p] = __ffma2_rn(c, xi[p], im[r][p]); }
#pragma unroll
                for (int p = 0; p < 4; ++p) { re[r][p] = __ffma2_rn(b, xi[p], re[r][p]); im[r][p] = __ffma2_rn(d, xr[p], im[r][p]); }
            }
        }
        // rotate the operands a little so that nothing is loop-invariant
        kk[0] += 1.0f; xr[0].x += 1.0f;
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int r = 0; r < 2; ++r) for (int p = 0; p < 4; ++p) s += re[r][p].x + re[r][p].y + im[r][p].x + im[r][p].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) reinterpret_cast<long long *>(out + (1 << 20))[blockIdx.x] = t1 - t0;
}

int main()
{
    float *out, *in;
    cudaMalloc(&out, (2 << 20) * sizeof(float));
    cudaMalloc(&in, 64 * sizeof(float));
    cudaMemset(in, 0, 64 * sizeof(float));
    for (int order = 0; order < 2; ++order)
        for (int warps : {1, 4, 8, 10, 12, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (order == 0) k<0><<<148, 512>>>(out, in, warps); else k<1><<<148, 512>>>(out, in, warps);
                cudaDeviceSynchronize();
            }
            long long cyc;
            cudaMemcpy(&cyc, reinterpret_cast<long long *>(out + (1 << 20)), sizeof(cyc), cudaMemcpyDeviceToHost);
            printf("order %d, %2d warps per SM: %.1f cycles per slot (32 FFMA2) per warp; %.2f cycles per FFMA2 per sub-partition (busiest)\n",
                   order, warps, (double)cyc / kIters, (double)cyc / kIters / 32.0 / ((warps + 3) / 4));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

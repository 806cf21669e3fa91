// microbench_pipes.cu -- issue / pipe rates of the instructions the VQT kernels are made of, on the box's B200:
// FFMA vs FFMA2 (packed f32x2) with different operand patterns, FADD2, LDS.64 / LDS.128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_pipes scripts/microbench_pipes.cu && ./microbench_pipes
// Prints cycles per warp-instruction per SM sub-partition at 1, 2, 4, 8 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float *out, const float *in, int warps)
{
    if ((int)(threadIdx.x >> 5) >= warps) return;
    float2 a[8], x[4], c0 = make_float2(in[0] + threadIdx.x, in[1] - threadIdx.x), c1 = make_float2(in[2] * threadIdx.x, in[3] + 2 * threadIdx.x);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(in[4 + i] + threadIdx.x, in[12 + i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = make_float2(in[20 + i], in[24 + i] + threadIdx.x);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0) {          // FFMA2: 8 independent accumulators, shared multiplier (the SpMM walk's pattern)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(c0, x[i & 3], a[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(c1, x[(i + 1) & 3], a[i]);
        } else if (MODE == 1) {   // scalar FFMA, the same arithmetic: 32 instructions
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(c0.x, x[i & 3].x, a[i].x); a[i].y = fmaf(c0.y, x[i & 3].y, a[i].y); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(c1.x, x[(i + 1) & 3].x, a[i].x); a[i].y = fmaf(c1.y, x[(i + 1) & 3].y, a[i].y); }
        } else if (MODE == 2) {   // FADD2: 16 per iteration (FFT butterflies)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fadd2_rn(a[i], x[i & 3]);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fadd2_rn(a[i], c0);
        } else if (MODE == 3) {   // scalar FADD: 32
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x += x[i & 3].x; a[i].y += x[i & 3].y; }
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x += c0.x; a[i].y += c0.y; }
        } else if (MODE == 5) {   // FFMA2 with a scalar multiplier broadcast to both halves (the SpMM walk's actual form: R.F32)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(make_float2(c0.x, c0.x), x[i & 3], a[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(make_float2(c1.y, c1.y), x[(i + 1) & 3], a[i]);
        } else if (MODE == 4) {   // FMUL2
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fmul2_rn(a[i], x[i & 3]);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fmul2_rn(a[i], c0);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(t1 - t0) * 1e-30f;
    if (threadIdx.x == 0) reinterpret_cast<long long *>(out + (1 << 20))[blockIdx.x] = t1 - t0;
}

template <int BYTES>
__global__ void __launch_bounds__(1024) lds(float *out, int warps)
{
    __shared__ __align__(16) float4 buf[2048];
    if ((int)(threadIdx.x >> 5) >= warps) return;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float acc = 0.f;
    int idx = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (BYTES == 16) { const float4 v = buf[(idx + 32 * u) & 2047]; acc += (v.x + v.y) + (v.z + v.w); }
            else { const float2 v = reinterpret_cast<const float2 *>(buf)[(idx + 32 * u) & 4095]; acc += v.x + v.y; }
        }
        idx += 7;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) reinterpret_cast<long long *>(out + (1 << 20))[blockIdx.x] = t1 - t0;
}

int main()
{
    float *out, *in;
    cudaMalloc(&out, (2 << 20) * sizeof(float));
    cudaMalloc(&in, 64 * sizeof(float));
    cudaMemset(in, 0, 64 * sizeof(float));
    const char *names[] = {"FFMA2 (16/iter)", "FFMA (32/iter)", "FADD2 (16/iter)", "FADD (32/iter)", "FMUL2 (16/iter)", "FFMA2 bcast (16/iter)"};
    const int per_iter[] = {16, 32, 16, 32, 16, 16};
    for (int mode = 0; mode < 6; ++mode)
        for (int wps : {1, 2, 4, 8}) {
            const int warps = 4 * wps;
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                case 0: k<0><<<148, 1024>>>(out, in, warps); break;
                case 1: k<1><<<148, 1024>>>(out, in, warps); break;
                case 2: k<2><<<148, 1024>>>(out, in, warps); break;
                case 3: k<3><<<148, 1024>>>(out, in, warps); break;
                case 4: k<4><<<148, 1024>>>(out, in, warps); break;
                case 5: k<5><<<148, 1024>>>(out, in, warps); break;
                }
                cudaDeviceSynchronize();
            }
            long long cyc;
            cudaMemcpy(&cyc, reinterpret_cast<long long *>(out + (1 << 20)), sizeof(cyc), cudaMemcpyDeviceToHost);
            // every sub-partition runs `wps` warps: cycles per warp-instruction per sub-partition
            printf("%-18s %d warps/SMSP: %.3f cycles per warp-instruction per SMSP  (%.1f lane-ops/clk/SM)\n", names[mode], wps,
                   (double)cyc / ((double)kIters * per_iter[mode] * wps),
                   4.0 * 32.0 * (mode == 0 || mode == 2 || mode == 4 || mode == 5 ? 2 : 1) * kIters * per_iter[mode] * wps / (double)cyc);
        }
    for (int bytes : {8, 16})
        for (int wps : {1, 2, 4, 8}) {
            const int warps = 4 * wps;
            for (int rep = 0; rep < 2; ++rep) {
                if (bytes == 16) lds<16><<<148, 1024>>>(out, warps); else lds<8><<<148, 1024>>>(out, warps);
                cudaDeviceSynchronize();
            }
            long long cyc;
            cudaMemcpy(&cyc, reinterpret_cast<long long *>(out + (1 << 20)), sizeof(cyc), cudaMemcpyDeviceToHost);
            printf("LDS.%d conflict-free %d warps/SMSP: %.2f cycles per warp-LDS per SM (%.1f B/clk/SM)\n", bytes * 8, wps,
                   (double)cyc / ((double)kIters * 8 * warps), 32.0 * bytes * kIters * 8 * warps / (double)cyc);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

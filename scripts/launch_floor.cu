// Diagnostic (GPU): what a back-to-back pair of small kernels with different shared-memory footprints costs,
// with and without a common shared-memory carve-out.  nvcc -arch=sm_100a -O3 scripts/launch_floor.cu -o /tmp/lf
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) ka(float *p, int iters)
{
    extern __shared__ float sm[];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    float a = sm[(threadIdx.x + 1) & 255];
    for (int i = 0; i < iters; ++i) a = a * 1.0001f + 0.5f;
    if (a == 12345.f) p[0] = a;
}
__global__ void __launch_bounds__(320, 3) kb(float *p, int iters)
{
    extern __shared__ float sm[];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    float a = sm[(threadIdx.x + 1) % 320];
    for (int i = 0; i < iters; ++i) a = a * 1.0001f + 0.5f;
    if (a == 12345.f) p[0] = a;
}
static float run(int na, int sa, int nb, int sb, int iters, int reps, float *d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) { ka<<<na, 256, sa>>>(d, iters); kb<<<nb, 320, sb>>>(d, iters); }
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) { ka<<<na, 256, sa>>>(d, iters); kb<<<nb, 320, sb>>>(d, iters); }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / reps;
}
int main()
{
    float *d; cudaMalloc(&d, 1024);
    cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int iters : {0, 2000}) {
        printf("iters %d\n", iters);
        printf("  A(3070 x 34KB) + B(439 x 72KB), default carve-out : %.1f us per pair\n", run(3070, 34816, 439, 73728, iters, 50, d));
        printf("  A(3070 x 34KB) + B(439 x 34KB)                     : %.1f us per pair\n", run(3070, 34816, 439, 34816, iters, 50, d));
        printf("  A(3070 x 1KB)  + B(439 x 1KB)                      : %.1f us per pair\n", run(3070, 1024, 439, 1024, iters, 50, d));
        cudaFuncSetAttribute(ka, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(kb, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        printf("  A(3070 x 34KB) + B(439 x 72KB), carve-out 100 both : %.1f us per pair\n", run(3070, 34816, 439, 73728, iters, 50, d));
        cudaFuncSetAttribute(ka, cudaFuncAttributePreferredSharedMemoryCarveout, -1);
        cudaFuncSetAttribute(kb, cudaFuncAttributePreferredSharedMemoryCarveout, -1);
        printf("  B alone (439 x 72KB) x2                            : %.1f us per pair\n", run(0 + 1, 1024, 439, 73728, iters, 50, d));
    }
    return 0;
}

// chroma_kernels.cu -- K-chroma: per-frame pitch-class energies of a dB spectrum (include/pvqt_analysis.h).
// Replaces pitchvis_viewer/src/display_system/update.rs:1104-1131.  One warp per frame: lane c < 12 owns pitch class c
// and adds its bins in ascending order (the reference's summation order per class), then the 12 sums are divided by
// their maximum.  Un-contracted f32 like the reference; powf is the device's (within an ulp or two of glibc's).
#include <cmath>
#include <string>

#include <cuda_runtime.h>

#include "last_error.hpp"
#include "pvqt_analysis.h"

namespace {

__global__ void __launch_bounds__(128) chroma_kernel(const float *db, float *out, unsigned n_frames, int n_buckets,
                                                     int buckets_per_octave, int pc0)
{
    const int lane = threadIdx.x & 31;
    const unsigned frame = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (frame >= n_frames) return;
    const float *x = db + (size_t)frame * n_buckets;
    float sum = 0.0f;
    if (lane < 12) {
        for (int bin = 0; bin < n_buckets; ++bin) {
            const int semitone = (int)roundf(__fdiv_rn((float)(bin * 12), (float)buckets_per_octave));
            if ((semitone + pc0) % 12 == lane) sum = __fadd_rn(sum, powf(10.0f, __fdiv_rn(x[bin], 10.0f)));
        }
    }
    float mx = sum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane < 12) out[(size_t)frame * 12 + lane] = mx > 0.0f ? __fdiv_rn(sum, mx) : sum;
}

int cfail(int st, const std::string &m)
{
    pvqt_detail::set_last_error(m);
    return st;
}
int ccuda(cudaError_t e, const char *what)
{
    pvqt_detail::set_last_error(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
    return PVQT_CUDA_ERROR;
}

}  // namespace

extern "C" {

int pvqt_chroma_device(const pvqt_range *range, int device, const float *d_db, size_t n_frames, float *d_out,
                       void *cuda_stream)
{
    if (!range || !d_db || !d_out) return cfail(PVQT_INVALID_ARGUMENT, "null argument");
    if (range->buckets_per_octave == 0 || range->octaves == 0) return cfail(PVQT_INVALID_ARGUMENT, "empty range");
    if (n_frames == 0) return PVQT_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return ccuda(e, "cudaSetDevice");
    // update.rs:1110-1112
    const float semitones_from_c4 = 12.0f * std::log2(range->min_freq / 261.626f);
    const int pc0 = (((int)std::round(semitones_from_c4) % 12) + 12) % 12;
    const int nb = (int)(range->buckets_per_octave * range->octaves);
    chroma_kernel<<<(unsigned)((n_frames + 3) / 4), 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_db, d_out, (unsigned)n_frames, nb, (int)range->buckets_per_octave, pc0);
    e = cudaGetLastError();
    return e == cudaSuccess ? PVQT_OK : ccuda(e, "launch chroma_kernel");
}

int pvqt_chroma(const pvqt_range *range, int device, const float *db, size_t n_frames, float *out)
{
    if (!range || !db || !out) return cfail(PVQT_INVALID_ARGUMENT, "null argument");
    if (n_frames == 0) return PVQT_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return ccuda(e, "cudaSetDevice");
    const size_t nb = (size_t)range->buckets_per_octave * range->octaves;
    float *d_db = nullptr, *d_out = nullptr;
    if ((e = cudaMalloc(&d_db, n_frames * nb * sizeof(float))) != cudaSuccess) return ccuda(e, "cudaMalloc");
    if ((e = cudaMalloc(&d_out, n_frames * 12 * sizeof(float))) != cudaSuccess) { cudaFree(d_db); return ccuda(e, "cudaMalloc"); }
    int rc = PVQT_OK;
    if ((e = cudaMemcpy(d_db, db, n_frames * nb * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) rc = ccuda(e, "copy in");
    if (rc == PVQT_OK) rc = pvqt_chroma_device(range, device, d_db, n_frames, d_out, nullptr);
    if (rc == PVQT_OK && (e = cudaMemcpy(out, d_out, n_frames * 12 * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess)
        rc = ccuda(e, "copy out");
    cudaFree(d_db);
    cudaFree(d_out);
    return rc;
}

}  // extern "C"

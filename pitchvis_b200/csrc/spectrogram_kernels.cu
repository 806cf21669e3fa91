// spectrogram_kernels.cu -- K-spectrogram: the viewer's spectrogram ring, SpectrogramMode::VQT (include/pvqt_analysis.h).
// Replaces pitchvis_viewer/src/display_system/update.rs:930-1088 for T frames at once: the image ends as if the
// reference had processed the frames one by one -- the last min(T, height - 1) frames own a row each (row
// height - 1 - index), the row after the last one is cleared.  One CTA per surviving row: frame maximum by a block
// reduction (update.rs:965), then one uchar4 per bin (coalesced 4-byte stores).  Un-contracted f32 like the reference.
#include <algorithm>
#include <cmath>
#include <string>

#include <cuda_runtime.h>

#include "last_error.hpp"
#include "pvqt_analysis.h"

namespace {

__device__ __forceinline__ unsigned char to_u8(float x)   // `(x).clamp(0.0, 255.0) as u8`
{
    if (!(x > 0.0f)) return 0;
    if (x >= 255.0f) return 255;
    return (unsigned char)x;
}

// update.rs:965-975 + :989
__device__ __forceinline__ unsigned char alpha_byte(float value_db, float max_val)
{
    float brightness = 0.0f;
    if (max_val > 0.0f) {
        const float normalized = __fdiv_rn(value_db, __fadd_rn(max_val, 0.001f));
        const float d = __fsub_rn(1.0f, normalized);
        brightness = __fmul_rn(__fsub_rn(1.0f, __fmul_rn(d, d)), 1.5f);
        brightness = brightness < 0.0f ? 0.0f : (brightness > 1.0f ? 1.0f : brightness);
        if (brightness != brightness) brightness = 0.0f;
    }
    return to_u8(__fmul_rn(__fmul_rn(brightness, 255.0f), 1.2f));
}

// blockIdx.x < n_rows: frame first_frame + blockIdx.x into its ring row; blockIdx.x == n_rows: clear the next row
__global__ void __launch_bounds__(256) spectrogram_vqt_kernel(const float *smoothed, const unsigned char *bin_rgb, uchar4 *image,
                                                              unsigned first_frame, unsigned n_rows, unsigned width, unsigned height,
                                                              unsigned index0 /* ring index of frame first_frame */)
{
    __shared__ float warp_max[8];
    const unsigned j = blockIdx.x;
    const unsigned idx = (unsigned)(((unsigned long long)index0 + j) % height);
    uchar4 *row = image + (size_t)(height - 1 - idx) * width;
    if (j == n_rows) {
        for (unsigned b = threadIdx.x; b < width; b += blockDim.x) row[b] = make_uchar4(0, 0, 0, 0);
        return;
    }
    const float *x = smoothed + (size_t)(first_frame + j) * width;
    float mx = 0.0f;
    for (unsigned b = threadIdx.x; b < width; b += blockDim.x) mx = fmaxf(mx, x[b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = warp_max[0];
    for (unsigned w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, warp_max[w]);
    for (unsigned b = threadIdx.x; b < width; b += blockDim.x)
        row[b] = make_uchar4(bin_rgb[3 * b], bin_rgb[3 * b + 1], bin_rgb[3 * b + 2], alpha_byte(x[b], mx));
}

int sfail(int st, const std::string &m)
{
    pvqt_detail::set_last_error(m);
    return st;
}
int scuda(cudaError_t e, const char *what)
{
    pvqt_detail::set_last_error(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
    return PVQT_CUDA_ERROR;
}

}  // namespace

extern "C" {

int pvqt_spectrogram_vqt_device(int device, const float *d_smoothed, size_t n_frames, size_t n_buckets, const uint8_t *d_bin_rgb,
                                uint8_t *d_image, size_t height, size_t *write_index, void *cuda_stream)
{
    if (!d_smoothed || !d_bin_rgb || !d_image || !write_index) return sfail(PVQT_INVALID_ARGUMENT, "null argument");
    if (n_buckets == 0 || height == 0 || *write_index >= height || n_buckets > 0xffffffffull || height > 0x7fffffffull ||
        n_frames > 0xffffffffull)
        return sfail(PVQT_INVALID_ARGUMENT, "spectrogram: empty image, write_index outside the ring or sizes beyond 32 bits");
    if ((reinterpret_cast<uintptr_t>(d_image) & 3) != 0) return sfail(PVQT_INVALID_ARGUMENT, "spectrogram: image must be 4-byte aligned");
    if (n_frames == 0) return PVQT_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return scuda(e, "cudaSetDevice");
    const size_t n_rows = std::min(n_frames, height - 1);      // frames whose row survives
    const size_t first = n_frames - n_rows;
    const size_t index0 = (*write_index + first) % height;
    spectrogram_vqt_kernel<<<(unsigned)(n_rows + 1), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_smoothed, d_bin_rgb, reinterpret_cast<uchar4 *>(d_image), (unsigned)first, (unsigned)n_rows, (unsigned)n_buckets,
        (unsigned)height, (unsigned)index0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return scuda(e, "launch spectrogram_vqt_kernel");
    *write_index = (*write_index + n_frames) % height;
    return PVQT_OK;
}

int pvqt_spectrogram_vqt(int device, const float *smoothed, size_t n_frames, size_t n_buckets, const uint8_t *bin_rgb, uint8_t *image,
                         size_t height, size_t *write_index)
{
    if (!smoothed || !bin_rgb || !image || !write_index) return sfail(PVQT_INVALID_ARGUMENT, "null argument");
    if (n_buckets == 0 || height == 0 || *write_index >= height)
        return sfail(PVQT_INVALID_ARGUMENT, "spectrogram: empty image or write_index outside the ring");
    if (n_frames == 0) return PVQT_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return scuda(e, "cudaSetDevice");
    // only the frames whose row survives travel to the device
    const size_t n_rows = std::min(n_frames, height - 1), first = n_frames - n_rows;
    const size_t img_bytes = height * n_buckets * 4;
    float *d_x = nullptr;
    uint8_t *d_rgb = nullptr, *d_img = nullptr;
    int rc = PVQT_OK;
    if ((e = cudaMalloc(&d_x, std::max<size_t>(n_rows, 1) * n_buckets * sizeof(float))) != cudaSuccess) rc = scuda(e, "cudaMalloc");
    if (rc == PVQT_OK && (e = cudaMalloc(&d_rgb, n_buckets * 3)) != cudaSuccess) rc = scuda(e, "cudaMalloc");
    if (rc == PVQT_OK && (e = cudaMalloc(&d_img, img_bytes)) != cudaSuccess) rc = scuda(e, "cudaMalloc");
    if (rc == PVQT_OK && n_rows > 0 &&
        (e = cudaMemcpy(d_x, smoothed + first * n_buckets, n_rows * n_buckets * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
        rc = scuda(e, "copy in");
    if (rc == PVQT_OK && (e = cudaMemcpy(d_rgb, bin_rgb, n_buckets * 3, cudaMemcpyHostToDevice)) != cudaSuccess) rc = scuda(e, "copy in");
    if (rc == PVQT_OK && (e = cudaMemcpy(d_img, image, img_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) rc = scuda(e, "copy in");
    if (rc == PVQT_OK) {
        size_t w = (*write_index + first) % height;   // the skipped frames only move the index
        rc = pvqt_spectrogram_vqt_device(device, d_x, n_rows, n_buckets, d_rgb, d_img, height, &w, nullptr);
        if (rc == PVQT_OK && n_rows == 0) {           // height 1: every frame's row is cleared again
            if ((e = cudaMemset(d_img, 0, img_bytes)) != cudaSuccess) rc = scuda(e, "clear");
        }
        if (rc == PVQT_OK && (e = cudaMemcpy(image, d_img, img_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = scuda(e, "copy out");
        if (rc == PVQT_OK) *write_index = (*write_index + n_frames) % height;
    }
    cudaFree(d_x);
    cudaFree(d_rgb);
    cudaFree(d_img);
    return rc;
}

}  // extern "C"

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

// ---- SpectrogramMode::Peaks (update.rs:997-1062) -------------------------------------------------------------------
// pitchvis_colors::calculate_color (pitchvis_colors/src/lib.rs:93-119) runs per peak; it goes through the `lab` crate
// 0.11.0 (absent from the reference tree): sRGB <-> XYZ <-> L*a*b* <-> LCh restated from the crate's published formulas.
// f32 like the crate; the transcendentals are evaluated in f64 and rounded once, which gives libm's (correctly
// rounded) f32 results -- the same device as in analysis_kernels.cu.
__constant__ float c_colors[12][3] = {   // pitchvis_colors/src/lib.rs:19-34
    {0.85f, 0.36f, 0.36f}, {0.01f, 0.52f, 0.71f}, {0.97f, 0.76f, 0.05f}, {0.45f, 0.34f, 0.63f},
    {0.47f, 0.77f, 0.22f}, {0.78f, 0.32f, 0.52f}, {0.00f, 0.64f, 0.56f}, {0.95f, 0.54f, 0.23f},
    {0.30f, 0.37f, 0.64f}, {1.00f, 0.96f, 0.03f}, {0.57f, 0.30f, 0.55f}, {0.12f, 0.71f, 0.34f}};
constexpr float kGrayLevel = 60.0f, kEasingPow = 1.3f;   // lib.rs:56-57
constexpr float kKappa = 24389.0f / 27.0f, kEpsilon = 216.0f / 24389.0f, kCbrtEpsilon = 6.0f / 29.0f;
constexpr float kS0 = 0.003130668442500564f, kWhiteX = 0.9504492182750991f, kWhiteZ = 1.0889166484304715f;

__device__ __forceinline__ float cr_powf(float a, float b) { return (float)pow((double)a, (double)b); }
__device__ __forceinline__ float srgb_to_linear_255(float c)
{
    const float e0_255 = 12.92f * kS0 * 255.0f;
    if (c > e0_255) return cr_powf((c + 0.055f * 255.0f) / (1.055f * 255.0f), 2.4f);
    return c / (12.92f * 255.0f);
}
__device__ __forceinline__ float xyz_to_lab_map(float c)
{
    return c > kEpsilon ? cr_powf(c, 1.0f / 3.0f) : (kKappa * c + 16.0f) / 116.0f;
}
__device__ __forceinline__ float linear_to_srgb(float c)
{
    float v = c > kS0 ? 1.055f * cr_powf(c, 1.0f / 2.4f) - 0.055f : 12.92f * c;
    return fmaxf(fminf(v, 1.0f), 0.0f);
}
// rgb bytes of the pixel, already scaled as update.rs:1046-1051 does: (c * 255 * 1.2).clamp(0, 255) as u8
__device__ uchar4 peak_color(unsigned bpo, float bucket)
{
    const float pitch_continuous = 12.0f * bucket / (float)bpo;
    const float rounded = roundf(pitch_continuous);
    const unsigned long long ri = rounded > 0.0f ? (rounded >= 1.8446744e19f ? 0xffffffffffffffffull : (unsigned long long)rounded) : 0ull;
    const int semitone = (int)(ri % 12ull);
    const float inaccuracy_cents = fabsf(pitch_continuous - rounded);
    const float r = srgb_to_linear_255((float)to_u8(c_colors[semitone][0] * 255.0f));
    const float g = srgb_to_linear_255((float)to_u8(c_colors[semitone][1] * 255.0f));
    const float b = srgb_to_linear_255((float)to_u8(c_colors[semitone][2] * 255.0f));
    const float X = r * 0.4124108464885388f + g * 0.3575845678529519f + b * 0.18045380393360833f;
    const float Y = r * 0.21264934272065283f + g * 0.7151691357059038f + b * 0.07218152157344333f;
    const float Z = r * 0.019331758429150258f + g * 0.11919485595098397f + b * 0.9503900340503373f;
    const float fx = xyz_to_lab_map(X / kWhiteX), fy = xyz_to_lab_map(Y), fz = xyz_to_lab_map(Z / kWhiteZ);
    float L = 116.0f * fy - 16.0f;
    const float A = 500.0f * (fx - fy), B = 200.0f * (fy - fz);
    float C = (float)hypot((double)A, (double)B);
    const float H = (float)atan2((double)B, (double)A);
    const float saturation = 1.0f - cr_powf(2.0f * inaccuracy_cents, kEasingPow);
    C *= saturation;
    L = saturation * L + (1.0f - saturation) * kGrayLevel;
    const float a2 = C * (float)cos((double)H), b2 = C * (float)sin((double)H);
    const float gy = (L + 16.0f) / 116.0f, gx = a2 / 500.0f + gy, gz = gy - b2 / 200.0f;
    const float xr = gx > kCbrtEpsilon ? gx * gx * gx : (gx * 116.0f - 16.0f) / kKappa;
    const float yr = L > kEpsilon * kKappa ? gy * gy * gy : L / kKappa;
    const float zr = gz > kCbrtEpsilon ? gz * gz * gz : (gz * 116.0f - 16.0f) / kKappa;
    const float x = xr * kWhiteX, y = yr, z = zr * kWhiteZ;
    const float lr = x * 3.240812398895283f - y * 1.5373084456298136f - z * 0.4985865229069666f;
    const float lg = x * -0.9692430170086407f + y * 1.8759663029085742f + z * 0.04155503085668564f;
    const float lb = x * 0.055638398436112804f - y * 0.20400746093241362f + z * 1.0571295702861434f;
    const float cr = (float)to_u8(roundf(linear_to_srgb(lr) * 255.0f)) / 255.0f;
    const float cg = (float)to_u8(roundf(linear_to_srgb(lg) * 255.0f)) / 255.0f;
    const float cb = (float)to_u8(roundf(linear_to_srgb(lb) * 255.0f)) / 255.0f;
    return make_uchar4(to_u8(cr * 255.0f * 1.2f), to_u8(cg * 255.0f * 1.2f), to_u8(cb * 255.0f * 1.2f), 0);
}

constexpr int kMaxRowPeaks = 256;   // K-analysis keeps at most this many peaks per frame
// One CTA per surviving row (blockIdx.x < n_rows) + one that clears the row after the last.  The reference draws the
// peaks of a frame in list order, later peaks overwriting earlier ones: every bin takes the LAST peak that covers it.
__global__ void __launch_bounds__(256) spectrogram_peaks_kernel(const pvqt_continuous_peak *peaks, const unsigned *peak_count,
                                                                unsigned max_peaks, uchar4 *image, unsigned first_frame,
                                                                unsigned n_rows, unsigned width, unsigned height, unsigned index0,
                                                                unsigned bpo, int keep_first_row)
{
    __shared__ float s_center[kMaxRowPeaks], s_bright[kMaxRowPeaks];
    __shared__ uchar4 s_rgb[kMaxRowPeaks];
    __shared__ float warp_max[8];
    const unsigned j = blockIdx.x;
    const unsigned idx = (unsigned)(((unsigned long long)index0 + j) % height);
    uchar4 *row = image + (size_t)(height - 1 - idx) * width;
    if (j == n_rows) {
        for (unsigned b = threadIdx.x; b < width; b += blockDim.x) row[b] = make_uchar4(0, 0, 0, 0);
        return;
    }
    const unsigned frame = first_frame + j;
    const unsigned n = min(min(peak_count[frame], max_peaks), (unsigned)kMaxRowPeaks);
    const pvqt_continuous_peak *pk = peaks + (size_t)frame * max_peaks;
    float mx = 0.0f;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, pk[i].size);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = warp_max[0];
    for (unsigned w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, warp_max[w]);
    const float semitone_offset = (float)(bpo - 3 * (bpo / 12));
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const float center = pk[i].center, size = pk[i].size;
        const float d = 1.0f - size / mx;
        float brightness = (1.0f - d * d) * 1.5f;
        brightness = brightness < 0.0f ? 0.0f : (brightness > 1.0f ? 1.0f : brightness);
        s_center[i] = center;
        s_bright[i] = brightness;
        s_rgb[i] = peak_color(bpo, fmodf(center + semitone_offset, (float)bpo));
    }
    __syncthreads();
    // a row written for the first time in this call was cleared by the step before it (update.rs:1065-1078); inside a
    // batch that step is part of this launch, so the row starts from zeros -- except the very first row of the call
    const bool clear = !(keep_first_row && j == 0);
    const float radius = 2.0f;
    for (unsigned b = threadIdx.x; b < width; b += blockDim.x) {
        int hit = -1;
        if (mx > 0.0f) {
            for (unsigned i = 0; i < n; ++i) {
                const float c = s_center[i];
                const float lo = fmaxf(floorf(c - radius), 0.0f), hi = fminf(ceilf(c + radius), (float)width);
                if ((float)b >= lo && (float)b < hi && fabsf((float)b - c) <= radius) hit = (int)i;
            }
        }
        if (hit >= 0) {
            const float distance = fabsf((float)b - s_center[hit]);
            const float falloff = (float)exp((double)(-distance * distance / (radius * radius * 0.5f)));
            uchar4 px = s_rgb[hit];
            px.w = to_u8(s_bright[hit] * falloff * 255.0f * 1.2f);
            row[b] = px;
        } else if (clear) {
            row[b] = make_uchar4(0, 0, 0, 0);
        }
    }
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

int pvqt_spectrogram_peaks_device(int device, const pvqt_range *range, const pvqt_continuous_peak *d_peaks,
                                  const uint32_t *d_peak_count, size_t max_peaks, size_t n_frames, uint8_t *d_image, size_t height,
                                  size_t *write_index, void *cuda_stream)
{
    if (!range || !d_peaks || !d_peak_count || !d_image || !write_index) return sfail(PVQT_INVALID_ARGUMENT, "null argument");
    const size_t width = (size_t)range->octaves * range->buckets_per_octave;
    if (width == 0 || range->buckets_per_octave < 12 || height == 0 || *write_index >= height || width > 0xffffffffull ||
        height > 0x7fffffffull || n_frames > 0xffffffffull || max_peaks == 0 || max_peaks > 0xffffffffull)
        return sfail(PVQT_INVALID_ARGUMENT, "spectrogram: empty image, write_index outside the ring or sizes beyond 32 bits");
    if ((reinterpret_cast<uintptr_t>(d_image) & 3) != 0) return sfail(PVQT_INVALID_ARGUMENT, "spectrogram: image must be 4-byte aligned");
    if (n_frames == 0) return PVQT_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return scuda(e, "cudaSetDevice");
    const size_t n_rows = std::min(n_frames, height - 1);
    const size_t first = n_frames - n_rows;
    const size_t index0 = (*write_index + first) % height;
    spectrogram_peaks_kernel<<<(unsigned)(n_rows + 1), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        d_peaks, d_peak_count, (unsigned)max_peaks, reinterpret_cast<uchar4 *>(d_image), (unsigned)first, (unsigned)n_rows,
        (unsigned)width, (unsigned)height, (unsigned)index0, range->buckets_per_octave, first == 0 ? 1 : 0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return scuda(e, "launch spectrogram_peaks_kernel");
    *write_index = (*write_index + n_frames) % height;
    return PVQT_OK;
}

int pvqt_spectrogram_peaks(int device, const pvqt_range *range, const pvqt_continuous_peak *peaks, const uint32_t *peak_count,
                           size_t max_peaks, size_t n_frames, uint8_t *image, size_t height, size_t *write_index)
{
    if (!range || !peaks || !peak_count || !image || !write_index) return sfail(PVQT_INVALID_ARGUMENT, "null argument");
    const size_t width = (size_t)range->octaves * range->buckets_per_octave;
    if (width == 0 || height == 0 || *write_index >= height || max_peaks == 0)
        return sfail(PVQT_INVALID_ARGUMENT, "spectrogram: empty image or write_index outside the ring");
    if (n_frames == 0) return PVQT_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return scuda(e, "cudaSetDevice");
    const size_t img_bytes = height * width * 4;
    pvqt_continuous_peak *d_pk = nullptr;
    uint32_t *d_cnt = nullptr;
    uint8_t *d_img = nullptr;
    int rc = PVQT_OK;
    if ((e = cudaMalloc(&d_pk, n_frames * max_peaks * sizeof(pvqt_continuous_peak))) != cudaSuccess) rc = scuda(e, "cudaMalloc");
    if (rc == PVQT_OK && (e = cudaMalloc(&d_cnt, n_frames * sizeof(uint32_t))) != cudaSuccess) rc = scuda(e, "cudaMalloc");
    if (rc == PVQT_OK && (e = cudaMalloc(&d_img, img_bytes)) != cudaSuccess) rc = scuda(e, "cudaMalloc");
    if (rc == PVQT_OK && (e = cudaMemcpy(d_pk, peaks, n_frames * max_peaks * sizeof(pvqt_continuous_peak), cudaMemcpyHostToDevice)) != cudaSuccess)
        rc = scuda(e, "copy in");
    if (rc == PVQT_OK && (e = cudaMemcpy(d_cnt, peak_count, n_frames * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) rc = scuda(e, "copy in");
    if (rc == PVQT_OK && (e = cudaMemcpy(d_img, image, img_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) rc = scuda(e, "copy in");
    if (rc == PVQT_OK) {
        size_t w = *write_index;
        rc = pvqt_spectrogram_peaks_device(device, range, d_pk, d_cnt, max_peaks, n_frames, d_img, height, &w, nullptr);
        if (rc == PVQT_OK && height == 1 && (e = cudaMemset(d_img, 0, img_bytes)) != cudaSuccess) rc = scuda(e, "clear");
        if (rc == PVQT_OK && (e = cudaMemcpy(image, d_img, img_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = scuda(e, "copy out");
        if (rc == PVQT_OK) *write_index = w;
    }
    cudaFree(d_pk);
    cudaFree(d_cnt);
    cudaFree(d_img);
    return rc;
}

}  // extern "C"

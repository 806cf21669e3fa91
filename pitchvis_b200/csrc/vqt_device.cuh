// vqt_device.cuh -- device-side descriptors shared by the kernels and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pvqt_dev {

constexpr int kMaxGroups = 8;          // window groups per Vqt (4 at the defaults, 5 hi-res)
constexpr int kMaxFftPasses = 4;       // radix passes of the largest plan (N_c = 16384)
constexpr int kPointsPerThread = 16;   // complex points each FFT thread keeps in registers
constexpr int kSpmmRowsPerBlock = 32;  // rows of one sliced-ELL block (one row per lane)

// One window group (WindowGroup, vqt.rs:388-404) as the FFT kernel sees it.
struct FftGroup {
    int32_t  window_begin;   // first sample of the window inside an n_fft frame
    int32_t  log2_nc;        // N_c = window_size / 2 complex points
    int32_t  col_lo, col_hi; // consumed real-FFT bins [col_lo, col_hi]
    int32_t  spec_offset;    // position of col_lo inside a frame's spectrum row
    int32_t  cta_begin;      // first CTA of this group in the fused launch
    int32_t  frames_per_cta;
    int32_t  _pad;
    const float2 *twiddle[kMaxFftPasses];  // per-pass tables, [ (r-1)*Ns + k ]
    const float2 *split_twiddle;           // exp(-2 pi i c / N), c = col_lo..col_hi
};

// Where frame f lives: audio[(f / frames_per_stream) * stream_stride + (f % frames_per_stream) * hop]
struct FrameLayout {
    const float *audio;
    uint64_t stream_stride;
    uint64_t hop;
    uint32_t frames_per_stream;
    uint32_t n_frames;        // frames in this launch
    uint64_t first_frame;     // global index of local frame 0 (for addressing `audio`)
};

struct FftParams {
    FftGroup    group[kMaxGroups];
    int32_t     n_groups;
    int32_t     spec_stride;  // complex elements per frame in `spec`
    FrameLayout frames;
    float2     *spec;         // [n_frames][spec_stride]
};

// Sliced-ELL, zero-filled band layout of the spectral kernel (all groups concatenated,
// rows in ascending frequency = output order).
struct SpmmBlock {
    int32_t val_base;   // first ELL slot of the positive band  (slot = 32 float2)
    int32_t width;      // slots in the positive band
    int32_t nval_base;  // first ELL slot of the conjugate-part band
    int32_t nwidth;     // slots in the conjugate-part band (0 for most blocks)
};

struct SpmmParams {
    const SpmmBlock *blocks;     // n_blocks
    const int2      *row_cols;   // per padded row: (first column of band, first column of conj band)
    const float2    *values;     // ELL slots
    int32_t  n_blocks;
    int32_t  n_buckets;
    int32_t  spec_stride;
    uint32_t n_frames;
    const float2 *spec;          // [n_frames][spec_stride]
    float   *out_db;             // [n_frames][n_buckets]
    float   *out_power;          // optional
    float    ref_db;             // 10 * log10(0.3 * 0.3), vqt.rs:923,927
};

cudaError_t launch_fft(const FftParams &p, int total_ctas, int block_threads, cudaStream_t stream);
cudaError_t launch_spmm_db(const SpmmParams &p, int frames_per_cta, cudaStream_t stream);
cudaError_t configure_kernels(int spec_stride, int n_buckets, int *spmm_frames_per_cta);
size_t fft_smem_bytes(int block_threads);
size_t spmm_smem_bytes(int frames_per_cta, int spec_stride, int n_buckets);

}  // namespace pvqt_dev

// vqt_device.cuh -- device-side descriptors shared by the kernels and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pvqt_dev {

constexpr int kMaxGroups = 8;          // window groups per Vqt (4 at the defaults, 5 hi-res)
constexpr int kMaxSdft = 2;            // window groups on the sliding partial-DFT path per launch
constexpr int kMaxFftPasses = 4;       // radix passes of the largest plan (N_c = 16384)
constexpr int kPointsPerThread = 16;   // complex points each FFT thread keeps in registers
constexpr int kTileFrames = 8;         // frames per spectrum tile (one 64-byte record per column)
constexpr int kRowsPerLane = 2;        // adjacent kernel rows one SpMM lane accumulates
constexpr int kRowsPerBlock = 32 * kRowsPerLane;

// Spectrum scratch layout ("tiled, planar"): [tile = frame / 8][column][re f0..f7 | im f0..f7] floats,
// i.e. one 64-byte record per (tile, column) holding the real parts of the tile's 8 frames followed by
// the imaginary parts.  SpMM lanes fetch a record with four LDS.128 and feed frame *pairs* to packed
// FFMA2s with the kernel coefficient as the broadcast scalar -- no register shuffling.  The four
// 16-byte chunks of a record are XOR-swizzled with bits 1..2 of the column so that lanes reading
// consecutive columns never collide on a bank.  Position of (frame f, column c), in floats:
__host__ __device__ inline size_t spec_index_re(size_t frame, int col, int spec_stride)
{
    const size_t tile = frame / kTileFrames;
    const int fi = (int)(frame % kTileFrames);
    const int chunk = (fi >> 2) ^ ((col >> 1) & 3);              // re: logical chunks 0,1; im: 2,3
    return (tile * spec_stride + col) * (2 * kTileFrames) + (chunk << 2) + (fi & 3);
}
__host__ __device__ inline size_t spec_index_im(size_t frame, int col, int spec_stride)
{
    const size_t tile = frame / kTileFrames;
    const int fi = (int)(frame % kTileFrames);
    const int chunk = (2 + (fi >> 2)) ^ ((col >> 1) & 3);
    return (tile * spec_stride + col) * (2 * kTileFrames) + (chunk << 2) + (fi & 3);
}

// Spectrum scratch layout "planes" (the pipeline form of K-spmm-db, spmm_pipe.cu): [tile][chunk 0..3][column < plane_stride]
// [4 floats] -- chunk 0,1 = Re of frames 0-3, 4-7; 2,3 = Im -- exactly the shared-memory image K-spmm-db walks, so a
// tile is staged with ONE bulk copy (cp.async.bulk) instead of 3300 16-byte cp.asyncs.  Columns nobody writes stay zero.
__host__ __device__ inline size_t plane_index(size_t frame, int col, int plane_stride, int imag)
{
    const size_t tile = frame / kTileFrames;
    const int fi = (int)(frame % kTileFrames);
    return ((tile * 4 + (size_t)(2 * imag + (fi >> 2))) * plane_stride + col) * 4 + (fi & 3);
}

// One window group (WindowGroup, vqt.rs:388-404) as the FFT kernel sees it.
struct FftGroup {
    int32_t  window_begin;   // first sample of the window inside an n_fft frame
    int32_t  log2_nc;        // N_c = window_size / 2 complex points
    int32_t  col_lo, col_hi; // consumed real-FFT bins [col_lo, col_hi]
    int32_t  spec_offset;    // position of col_lo inside a frame's spectrum row
    int32_t  cta_begin;      // first CTA of this group in the fused launch
    int32_t  frames_per_cta; // frames one work item (one trip of a CTA) transforms
    int32_t  n_ctas;         // CTAs of this group; CTA c works on items c, c + n_ctas, c + 2 n_ctas, ...
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
    uint32_t first_stream;    // first_frame = first_stream * frames_per_stream + first_t (32-bit arithmetic in the kernels)
    uint32_t first_t;
};

// Sliding partial-DFT path of one window group ("K-sdft").  When consecutive frames overlap (hop H much
// smaller than the window N) and the sparse kernel consumes only a few low bins of the group's FFT,
// the consumed bins are cheaper as sums of hop-sized partial DFTs that neighbouring frames share:
//   N = q H + rem,   X_t[k] = sum_{i<q} e^{-2 pi i k i H / N} C[t+i][k] + e^{-2 pi i k q H / N} R[t+q][k]
//   C[c][k] = sum_{j<H}   x[c H + window_begin + j] e^{-2 pi i k j / N}     (one chunk, computed once)
//   R[c][k] = sum_{j<rem} x[c H + window_begin + j] e^{-2 pi i k j / N}     (prefix of the same sum)
// Same unnormalised forward DFT as the FFT path (vqt.rs:884-887 / :1087-1128), evaluated in another
// order; each sample is read once instead of q times.  The chunk twiddle is split in two levels,
// e^{-2 pi i k (16 a + b) / N} = A[a][k] B[b][k], so the 16 B's stay in registers.
struct SdftGroup {
    int32_t window_begin;   // first sample of the window inside an n_fft frame
    int32_t n_window;       // N
    int32_t k_lo, nk;       // consumed bins k_lo .. k_lo + nk - 1
    int32_t spec_offset;    // position of k_lo inside a frame's spectrum row
    int32_t hop, q, rem;    // N = q * hop + rem
    int32_t n_blocks;       // ceil(hop / 16)
    int32_t hop_pad;        // 16 * n_blocks
    const float2 *tw_a;     // [n_blocks][nk]
    const float2 *tw_b;     // [16][nk]
    const float2 *phase;    // [q + 1][nk]
};

struct SdftParams {
    SdftGroup    g;
    const float *audio;
    uint64_t     stream_stride;
    uint64_t     valid_samples;    // samples every stream holds: (frames_per_stream - 1) * hop + n_fft
    uint32_t     first_stream;     // streams [first_stream, first_stream + n_streams) ...
    uint32_t     n_streams;
    uint32_t     first_frame;      // ... frames [first_frame, first_frame + frames) of each
    uint32_t     frames;
    uint32_t     rows_per_stream;  // chunk rows per stream: frames + q
    int32_t      spec_stride;
    float2      *partial_c;        // [n_streams * rows_per_stream][nk]
    float2      *partial_r;
    float       *spec;             // tiled planar layout; local frame = stream_local * frames + t
    unsigned    *done_counter;     // optional: every CTA of the partial-sum kernel adds 1 when its sums are written
                                   // (K-spmm-db starts its combine step on it, before K-fft has finished)
};

// Plan of the tcgen05 form of the partial sums (sdft_tc_kernel.cu): accumulation groups of 16-sample blocks.
constexpr int kTcMaxGroups = 32;
struct SdftTcPlan {
    const float2 *tw_b;        // [16 * group16][nk]: e^{-2 pi i k b / N}, b inside a group
    const float2 *tw_g;        // [n_groups][nk]:     e^{-2 pi i k s_g / N}, s_g the group's first sample
    int32_t group16;           // 16-sample blocks per full group (1, 2 or 4)
    int32_t n_groups;
    int32_t r_groups;          // the first r_groups groups cover exactly the first `rem` samples (0: rem = 0)
    uint8_t start16[kTcMaxGroups], len16[kTcMaxGroups];
};

struct FftParams {
    FftGroup    group[kMaxGroups];
    int32_t     n_groups;
    int32_t     spec_stride;  // columns per tile (multiple of 8)
    FrameLayout frames;
    float      *spec;         // tiled planar layout, ceil(n_frames / 8) tiles
    int32_t     plane_stride; // > 0: write the "planes" layout with this many columns per plane (spec_stride unused)
    int32_t     wait_prior;   // 1: launched programmatically behind K-sdft, wait for it before exiting
    int32_t     n_sdft;       // K-sdft groups whose combine step the CTAs of FFT group `combine_group` run
    int32_t     combine_group;
    unsigned   *tile_ready;   // optional [ceil(n_frames / 8)]: every CTA adds 1 to each tile it has written its frames of
                              // (K-spmm-db starts on a tile's count instead of on the whole grid; it resets the count)
    SdftParams  sdft[kMaxSdft];     // (see sdft_combine.cuh)
};

// shared memory of the stand-alone combine kernel: the chunk rows of one 8-frame tile
__host__ __device__ inline size_t sdft_combine_smem_bytes_dev(int q, int nk)
{
    return (size_t)(2 * q + 2 * kTileFrames + 2) * nk * sizeof(float2);  // C rows, R rows, phase table, slack
}

constexpr int kSdftThreads = 256;     // 4 row groups x 2 bin halves (64 bins per CTA)
constexpr int kSdftRowsPerWarp = 4;   // chunk rows a lane accumulates concurrently
constexpr int kSdftRowsPerCta = 16;   // few, fat CTAs: all resident at once, so K-fft can start beside them

// The spectral kernel as the SpMM sees it: blocks of 64 consecutive rows of one window group, two
// adjacent rows per lane.  A lane's two rows share one zero-filled band of columns
// [pair_col0, pair_col0 + pair_len); slot j of a block holds, for each lane, the float4
// (K[row0, c].re, K[row0, c].im, K[row1, c].re, K[row1, c].im) with c = pair_col0 + j.
struct SpmmRowBlock {
    int32_t first_row;   // output row of lane 0's first row
    int32_t n_rows;      // valid rows in this block (<= 64)
    int32_t width;       // slots of the band
    int32_t nwidth;      // slots of the conjugate-part band
    int32_t val_base;    // first slot of the band in `values` (slot = 32 float4)
    int32_t nval_base;   // first slot of the conjugate-part band
    int32_t col_lo;      // first spectrum column staged for this block (multiple of 8)
    int32_t n_cols;      // columns staged
};

struct SpmmParams {
    const SpmmRowBlock *blocks;
    const int4   *lane_meta;   // [block][lane]: (pair_col0 - col_lo, pair_len, npair_col0 - col_lo, npair_len)
    const float4 *values;
    int32_t  n_blocks;
    int32_t  n_buckets;
    int32_t  spec_stride;
    int32_t  max_cols;         // largest n_cols over blocks (sizes the per-warp staging buffer)
    const int32_t *block_order; // row blocks, widest band first (longest-processing-time-first dispatch)
    uint32_t n_frames;
    uint32_t n_tiles;
    const float  *spec;        // tiled planar layout
    float   *power;            // [n_frames][n_buckets]: |z|^2 (norm_sqr, vqt.rs:930)
};

struct DbParams {
    const float *power;        // [n_frames][n_buckets]
    float   *out_db;           // [n_frames][n_buckets]
    uint32_t n_frames;
    int32_t  n_buckets;
    float    ref_db;           // 10 * log10(0.3 * 0.3), vqt.rs:923,927
};

// Fused K-spmm-db: one CTA per 8-frame tile owns every kernel row, so power_to_db's frame-wise
// max / min never leave the SM.  The row pairs ("units") are sorted by band length and dealt to
// warps, so the lanes of a warp run (almost) the same number of band slots; inside a warp the units
// are placed so that the spectrum records of a quarter-warp fall on distinct bank groups.
struct FusedWarp {
    int32_t width;       // band slots the warp walks (the longest unit of the warp)
    int32_t nwidth;      // conjugate-part slots
    int32_t val_base;    // first slot in `values` (slot = 32 float4)
    int32_t nval_base;
};

struct FusedParams {
    FusedWarp warp[32];        // per-warp walk descriptors, by value: the coefficient prefetch at kernel start must not
                               // wait for a (cold) global load
    const int4   *lane_meta;   // [warp][lane]: (col0, len, ncol0, nlen) in spectrum columns
    const int2   *lane_rows;   // [warp][lane]: (first output row, valid rows 0..rows_per_lane)
    const float4 *values;      // [copy][slot][row pair][lane]: (K[row0].re, K[row0].im, K[row1].re, K[row1].im)
    uint32_t values_stride;    // float4 entries per copy
    uint32_t values_copies;    // identical copies: CTA b streams copy b % copies (see build_fused_plan)
    int32_t  n_warps;
    int32_t  rows_per_lane;    // 2 or 4 adjacent kernel rows per lane
    int32_t  n_buckets;
    int32_t  spec_stride;      // columns per tile
    int32_t  n_cols;           // columns staged per tile (<= spec_stride)
    int32_t  cols_touched;     // columns the band walk may read (>= n_cols; the excess is zero-filled)
    uint32_t n_frames;
    uint32_t n_tiles;
    const float *spec;         // tiled planar layout
    float   *out_db;           // [n_frames][n_buckets]
    float   *power;            // optional [n_frames][n_buckets]
    float    ref_db;
    int32_t  plane_stride;     // pipeline form: columns per plane of the "planes" scratch layout (= the kernel's PLANE)
    int32_t  n_sdft;           // K-sdft groups whose combine step this kernel runs while it stages the tile
    const unsigned *sdft_done; // completion counter of the partial-sum launches (nullptr: combine after the grid wait)
    uint32_t sdft_expected;    // counter value once every partial-sum CTA of this launch has finished (modulo 2^32)
    unsigned *tile_ready;      // per-tile count of K-fft CTAs that have written the tile (nullptr: wait for the grid)
    int32_t  n_ready_groups;   // window groups on the FFT path and the frames one K-fft CTA covers in each
    int32_t  ready_fpc[kMaxGroups];
    SdftParams sdft[kMaxSdft];
};

// K-spmm-db, cluster form ("coefficient-stationary"): a thread-block cluster of CS CTAs splits the kernel
// rows into CS contiguous parts; each CTA keeps its part's coefficients in shared memory for the whole
// launch and streams 16-frame rounds (two tiles) through; the frame-wise max / min of power_to_db are
// exchanged between the CTAs of the cluster through distributed shared memory.
struct ClusterPart {
    int32_t n_warps;     // warps of this part that own band slots
    int32_t col_lo;      // first spectrum column staged (multiple of 8)
    int32_t n_cols;      // columns staged per tile (clipped to the tile's spec_stride)
    int32_t cols_touched; // columns the band halves may read (>= n_cols; the excess is zero-filled)
    int32_t row_lo;      // first output row owned
    int32_t n_rows;      // rows owned (contiguous)
    int32_t coef_base;   // first coefficient slot of the part in `coef` (slot = 16 float4)
    int32_t coef_slots;  // slots of the part
    int32_t desc_base;   // first warp descriptor / 16-lane group of the part
};
struct ClusterWarp {
    int32_t width;       // band slots per lane: half the longest band of the warp's 8 row pairs
    int32_t nwidth;      // conjugate-part slots
    int32_t slot_base;   // first slot of the warp, relative to the part's coef_base
    int32_t nslot_base;
};
struct ClusterLane {     // per (warp, half h, unit u): 16 per warp
    int32_t col;         // first column of this lane's half band, relative to col_lo
    int32_t ncol;        // first column of the conjugate-part band
    int32_t row;         // first output row relative to row_lo (-1: none)
    int32_t n_rows;      // 0..2
};
struct ClusterParams {
    const ClusterPart *parts;
    const ClusterWarp *warps;
    const ClusterLane *lanes;
    const float4 *coef;
    int32_t  cluster_size;
    int32_t  n_buckets;
    int32_t  spec_stride;
    int32_t  max_rows;      // largest n_rows over parts (the log-spectrum buffer aliases the staged tiles)
    int32_t  coef_bytes;    // shared memory reserved for the coefficients (largest part)
    int32_t  mm_offset;     // byte offset of the max / min exchange slots (after the tiles / log-spectrum buffer)
    uint32_t n_frames;
    uint32_t n_tiles;
    const float *spec;
    float   *out_db;
    float   *power;
    float    ref_db;
};
constexpr int kClusterPlaneCols = 384;   // columns per chunk plane of a staged tile (compile-time: immediates)
constexpr int kClusterThreads = 384;     // 12 warps: up to 96 row pairs per part; two CTAs share an SM
constexpr int kClusterRoundFrames = 2 * kTileFrames;

#ifndef PVQT_STEP_CARVEOUT_PCT
#define PVQT_STEP_CARVEOUT_PCT 72
#endif
constexpr int kStepCarveoutPct = PVQT_STEP_CARVEOUT_PCT;   // shared-memory carve-out (% of 228 KB) of K-sdft and K-fft; -1: driver's choice
constexpr int kFusedRing = 4;   // coefficient slots the one-CTA-per-tile K-spmm-db keeps in flight per lane
constexpr int kSpmmWarps = 4;   // warps (= tiles) per SpMM CTA
constexpr int kSpmmUnroll = 4;  // band slots per software-pipeline group (band widths are padded to it)

cudaError_t launch_fft(const FftParams &p, int total_ctas, int block_threads, cudaStream_t stream);
cudaError_t launch_sdft_partial(const SdftParams &p, bool tensor_cores, cudaStream_t stream);
uint32_t    sdft_partial_ctas(const SdftParams &p, bool tensor_cores);   // CTAs of that launch (each adds 1 to done_counter)
cudaError_t launch_sdft_combine(const SdftParams &p, cudaStream_t stream);
bool        sdft_tc_supported(const SdftGroup &g);          // sdft_tc_kernel.cu: the tcgen05 form of the partial sums
void        sdft_tc_make_plan(const SdftGroup &g, int group16, SdftTcPlan *t);
cudaError_t configure_sdft_tc(const SdftTcPlan &t);
cudaError_t launch_sdft_partial_tc(const SdftParams &p, const SdftTcPlan &t, cudaStream_t stream);
cudaError_t configure_sdft(int hop_pad);
cudaError_t configure_sdft_combine(int q, int nk);
size_t sdft_combine_smem_bytes(int q, int nk);
cudaError_t launch_read_sweep(const void *p, size_t bytes, unsigned *sink, cudaStream_t stream);
size_t sdft_smem_bytes(int hop_pad);
cudaError_t launch_spmm(const SpmmParams &p, cudaStream_t stream);
cudaError_t launch_power_to_db(const DbParams &p, cudaStream_t stream);
cudaError_t launch_spmm_db_fused(const FusedParams &p, cudaStream_t stream);
size_t      cluster_smem_bytes(int coef_bytes, int max_rows, int cluster_size);
cudaError_t configure_cluster(int coef_bytes, int max_rows, int cluster_size, int *max_clusters);
cudaError_t launch_spmm_db_cluster(const ClusterParams &p, int n_clusters, cudaStream_t stream);
bool        fused_supported(int n_warps, int cols_touched, int n_buckets, int rows_per_lane);
size_t      fused_smem_bytes(int cols_touched, int n_buckets, int n_warps, int rows_per_lane);
cudaError_t configure_fused(int n_warps, int cols_touched, int n_buckets, int rows_per_lane);
// K-spmm-db, persistent warp-specialised pipeline form (spmm_pipe.cu): same FusedParams plan, one CTA per SM
bool        pipe_supported(int n_warps, int cols_touched, int n_buckets, int rows_per_lane, int min_slots);
int         pipe_plane_stride(int cols_touched);          // columns per plane of the scratch layout the kernel stages
size_t      pipe_smem_bytes(int cols_touched, int n_buckets, int n_warps, int sdft_floats2, int n_sdft);
size_t      pipe_sdft_floats2(int q, int nk);             // float2 entries of the combine's staging for one K-sdft group
cudaError_t configure_pipe(int n_warps, int cols_touched, int n_buckets, int sdft_floats2, int n_sdft);
cudaError_t launch_spmm_db_pipe(const FusedParams &p, int n_ctas, int sdft_floats2, cudaStream_t stream);
cudaError_t configure_kernels(int max_cols);
size_t fft_smem_bytes(int block_threads);
size_t spmm_smem_bytes(int max_cols);

}  // namespace pvqt_dev

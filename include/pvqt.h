/*
 * pvqt.h -- C ABI of the B200-native VQT hot path (libpvqt.so).
 *
 * This is the drop-in boundary for pitchvis_analysis's variable-Q transform.  The
 * reference has no FFI layer: the boundary is the Rust type
 * `pitchvis_analysis::vqt::Vqt` (+ `analysis::AnalysisState`).  Each entry point
 * below cites the reference interface it replaces (paths relative to the upstream
 * repository heinzelotto/pitchvis); INTEGRATION.md shows the Rust `extern "C"`
 * binding and the shim crate that re-exports the reference's type names on top.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types, never throws or aborts;
 *   - every function returns a pvqt_status (0 = OK) unless stated otherwise;
 *     pvqt_last_error_string() gives a thread-local human-readable message;
 *   - handles are not re-entrant (the reference takes `&mut self`, vqt.rs:866);
 *     distinct handles may be used concurrently from different threads;
 *   - there is NO CPU fallback: if no CUDA device / kernel image is usable,
 *     pvqt_create fails with PVQT_CUDA_ERROR.
 */
#ifndef PVQT_H
#define PVQT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVQT_ABI_VERSION 1

typedef enum pvqt_status {
    PVQT_OK = 0,
    PVQT_ABOVE_NYQUIST = 1,        /* VqtError::AboveNyquist      vqt.rs:357-360 */
    PVQT_WINDOW_EXCEEDS_NFFT = 2,  /* VqtError::WindowExceedsNFft vqt.rs:365     */
    PVQT_PANIC = 3,                /* an assert!/expect inside Vqt::new (vqt.rs:785-792) */
    PVQT_BAD_LENGTH = 4,           /* assert_eq!(x.len(), n_fft)  vqt.rs:867-871; analysis.rs:289 */
    PVQT_INVALID_ARGUMENT = 5,
    PVQT_UNSUPPORTED = 6,          /* window sizes the sm_100a FFT plans do not cover */
    PVQT_CUDA_ERROR = 7,
    PVQT_OUT_OF_MEMORY = 8
} pvqt_status;

/* VqtRange (vqt.rs:239-255) + VqtParameters (vqt.rs:279-331), flattened, same field
 * meaning and units. */
typedef struct pvqt_params {
    float    sr;
    uint64_t n_fft;
    float    min_freq;
    uint32_t octaves;             /* u8  in the reference */
    uint32_t buckets_per_octave;  /* u16 in the reference */
    float    sparsity_quantile;
    float    quality;
    float    gamma;
} pvqt_params;

/* Error payload of pvqt_create (the fields of VqtError, vqt.rs:352-366). */
typedef struct pvqt_error {
    int32_t  status;              /* pvqt_status */
    float    highest_frequency;   /* AboveNyquist */
    float    nyquist_frequency;   /* AboveNyquist */
    float    window_length;       /* WindowExceedsNFft */
    uint64_t n_fft;               /* WindowExceedsNFft */
    int32_t  cuda_error;          /* cudaError_t when status == PVQT_CUDA_ERROR */
} pvqt_error;

/* Read-only view of one sprs::CsMat<Complex32> (WindowGroup::filter_bank /
 * negative_filter_bank, vqt.rs:396,403).  Host pointers owned by the handle. */
typedef struct pvqt_csr_view {
    int32_t        rows, cols;
    int64_t        nnz;
    const int32_t *indptr;   /* rows + 1 */
    const int32_t *indices;  /* nnz, ascending within a row */
    const float   *data;     /* 2 * nnz, interleaved (re, im) */
} pvqt_csr_view;

typedef struct pvqt pvqt;              /* one Vqt on one device            */
typedef struct pvqt_multi pvqt_multi;  /* one Vqt replicated over N devices */

/* ---- library ---------------------------------------------------------------- */
int         pvqt_abi_version(void);
const char *pvqt_last_error_string(void);
int         pvqt_device_count(int *count);

/* Log sink.  The reference logs through the `log` crate: info! the analysis delay (vqt.rs:468), warn! a coverage gap
 * between neighbouring filters' -3 dB bands (vqt.rs:695-710), debug! the structure of every window group and filter
 * (vqt.rs:661-667, :688-694, :741-746, :843-846).  The sink receives the same lines from pvqt_kernel_create /
 * pvqt_create; level: 1 = warn, 2 = info, 3 = debug; lines above max_level are not formatted.  fn == NULL removes
 * the sink.  Process-wide, like the crate's global logger; the callback runs on the calling thread. */
typedef void (*pvqt_log_fn)(int level, const char *message, void *user);
int         pvqt_set_log_callback(pvqt_log_fn fn, void *user, int max_level);

/* SM count, SM clock (kHz), L2 size (bytes) and host NUMA node (-1: unknown) of a device; any pointer may be NULL.  For
 * roofline reports (bench.py), so that hosts need no CUDA runtime binding. */
int         pvqt_device_attributes(int device, int32_t *sm_count, int32_t *sm_clock_khz, int32_t *l2_bytes, int32_t *numa_node);

/* ---- parameters ------------------------------------------------------------- */
/* `impl Default for VqtParameters` (vqt.rs:333-348, constants vqt.rs:180-214) */
int pvqt_default_params(pvqt_params *out);
/* VqtRange::n_buckets (vqt.rs:259-261) */
size_t pvqt_params_n_buckets(const pvqt_params *p);

/* ---- host-only kernel construction (Vqt::vqt_kernel, vqt.rs:599-759) ----------
 * Builds the sparse spectral kernel without touching a GPU: what Vqt::new computes
 * before planning FFTs.  pvqt_create runs the same builder and then uploads it. */
typedef struct pvqt_kernel pvqt_kernel;
/* FilterParams (vqt.rs:370-384) */
typedef struct pvqt_filter_params {
    float    freq;
    float    window_length;
    uint64_t sr_downscaling_factor;
    uint64_t minimum_needed_window_size;
} pvqt_filter_params;
/* Vqt::filter_bank_params (vqt.rs:517-587); out: n == n_buckets entries */
int    pvqt_filter_bank_params(const pvqt_params *params, pvqt_filter_params *out, size_t n, pvqt_error *err);
int    pvqt_kernel_create(const pvqt_params *params, pvqt_kernel **out, pvqt_error *err);
void   pvqt_kernel_destroy(pvqt_kernel *k);
size_t pvqt_kernel_n_buckets(const pvqt_kernel *k);
double pvqt_kernel_delay_seconds(const pvqt_kernel *k);
size_t pvqt_kernel_num_window_groups(const pvqt_kernel *k);
int    pvqt_kernel_group_window(const pvqt_kernel *k, size_t group, uint64_t *begin, uint64_t *end);
int    pvqt_kernel_group_csr(const pvqt_kernel *k, size_t group, int negative, pvqt_csr_view *out);
/* Diagnostic (Filter::bandwidth_3db_in_hz, vqt.rs:417-420, calculate_bandwidth vqt.rs:962-989): the -3 dB band of every
 * filter in Hz, lo / hi: n == n_buckets entries each (either may be NULL). */
int    pvqt_kernel_filter_bandwidths(const pvqt_kernel *k, float *lo_hz, float *hi_hz, size_t n);
/* The filters below which the reference warns about a coverage gap (vqt.rs:695-710: the band of filter i starts above
 * the end of the band of filter i - 1).  *count receives their number; up to `capacity` indices are stored in `out`
 * (may be NULL).  None at the default parameters. */
int    pvqt_kernel_coverage_gaps(const pvqt_kernel *k, uint32_t *out, size_t capacity, size_t *count);

/* ---- construction (Vqt::new, vqt.rs:465-505) -------------------------------- */
int  pvqt_create(const pvqt_params *params, int device, pvqt **out, pvqt_error *err);
void pvqt_destroy(pvqt *v); /* Drop */

/* ---- introspection (Vqt::params/kernel/delay, vqt.rs:449,507-513) ------------ */
int    pvqt_get_params(const pvqt *v, pvqt_params *out);
size_t pvqt_n_buckets(const pvqt *v);
size_t pvqt_n_fft(const pvqt *v);
double pvqt_delay_seconds(const pvqt *v);  /* Duration::from_secs_f32(...) vqt.rs:756 */
size_t pvqt_num_window_groups(const pvqt *v);
int    pvqt_group_window(const pvqt *v, size_t group, uint64_t *begin, uint64_t *end); /* WindowGroup::window */
int    pvqt_group_csr(const pvqt *v, size_t group, int negative, pvqt_csr_view *out);
int    pvqt_device(const pvqt *v);
/* Smallest sample offset inside an n_fft frame that the transform reads (the union of
 * the window groups starts here; 24576 at the defaults). */
size_t pvqt_first_sample_used(const pvqt *v);

/* ---- per-frame entry (Vqt::calculate_vqt_instant_in_db, vqt.rs:866-916) ------ */
/* x: n == n_fft host samples, the last one is "now"; out: n_buckets host floats.
 * Returns PVQT_BAD_LENGTH where the reference panics. */
int pvqt_calc_instant_db(pvqt *v, const float *x, size_t n, float *out);

/* ---- batched entries (new; sliding-window template: pitchvis_train/src/train.rs:276-341)
 * Host buffers in, host buffers out; H2D/D2H copies are done inside. -------------------- */
/* frame t = audio[t*hop .. t*hop + n_fft), t < n_frames; out[n_frames][n_buckets]. */
int pvqt_calc_batch_db(pvqt *v, const float *audio, size_t n_samples, size_t hop, size_t n_frames,
                       float *out);
/* frames[B][n_fft] independent frames. */
int pvqt_calc_frames_db(pvqt *v, const float *frames, size_t n_frames, float *out);
/* n_streams independent recordings, stream s at audio + s*stream_stride, n_samples each;
 * out[n_streams][frames_per_stream][n_buckets]. */
int pvqt_calc_streams_db(pvqt *v, const float *audio, size_t n_streams, size_t stream_stride,
                         size_t n_samples, size_t hop, size_t frames_per_stream, float *out);
/* Number of whole frames in n_samples: 0 if n_samples < n_fft, else (n_samples-n_fft)/hop + 1. */
size_t pvqt_frames_in(const pvqt *v, size_t n_samples, size_t hop);

/* ---- device-pointer entry (inputs already resident in HBM) ------------------ */
/* d_audio / d_out are device pointers on pvqt_device(v).  Asynchronous on `cuda_stream`
 * (a cudaStream_t, NULL = the handle's own stream).  Frame f (0 <= f < n_streams *
 * frames_per_stream) reads d_audio[(f / frames_per_stream) * stream_stride +
 * (f % frames_per_stream) * hop + i], i < n_fft.  d_power (optional, may be NULL)
 * receives |z|^2 before power_to_db, same shape as d_out. */
int pvqt_calc_db_device(pvqt *v, const float *d_audio, size_t n_streams, size_t stream_stride,
                        size_t hop, size_t frames_per_stream, float *d_out, float *d_power,
                        void *cuda_stream);
/* Test hook: spectra of the window groups restricted to the consumed columns, in the library's
 * tiled planar scratch layout: [tile = frame / 8][column < pvqt_spec_stride(v)][16 floats], a 64-byte
 * record per (tile, column) with logical 16-byte chunks 0,1 = Re of frames 0-3, 4-7 and 2,3 = Im of
 * frames 0-3, 4-7; logical chunk q is stored at physical chunk q ^ ((column >> 1) & 3).
 * d_spec: ceil(n_frames / 8) tiles. */
int    pvqt_fft_device(pvqt *v, const float *d_audio, size_t n_streams, size_t stream_stride, size_t hop,
                       size_t frames_per_stream, float *d_spec, void *cuda_stream);
size_t pvqt_spec_stride(const pvqt *v);                 /* columns per tile */
int    pvqt_group_columns(const pvqt *v, size_t group, uint32_t *first_col, uint32_t *n_cols,
                          uint32_t *spec_offset);       /* consumed FFT bins of a group */

/* ---- device memory / timing helpers (so hosts need no CUDA runtime binding) -- */
int pvqt_dev_alloc(int device, size_t bytes, void **out);
int pvqt_dev_free(int device, void *p);
int pvqt_host_alloc_pinned(size_t bytes, void **out);
int pvqt_host_free_pinned(void *p);
int pvqt_memcpy_h2d(pvqt *v, void *dst, const void *src, size_t bytes, int async);
int pvqt_memcpy_d2h(pvqt *v, void *dst, const void *src, size_t bytes, int async);
int pvqt_dev_memset(pvqt *v, void *dst, int value, size_t bytes);
/* Benchmark hygiene: evict everything from L2 by writing `bytes` (> L2 size) of zeros to `scratch` and
 * reading them back, so that no dirty lines are left for the next kernel to write back.  Asynchronous on
 * the handle's stream. */
int pvqt_dev_flush_l2(pvqt *v, void *scratch, size_t bytes);
int pvqt_synchronize(pvqt *v);
int pvqt_event_create(pvqt *v, void **out_event);
int pvqt_event_destroy(pvqt *v, void *event);
int pvqt_event_record(pvqt *v, void *event);                      /* on the handle's stream */
int pvqt_event_elapsed_ms(pvqt *v, void *start, void *stop, float *ms); /* synchronises on stop */
/* Number of kernel launches this handle has issued since creation. */
uint64_t pvqt_launch_count(const pvqt *v);
/* Per-kernel device timing: while enabled, every launch is bracketed by CUDA events on the
 * launching stream.  pvqt_get_profile synchronises and returns the summed durations and launch
 * counts since the last reset, indexed by kernel kind: 0 = K-fft, 1 = K-spmm (unfused fallback),
 * 2 = K-db (unfused fallback), 3 = K-spmm-db (fused), 4 = K-sdft (sliding partial DFTs),
 * 5 = K-sdft-combine (stand-alone), 6 = K-spmm-db cluster form; both arrays hold PVQT_PROFILE_KINDS entries. */
#define PVQT_PROFILE_KINDS 8
int pvqt_set_profiling(pvqt *v, int enabled);
int pvqt_get_profile(pvqt *v, int reset, double *kernel_ms, uint64_t *kernel_launches);
/* Test / tuning switch for the SpMM + power_to_db stage: 0 = unfused K-spmm + K-db pair, 1 = K-spmm-db with
 * one CTA per tile, 2 = K-spmm-db in its cluster form (coefficients stationary in shared memory, frame max / min
 * exchanged through distributed shared memory; measured slower at the default parameters, kept selectable),
 * 3 (default where the kernel fits) = K-spmm-db as a persistent warp-specialised pipeline: one CTA per SM, helper
 * warps stage / combine the next tile and convert the previous one while the band walk runs.  0, 1 and 3 give the
 * same bits.  A mode the kernel does not fit falls back (3 -> 1 -> 0).  Returns the mode in effect. */
int pvqt_set_fused_epilogue(pvqt *v, int mode);
/* Plan introspection for tests and bench reports; out[0..n) (n <= 10; the 10th: bit mask of the window groups that took
 * the K-sdft path in the most recent batched launch): cluster size of K-spmm-db's cluster form
 * (0: not available), co-resident clusters, shared-memory bytes of its coefficients, rows of its largest
 * part, warps of the one-CTA-per-tile form (0: not available), K-fft block size, columns per spectrum tile,
 * K-sdft plans cached, band slots the warps of one K-spmm-db CTA walk per tile (padding included). */
int pvqt_plan_info(const pvqt *v, int32_t *out, size_t n);
/* Test / tuning switch: 0 keeps every window group on the per-frame FFT path; 1 and 2 (default) let groups whose
 * consumed bins are cheaper as sums of hop-sized partial DFTs shared between overlapping frames take the K-sdft
 * path in the batched entries (never in the per-frame / independent-frames entries), with the partial sums on the
 * FP32 pipe (1) or on the tensor cores (2: mma.sync TF32 with the 3xTF32 split; 3, opt-in: tcgen05.mma with the
 * accumulators in TMEM, for hops whose window remainder is a multiple of 16 samples, else as 2).  All evaluate the
 * same DFT; they differ by f32 rounding only.  Returns the mode in effect. */
int pvqt_set_sliding_dft(pvqt *v, int mode);

/* ---- sharding (frame ranges / streams, no collective) ---------- */
/* Split `n_units` (frames or streams) into `n_parts` contiguous, balanced ranges.  Pure
 * host arithmetic; part p covers [begin, end). */
int pvqt_shard_range(size_t n_units, size_t n_parts, size_t part, size_t *begin, size_t *end);
/* For a frame range [f0, f1) of one recording: the sample range [s0, s1) it reads
 * (halo included): s0 = f0*hop, s1 = (f1-1)*hop + n_fft (s0 == s1 == 0 if empty). */
int pvqt_frame_range_samples(size_t n_fft, size_t hop, size_t f0, size_t f1, size_t *s0, size_t *s1);

/* ---- multi-GPU, single process (one host thread per device, pinned gather) ---- */
int  pvqt_multi_create(const pvqt_params *params, int n_devices, const int *device_ids, pvqt_multi **out,
                       pvqt_error *err);
void pvqt_multi_destroy(pvqt_multi *m);
int  pvqt_multi_num_devices(const pvqt_multi *m);
pvqt *pvqt_multi_handle(pvqt_multi *m, int index);
/* One long recording, frame-range sharded with an (n_fft - hop)-sample halo. */
int  pvqt_multi_calc_batch_db(pvqt_multi *m, const float *audio, size_t n_samples, size_t hop,
                              size_t n_frames, float *out);
/* Independent streams, contiguous blocks of streams per device. */
int  pvqt_multi_calc_streams_db(pvqt_multi *m, const float *audio, size_t n_streams, size_t stream_stride,
                                size_t n_samples, size_t hop, size_t frames_per_stream, float *out);

/* Measurement aid: the host <-> device copy rate of the box with every device of `m` copying at once -- h2d_bytes in
 * and d2h_bytes out per device and repetition, pinned host memory, both directions in flight, no kernel.  It is the
 * ceiling of the host-buffer entries above (their end-to-end rate is PCIe-bound).  *seconds: wall time of `reps`
 * repetitions on all devices. */
int  pvqt_multi_pcie_probe(pvqt_multi *m, size_t h2d_bytes, size_t d2h_bytes, int reps, double *seconds);

#ifdef __cplusplus
}
#endif
#endif /* PVQT_H */

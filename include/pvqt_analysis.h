/*
 * pvqt_analysis.h -- C ABI of the AnalysisState epilogue (libpvqt.so), the optional stage after the
 * VQT: per-bin calmness-adaptive EMA smoothing, peak detection, continuous peak refinement, bass
 * promotion, afterglow, calmness and tuning measures.
 *
 * Replaces pitchvis_analysis::analysis::AnalysisState (analysis.rs:119-410) and its modules
 * (analysis_modules/{peak_detection,calmness,afterglow,pitch_analysis}.rs, util.rs:91-137).
 * One state object holds `n_streams` independent AnalysisStates (one per audio stream); each is a
 * recurrence in time, so a batch call advances every stream by `n_frames` frames, streams in
 * parallel on the GPU (one CTA per stream), frames in order.
 *
 * Peak semantics: the reference delegates to the crates.io crate find_peaks 0.1.5, which is not part
 * of the reference tree; this library implements strict local maxima (plateaus at their middle),
 * then min_height, then min_distance (taller peaks win), then min_prominence -- the algorithm of
 * scipy.signal.find_peaks, which the crate advertises compatibility with ("peaks").
 */
#ifndef PVQT_ANALYSIS_H
#define PVQT_ANALYSIS_H

#include "pvqt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* PeakDetectionParameters (peak_detection.rs:10-15) */
typedef struct pvqt_peak_params {
    float min_prominence;
    float min_height;
} pvqt_peak_params;

/* AnalysisParameters (analysis.rs:36-65); std::time::Duration fields in nanoseconds */
typedef struct pvqt_analysis_params {
    uint64_t         spectrogram_length;
    pvqt_peak_params peak_config;
    pvqt_peak_params bassline_peak_config;
    uint64_t         highest_bassnote;
    uint64_t         vqt_smoothing_duration_base_ns;
    float            vqt_smoothing_calmness_min;
    float            vqt_smoothing_calmness_max;
    uint64_t         note_calmness_smoothing_duration_ns;
    uint64_t         scene_calmness_smoothing_duration_ns;
    uint64_t         tuning_inaccuracy_smoothing_duration_ns;
    float            harmonic_threshold;
} pvqt_analysis_params;

/* VqtRange (vqt.rs:239-255) */
typedef struct pvqt_range {
    float    min_freq;
    uint32_t octaves;
    uint32_t buckets_per_octave;
} pvqt_range;

/* ContinuousPeak (peak_detection.rs:17-23) */
typedef struct pvqt_continuous_peak {
    float center;
    float size;
} pvqt_continuous_peak;

/* Per-frame results of a batch call: the public fields of AnalysisState (analysis.rs:119-177) after
 * each preprocess().  Every pointer may be NULL (that output is skipped).  S = n_streams, T = n_frames,
 * NB = n_buckets, P = max_peaks (peaks beyond P are counted but not stored). */
typedef struct pvqt_analysis_outputs {
    uint32_t              max_peaks;
    uint32_t             *peak_count;        /* [S][T]       |peaks| */
    uint32_t             *peak_indices;      /* [S][T][P]    peaks, ascending */
    pvqt_continuous_peak *peaks_continuous;  /* [S][T][P]    sorted by center (peak_detection.rs:145) */
    float                *x_vqt_smoothed;    /* [S][T][NB] */
    float                *x_vqt_peakfiltered;/* [S][T][NB] */
    float                *x_vqt_afterglow;   /* [S][T][NB] */
    float                *calmness;          /* [S][T][NB] */
    float                *pitch_accuracy;    /* [S][T][NB] */
    float                *pitch_deviation;   /* [S][T][NB] */
    float                *smoothed_scene_calmness;          /* [S][T] */
    float                *smoothed_tuning_grid_inaccuracy;  /* [S][T] */
} pvqt_analysis_outputs;

typedef struct pvqt_analysis pvqt_analysis;

/* `impl Default for AnalysisParameters` (analysis.rs:72-98) */
int  pvqt_analysis_default_params(pvqt_analysis_params *out);
/* AnalysisState::new (analysis.rs:192-241), n_streams independent states on `device` */
int  pvqt_analysis_create(const pvqt_range *range, const pvqt_analysis_params *params, size_t n_streams, int device,
                          pvqt_analysis **out);
void pvqt_analysis_destroy(pvqt_analysis *a);
size_t pvqt_analysis_n_buckets(const pvqt_analysis *a);
size_t pvqt_analysis_n_streams(const pvqt_analysis *a);
/* AnalysisState::update_vqt_smoothing_duration (analysis.rs:251-270); has_duration == 0 <=> None */
int  pvqt_analysis_update_vqt_smoothing_duration(pvqt_analysis *a, int has_duration, uint64_t duration_ns);

/* AnalysisState::preprocess (analysis.rs:288-404) for T consecutive frames of every stream.
 * db: host [S][T][NB] dB spectra (the layout pvqt_calc_streams_db writes); n_buckets must equal NB
 * (PVQT_BAD_LENGTH where the reference asserts, analysis.rs:289).  frame_time_ns: Duration per frame.
 * `out` holds host pointers. */
int  pvqt_analysis_preprocess_batch(pvqt_analysis *a, const float *db, size_t n_buckets, size_t n_frames,
                                    uint64_t frame_time_ns, const pvqt_analysis_outputs *out);
/* Same with device pointers for db and for every non-NULL member of `out` (fused after
 * pvqt_calc_db_device on the same device); asynchronous on cuda_stream (NULL = the state's stream). */
int  pvqt_analysis_preprocess_device(pvqt_analysis *a, const float *d_db, size_t n_buckets, size_t n_frames,
                                     uint64_t frame_time_ns, const pvqt_analysis_outputs *d_out, void *cuda_stream);
int  pvqt_analysis_synchronize(pvqt_analysis *a);

/* ---- chroma reduction ------------------------------------------------------
 * pitchvis_viewer/src/display_system/update.rs:1104-1131: per frame, the sum of 10^(dB/10) per pitch class (bin b
 * belongs to class (round(12 b / buckets_per_octave) + class of min_freq relative to C4 = 261.626 Hz) mod 12),
 * divided by the largest of the 12 sums when that is positive.  The reference feeds x_vqt_smoothed; any
 * [n_frames][n_buckets] dB array works.  out: [n_frames][12].  Host and device-pointer forms. */
int  pvqt_chroma(const pvqt_range *range, int device, const float *db, size_t n_frames, float *out);
int  pvqt_chroma_device(const pvqt_range *range, int device, const float *d_db, size_t n_frames, float *d_out,
                        void *cuda_stream);

/* ---- spectrogram ring -------------------------------------------------------
 * pitchvis_viewer/src/display_system/update.rs:930-1088, SpectrogramMode::VQT: every frame writes one RGBA8 row of
 * width = n_buckets pixels into a ring image [height][n_buckets][4] -- row height-1-write_index --, clears the next
 * row and advances write_index.  Alpha is the frame-normalised brightness of x_vqt_smoothed (update.rs:965-975,
 * :989); RGB is a property of the bin alone (pitchvis_colors::calculate_color, an LCh round trip through the `lab`
 * crate: a display concern) and is passed in as bytes already scaled as update.rs:986-988 does,
 * bin_rgb[3 b + c] = (c * 255 * 1.2).clamp(0, 255) as u8.  n_frames frames at once: the image ends as if the
 * reference had processed them one by one (the last min(n_frames, height-1) frames own a row each, the row after
 * the last is cleared); *write_index is advanced by n_frames modulo height.  Host and device-pointer forms; the device
 * form with height 1 is the caller's memset. */
int  pvqt_spectrogram_vqt(int device, const float *smoothed, size_t n_frames, size_t n_buckets, const uint8_t *bin_rgb,
                          uint8_t *image, size_t height, size_t *write_index);
int  pvqt_spectrogram_vqt_device(int device, const float *d_smoothed, size_t n_frames, size_t n_buckets,
                                 const uint8_t *d_bin_rgb, uint8_t *d_image, size_t height, size_t *write_index,
                                 void *cuda_stream);

/* SpectrogramMode::Peaks of the same system (update.rs:997-1062): a frame's row shows only its continuous peaks --
 * every bin within PEAK_RADIUS = 2 bins of a peak centre gets the peak's pitch colour (pitchvis_colors::calculate_color
 * of the FRACTIONAL bin, pitchvis_colors/src/lib.rs:93-119, computed here: it depends on the peak) and alpha =
 * brightness(size / max size) * exp(-d^2 / 2); peaks are drawn in list order, later ones over earlier ones; all other
 * pixels of the row keep the zeros the previous step left.  peaks / peak_count are exactly what K-analysis writes
 * (pvqt_analysis_outputs::peaks_continuous [n_frames][max_peaks], ::peak_count [n_frames]), so the device form chains
 * behind pvqt_analysis_preprocess_device without a host round trip.  width = range->octaves * buckets_per_octave. */
int  pvqt_spectrogram_peaks(int device, const pvqt_range *range, const pvqt_continuous_peak *peaks, const uint32_t *peak_count,
                            size_t max_peaks, size_t n_frames, uint8_t *image, size_t height, size_t *write_index);
int  pvqt_spectrogram_peaks_device(int device, const pvqt_range *range, const pvqt_continuous_peak *d_peaks,
                                   const uint32_t *d_peak_count, size_t max_peaks, size_t n_frames, uint8_t *d_image,
                                   size_t height, size_t *write_index, void *cuda_stream);

/* ---- VQT + AnalysisState in one call (BASELINE.json configs[4]) ----------------------------------------------
 * Every caller of the reference runs the two back to back, per frame: Vqt::calculate_vqt_instant_in_db then
 * AnalysisState::preprocess (pitchvis_viewer/src/vqt_system.rs:40-68 + analysis_system.rs:10-20;
 * pitchvis_serial/src/main.rs:206-230).  These entries do it for whole recordings: host audio in (same layout and
 * frame rule as pvqt_calc_batch_db / pvqt_calc_streams_db), the dB spectra go from the transform to K-analysis
 * through HBM, and only the results `out` names travel back (host pointers, [S][T]... as in
 * pvqt_analysis_preprocess_batch).  out_db (optional, may be NULL) also returns the spectra [S][T][NB].
 * `a` must live on v's device, hold n_streams states and the same n_buckets; its states advance by
 * frames_per_stream frames.  *d2h_bytes (optional) receives the bytes copied device -> host.  One recording must
 * fit one staging batch (2^28 samples); more streams than fit are processed in groups. */
int  pvqt_calc_batch_analysis(pvqt *v, pvqt_analysis *a, const float *audio, size_t n_samples, size_t hop, size_t n_frames,
                              uint64_t frame_time_ns, const pvqt_analysis_outputs *out, float *out_db, uint64_t *d2h_bytes);
int  pvqt_calc_streams_analysis(pvqt *v, pvqt_analysis *a, const float *audio, size_t n_streams, size_t stream_stride,
                                size_t n_samples, size_t hop, size_t frames_per_stream, uint64_t frame_time_ns,
                                const pvqt_analysis_outputs *out, float *out_db, uint64_t *d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* PVQT_ANALYSIS_H */

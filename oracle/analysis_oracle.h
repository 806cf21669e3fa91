/*
 * analysis_oracle.h -- CPU restatement of pitchvis_analysis::analysis::AnalysisState and its
 * analysis_modules (peak detection, calmness, afterglow, pitch accuracy).
 *
 * TEST INFRASTRUCTURE ONLY (see vqt_oracle.h).
 *
 * Pinning status: the only reference test that crosses this code with real data is
 * test_vqt_close_frequencies (lib.rs:16-48, a peak *count*), plus test_analysis_does_something
 * (analysis.rs:415-428) and the EMA tests (util.rs:143-225); all three are reproduced in
 * tests/test_analysis_oracle.py.  The third-party crate find_peaks 0.1.5 is not in the reference
 * tree; its semantics are ASSUMED to be those of scipy.signal.find_peaks(height, distance,
 * prominence), which the crate advertises compatibility with and which the tests cross-check
 * against scipy itself.  Peak-index parity with the Rust binary is therefore "unpinned".
 */
#ifndef ANALYSIS_ORACLE_H
#define ANALYSIS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PeakDetectionParameters, peak_detection.rs:10-15 */
typedef struct orc_peak_params {
    float min_prominence;
    float min_height;
} orc_peak_params;

/* AnalysisParameters, analysis.rs:36-65 (Durations in nanoseconds) */
typedef struct orc_analysis_params {
    uint64_t        spectrogram_length;
    orc_peak_params peak_config;
    orc_peak_params bassline_peak_config;
    uint64_t        highest_bassnote;
    uint64_t        vqt_smoothing_duration_base_ns;
    float           vqt_smoothing_calmness_min;
    float           vqt_smoothing_calmness_max;
    uint64_t        note_calmness_smoothing_duration_ns;
    uint64_t        scene_calmness_smoothing_duration_ns;
    uint64_t        tuning_inaccuracy_smoothing_duration_ns;
    float           harmonic_threshold;
} orc_analysis_params;

/* ContinuousPeak, peak_detection.rs:17-23 */
typedef struct orc_continuous_peak {
    float center;
    float size;
} orc_continuous_peak;

typedef struct orc_analysis orc_analysis;

void orc_analysis_default_params(orc_analysis_params *p);                 /* analysis.rs:72-98 */
orc_analysis *orc_analysis_new(float min_freq, uint32_t octaves, uint32_t buckets_per_octave,
                               const orc_analysis_params *p);             /* analysis.rs:192-241 */
void orc_analysis_free(orc_analysis *a);
/* analysis.rs:251-270; has_duration == 0 <=> None */
void orc_analysis_update_vqt_smoothing_duration(orc_analysis *a, int has_duration, uint64_t duration_ns);
/* analysis.rs:288-404; returns 0, or 4 (ORC_BAD_LENGTH) where the reference asserts */
int  orc_analysis_preprocess(orc_analysis *a, const float *x_vqt, size_t n, uint64_t frame_time_ns);

/* results of the last preprocess call (public fields, analysis.rs:119-177) */
size_t orc_analysis_n_buckets(const orc_analysis *a);
size_t orc_analysis_peaks(const orc_analysis *a, uint32_t *out, size_t cap);      /* ascending indices */
size_t orc_analysis_peaks_continuous(const orc_analysis *a, orc_continuous_peak *out, size_t cap);
void   orc_analysis_vectors(const orc_analysis *a, float *smoothed, float *peakfiltered, float *afterglow,
                            float *calmness, float *pitch_accuracy, float *pitch_deviation); /* any may be NULL */
float  orc_analysis_scene_calmness(const orc_analysis *a);
float  orc_analysis_tuning_inaccuracy(const orc_analysis *a);

/* The find_peaks wrapper alone, peak_detection.rs:26-51.  order: 0 = height, distance, prominence
 * (scipy's order, the assumed one); 1 = height, prominence, distance (the alternative reading of the
 * crate) -- the tests assert that both give identical sets on the whole corpus.
 * Returns the number of peaks, ascending indices in out. */
size_t orc_find_peaks(const float *x, size_t n, float min_prominence, float min_height,
                      uint32_t buckets_per_octave, int order, uint32_t *out, size_t cap);

/* EmaMeasurement (util.rs:91-137) for the EMA unit tests: returns the updated y. */
float orc_ema_update(float y, int has_horizon, uint64_t horizon_ns, float new_value, uint64_t timestep_ns);

#ifdef __cplusplus
}
#endif
#endif

/*
 * pvqt_agc.h -- C ABI of the AGC pre-stage (libpvqt.so): the step immediately in front of the VQT path.
 *
 * Replaces dagc_fork::MonoAgc (dagc_fork/src/lib.rs:19-87) as its callers drive it, one audio chunk at a time
 * (pitchvis_audio/src/audio_desktop.rs:97-117, pitchvis_train/src/train.rs:271,296-310): the gain is frozen for a
 * chunk whose sum of squares is below a threshold, then every sample is multiplied by the running gain and the
 * gain updated per sample (a nonlinear recurrence: sequential in time per stream, parallel across streams).
 * One object holds `n_streams` independent MonoAgc states.
 */
#ifndef PVQT_AGC_H
#define PVQT_AGC_H

#include "pvqt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pvqt_agc pvqt_agc;

/* MonoAgc::new (lib.rs:35-53) for n_streams streams on `device`.  PVQT_INVALID_ARGUMENT where the reference
 * returns Error::InvalidDesiredOutputRms / Error::InvalidDistortionFactor (pvqt_last_error_string names which). */
int  pvqt_agc_create(float desired_output_rms, float distortion_factor, size_t n_streams, int device, pvqt_agc **out);
void pvqt_agc_destroy(pvqt_agc *a);
size_t pvqt_agc_n_streams(const pvqt_agc *a);
/* MonoAgc::gain (lib.rs:69-71) of every stream; out: n_streams floats. */
int  pvqt_agc_gains(pvqt_agc *a, float *out);
/* MonoAgc::freeze_gain (lib.rs:59-61) for every stream; used by pvqt_agc_process when silence_threshold is NaN. */
int  pvqt_agc_freeze_gain(pvqt_agc *a, int freeze);

/* MonoAgc::process (lib.rs:76-86) driven chunk by chunk, in place, for every stream:
 *   for each chunk of `chunk` samples (the last may be shorter; chunk == 0: one chunk of n_samples):
 *       frozen = sum_i x[i]^2 < silence_threshold     (sequential f32 sum over the chunk, audio_desktop.rs:106-107)
 *       x[i] *= gain;  if !frozen { gain *= max(1 + d (1 - x[i]^2 / rms), d) }
 * silence_threshold < 0 never freezes; NaN uses the flag set by pvqt_agc_freeze_gain instead.
 * audio: host [n_streams][stream_stride] (n_samples used per stream); state carries over between calls. */
int  pvqt_agc_process(pvqt_agc *a, float *audio, size_t stream_stride, size_t n_samples, size_t chunk,
                      float silence_threshold);
/* Same on device memory of the object's device, asynchronous on cuda_stream (NULL = the object's stream). */
int  pvqt_agc_process_device(pvqt_agc *a, float *d_audio, size_t stream_stride, size_t n_samples, size_t chunk,
                             float silence_threshold, void *cuda_stream);
int  pvqt_agc_synchronize(pvqt_agc *a);

#ifdef __cplusplus
}
#endif
#endif /* PVQT_AGC_H */

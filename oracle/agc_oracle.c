/*
 * agc_oracle.c -- CPU restatement of dagc_fork::MonoAgc (the AGC stage in front of the VQT path).
 *
 * TEST INFRASTRUCTURE ONLY (see vqt_oracle.h).  Follows dagc_fork/src/lib.rs:19-87 and the way the callers
 * drive it per audio chunk (pitchvis_audio/src/audio_desktop.rs:101-117, pitchvis_train/src/train.rs:296-310):
 * freeze the gain when the chunk's sum of squares is below a threshold, then process the chunk in place.
 * f32 arithmetic in the order of the Rust source; compile with -ffp-contract=off.
 */
#include "agc_oracle.h"

#include <math.h>

/* MonoAgc::new, lib.rs:35-53: 0 ok, 1 InvalidDesiredOutputRms, 2 InvalidDistortionFactor */
int orc_agc_check(float desired_output_rms, float distortion_factor)
{
    if (!(desired_output_rms > 0.0f && isfinite(desired_output_rms))) return 1;
    if (!(distortion_factor >= 0.0f && distortion_factor <= 1.0f)) return 2;
    return 0;
}

/* MonoAgc::process, lib.rs:76-86 */
void orc_agc_process(float *samples, size_t n, float desired_output_rms, float distortion_factor, float *gain,
                     int frozen)
{
    float g_state = *gain;
    for (size_t i = 0; i < n; ++i) {
        samples[i] *= g_state;
        if (!frozen) {
            const float x = samples[i];
            const float y = (x * x) / desired_output_rms;          /* x.powi(2) / desired_output_rms */
            float g = 1.0f + (distortion_factor * (1.0f - y));
            g = g > distortion_factor ? g : distortion_factor;      /* g.max(distortion_factor) */
            g_state *= g;
        }
    }
    *gain = g_state;
}

/* The callers' chunk loop: sample_sq_sum = sum x^2 (sequential f32, before the gain is applied);
 * freeze_gain(sample_sq_sum < threshold); process(chunk).  A negative threshold never freezes. */
void orc_agc_process_chunks(float *samples, size_t n, size_t chunk, float desired_output_rms, float distortion_factor,
                            float silence_threshold, float *gain)
{
    if (chunk == 0) chunk = n;
    for (size_t b = 0; b < n; b += chunk) {
        const size_t m = n - b < chunk ? n - b : chunk;
        float sq = 0.0f;
        for (size_t i = 0; i < m; ++i) sq += samples[b + i] * samples[b + i];
        orc_agc_process(samples + b, m, desired_output_rms, distortion_factor, gain, sq < silence_threshold);
    }
}

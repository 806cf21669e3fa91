/*
 * chroma_oracle.c -- CPU restatement of the chroma reduction downstream of AnalysisState
 * (pitchvis_viewer/src/display_system/update.rs:1104-1131).  TEST INFRASTRUCTURE ONLY (see vqt_oracle.h).
 * No test of the reference covers it: parity unpinned beyond this line-by-line restatement.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* update.rs:1110-1112: pitch class of bin 0 relative to C4 = 261.626 Hz */
int orc_chroma_bin0_pitch_class(float min_freq)
{
    const float semitones_from_c4 = 12.0f * log2f(min_freq / 261.626f);
    return (((int)roundf(semitones_from_c4) % 12) + 12) % 12;
}

/* update.rs:1114-1131: out[12] = per-pitch-class sum of 10^(dB/10), divided by its maximum if that is > 0 */
void orc_chroma(const float *db, size_t n_buckets, float min_freq, uint32_t buckets_per_octave, float *out)
{
    const int pc0 = orc_chroma_bin0_pitch_class(min_freq);
    for (int c = 0; c < 12; ++c) out[c] = 0.0f;
    for (size_t bin = 0; bin < n_buckets; ++bin) {
        const size_t semitone = (size_t)roundf((float)(bin * 12) / (float)buckets_per_octave);
        const int pc = ((int)semitone + pc0) % 12;
        out[pc] += powf(10.0f, db[bin] / 10.0f);
    }
    float mx = 0.0f;
    for (int c = 0; c < 12; ++c) mx = fmaxf(mx, out[c]);
    if (mx > 0.0f)
        for (int c = 0; c < 12; ++c) out[c] /= mx;
}

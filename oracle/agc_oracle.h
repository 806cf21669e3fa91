/* agc_oracle.h -- CPU restatement of dagc_fork::MonoAgc.  TEST INFRASTRUCTURE ONLY (see vqt_oracle.h).
 * Pinning: the reference's only test for it (dagc_fork/src/lib.rs:93-108, frozen gain stays 1.0, unfrozen gain
 * moves) is restated in tests/test_agc.py; there are no golden vectors. */
#ifndef AGC_ORACLE_H
#define AGC_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
int  orc_agc_check(float desired_output_rms, float distortion_factor);
void orc_agc_process(float *samples, size_t n, float desired_output_rms, float distortion_factor, float *gain,
                     int frozen);
void orc_agc_process_chunks(float *samples, size_t n, size_t chunk, float desired_output_rms, float distortion_factor,
                            float silence_threshold, float *gain);
#ifdef __cplusplus
}
#endif
#endif

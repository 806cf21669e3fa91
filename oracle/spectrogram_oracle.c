/*
 * spectrogram_oracle.c -- CPU restatement of the spectrogram ring, SpectrogramMode::VQT
 * (pitchvis_viewer/src/display_system/update.rs:930-1088).  TEST INFRASTRUCTURE ONLY (see vqt_oracle.h).
 * No test of the reference covers it: parity unpinned beyond this line-by-line restatement.  The per-bin RGB bytes
 * come from pitchvis_colors::calculate_color through the `lab` crate (absent from the tree): an input here.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* `(x).clamp(0.0, 255.0) as u8`: saturating, truncating, NaN -> 0 */
static uint8_t to_u8(float x)
{
    if (!(x > 0.0f)) return 0;
    if (x >= 255.0f) return 255;
    return (uint8_t)x;
}

/* update.rs:965-975 + :989: the alpha byte of one bin */
uint8_t orc_spectrogram_alpha(float value_db, float max_val)
{
    float brightness = 0.0f;
    if (max_val > 0.0f) {
        const float normalized = value_db / (max_val + 0.001f);
        const float d = 1.0f - normalized;
        brightness = (1.0f - d * d) * 1.5f;                 /* powf(2.0): the correctly rounded square */
        brightness = brightness < 0.0f ? 0.0f : (brightness > 1.0f ? 1.0f : brightness);
        if (brightness != brightness) brightness = 0.0f;   /* f32::clamp passes NaN through; `as u8` maps it to 0 */
    }
    return to_u8(brightness * 255.0f * 1.2f);
}

/* One call of update_spectrogram_system in VQT mode: write the row of `write_index`, clear the next one, advance. */
void orc_spectrogram_vqt_step(const float *smoothed, size_t width, const uint8_t *bin_rgb, uint8_t *image, size_t height,
                              size_t *write_index)
{
    const size_t w = *write_index;
    float max_val = 0.0f;
    for (size_t b = 0; b < width; ++b) max_val = fmaxf(max_val, smoothed[b]);   /* update.rs:965 */
    for (size_t b = 0; b < width; ++b) {
        uint8_t *px = image + ((height - 1 - w) * width + b) * 4;                 /* update.rs:983 */
        px[0] = bin_rgb[3 * b];
        px[1] = bin_rgb[3 * b + 1];
        px[2] = bin_rgb[3 * b + 2];
        px[3] = orc_spectrogram_alpha(smoothed[b], max_val);
    }
    const size_t next = (w + 1) % height;                                        /* update.rs:1065-1078 */
    memset(image + (height - 1 - next) * width * 4, 0, width * 4);
    *write_index = next;
}

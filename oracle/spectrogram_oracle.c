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

/* ------------------------------------------------------------------------------------------------------------
 * SpectrogramMode::Peaks (update.rs:997-1062) and the colour function it calls per peak,
 * pitchvis_colors::calculate_color (pitchvis_colors/src/lib.rs:93-119).  calculate_color goes through the `lab`
 * crate (Cargo.lock: lab 0.11.0), which is NOT in the reference tree: its sRGB <-> XYZ <-> L*a*b* <-> LCh formulas
 * are restated here from the crate's published algorithm (Bruce Lindbloom's kappa / epsilon, D65 white from the
 * sRGB primaries) -- PARITY UNPINNED for the colour bytes.
 * ------------------------------------------------------------------------------------------------------------ */
static const float kColors[12][3] = {   /* pitchvis_colors/src/lib.rs:19-34 */
    {0.85f, 0.36f, 0.36f}, {0.01f, 0.52f, 0.71f}, {0.97f, 0.76f, 0.05f}, {0.45f, 0.34f, 0.63f},
    {0.47f, 0.77f, 0.22f}, {0.78f, 0.32f, 0.52f}, {0.00f, 0.64f, 0.56f}, {0.95f, 0.54f, 0.23f},
    {0.30f, 0.37f, 0.64f}, {1.00f, 0.96f, 0.03f}, {0.57f, 0.30f, 0.55f}, {0.12f, 0.71f, 0.34f}};
static const float kGrayLevel = 60.0f;   /* lib.rs:56 */
static const float kEasingPow = 1.3f;    /* lib.rs:57 */

static const float kKappa = 24389.0f / 27.0f, kEpsilon = 216.0f / 24389.0f, kCbrtEpsilon = 6.0f / 29.0f;
static const float kS0 = 0.003130668442500564f;
static const float kWhiteX = 0.9504492182750991f, kWhiteZ = 1.0889166484304715f;

static float srgb_to_linear_255(float c)
{
    const float e0_255 = 12.92f * kS0 * 255.0f;
    if (c > e0_255) return powf((c + 0.055f * 255.0f) / (1.055f * 255.0f), 2.4f);
    return c / (12.92f * 255.0f);
}
static float xyz_to_lab_map(float c) { return c > kEpsilon ? powf(c, 1.0f / 3.0f) : (kKappa * c + 16.0f) / 116.0f; }
static float linear_to_srgb(float c)
{
    float v = c > kS0 ? 1.055f * powf(c, 1.0f / 2.4f) - 0.055f : 12.92f * c;
    v = fminf(v, 1.0f);     /* .min(1.0).max(0.0): NaN falls to the other operand */
    v = fmaxf(v, 0.0f);
    return v;
}

/* calculate_color(buckets_per_octave, bucket, COLORS, GRAY_LEVEL, EASING_POW) -> rgb in [0, 1] (lib.rs:93-119) */
void orc_calculate_color(uint32_t buckets_per_octave, float bucket, float rgb_out[3])
{
    const float pitch_continuous = 12.0f * bucket / (float)buckets_per_octave;
    const float rounded = roundf(pitch_continuous);
    const size_t semitone = (size_t)(rounded > 0.0f ? (rounded >= 1.8446744e19f ? 0xffffffffffffffffull : (unsigned long long)rounded) : 0) % 12;
    uint8_t base[3];
    for (int i = 0; i < 3; ++i) base[i] = to_u8(kColors[semitone][i] * 255.0f);
    const float inaccuracy_cents = fabsf(pitch_continuous - rounded);
    /* LCh::from_rgb */
    const float r = srgb_to_linear_255((float)base[0]), g = srgb_to_linear_255((float)base[1]), b = srgb_to_linear_255((float)base[2]);
    const float X = r * 0.4124108464885388f + g * 0.3575845678529519f + b * 0.18045380393360833f;
    const float Y = r * 0.21264934272065283f + g * 0.7151691357059038f + b * 0.07218152157344333f;
    const float Z = r * 0.019331758429150258f + g * 0.11919485595098397f + b * 0.9503900340503373f;
    const float fx = xyz_to_lab_map(X / kWhiteX), fy = xyz_to_lab_map(Y), fz = xyz_to_lab_map(Z / kWhiteZ);
    float L = 116.0f * fy - 16.0f;
    const float A = 500.0f * (fx - fy), B = 200.0f * (fy - fz);
    float C = hypotf(A, B);
    const float H = atan2f(B, A);
    /* lib.rs:109-115 */
    const float saturation = 1.0f - powf(2.0f * inaccuracy_cents, kEasingPow);
    C *= saturation;
    L = saturation * L + (1.0f - saturation) * kGrayLevel;
    /* LCh::to_rgb */
    const float a2 = C * cosf(H), b2 = C * sinf(H);
    const float gy = (L + 16.0f) / 116.0f, gx = a2 / 500.0f + gy, gz = gy - b2 / 200.0f;
    const float xr = gx > kCbrtEpsilon ? gx * gx * gx : (gx * 116.0f - 16.0f) / kKappa;
    const float yr = L > kEpsilon * kKappa ? gy * gy * gy : L / kKappa;
    const float zr = gz > kCbrtEpsilon ? gz * gz * gz : (gz * 116.0f - 16.0f) / kKappa;
    const float x = xr * kWhiteX, y = yr, z = zr * kWhiteZ;
    const float lr = x * 3.240812398895283f - y * 1.5373084456298136f - z * 0.4985865229069666f;
    const float lg = x * -0.9692430170086407f + y * 1.8759663029085742f + z * 0.04155503085668564f;
    const float lb = x * 0.055638398436112804f - y * 0.20400746093241362f + z * 1.0571295702861434f;
    const float s[3] = {linear_to_srgb(lr), linear_to_srgb(lg), linear_to_srgb(lb)};
    for (int i = 0; i < 3; ++i) rgb_out[i] = (float)to_u8(roundf(s[i] * 255.0f)) / 255.0f;   /* [u8; 3] -> f32 / 255 */
}

/* One call of update_spectrogram_system in Peaks mode (update.rs:997-1062 + the shared tail :1065-1081).
 * peaks: n_peaks (center, size) pairs in the order of AnalysisState::peaks_continuous. */
void orc_spectrogram_peaks_step(const float *peaks, size_t n_peaks, uint32_t buckets_per_octave, size_t width, uint8_t *image,
                                size_t height, size_t *write_index)
{
    const size_t w = *write_index;
    const float radius = 2.0f;                                                    /* PEAK_RADIUS */
    float max_size = 0.0f;
    for (size_t i = 0; i < n_peaks; ++i) max_size = fmaxf(max_size, peaks[2 * i + 1]);
    if (max_size > 0.0f) {
        for (size_t i = 0; i < n_peaks; ++i) {
            const float center = peaks[2 * i], size = peaks[2 * i + 1];
            const float d = 1.0f - size / max_size;
            float brightness = (1.0f - d * d) * 1.5f;
            brightness = brightness < 0.0f ? 0.0f : (brightness > 1.0f ? 1.0f : brightness);
            const uint32_t bps = buckets_per_octave / 12;
            const float semitone_offset = (float)(buckets_per_octave - 3 * bps);
            float rgb[3];
            orc_calculate_color(buckets_per_octave, fmodf(center + semitone_offset, (float)buckets_per_octave), rgb);
            const float lo = fmaxf(floorf(center - radius), 0.0f), hi = fminf(ceilf(center + radius), (float)width);
            const size_t min_bin = lo > 0.0f ? (size_t)lo : 0, max_bin = hi > 0.0f ? (size_t)hi : 0;   /* `as usize` */
            for (size_t bin = min_bin; bin < max_bin; ++bin) {
                const float distance = fabsf((float)bin - center);
                if (distance <= radius) {
                    const float falloff = expf(-distance * distance / (radius * radius * 0.5f));
                    const float pixel_brightness = brightness * falloff;
                    uint8_t *px = image + ((height - 1 - w) * width + bin) * 4;
                    px[0] = to_u8(rgb[0] * 255.0f * 1.2f);
                    px[1] = to_u8(rgb[1] * 255.0f * 1.2f);
                    px[2] = to_u8(rgb[2] * 255.0f * 1.2f);
                    px[3] = to_u8(pixel_brightness * 255.0f * 1.2f);
                }
            }
        }
    }
    const size_t next = (w + 1) % height;
    memset(image + (height - 1 - next) * width * 4, 0, width * 4);
    *write_index = next;
}

// kernel_builder.cpp -- host construction of the VQT spectral kernel (product code).
//
// Follows the arithmetic of pitchvis_analysis/src/vqt.rs:517-852 so that the kernel
// *values* and the sparsity pattern are the reference's.  All f32 expressions are
// evaluated in the reference's order (this file is built with -ffp-contract=off).
// The reference FFTs each wavelet with rustfft in f32 (vqt.rs:808); here the
// transform runs in f64 and is rounded to f32 once, which is the same value up to
// rustfft's own rounding noise ("kernel values").
#include "kernel_builder.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>

namespace pvqt_host {

// ---- log sink (the `log` crate's global logger, as a C callback) ----------------------------------------
namespace {
std::mutex g_log_mutex;
pvqt_log_fn g_log_fn = nullptr;
void *g_log_user = nullptr;
std::atomic<int> g_log_level{0};
}  // namespace

bool log_enabled(int level) { return level <= g_log_level.load(std::memory_order_relaxed); }

void log_line(int level, const std::string &message)
{
    std::lock_guard<std::mutex> lock(g_log_mutex);
    if (g_log_fn && level <= g_log_level.load(std::memory_order_relaxed)) g_log_fn(level, message.c_str(), g_log_user);
}

namespace {

constexpr float  kPiF = 3.14159274101257324219f;  // std::f32::consts::PI
constexpr double kPiD = 3.14159265358979323846;

using cd = std::complex<double>;
using cf = std::complex<float>;

// Rust `as usize` / `as u32` on f32: truncation toward zero, saturating, NaN -> 0.
inline uint64_t trunc_to_u64(float x)
{
    if (!(x > 0.0f)) return 0;
    if (x >= 18446744073709551615.0f) return UINT64_MAX;
    return static_cast<uint64_t>(x);
}
inline uint32_t trunc_to_u32(float x)
{
    uint64_t v = trunc_to_u64(x);
    return v > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(v);
}

// Forward DFT in double precision, X[k] = sum_n x[n] exp(-2 pi i k n / N).
// Power-of-two sizes: Stockham autosort radix-2; other sizes: direct O(N^2) sum.
class Dft64 {
public:
    explicit Dft64(size_t n) : n_(n), pow2_(n && !(n & (n - 1))), tw_(n), tmp_(n)
    {
        for (size_t k = 0; k < n; ++k) {
            double a = -2.0 * kPiD * static_cast<double>(k) / static_cast<double>(n);
            tw_[k] = cd(std::cos(a), std::sin(a));
        }
    }
    size_t size() const { return n_; }
    void forward(std::vector<cd> &x)
    {
        if (!pow2_) { direct(x); return; }
        cd *a = x.data(), *b = tmp_.data();
        size_t s = 1;
        for (size_t n = n_; n > 1; n >>= 1, s <<= 1) {
            size_t m = n >> 1, tstep = n_ / n;
            for (size_t p = 0; p < m; ++p) {
                cd w = tw_[p * tstep];
                for (size_t q = 0; q < s; ++q) {
                    cd u = a[q + s * p], v = a[q + s * (p + m)];
                    b[q + s * (2 * p)] = u + v;
                    b[q + s * (2 * p + 1)] = (u - v) * w;
                }
            }
            std::swap(a, b);
        }
        if (a != x.data()) std::memcpy(x.data(), a, sizeof(cd) * n_);
    }

private:
    void direct(std::vector<cd> &x)
    {
        for (size_t k = 0; k < n_; ++k) {
            cd acc(0, 0);
            for (size_t m = 0; m < n_; ++m) acc += x[m] * tw_[(k * m) % n_];
            tmp_[k] = acc;
        }
        x = tmp_;
    }
    size_t n_;
    bool pow2_;
    std::vector<cd> tw_, tmp_;
};

struct PanicMessage {
    const char *text;
};

// Vqt::calculate_filter, vqt.rs:769-852.  `v` receives scaled_n_fft coefficients.
// calculate_bandwidth + find_3db_points + util::arg_max (vqt.rs:956-989, util.rs:49-58): the -3 dB points of the
// (decimated) frequency response, a crude diagnostic the reference computes for its coverage-gap warning
void bandwidth_3db(const std::vector<float> &response, float scaled_sr, float &lo_hz, float &hi_hz)
{
    size_t center = 0;
    float best = -std::numeric_limits<float>::max();          // fold from f32::MIN, strict >
    for (size_t i = 0; i < response.size(); ++i)
        if (response[i] > best) { best = response[i]; center = i; }
    const float threshold = response[center] / std::sqrt(2.0f);
    size_t lower = center, upper = center;
    while (lower > 0 && response[lower] > threshold) --lower;
    while (upper < response.size() - 1 && response[upper] > threshold) ++upper;
    lo_hz = static_cast<float>(lower) * scaled_sr / static_cast<float>(response.size());
    hi_hz = static_cast<float>(upper) * scaled_sr / static_cast<float>(response.size());
}

std::string fmt(const char *format, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, format);
    std::vsnprintf(buf, sizeof(buf), format, ap);
    va_end(ap);
    return buf;
}

bool calculate_filter(float sr, float sparsity_quantile, uint64_t sr_scaling, const FilterParams &fp,
                      uint64_t win_begin, uint64_t win_end, float window_center, Dft64 &fft,
                      std::vector<cf> &v, std::vector<cd> &work, std::vector<float> &mags, PanicMessage &panic,
                      float &band_lo_hz, float &band_hi_hz)
{
    const float m = static_cast<float>(sr_scaling);
    const float scaled_freq = fp.freq * m;                                              // :778
    const float scaled_window_length = fp.window_length / m;                            // :779
    const uint64_t len = trunc_to_u64(std::round(scaled_window_length));                // :780
    const float scaled_window_center = (window_center - static_cast<float>(win_begin)) / m;  // :781
    const uint64_t center = trunc_to_u64(std::floor(scaled_window_center));             // :782
    const size_t scaled_n_fft = static_cast<size_t>((win_end - win_begin) / sr_scaling);  // :783

    if (len > scaled_n_fft) {                                                           // :785
        panic.text = "assertion failed: scaled_window_length_rounded <= scaled_n_fft";
        return false;
    }
    if (center < len / 2) {                                                             // :786-788
        panic.text = "filter window must fit between the start of its group window and the common window center";
        return false;
    }
    const uint64_t filter_begin = center - len / 2;
    if (filter_begin + len > scaled_n_fft) {                                            // :789-792
        panic.text = "filter window must end before the end of its group window";
        return false;
    }

    v.assign(scaled_n_fft, cf(0.0f, 0.0f));                                             // :796
    const double hann_den = len > 1 ? static_cast<double>(len - 1) : 1.0;
    for (uint64_t i = 0; i < len; ++i) {                                                // :797-800
        // apodize::hanning_iter(len): symmetric Hann evaluated in f64, cast to f32
        const double w = len > 1 ? 0.5 - 0.5 * std::cos(2.0 * kPiD * static_cast<double>(i) / hann_den) : 1.0;
        // Complex32::i() * 2.0 * PI * (i as f32) * scaled_freq / sr -- strictly left to right
        float phase = 2.0f;
        phase *= kPiF;
        phase *= static_cast<float>(i);
        phase *= scaled_freq;
        phase /= sr;
        const float wf = static_cast<float>(w);
        v[filter_begin + i] = cf(wf * std::cos(phase), wf * std::sin(phase));
    }

    float norm_1 = 0.0f;                                                                // :804
    for (const cf &z : v) norm_1 += std::hypot(z.real(), z.imag());
    for (cf &z : v) z = cf(z.real() / norm_1, z.imag() / norm_1);                       // :805

    work.resize(scaled_n_fft);
    for (size_t i = 0; i < scaled_n_fft; ++i) work[i] = cd(v[i].real(), v[i].imag());
    fft.forward(work);                                                                  // :808
    for (size_t i = 0; i < scaled_n_fft; ++i)                                           // :811 conj
        v[i] = cf(static_cast<float>(work[i].real()), -static_cast<float>(work[i].imag()));

    // :813-842 -- drop the smallest coefficients carrying (1 - q) of the L1 mass
    mags.resize(scaled_n_fft);
    for (size_t i = 0; i < scaled_n_fft; ++i) mags[i] = std::hypot(v[i].real(), v[i].imag());
    bandwidth_3db(mags, sr / m, band_lo_hz, band_hi_hz);                               // :818-819
    std::vector<float> &sorted = mags;
    std::sort(sorted.begin(), sorted.end());                                            // :823
    float v_abs_sum = 0.0f;
    for (float a : sorted) v_abs_sum += a;                                              // :824
    const float limit = (1.0f - sparsity_quantile) * v_abs_sum;                         // :827
    float accum = 0.0f;
    size_t cutoff_idx = 0;
    while (accum < limit) {
        if (cutoff_idx >= scaled_n_fft) {
            panic.text = "index out of bounds while accumulating the sparsity cutoff";
            return false;
        }
        accum += sorted[cutoff_idx];
        ++cutoff_idx;
    }
    const float cutoff_value = cutoff_idx == 0 ? 0.0f : sorted[cutoff_idx - 1];         // :831-835
    size_t erased = 0;
    for (cf &z : v)                                                                     // :837-842
        if (std::hypot(z.real(), z.imag()) < cutoff_value) { z = cf(0.0f, 0.0f); ++erased; }
    if (log_enabled(kLogDebug))                                                         // :843-846
        log_line(kLogDebug, fmt("for freq %g erased %zu points below %g with sum %g out of total %g", fp.freq, erased,
                                cutoff_value, accum, v_abs_sum));
    return true;
}

struct Triplet {
    int32_t row, col;
    cf value;
};

void to_csr(std::vector<Triplet> &t, int32_t rows, int32_t cols, Csr &m)
{
    // sprs TriMat::to_csr: rows with ascending column indices
    std::stable_sort(t.begin(), t.end(), [](const Triplet &a, const Triplet &b) {
        return a.row != b.row ? a.row < b.row : a.col < b.col;
    });
    m.rows = rows;
    m.cols = cols;
    m.indptr.assign(static_cast<size_t>(rows) + 1, 0);
    m.indices.resize(t.size());
    m.data.resize(t.size());
    for (size_t i = 0; i < t.size(); ++i) {
        m.indptr[t[i].row + 1] += 1;
        m.indices[i] = t[i].col;
        m.data[i] = t[i].value;
    }
    for (int32_t r = 0; r < rows; ++r) m.indptr[r + 1] += m.indptr[r];
}

}  // namespace

bool filter_bank_params(const pvqt_params &p, std::vector<FilterParams> &out, BuildError &err)
{
    const size_t nb = static_cast<size_t>(p.buckets_per_octave) * p.octaves;
    const float bpo = static_cast<float>(p.buckets_per_octave);
    if (nb == 0 || !(p.sr > 0.0f) || p.n_fft == 0) {
        err.status = PVQT_INVALID_ARGUMENT;
        err.message = "n_buckets, sr and n_fft must be positive";
        return false;
    }
    const float highest_frequency = p.min_freq * std::pow(2.0f, static_cast<float>(nb - 1) / bpo);  // :518-521
    const float nyquist_frequency = p.sr / 2.0f;                                                    // :522
    if (highest_frequency > nyquist_frequency) {                                                    // :523-528
        err.status = PVQT_ABOVE_NYQUIST;
        err.highest_frequency = highest_frequency;
        err.nyquist_frequency = nyquist_frequency;
        err.message = "the highest VQT bin frequency (" + std::to_string(highest_frequency) +
                      " Hz) exceeds the Nyquist frequency (" + std::to_string(nyquist_frequency) +
                      " Hz); reduce octaves or increase the sample rate";
        return false;
    }
    const float r = std::pow(2.0f, 1.0f / bpo);                     // :532
    const float alpha = (r * r - 1.0f) / (r * r + 1.0f);            // :533

    out.resize(nb);
    for (size_t k = 0; k < nb; ++k) {
        FilterParams &f = out[k];
        f.freq = p.min_freq * std::pow(2.0f, static_cast<float>(k) / bpo);             // :537-538
        f.window_length = p.quality * p.sr / (alpha * f.freq + p.gamma);               // :539
        constexpr float kGraceFactor = 1.15f;                                          // :545
        const float minimum_scaled_sr = std::ceil(f.freq * 2.0f * kGraceFactor);       // :546
        const uint32_t shift = trunc_to_u32(std::floor(std::log2(p.sr / minimum_scaled_sr)));  // :548
        f.sr_downscaling_factor = shift < 64 ? (uint64_t{1} << shift) : 0;             // :549
        const uint32_t wshift = trunc_to_u32(std::floor(std::log2(static_cast<float>(p.n_fft) / f.window_length)));  // :554
        f.minimum_needed_window_size = wshift < 64 ? (p.n_fft >> wshift) : 0;          // :555
    }
    const float longest_window = out[0].window_length;              // :567
    if (longest_window > static_cast<float>(p.n_fft)) {             // :568-573
        err.status = PVQT_WINDOW_EXCEEDS_NFFT;
        err.window_length = longest_window;
        err.n_fft = p.n_fft;
        err.message = "the longest filter window (" + std::to_string(longest_window) + " samples) exceeds n_fft (" +
                      std::to_string(p.n_fft) + " samples); increase n_fft or gamma, or decrease quality";
        return false;
    }
    return true;
}

bool build_kernel(const pvqt_params &p, Kernel &out, BuildError &err)
{
    std::vector<FilterParams> filters;
    if (!filter_bank_params(p, filters, err)) return false;
    const size_t nb = filters.size();

    const float max_window_length = filters[0].window_length;                         // :604
    const float window_center = static_cast<float>(p.n_fft) - max_window_length / 2.0f;  // :605

    struct RateGroup {
        uint64_t factor, begin, end;
        size_t first, count;
    };
    std::vector<RateGroup> rate_groups;                                               // :616-642
    for (size_t i = 0; i < nb;) {
        size_t j = i;
        uint64_t window_size = 0;
        while (j < nb && filters[j].sr_downscaling_factor == filters[i].sr_downscaling_factor) {
            window_size = std::max(window_size, filters[j].minimum_needed_window_size);
            ++j;
        }
        RateGroup rg{filters[i].sr_downscaling_factor, 0, 0, i, j - i};
        const float half = static_cast<float>(window_size) / 2.0f;
        if ((window_center + half) < static_cast<float>(p.n_fft)) {                   // :627
            rg.begin = trunc_to_u64(window_center - half);                            // :630
            rg.end = trunc_to_u64(window_center + half);                              // :631
        } else {
            rg.begin = p.n_fft - window_size;                                         // :634
            rg.end = p.n_fft;
        }
        if (rg.factor == 0 || rg.end <= rg.begin) {
            err.status = PVQT_PANIC;
            err.message = "degenerate rate group (window size or downscaling factor is zero)";
            return false;
        }
        rate_groups.push_back(rg);
        i = j;
    }

    const float kernel_gain = std::sqrt(p.sr);                                        // :646

    out = Kernel{};
    out.n_buckets = nb;
    out.band_lo_hz.assign(nb, 0.0f);
    out.band_hi_hz.assign(nb, 0.0f);
    float last_upper_bandwidth = 0.0f;                                                // :650
    std::vector<cf> v;
    std::vector<cd> work;
    std::vector<float> mags;
    for (size_t a = 0; a < rate_groups.size();) {                                     // :653-754
        size_t b = a, n_filters = 0;
        while (b < rate_groups.size() && rate_groups[b].begin == rate_groups[a].begin &&
               rate_groups[b].end == rate_groups[a].end) {
            n_filters += rate_groups[b].count;
            ++b;
        }
        WindowGroup wg;
        wg.window_begin = rate_groups[a].begin;
        wg.window_end = rate_groups[a].end;
        const uint64_t window_size = wg.window_size();                                // :657
        const size_t n_spectrum = static_cast<size_t>(window_size / 2 + 1);           // :658
        if (log_enabled(kLogDebug))                                                   // :661-667
            log_line(kLogDebug, fmt("window (%llu, %llu) (%llu samples): %zu filters in %zu rate group(s)",
                                    (unsigned long long)wg.window_begin, (unsigned long long)wg.window_end,
                                    (unsigned long long)window_size, n_filters, b - a));
        std::vector<Triplet> pos, neg;
        int32_t row = 0;
        for (size_t g = a; g < b; ++g) {
            const RateGroup &rg = rate_groups[g];
            const size_t scaled_n_fft = static_cast<size_t>(window_size / rg.factor); // :674
            if (scaled_n_fft == 0) {
                err.status = PVQT_PANIC;
                err.message = "rate group window is shorter than its downscaling factor";
                return false;
            }
            Dft64 fft(scaled_n_fft);                                                  // :675
            for (size_t f = 0; f < rg.count; ++f) {
                PanicMessage panic{nullptr};
                const FilterParams &fp = filters[rg.first + f];
                float lo_hz = 0.0f, hi_hz = 0.0f;
                if (!calculate_filter(p.sr, p.sparsity_quantile, rg.factor, fp, wg.window_begin,
                                      wg.window_end, window_center, fft, v, work, mags, panic, lo_hz, hi_hz)) {
                    err.status = PVQT_PANIC;
                    err.message = panic.text ? panic.text : "panic in calculate_filter";
                    return false;
                }
                out.band_lo_hz[rg.first + f] = lo_hz;
                out.band_hi_hz[rg.first + f] = hi_hz;
                if (log_enabled(kLogDebug))                                           // :688-694
                    log_line(kLogDebug, fmt("filter at %.1f Hz: window %.1f samples, -3 dB band (%.2f, %.2f) Hz", fp.freq,
                                            fp.window_length, lo_hz, hi_hz));
                if (last_upper_bandwidth > 0.0f && lo_hz > last_upper_bandwidth) {    // :695-710
                    out.coverage_gaps.push_back(static_cast<uint32_t>(rg.first + f));
                    if (log_enabled(kLogWarn))
                        log_line(kLogWarn,
                                 fmt("coverage gap below the filter at %.1f Hz: its -3 dB band starts at %.2f Hz but the "
                                     "previous filter's band ends at %.2f Hz (%.1f%% of this filter's bandwidth); decrease "
                                     "quality to close the gap",
                                     fp.freq, lo_hz, last_upper_bandwidth,
                                     100.0f * (lo_hz - last_upper_bandwidth) / (hi_hz - lo_hz)));
                }
                last_upper_bandwidth = hi_hz;                                         // :711
                for (size_t j = 0; j < scaled_n_fft; ++j) {                           // :725-735
                    const cf z = v[j];
                    if (z.real() == 0.0f && z.imag() == 0.0f) continue;               // z.is_zero()
                    // *z * kernel_gain / window_size as f32 (two roundings per component)
                    float re = z.real() * kernel_gain, im = z.imag() * kernel_gain;
                    re = re / static_cast<float>(window_size);
                    im = im / static_cast<float>(window_size);
                    if (j <= scaled_n_fft / 2)
                        pos.push_back({row, static_cast<int32_t>(j), cf(re, im)});
                    else
                        neg.push_back({row, static_cast<int32_t>(scaled_n_fft - j), cf(re, -im)});
                }
                ++row;
            }
        }
        to_csr(pos, static_cast<int32_t>(n_filters), static_cast<int32_t>(n_spectrum), wg.filter_bank);
        to_csr(neg, static_cast<int32_t>(n_filters), static_cast<int32_t>(n_spectrum), wg.negative_filter_bank);
        if (log_enabled(kLogDebug))                                                   // :741-746
            log_line(kLogDebug, fmt("window (%llu, %llu): kernel nnz %lld, conjugate-part nnz %lld",
                                    (unsigned long long)wg.window_begin, (unsigned long long)wg.window_end,
                                    (long long)wg.filter_bank.nnz(), (long long)wg.negative_filter_bank.nnz()));
        out.window_groups.push_back(std::move(wg));
        a = b;
    }
    // :756 Duration::from_secs_f32((n_fft as f32 - window_center) / sr)
    out.delay_seconds = static_cast<double>((static_cast<float>(p.n_fft) - window_center) / p.sr);
    if (log_enabled(kLogInfo))                                                        // vqt.rs:468 (Vqt::new)
        log_line(kLogInfo, fmt("VQT analysis delay: %llu ms.", (unsigned long long)(out.delay_seconds * 1000.0)));
    return true;
}

}  // namespace pvqt_host

extern "C" int pvqt_set_log_callback(pvqt_log_fn fn, void *user, int max_level)
{
    std::lock_guard<std::mutex> lock(pvqt_host::g_log_mutex);
    pvqt_host::g_log_fn = fn;
    pvqt_host::g_log_user = user;
    pvqt_host::g_log_level.store(fn ? std::max(0, std::min(max_level, 3)) : 0);
    return PVQT_OK;
}

// kernel_builder.hpp -- host-side construction of the sparse multi-rate VQT kernel.
//
// Product code (runs once per parameter set, on the host, as in the reference:
// Vqt::vqt_kernel / calculate_filter, pitchvis_analysis/src/vqt.rs:599-852).
// The result feeds both the C-ABI introspection calls (Vqt::kernel()) and the
// device-side banded layout built in device_plan.cu.
#pragma once

#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "pvqt.h"

namespace pvqt_host {

// FilterParams, vqt.rs:370-384
struct FilterParams {
    float    freq;
    float    window_length;
    uint64_t sr_downscaling_factor;
    uint64_t minimum_needed_window_size;
};

// sprs::CsMat<Complex32> with column-sorted rows
struct Csr {
    int32_t rows = 0, cols = 0;
    std::vector<int32_t> indptr;             // rows + 1
    std::vector<int32_t> indices;            // nnz
    std::vector<std::complex<float>> data;   // nnz
    int64_t nnz() const { return static_cast<int64_t>(indices.size()); }
};

// WindowGroup, vqt.rs:388-404
struct WindowGroup {
    uint64_t window_begin = 0, window_end = 0;
    Csr filter_bank;
    Csr negative_filter_bank;  // nnz()==0 <=> None
    uint64_t window_size() const { return window_end - window_begin; }
};

struct Kernel {
    std::vector<WindowGroup> window_groups;
    double delay_seconds = 0.0;
    size_t n_buckets = 0;
    // diagnostic only (Filter::bandwidth_3db_in_hz, vqt.rs:417-420, :818-819): -3 dB band of every filter, and the
    // filters below which the reference warns about a coverage gap (vqt.rs:695-710)
    std::vector<float> band_lo_hz, band_hi_hz;
    std::vector<uint32_t> coverage_gaps;
};

struct BuildError {
    pvqt_status status = PVQT_OK;
    float highest_frequency = 0, nyquist_frequency = 0, window_length = 0;
    uint64_t n_fft = 0;
    std::string message;
};

// The reference logs through the `log` crate: info! the delay (vqt.rs:468), warn! coverage gaps between neighbouring
// filters' -3 dB bands (vqt.rs:695-710), debug! the structure of every window group and filter (vqt.rs:661-667,
// :688-694, :741-746, :843-846).  A sink installed with pvqt_set_log_callback receives the same lines; messages of a
// level the sink did not ask for are not even formatted.
enum LogLevel { kLogWarn = 1, kLogInfo = 2, kLogDebug = 3 };
bool log_enabled(int level);
void log_line(int level, const std::string &message);

// Vqt::filter_bank_params, vqt.rs:517-587
bool filter_bank_params(const pvqt_params &p, std::vector<FilterParams> &out, BuildError &err);

// Vqt::vqt_kernel, vqt.rs:599-759 (calls calculate_filter, vqt.rs:769-852)
bool build_kernel(const pvqt_params &p, Kernel &out, BuildError &err);

}  // namespace pvqt_host

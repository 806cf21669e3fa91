// pvqt_api.cu -- C-ABI layer of libpvqt.so (include/pvqt.h): handle management, the
// device-side plan (twiddles, banded kernel layout), launch orchestration, host<->device
// staging and the single-process multi-GPU dispatcher.
//
// There is no CPU compute path in this library: every calc entry point runs the sm_100a
// kernels of vqt_kernels.cu or fails with PVQT_CUDA_ERROR.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "kernel_builder.hpp"
#include "last_error.hpp"
#include "pvqt.h"
#include "pvqt_internal.hpp"
#include "vqt_device.cuh"

using namespace pvqt_dev;

namespace {

thread_local std::string g_last_error;

}  // namespace

void pvqt_detail::set_last_error(const std::string &message) { g_last_error = message; }

namespace {

int fail(pvqt_status st, const std::string &msg)
{
    g_last_error = msg;
    return st;
}

int cuda_fail(cudaError_t e, const char *what)
{
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return PVQT_CUDA_ERROR;
}

#define PVQT_CUDA(call)                                             \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);         \
    } while (0)

constexpr double kPi = 3.14159265358979323846;

// a * b without wrapping: false if the product does not fit size_t
bool mul_ok(size_t a, size_t b, size_t *out)
{
    if (a != 0 && b > SIZE_MAX / a) return false;
    *out = a * b;
    return true;
}

// (frames - 1) * hop + n_fft <= n_samples, evaluated without overflow (a huge hop or frame count must fail the
// length check, not wrap past it)
bool frames_fit(size_t frames, size_t hop, size_t n_fft, size_t n_samples)
{
    if (frames == 0) return true;
    if (n_samples < n_fft) return false;
    if (frames == 1 || hop == 0) return true;
    return (n_samples - n_fft) / hop >= frames - 1;
}

// mirrors plan_radix() of vqt_kernels.cu (host copy used to size the twiddle tables)
int host_plan_radix(int nc, int pass)
{
    static const int plans[][5] = {
        {32, 16, 2, 1, 1},    {64, 16, 4, 1, 1},     {128, 16, 8, 1, 1},    {256, 16, 16, 1, 1},
        {512, 16, 8, 4, 1},   {1024, 16, 16, 4, 1},  {2048, 16, 16, 8, 1},  {4096, 16, 16, 16, 1},
        {8192, 16, 16, 8, 4}, {16384, 16, 16, 16, 4},
    };
    for (const auto &p : plans)
        if (p[0] == nc) return pass < 4 ? p[1 + pass] : 1;
    return 1;
}

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    uint64_t generation = 0;  // bumped on every reallocation (cached CUDA graphs bake the pointer in)
    cudaError_t reserve(size_t want)
    {
        if (want <= bytes) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        ++generation;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

}  // namespace

struct pvqt {
    pvqt_params params{};
    pvqt_host::Kernel kernel;
    int device = 0;
    cudaStream_t stream = nullptr;           // compute (and device-pointer entry points)
    cudaStream_t s_in = nullptr, s_out = nullptr;  // host->device / device->host copies of the host entry points
    std::vector<cudaEvent_t> events;         // pipeline edges, reused across calls
    size_t staging_budget_samples = (size_t)256 << 20;  // device audio per super-batch (1 GiB)
    // The pipelined host entry of one super-batch is captured into a CUDA graph and replayed while
    // the caller keeps passing the same buffers and shapes (one launch instead of ~60 API calls).
    struct HostCallKey {
        const float *audio = nullptr; float *out = nullptr;
        size_t n_streams = 0, stream_stride = 0, n_samples = 0, hop = 0, frames_per_stream = 0;
        uint64_t generations = 0;
        bool operator==(const HostCallKey &o) const
        {
            return audio == o.audio && out == o.out && n_streams == o.n_streams && stream_stride == o.stream_stride &&
                   n_samples == o.n_samples && hop == o.hop && frames_per_stream == o.frames_per_stream &&
                   generations == o.generations;
        }
    } graph_key, graph_warm_key;   // warm: this call shape has run eagerly once (every plan built, every buffer reserved)
    bool graph_warm = false;
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_launches = 0;  // kernel launches one replay performs
    size_t segments_per_batch = 3;  // measured best on B200 + PCIe Gen5 (scripts/pcie_probe.py): 2-4 equal, 6+ slower
    bool use_graphs = false;  // PVQT_GRAPHS=1: replay buys nothing at 3 segments (the path is PCIe-bound), keep eager

    // device plan
    std::vector<void *> owned;          // device allocations freed on destroy
    FftParams fft{};                    // template (frames/spec filled per launch)
    SpmmParams spmm{};
    FusedParams fused{};                // K-spmm-db plan (one CTA per tile owns every row)
    bool fused_capable = false;         // false: the kernel is too large for one CTA -> K-spmm + K-db
    bool fused_ok = false;              // fused_capable and not switched off (pvqt_set_fused_epilogue)
    bool pipe_capable = false, pipe_ok = false;   // K-spmm-db as a persistent warp-specialised pipeline (spmm_pipe.cu)
    int pipe_sdft_configured = -1;                // combine staging (float2 per K-sdft group) the kernel's smem limit covers
    int sm_count = 148;
    ClusterParams cluster{};            // K-spmm-db, cluster form (coefficients stationary in shared memory)
    bool cluster_capable = false, cluster_ok = false;
    int cluster_max_active = 0;         // co-resident clusters (cudaOccupancyMaxActiveClusters)
    int fft_block_threads = 256;
    int fft_wave_ctas = 0;              // > 0: K-fft as at most this many persistent CTAs (PVQT_FFT_WAVE); 0: one item per CTA
    float ref_db = 0.0f;
    std::vector<uint32_t> col_lo, n_cols, spec_off;
    size_t first_sample_used = 0;
    size_t last_sample_used = 0;        // one past
    size_t upload_skip = 0;             // leading samples of a stream no kernel reads (multiple of 4: alignment is kept)
    uint32_t segment_cap = 16384;       // frames per copy/compute segment of the host-buffer entries (PVQT_SEGMENT_FRAMES)
    uint32_t chunk_frames = 131072;     // frames per launch (PVQT_CHUNK_FRAMES): measured 42.4 M frames/s at 8192, 46.7 M at
                                        // 131072 on 1024 streams x 511 frames (scripts/chunk_sweep.py); 850 MB of scratch

    // K-sdft plans, one per hop seen (tables depend on the hop); see select_sdft()
    struct SdftPlan {
        size_t hop = 0;
        std::vector<int> group_index;      // window groups this hop has tables for
        std::vector<SdftGroup> groups;     // device tables
        std::vector<SdftTcPlan> tc;        // per group: plan of the tcgen05 form (n_groups = 0: not available)
    };
    std::vector<SdftPlan> sdft_plans;
    // Which groups take the K-sdft path is decided from the frames per stream of the WHOLE job, never from the piece a
    // launch sees: the pipelined host entries cut a job into segments and pvqt_multi_* into shards, and a frame must
    // get the same bits whatever piece it lands in.  0: no enclosing job, use the launch's own count.
    size_t sdft_job_frames = 0;
    uint32_t last_sdft_mask = 0;           // window groups on the K-sdft path in the most recent launch (plan_info)
    bool sdft_enabled = true;              // pvqt_set_sliding_dft
    bool sdft_tensor_cores = true;         // mode 2 of pvqt_set_sliding_dft: partial sums on mma.sync (3xTF32)
    int sdft_tc = 0;                       // 1 (mode 3): tcgen05 partial sums for the groups mode 2 selects; 2 (PVQT_SDFT_TC=2,
                                           //    experiment): also the larger groups sdft_tc_worthwhile() accepts

    // scratch
    // One scratch set per launch lane: consecutive launch chains of a call alternate between kLanes internal
    // streams, so the tail of one chain (few CTAs left) overlaps the head of the next.
    static constexpr int kLanes = 3;
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        DeviceBuffer spec, power, sdft_c, sdft_r, tile_ready;
        // The scratch above is reused by every launch chain of the lane.  Chains on one stream are ordered by the
        // stream; when a chain is enqueued on another stream than the lane's previous one (a device-pointer entry
        // on a caller stream after a host entry, two caller streams, ...) it first waits for `scratch_free`, which
        // the previous chain recorded behind its last kernel.
        cudaEvent_t scratch_free = nullptr;
        cudaStream_t last_stream = nullptr;
        bool used = false;
        unsigned *sdft_done = nullptr;   // completion counter of the lane's K-sdft launches (device, 4 bytes)
        uint32_t sdft_expected = 0;      // its value once every launch issued so far has finished
    } lane[kLanes];
    bool tile_flags = false;            // PVQT_TILE_FLAGS=1: K-spmm-db starts a tile on the completion counters of K-sdft and of
                                        // the K-fft CTAs that write the tile, instead of on the whole K-fft grid
    cudaEvent_t lane_fork = nullptr;
    int host_lanes = 2;  // compute lanes of the pipelined host entries (PVQT_HOST_LANES)
    int n_lanes = 1;  // PVQT_LANES; measured on B200: 2 lanes -17 %, 3 lanes -29 % at 3507 frames
    DeviceBuffer d_audio, d_out;
    std::atomic<uint64_t> launches{0};
    bool capturing = false;             // a stream capture is in progress: no cross-stream scratch events

    // Per-frame entry (pvqt_calc_instant_db, the reference's real-time call at 60 FPS): persistent pinned staging,
    // only the samples the windows read are uploaded, and the whole call -- H2D, K-fft, K-spmm-db, D2H -- is one
    // captured graph; nothing is allocated and no event is created on the call path after the second call.
    struct Instant {
        float *h_in = nullptr, *h_out = nullptr;   // pinned
        float *d_in = nullptr, *d_out = nullptr;
        cudaGraphExec_t exec = nullptr;
        uint64_t scratch_generation = 0;           // of lane[0].spec when `exec` was captured
        int config = -1;                           // fused / sdft switches when `exec` was captured
        uint64_t graph_launches = 0;
        int calls = 0;
    } instant;

    // pvqt_calc_batch_analysis with one frame per call (what the reference's viewer does 60 times a second): pinned staging
    // both ways and one captured graph -- H2D, K-fft, K-spmm-db, K-analysis, D2H of the requested results
    struct FramePipe {
        float *h_in = nullptr, *d_in = nullptr, *d_db = nullptr;
        char *h_res = nullptr, *d_res = nullptr;   // pinned / device: the requested result arrays of one frame, back to back
        size_t h_res_bytes = 0;
        cudaGraphExec_t exec = nullptr;
        const void *analysis = nullptr;            // what `exec` was captured for
        uint64_t analysis_generation = 0, frame_time_ns = 0, scratch_generation = 0, max_peaks = 0;
        uint32_t wanted = 0;                       // bit i: result array i requested
        int config = -1;
        uint64_t graph_launches = 0;
        int calls = 0;
    } frame_pipe;

    // optional per-kernel timing (pvqt_set_profiling): event pairs around every launch
    bool frame_pipe_enabled = true;     // PVQT_FRAME_PIPE=0: one-frame pipeline calls take the general path
    bool profiling = false;
    struct Timed { cudaEvent_t a, b; int kind; };
    std::vector<Timed> timed;
};

struct pvqt_kernel {
    pvqt_host::Kernel kernel;
};

namespace {

template <typename T>
cudaError_t upload(pvqt *v, const std::vector<T> &host, const T **dev)
{
    void *p = nullptr;
    size_t bytes = std::max<size_t>(host.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    v->owned.push_back(p);
    if (!host.empty()) {
        e = cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
    }
    *dev = static_cast<const T *>(p);
    return cudaSuccess;
}

int build_fused_plan(pvqt *v);
int build_cluster_plan(pvqt *v);

int build_device_plan(pvqt *v)
{
    const auto &groups = v->kernel.window_groups;
    if (groups.empty() || groups.size() > (size_t)kMaxGroups)
        return fail(PVQT_UNSUPPORTED, "unsupported number of window groups: " + std::to_string(groups.size()));

    int max_threads_per_fft = 0;
    v->first_sample_used = v->params.n_fft;
    v->last_sample_used = 0;
    for (const auto &g : groups) {
        const uint64_t n = g.window_size();
        if (n < 64 || n > 32768 || (n & (n - 1)) != 0)
            return fail(PVQT_UNSUPPORTED, "window group of " + std::to_string(n) +
                                              " samples: the sm_100a FFT plans cover powers of two in [64, 32768]");
        max_threads_per_fft = std::max<int>(max_threads_per_fft, (int)(n / 2 / kPointsPerThread));
        v->first_sample_used = std::min<size_t>(v->first_sample_used, g.window_begin);
        v->last_sample_used = std::max<size_t>(v->last_sample_used, g.window_end);
    }
    v->fft_block_threads = max_threads_per_fft <= 256 ? 256 : (max_threads_per_fft <= 512 ? 512 : 1024);
    // K-fft transforms a window that begins on an odd sample one sample lower (vqt_kernels.cu, fft_group_body)
    v->upload_skip = (v->first_sample_used > 0 ? v->first_sample_used - 1 : 0) & ~(size_t)3;

    // ---- FFT descriptors -------------------------------------------------------------
    FftParams &F = v->fft;
    std::memset(&F, 0, sizeof(F));
    F.n_groups = (int)groups.size();
    int spec_cursor = 0;
    v->col_lo.clear(); v->n_cols.clear(); v->spec_off.clear();
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const auto &g = groups[gi];
        const int nc = (int)(g.window_size() / 2);
        int lo = nc, hi = 0;
        for (const pvqt_host::Csr *m : {&g.filter_bank, &g.negative_filter_bank})
            for (int32_t c : m->indices) { lo = std::min(lo, c); hi = std::max(hi, c); }
        if (hi < lo) { lo = 0; hi = 0; }  // group without coefficients: keep one (unused) column
        FftGroup &d = F.group[gi];
        d.window_begin = (int32_t)g.window_begin;
        d.log2_nc = 0;
        while ((1 << d.log2_nc) < nc) ++d.log2_nc;
        d.col_lo = lo;
        d.col_hi = hi;
        d.spec_offset = spec_cursor;
        d.frames_per_cta = v->fft_block_threads / (nc / kPointsPerThread);
        spec_cursor += hi - lo + 1;
        v->col_lo.push_back(lo); v->n_cols.push_back(hi - lo + 1); v->spec_off.push_back(d.spec_offset);

        int ns = 1;
        for (int pass = 0; pass < kMaxFftPasses; ++pass) {
            const int r = host_plan_radix(nc, pass);
            if (r == 1) break;
            if (pass > 0) {
                std::vector<float2> tw((size_t)(r - 1) * ns);
                for (int rr = 1; rr < r; ++rr)
                    for (int k = 0; k < ns; ++k) {
                        const double a = -2.0 * kPi * (double)rr * (double)k / ((double)ns * r);
                        tw[(size_t)(rr - 1) * ns + k] = make_float2((float)std::cos(a), (float)std::sin(a));
                    }
                cudaError_t e = upload(v, tw, &d.twiddle[pass]);
                if (e != cudaSuccess) return cuda_fail(e, "upload twiddles");
            }
            ns *= r;
        }
        std::vector<float2> st((size_t)(hi - lo + 1));
        for (int c = lo; c <= hi; ++c) {
            const double a = -2.0 * kPi * (double)c / (double)(2 * nc);
            st[c - lo] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
        cudaError_t e = upload(v, st, &d.split_twiddle);
        if (e != cudaSuccess) return cuda_fail(e, "upload split twiddles");
    }
    F.spec_stride = (spec_cursor + 7) & ~7;  // whole 8-column groups: row blocks stage from a multiple of 8

    // ---- banded layout of the spectral kernel: 64-row blocks, two adjacent rows per lane ----
    struct Row { int col0 = 0, len = 0, ncol0 = 0, nlen = 0; const pvqt_host::Csr *m = nullptr, *mn = nullptr; int local = 0; };
    const int nb = (int)v->kernel.n_buckets;
    std::vector<SpmmRowBlock> blocks;
    std::vector<int4> lane_meta;
    std::vector<float4> values;
    int max_cols = 8;
    int first_row = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const auto &g = groups[gi];
        const int spec = F.group[gi].spec_offset, lo = F.group[gi].col_lo;
        const int rows_in_group = g.filter_bank.rows;
        for (int r0 = 0; r0 < rows_in_group; r0 += kRowsPerBlock) {
            const int n_rows = std::min(kRowsPerBlock, rows_in_group - r0);
            Row rows[kRowsPerBlock];
            for (int i = 0; i < n_rows; ++i) {
                Row &rw = rows[i];
                const int r = r0 + i;
                rw.m = &g.filter_bank; rw.mn = &g.negative_filter_bank; rw.local = r;
                const int s = g.filter_bank.indptr[r], e = g.filter_bank.indptr[r + 1];
                if (e > s) { rw.col0 = spec + g.filter_bank.indices[s] - lo; rw.len = g.filter_bank.indices[e - 1] - g.filter_bank.indices[s] + 1; }
                if (g.negative_filter_bank.nnz() > 0) {
                    const int ns_ = g.negative_filter_bank.indptr[r], ne = g.negative_filter_bank.indptr[r + 1];
                    if (ne > ns_) {
                        rw.ncol0 = spec + g.negative_filter_bank.indices[ns_] - lo;
                        rw.nlen = g.negative_filter_bank.indices[ne - 1] - g.negative_filter_bank.indices[ns_] + 1;
                    }
                }
            }
            // per lane: the union band of its two rows
            struct Pair { int col0 = 0, len = 0, ncol0 = 0, nlen = 0; };
            Pair pairs[32];
            int width = 0, nwidth = 0, cmin = 1 << 30, cmax = -1;
            for (int l = 0; l < 32; ++l) {
                int a0 = 1 << 30, a1 = -1, n0 = 1 << 30, n1 = -1;
                for (int q = 0; q < kRowsPerLane; ++q) {
                    const Row &rw = rows[kRowsPerLane * l + q];
                    if (rw.len > 0) { a0 = std::min(a0, rw.col0); a1 = std::max(a1, rw.col0 + rw.len); }
                    if (rw.nlen > 0) { n0 = std::min(n0, rw.ncol0); n1 = std::max(n1, rw.ncol0 + rw.nlen); }
                }
                if (a1 >= 0) { pairs[l].col0 = a0; pairs[l].len = a1 - a0; cmin = std::min(cmin, a0); cmax = std::max(cmax, a1); }
                if (n1 >= 0) { pairs[l].ncol0 = n0; pairs[l].nlen = n1 - n0; cmin = std::min(cmin, n0); cmax = std::max(cmax, n1); }
                width = std::max(width, pairs[l].len);
                nwidth = std::max(nwidth, pairs[l].nlen);
            }
            if (cmax < 0) { cmin = 0; cmax = 1; }
            width = (width + kSpmmUnroll - 1) / kSpmmUnroll * kSpmmUnroll;  // zero-padded, see spmm_kernel
            SpmmRowBlock B{};
            B.first_row = first_row + r0;
            B.n_rows = n_rows;
            B.width = width;
            B.nwidth = nwidth;
            B.col_lo = cmin & ~7;
            B.n_cols = cmax - B.col_lo;
            B.val_base = (int)(values.size() / 32);
            values.resize(values.size() + (size_t)width * 32, make_float4(0.f, 0.f, 0.f, 0.f));
            B.nval_base = (int)(values.size() / 32);
            values.resize(values.size() + (size_t)nwidth * 32, make_float4(0.f, 0.f, 0.f, 0.f));
            max_cols = std::max(max_cols, B.n_cols);
            for (int l = 0; l < 32; ++l) {
                const Pair &pr = pairs[l];
                // lanes without a (conjugate) band run zero trips; keep their column inside the staged range
                lane_meta.push_back(make_int4(pr.len > 0 ? pr.col0 - B.col_lo : 0, pr.len,
                                              pr.nlen > 0 ? pr.ncol0 - B.col_lo : 0, pr.nlen));
                for (int q = 0; q < kRowsPerLane; ++q) {
                    const Row &rw = rows[kRowsPerLane * l + q];
                    if (!rw.m) continue;
                    for (int e = rw.m->indptr[rw.local]; e < rw.m->indptr[rw.local + 1]; ++e) {
                        const int j = spec + rw.m->indices[e] - lo - pr.col0;
                        float4 &slot = values[((size_t)B.val_base + j) * 32 + l];
                        (q == 0 ? slot.x : slot.z) = rw.m->data[e].real();
                        (q == 0 ? slot.y : slot.w) = rw.m->data[e].imag();
                    }
                    if (rw.nlen > 0)
                        for (int e = rw.mn->indptr[rw.local]; e < rw.mn->indptr[rw.local + 1]; ++e) {
                            const int j = spec + rw.mn->indices[e] - lo - pr.ncol0;
                            float4 &slot = values[((size_t)B.nval_base + j) * 32 + l];
                            // conj(Kneg X) = conj(Kneg) conj(X): keep conj(Kneg)
                            (q == 0 ? slot.x : slot.z) = rw.mn->data[e].real();
                            (q == 0 ? slot.y : slot.w) = -rw.mn->data[e].imag();
                        }
                }
            }
            blocks.push_back(B);
        }
        first_row += rows_in_group;
    }
    if (first_row != nb) return fail(PVQT_PANIC, "kernel rows do not add up to n_buckets");
    values.resize(values.size() + (size_t)kSpmmUnroll * 32, make_float4(0.f, 0.f, 0.f, 0.f));  // prefetch overrun

    SpmmParams &S = v->spmm;
    std::memset(&S, 0, sizeof(S));
    cudaError_t e;
    if ((e = upload(v, blocks, &S.blocks)) != cudaSuccess) return cuda_fail(e, "upload blocks");
    if ((e = upload(v, lane_meta, &S.lane_meta)) != cudaSuccess) return cuda_fail(e, "upload lane_meta");
    if ((e = upload(v, values, &S.values)) != cudaSuccess) return cuda_fail(e, "upload values");
    std::vector<int32_t> order(blocks.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return blocks[x].width + blocks[x].nwidth > blocks[y].width + blocks[y].nwidth;
    });
    if ((e = upload(v, order, &S.block_order)) != cudaSuccess) return cuda_fail(e, "upload block order");
    S.n_blocks = (int)blocks.size();
    S.n_buckets = nb;
    S.spec_stride = F.spec_stride;
    S.max_cols = max_cols;
    v->ref_db = 10.0f * std::log10(0.3f * 0.3f);  // vqt.rs:923,927

    if ((e = configure_kernels(max_cols)) != cudaSuccess)
        return cuda_fail(e, "configure kernels (is this an sm_100a device?)");
    int rc = build_fused_plan(v);
    if (rc != PVQT_OK) return rc;
    return build_cluster_plan(v);
}

// ---- K-spmm-db plan: row pairs sorted by band length, dealt to warps, bank-conflict-free lanes ----
int build_fused_plan(pvqt *v)
{
    // rows per lane: 4 would share every spectrum load between four kernel rows (half the shared-memory traffic per
    // FMA of 2) but leaves 5 warps per CTA: measured 68 us against 38 us on B200 (DESIGN.md, K-spmm-db), so 2 is
    // the default and PVQT_ROWS_PER_LANE=4 selects the four-row form.
    int R = 2;
    if (const char *e = std::getenv("PVQT_ROWS_PER_LANE")) R = std::atoi(e) == 4 ? 4 : 2;
    const int H = R / 2;
    const auto &groups = v->kernel.window_groups;
    const FftParams &F = v->fft;
    struct Unit {
        int col0 = 0, len = 0, ncol0 = 0, nlen = 0;   // bands in spectrum columns
        int first_row = 0, n_rows = 0;                // output rows
        int group = 0, local_row = 0;
    };
    std::vector<Unit> units;
    int first_row = 0, n_cols = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const auto &g = groups[gi];
        const int spec = F.group[gi].spec_offset, lo = F.group[gi].col_lo;
        n_cols = std::max(n_cols, spec + F.group[gi].col_hi - lo + 1);
        for (int r0 = 0; r0 < g.filter_bank.rows; r0 += R) {
            Unit u;
            u.group = (int)gi; u.local_row = r0; u.first_row = first_row + r0;
            u.n_rows = std::min(R, g.filter_bank.rows - r0);
            int a0 = 1 << 30, a1 = -1, n0 = 1 << 30, n1 = -1;
            for (int q = 0; q < u.n_rows; ++q) {
                const int r = r0 + q;
                const int s = g.filter_bank.indptr[r], e = g.filter_bank.indptr[r + 1];
                if (e > s) {
                    a0 = std::min(a0, spec + g.filter_bank.indices[s] - lo);
                    a1 = std::max(a1, spec + g.filter_bank.indices[e - 1] - lo + 1);
                }
                if (g.negative_filter_bank.nnz() > 0) {
                    const int ns = g.negative_filter_bank.indptr[r], ne = g.negative_filter_bank.indptr[r + 1];
                    if (ne > ns) {
                        n0 = std::min(n0, spec + g.negative_filter_bank.indices[ns] - lo);
                        n1 = std::max(n1, spec + g.negative_filter_bank.indices[ne - 1] - lo + 1);
                    }
                }
            }
            if (a1 >= 0) { u.col0 = a0; u.len = a1 - a0; }
            if (n1 >= 0) { u.ncol0 = n0; u.nlen = n1 - n0; }
            units.push_back(u);
        }
        first_row += g.filter_bank.rows;
    }
    // Order of the units = which 32 share a warp.  A warp walks max(len) + max(nlen) slots, so units are sorted by band
    // length, and the few units with a conjugate-part band (38 of 294 at the defaults) stay together: they go in as
    // one block, at the position of the length-sorted list that minimises the sum of the warps' walks (at the
    // defaults 408 slots per tile instead of the 435 of sorting by len + nlen).  Per-row arithmetic is unchanged.
    {
        std::vector<Unit> pos, neg;
        for (const Unit &u : units) (u.nlen > 0 ? neg : pos).push_back(u);
        auto by_len = [](const Unit &a, const Unit &b) { return a.len > b.len; };
        std::stable_sort(pos.begin(), pos.end(), by_len);
        std::stable_sort(neg.begin(), neg.end(), by_len);
        auto walk_slots = [&](size_t at) {
            long total = 0;
            int w = 0, nw = 0;
            const size_t n = pos.size() + neg.size();
            for (size_t i = 0; i < n; ++i) {
                const Unit &u = i < at ? pos[i] : (i < at + neg.size() ? neg[i - at] : pos[i - neg.size()]);
                w = std::max(w, u.len);
                nw = std::max(nw, u.nlen);
                if ((i & 31) == 31 || i + 1 == n) { total += w + nw; w = nw = 0; }
            }
            return total;
        };
        size_t best_at = 0;
        long best = -1;
        for (size_t at = 0; at <= pos.size(); ++at) {
            const long c = walk_slots(at);
            if (best < 0 || c < best) { best = c; best_at = at; }
        }
        units.assign(pos.begin(), pos.begin() + (long)best_at);
        units.insert(units.end(), neg.begin(), neg.end());
        units.insert(units.end(), pos.begin() + (long)best_at, pos.end());
    }
    const int n_warps = (int)((units.size() + 31) / 32);
    // Which warp walks which group of 32 units.  Warp w issues on SM sub-partition w % 4, and the walk is bound by the
    // FMA pipe of the busiest sub-partition: the groups (longest walk first) go to the sub-partition with the least
    // work so far that still has a free warp (groups sorted by length would put the three longest-but-one on
    // sub-partition 0: 128 / 115 / 86 / 80 slots at the defaults instead of ~102 each).
    {
        std::vector<int> walk((size_t)n_warps, 0);
        for (int w = 0; w < n_warps; ++w) {
            int a = 0, b = 0;
            for (size_t i = (size_t)w * 32; i < std::min(units.size(), (size_t)(w + 1) * 32); ++i) {
                a = std::max(a, units[i].len);
                b = std::max(b, units[i].nlen);
            }
            walk[(size_t)w] = a + b;
        }
        std::vector<int> order((size_t)n_warps);
        for (int w = 0; w < n_warps; ++w) order[(size_t)w] = w;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return walk[(size_t)x] > walk[(size_t)y]; });
        int cap[4];
        for (int s = 0; s < 4; ++s) cap[s] = (n_warps - s + 3) / 4;     // warps s, s + 4, s + 8, ...
        std::vector<int> group_of_warp((size_t)n_warps, -1), bin_of((size_t)n_warps, 0), best_bin;
        // exhaustive for the usual handful of warps (10 at the defaults: 25,200 assignments), greedy beyond
        long best_max = -1, nodes = 0;
        {
            int load[4] = {0, 0, 0, 0}, used[4] = {0, 0, 0, 0};
            std::function<void(int)> place = [&](int at) {
                if (++nodes > 4000000) return;
                const long cur = std::max(std::max(load[0], load[1]), std::max(load[2], load[3]));
                if (best_max >= 0 && cur >= best_max) return;
                if (at == n_warps) { best_max = cur; best_bin = bin_of; return; }
                const int g = order[(size_t)at];
                for (int s = 0; s < 4; ++s) {
                    if (used[s] >= cap[s]) continue;
                    bool twin = false;                                  // an equally filled, equally sized bin was tried already
                    for (int s2 = 0; s2 < s; ++s2) twin |= load[s2] == load[s] && used[s2] == used[s] && cap[s2] == cap[s];
                    if (twin) continue;
                    load[s] += walk[(size_t)g]; ++used[s]; bin_of[(size_t)g] = s;
                    place(at + 1);
                    load[s] -= walk[(size_t)g]; --used[s];
                }
            };
            place(0);
        }
        {
            int used[4] = {0, 0, 0, 0}, load[4] = {0, 0, 0, 0};
            for (int g : order) {
                int bin = -1;
                if (best_max >= 0) {
                    bin = best_bin[(size_t)g];
                } else {
                    for (int s = 0; s < 4; ++s)
                        if (used[s] < cap[s] && (bin < 0 || load[s] < load[bin])) bin = s;
                }
                group_of_warp[(size_t)(bin + 4 * used[bin])] = g;
                load[bin] += walk[(size_t)g];
                ++used[bin];
            }
        }
        std::vector<Unit> permuted;
        permuted.reserve((size_t)n_warps * 32);
        for (int w = 0; w < n_warps; ++w) {
            const size_t g = (size_t)group_of_warp[(size_t)w];
            for (size_t i = g * 32; i < (g + 1) * 32; ++i) permuted.push_back(i < units.size() ? units[i] : Unit{});
        }
        units.swap(permuted);   // trailing empty units of a short group walk nothing and own no rows
    }
    n_cols = std::min<int>((n_cols + 7) & ~7, F.spec_stride);
    // columns the unpredicated band walk may read: start column + the warp's (padded) width; bounded by
    // n_cols + the longest band + the 7-column placement shift
    int longest = 0;
    for (const Unit &u : units) longest = std::max(longest, std::max(u.len, u.nlen));
    const int touched_bound = n_cols + longest + 8;
    v->fused_capable = v->fused_ok = fused_supported(n_warps, touched_bound, (int)v->kernel.n_buckets, R);
    if (!v->fused_ok) return PVQT_OK;

    // columns a unit may start early (zero coefficients) to land on a free bank-group residue: 8 makes every LDS.128 of the
    // walk conflict-free but lengthens the walk by ~3.5 slots per unit (PVQT_CONFLICT_SHIFT)
    int max_shift = 8;
    if (const char *e = std::getenv("PVQT_CONFLICT_SHIFT")) max_shift = std::max(1, std::min(std::atoi(e), 8));
    std::vector<FusedWarp> warps((size_t)n_warps);
    std::vector<int4> lane_meta((size_t)n_warps * 32, make_int4(0, 0, 0, 0));
    std::vector<int2> lane_rows((size_t)n_warps * 32, make_int2(0, 0));
    std::vector<float4> values;
    for (int w = 0; w < n_warps; ++w) {
        // A lane reads the 64-byte record of column col0 + j at slot j; its bank group depends on the column
        // modulo 8 only, and all lanes advance together, so a quarter-warp (one LDS.128 wavefront) is
        // conflict-free for the whole band iff its lanes start on distinct residues (or the same column).
        // Units are placed greedily; a unit may start up to 7 columns early (zero coefficients) to get a
        // free residue.
        const int u0 = w * 32, u1 = std::min<int>((int)units.size(), u0 + 32);
        int residue_col[4][8];
        int count[4] = {0, 0, 0, 0};
        for (auto &q : residue_col) for (int &c : q) c = -1;
        int lane_of[32];
        for (int i = u0; i < u1; ++i) {
            Unit &u = units[i];
            int best_q = -1, best_shift = 0;
            for (int shift = 0; shift < max_shift && best_q < 0 && u.len > 0; ++shift) {
                const int c = u.col0 - shift;
                if (c < 0) break;
                for (int q = 0; q < 4; ++q) {
                    if (count[q] >= 8) continue;
                    const int held = residue_col[q][c & 7];
                    if (held != -1 && held != c) continue;
                    if (best_q < 0 || count[q] < count[best_q]) { best_q = q; best_shift = shift; }
                }
            }
            if (best_q < 0) {  // empty unit, or no conflict-free slot: any free lane
                for (int q = 0; q < 4; ++q)
                    if (count[q] < 8 && (best_q < 0 || count[q] < count[best_q])) best_q = q;
                best_shift = 0;
            }
            if (u.len > 0) {
                u.col0 -= best_shift;
                u.len += best_shift;
                if (residue_col[best_q][u.col0 & 7] == -1) residue_col[best_q][u.col0 & 7] = u.col0;
            }
            lane_of[i - u0] = best_q * 8 + count[best_q]++;
        }
        int width = 0, nwidth = 0;
        for (int i = u0; i < u1; ++i) { width = std::max(width, units[i].len); nwidth = std::max(nwidth, units[i].nlen); }
        FusedWarp &W = warps[(size_t)w];
        W.width = width;
        W.nwidth = nwidth;
        W.val_base = (int)(values.size() / (H * 32));
        values.resize(values.size() + (size_t)width * H * 32, make_float4(0.f, 0.f, 0.f, 0.f));
        W.nval_base = (int)(values.size() / (H * 32));
        values.resize(values.size() + (size_t)nwidth * H * 32, make_float4(0.f, 0.f, 0.f, 0.f));
        for (int i = u0; i < u1; ++i) {
            const Unit &u = units[i];
            const int l = lane_of[i - u0];
            const auto &g = groups[(size_t)u.group];
            const int spec = F.group[u.group].spec_offset, lo = F.group[u.group].col_lo;
            lane_meta[(size_t)w * 32 + l] = make_int4(u.col0, u.len, u.ncol0, u.nlen);
            lane_rows[(size_t)w * 32 + l] = make_int2(u.first_row, u.n_rows);
            for (int q = 0; q < u.n_rows; ++q) {
                const int r = u.local_row + q;
                for (int e = g.filter_bank.indptr[r]; e < g.filter_bank.indptr[r + 1]; ++e) {
                    const int j = spec + g.filter_bank.indices[e] - lo - u.col0;
                    float4 &slot = values[(((size_t)W.val_base + j) * H + q / 2) * 32 + l];
                    ((q & 1) == 0 ? slot.x : slot.z) = g.filter_bank.data[e].real();
                    ((q & 1) == 0 ? slot.y : slot.w) = g.filter_bank.data[e].imag();
                }
                if (u.nlen > 0)
                    for (int e = g.negative_filter_bank.indptr[r]; e < g.negative_filter_bank.indptr[r + 1]; ++e) {
                        const int j = spec + g.negative_filter_bank.indices[e] - lo - u.ncol0;
                        float4 &slot = values[(((size_t)W.nval_base + j) * H + q / 2) * 32 + l];
                        // conj(Kneg X) = conj(Kneg) conj(X): keep conj(Kneg)
                        ((q & 1) == 0 ? slot.x : slot.z) = g.negative_filter_bank.data[e].real();
                        ((q & 1) == 0 ? slot.y : slot.w) = -g.negative_filter_bank.data[e].imag();
                    }
            }
        }
    }
    if (std::getenv("PVQT_DEBUG_PLAN")) {
        int sub[4] = {0, 0, 0, 0};
        for (int w = 0; w < n_warps; ++w) {
            std::fprintf(stderr, "K-spmm-db warp %d: %d band + %d conjugate-part slots\n", w, warps[(size_t)w].width, warps[(size_t)w].nwidth);
            sub[w & 3] += warps[(size_t)w].width + warps[(size_t)w].nwidth;
        }
        std::fprintf(stderr, "slots per SM sub-partition: %d %d %d %d\n", sub[0], sub[1], sub[2], sub[3]);
    }
    values.resize(values.size() + (size_t)2 * kFusedRing * H * 32, make_float4(0.f, 0.f, 0.f, 0.f));  // prefetch overrun
    // Every CTA walks the same coefficients in the same order; identical copies at different addresses (CTA b
    // streams copy b % copies) would spread the requests over more L2 slices.  Measured with 16 copies on B200:
    // no change (DESIGN.md, K-spmm-db), so one copy is kept; the mechanism stays for larger kernels.
    uint32_t copies = 1;
    if (const char *e = std::getenv("PVQT_VALUE_COPIES")) copies = (uint32_t)std::max(1, std::min(std::atoi(e), 64));
    const size_t one = values.size();
    values.resize(one * copies);
    for (uint32_t c = 1; c < copies; ++c) std::copy(values.begin(), values.begin() + (long)one, values.begin() + (long)(one * c));

    FusedParams &P = v->fused;
    std::memset(&P, 0, sizeof(P));
    cudaError_t e;
    for (int w = 0; w < n_warps; ++w) P.warp[w] = warps[(size_t)w];
    if ((e = upload(v, lane_meta, &P.lane_meta)) != cudaSuccess) return cuda_fail(e, "upload fused lane_meta");
    if ((e = upload(v, lane_rows, &P.lane_rows)) != cudaSuccess) return cuda_fail(e, "upload fused lane_rows");
    if ((e = upload(v, values, &P.values)) != cudaSuccess) return cuda_fail(e, "upload fused values");
    P.n_warps = n_warps;
    P.rows_per_lane = R;
    P.values_stride = (uint32_t)one;
    P.values_copies = copies;
    P.n_buckets = (int32_t)v->kernel.n_buckets;
    P.spec_stride = F.spec_stride;
    P.n_cols = n_cols;
    P.cols_touched = n_cols;
    for (int w = 0; w < n_warps; ++w)
        for (int l = 0; l < 32; ++l) {
            const int4 m = lane_meta[(size_t)w * 32 + l];
            P.cols_touched = std::max(P.cols_touched, std::max(m.x + warps[(size_t)w].width, m.z + warps[(size_t)w].nwidth));
        }
    P.ref_db = v->ref_db;
    if ((e = configure_fused(n_warps, P.cols_touched, P.n_buckets, R)) != cudaSuccess)
        return cuda_fail(e, "configure spmm_db_fused_kernel");
    // the persistent pipeline form (default where it fits: two plane sets + two log-spectrum buffers in one CTA)
    int min_slots = 1 << 30;
    for (int w = 0; w < n_warps; ++w) min_slots = std::min(min_slots, warps[(size_t)w].width + warps[(size_t)w].nwidth);
    v->pipe_capable = pipe_supported(n_warps, P.cols_touched, P.n_buckets, R, min_slots);
    if (v->pipe_capable && configure_pipe(n_warps, P.cols_touched, P.n_buckets, 0, 0) != cudaSuccess) {
        cudaGetLastError();
        v->pipe_capable = false;
    }
    v->pipe_sdft_configured = 0;
    P.plane_stride = pipe_plane_stride(P.cols_touched);
    v->pipe_ok = v->pipe_capable;
    if (const char *m = std::getenv("PVQT_SPMM_PIPE")) v->pipe_ok = v->pipe_capable && std::atoi(m) != 0;
    if (cudaDeviceGetAttribute(&v->sm_count, cudaDevAttrMultiProcessorCount, v->device) != cudaSuccess) v->sm_count = 148;
    return PVQT_OK;
}

// ---- K-sdft: which window groups take the sliding partial-DFT path for this (hop, frames per stream) ----
// Per frame the FFT path costs about 2 N log2 N pipe operations and touches shared memory three times per
// point; the sliding path costs 2 * nk * hop FMAs (one chunk per new frame) plus the combine, and reads
// each sample once.  It pays when the kernel consumes few bins of a long window and frames overlap a lot
// (group 0 at the defaults: 61 bins of an 8192-point FFT, hop 368).
bool sdft_worthwhile(size_t n_window, size_t nk, size_t hop, size_t frames_per_stream)
{
    if (hop < 16 || hop * 4 > n_window || nk == 0 || nk > 1024) return false;
    const double q = (double)(n_window / hop);
    const double sdft = 2.4 * (double)nk * (double)hop * ((double)frames_per_stream + q) + 8.0 * nk * (q + 1) * frames_per_stream;
    const double fft = 1.6 * (double)n_window * std::log2((double)n_window) * (double)frames_per_stream;
    return sdft < 0.7 * fft;
}

// With the inner GEMM on tcgen05 the partial sums cost little; what remains is the combine step (q + 1 complex MACs
// per bin and frame, read from L2) against the FFT the group would otherwise run.
bool sdft_tc_worthwhile(size_t n_window, size_t nk, size_t hop)
{
    if (hop < 16 || hop * 4 > n_window || nk == 0 || nk > 1024) return false;
    const double q = (double)(n_window / hop);
    const double sdft = 0.3 * (double)nk * (double)hop + 8.0 * nk * (q + 1);
    const double fft = 1.6 * (double)n_window * std::log2((double)n_window);
    return sdft < 0.7 * fft;
}

const pvqt::SdftPlan *sdft_plan_for(pvqt *v, size_t hop, int *status)
{
    *status = PVQT_OK;
    for (const auto &p : v->sdft_plans)
        if (p.hop == hop) return &p;
    if (v->sdft_plans.size() >= 8) return nullptr;  // tables are never evicted: bound them
    pvqt::SdftPlan plan;
    plan.hop = hop;
    const auto &groups = v->kernel.window_groups;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const size_t n = groups[gi].window_size();
        const size_t nk = v->n_cols[gi];
        if (!sdft_worthwhile(n, nk, hop, 1u << 20) && !sdft_tc_worthwhile(n, nk, hop)) continue;  // not even for a long stream
        SdftGroup g{};
        g.window_begin = (int32_t)groups[gi].window_begin;
        g.n_window = (int32_t)n;
        g.k_lo = (int32_t)v->col_lo[gi];
        g.nk = (int32_t)nk;
        g.spec_offset = (int32_t)v->spec_off[gi];
        g.hop = (int32_t)hop;
        g.q = (int32_t)(n / hop);
        g.rem = (int32_t)(n % hop);
        g.n_blocks = (int32_t)((hop + 15) / 16);
        g.hop_pad = g.n_blocks * 16;
        if (sdft_smem_bytes(g.hop_pad) > 200 * 1024 || sdft_combine_smem_bytes(g.q, g.nk) > 200 * 1024) continue;
        auto tw = [&](uint64_t k, uint64_t m) {  // exp(-2 pi i k m / N), argument reduced exactly
            const uint64_t r = (k * m) % n;
            const double a = -2.0 * kPi * (double)r / (double)n;
            return make_float2((float)std::cos(a), (float)std::sin(a));
        };
        std::vector<float2> ta((size_t)g.n_blocks * nk), tb((size_t)16 * nk), ph((size_t)(g.q + 1) * nk);
        for (size_t k = 0; k < nk; ++k) {
            const uint64_t ka = g.k_lo + k;
            for (int a = 0; a < g.n_blocks; ++a) ta[(size_t)a * nk + k] = tw(ka, 16ull * a);
            for (int b = 0; b < 16; ++b) tb[(size_t)b * nk + k] = tw(ka, b);
            for (int i = 0; i <= g.q; ++i) ph[(size_t)i * nk + k] = tw(ka, (uint64_t)i * hop);
        }
        cudaError_t e;
        if ((e = upload(v, ta, &g.tw_a)) != cudaSuccess || (e = upload(v, tb, &g.tw_b)) != cudaSuccess ||
            (e = upload(v, ph, &g.phase)) != cudaSuccess || (e = configure_sdft(g.hop_pad)) != cudaSuccess ||
            (e = configure_sdft_combine(g.q, g.nk)) != cudaSuccess) {
            *status = cuda_fail(e, "build K-sdft plan");
            return nullptr;
        }
        SdftTcPlan tc{};
        if (sdft_tc_supported(g)) {
            int group16 = 1;
            if (const char *s = std::getenv("PVQT_TC_GROUP")) group16 = std::max(1, std::min(std::atoi(s) / 16, 4));
            sdft_tc_make_plan(g, group16, &tc);
            std::vector<float2> tgb((size_t)16 * group16 * nk), tgg((size_t)tc.n_groups * nk);
            for (size_t k = 0; k < nk; ++k) {
                const uint64_t ka = g.k_lo + k;
                for (int b = 0; b < 16 * group16; ++b) tgb[(size_t)b * nk + k] = tw(ka, b);
                for (int i = 0; i < tc.n_groups; ++i) tgg[(size_t)i * nk + k] = tw(ka, 16ull * tc.start16[i]);
            }
            if ((e = upload(v, tgb, &tc.tw_b)) != cudaSuccess || (e = upload(v, tgg, &tc.tw_g)) != cudaSuccess ||
                (e = configure_sdft_tc(tc)) != cudaSuccess) {
                *status = cuda_fail(e, "build K-sdft tcgen05 plan");
                return nullptr;
            }
        }
        plan.tc.push_back(tc);
        plan.group_index.push_back((int)gi);
        plan.groups.push_back(g);
    }
    v->sdft_plans.push_back(std::move(plan));
    return &v->sdft_plans.back();
}

// ---- K-spmm-db cluster plan: CS contiguous row parts, 8 row pairs per warp, bands split in two halves ----
int build_cluster_plan(pvqt *v)
{
    const auto &groups = v->kernel.window_groups;
    const FftParams &F = v->fft;
    struct Unit {
        int col0 = 0, len = 0, ncol0 = 0, nlen = 0;
        int first_row = 0, n_rows = 0, group = 0, local_row = 0;
    };
    std::vector<Unit> units;
    int first_row = 0;
    long total_work = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const auto &g = groups[gi];
        const int spec = F.group[gi].spec_offset, lo = F.group[gi].col_lo;
        for (int r0 = 0; r0 < g.filter_bank.rows; r0 += kRowsPerLane) {
            Unit u;
            u.group = (int)gi; u.local_row = r0; u.first_row = first_row + r0;
            u.n_rows = std::min(kRowsPerLane, g.filter_bank.rows - r0);
            int a0 = 1 << 30, a1 = -1, n0 = 1 << 30, n1 = -1;
            for (int q = 0; q < u.n_rows; ++q) {
                const int r = r0 + q;
                const int s = g.filter_bank.indptr[r], e = g.filter_bank.indptr[r + 1];
                if (e > s) {
                    a0 = std::min(a0, spec + g.filter_bank.indices[s] - lo);
                    a1 = std::max(a1, spec + g.filter_bank.indices[e - 1] - lo + 1);
                }
                if (g.negative_filter_bank.nnz() > 0) {
                    const int ns = g.negative_filter_bank.indptr[r], ne = g.negative_filter_bank.indptr[r + 1];
                    if (ne > ns) {
                        n0 = std::min(n0, spec + g.negative_filter_bank.indices[ns] - lo);
                        n1 = std::max(n1, spec + g.negative_filter_bank.indices[ne - 1] - lo + 1);
                    }
                }
            }
            if (a1 >= 0) { u.col0 = a0; u.len = a1 - a0; }
            if (n1 >= 0) { u.ncol0 = n0; u.nlen = n1 - n0; }
            total_work += u.len + u.nlen + 4;
            units.push_back(u);
        }
        first_row += g.filter_bank.rows;
    }
    if (units.empty()) return PVQT_OK;

    struct WarpPlan { int width = 0, nwidth = 0, slot_base = 0, nslot_base = 0; std::vector<Unit> u; };
    struct PartPlan { std::vector<WarpPlan> warps; int col_lo = 0, cols_needed = 0, n_cols = 0, row_lo = 0, n_rows = 0, slots = 0; };
    const int max_warps = kClusterThreads / 32;
    std::vector<PartPlan> parts;
    int cs = 0;
    for (int cand : {1, 2, 4, 8}) {
        parts.assign((size_t)cand, PartPlan());
        // contiguous split of the row-ordered units, balanced on the band slots the parts will walk (a part's
        // cost is the sum over its 8-unit warps of half the warp's longest band, estimated by plan_slots)
        auto plan_slots = [&](size_t b, size_t e) {
            std::vector<int> lens, nlens;
            for (size_t i = b; i < e; ++i) { lens.push_back(units[i].len); nlens.push_back(units[i].nlen); }
            std::sort(lens.begin(), lens.end(), std::greater<int>());
            std::sort(nlens.begin(), nlens.end(), std::greater<int>());
            long sl = 0;
            for (size_t i = 0; i < lens.size(); i += 8) sl += (lens[i] + 4) / 2 + nlens[i];
            return sl;
        };
        std::vector<size_t> bound((size_t)cand + 1);
        for (int p = 0; p <= cand; ++p) bound[(size_t)p] = units.size() * (size_t)p / (size_t)cand;
        for (int iter = 0; iter < 200 && cand > 1; ++iter) {
            int worst = 0;
            long worst_slots = -1;
            for (int p = 0; p < cand; ++p) {
                const long sl = plan_slots(bound[(size_t)p], bound[(size_t)p + 1]);
                if (sl > worst_slots) { worst_slots = sl; worst = p; }
            }
            // give one unit of the heaviest part to its lighter neighbour
            const long left = worst > 0 ? plan_slots(bound[(size_t)worst - 1], bound[(size_t)worst]) : (1L << 60);
            const long right = worst + 1 < cand ? plan_slots(bound[(size_t)worst + 1], bound[(size_t)worst + 2]) : (1L << 60);
            if (std::min(left, right) + 2 >= worst_slots) break;
            if (bound[(size_t)worst + 1] - bound[(size_t)worst] <= 1) break;
            if (left <= right) ++bound[(size_t)worst]; else --bound[(size_t)worst + 1];
        }
        // a part's staged column range must fit the compile-time plane width: shrink offenders
        auto col_span = [&](size_t b, size_t e) {
            int lo = 1 << 30, hi = 0;
            for (size_t i = b; i < e; ++i) {
                const Unit &u = units[i];
                if (u.len > 0) { lo = std::min(lo, u.col0 - 7); hi = std::max(hi, u.col0 + ((u.len + 8) / 2) * 2 + 2); }
                if (u.nlen > 0) { lo = std::min(lo, u.ncol0); hi = std::max(hi, u.ncol0 + u.nlen); }
            }
            return hi > 0 ? hi - (std::max(lo, 0) & ~7) : 0;
        };
        for (int iter = 0; iter < 400 && cand > 1; ++iter) {
            int bad = -1;
            for (int p = 0; p < cand; ++p)
                if (col_span(bound[(size_t)p], bound[(size_t)p + 1]) > kClusterPlaneCols) { bad = p; break; }
            if (bad < 0) break;
            if (bound[(size_t)bad + 1] - bound[(size_t)bad] <= 1) break;
            if (bad == cand - 1) ++bound[(size_t)bad]; else --bound[(size_t)bad + 1];
        }
        std::vector<std::vector<Unit>> pu((size_t)cand);
        for (int p = 0; p < cand; ++p)
            pu[(size_t)p].assign(units.begin() + (long)bound[(size_t)p], units.begin() + (long)bound[(size_t)p + 1]);
        bool ok = true;
        int coef_slots_max = 0, rows_max = 0;
        for (int p = 0; p < cand && ok; ++p) {
            PartPlan &pp = parts[(size_t)p];
            std::vector<Unit> us = pu[(size_t)p];
            if (us.empty()) { ok = false; break; }
            pp.row_lo = us.front().first_row;
            for (const Unit &u : us) pp.n_rows += u.n_rows;
            std::stable_sort(us.begin(), us.end(), [](const Unit &a, const Unit &b) { return a.len > b.len; });
            const int nw = (int)((us.size() + 7) / 8);
            if (nw > max_warps) { ok = false; break; }
            int cmin = 1 << 30, cmax = 0, slots = 0;
            for (int w = 0; w < nw; ++w) {
                WarpPlan wp;
                const size_t u0 = (size_t)w * 8, u1 = std::min(us.size(), u0 + 8);
                // distinct start residues modulo 8 inside the warp (or the same column): each quarter-warp
                // LDS.128 of the band walk is then conflict-free.  A unit may start up to 7 columns early.
                int residue_col[8];
                for (int &c : residue_col) c = -1;
                for (size_t i = u0; i < u1; ++i) {
                    Unit u = us[i];
                    if (u.len > 0) {
                        int shift = 0;
                        for (; shift < 8; ++shift) {
                            const int c = u.col0 - shift;
                            if (c < 0) { shift = 8; break; }
                            if (residue_col[c & 7] == -1 || residue_col[c & 7] == c) break;
                        }
                        if (shift < 8) { u.col0 -= shift; u.len += shift; }
                        if (residue_col[u.col0 & 7] == -1) residue_col[u.col0 & 7] = u.col0;
                    }
                    wp.u.push_back(u);
                    wp.width = std::max(wp.width, (u.len + 1) / 2);
                    wp.nwidth = std::max(wp.nwidth, u.nlen);
                }
                for (const Unit &u : wp.u) {
                    if (u.len > 0) { cmin = std::min(cmin, u.col0); cmax = std::max(cmax, u.col0 + 2 * wp.width); }
                    if (u.nlen > 0) { cmin = std::min(cmin, u.ncol0); cmax = std::max(cmax, u.ncol0 + wp.nwidth); }
                }
                wp.slot_base = slots; slots += wp.width;
                wp.nslot_base = slots; slots += wp.nwidth;
                pp.warps.push_back(std::move(wp));
            }
            if (cmax <= 0) { cmin = 0; cmax = 8; }
            pp.col_lo = cmin & ~7;
            pp.cols_needed = cmax - pp.col_lo;
            pp.n_cols = std::min(pp.cols_needed, F.spec_stride - pp.col_lo);
            pp.slots = slots;
            if (pp.cols_needed > kClusterPlaneCols) ok = false;
            coef_slots_max = std::max(coef_slots_max, slots);
            rows_max = std::max(rows_max, pp.n_rows);
        }
        if (!ok) continue;
        const int coef_bytes = (coef_slots_max * 256 + 127) & ~127;
        if (cluster_smem_bytes(coef_bytes, rows_max, cand) > 227 * 1024) continue;
        cs = cand;
        v->cluster.coef_bytes = coef_bytes;
        v->cluster.max_rows = rows_max;
        v->cluster.mm_offset = (int32_t)(cluster_smem_bytes(coef_bytes, rows_max, cand) -
                                         (size_t)2 * cand * kClusterRoundFrames * sizeof(float2));
        break;
    }
    if (cs == 0) return PVQT_OK;  // does not fit: the single-CTA or unfused kernels take over

    std::vector<ClusterPart> dparts;
    std::vector<ClusterWarp> dwarps;
    std::vector<ClusterLane> dlanes;
    std::vector<float4> coef;
    for (const PartPlan &pp : parts) {
        ClusterPart dp{};
        dp.n_warps = (int)pp.warps.size();
        dp.col_lo = pp.col_lo; dp.n_cols = pp.n_cols; dp.cols_touched = pp.cols_needed;
        dp.row_lo = pp.row_lo; dp.n_rows = pp.n_rows;
        dp.coef_base = (int)(coef.size() / 16); dp.coef_slots = pp.slots; dp.desc_base = (int)dwarps.size();
        coef.resize(coef.size() + (size_t)pp.slots * 16, make_float4(0.f, 0.f, 0.f, 0.f));
        for (const WarpPlan &wp : pp.warps) {
            dwarps.push_back(ClusterWarp{wp.width, wp.nwidth, wp.slot_base, wp.nslot_base});
            const size_t lane0 = dlanes.size();
            dlanes.resize(lane0 + 16, ClusterLane{0, 0, -1, 0});
            for (size_t ui = 0; ui < wp.u.size(); ++ui) {
                const Unit &u = wp.u[ui];
                const auto &g = groups[(size_t)u.group];
                const int spec = F.group[u.group].spec_offset, lo = F.group[u.group].col_lo;
                for (int h = 0; h < 2; ++h) {
                    ClusterLane &L = dlanes[lane0 + (size_t)h * 8 + ui];
                    L.col = (u.len > 0 ? u.col0 - pp.col_lo : 0) + h * wp.width;
                    L.ncol = u.nlen > 0 ? u.ncol0 - pp.col_lo : 0;
                    L.row = u.first_row - pp.row_lo;
                    L.n_rows = u.n_rows;
                }
                for (int q = 0; q < u.n_rows; ++q) {
                    const int r = u.local_row + q;
                    for (int e = g.filter_bank.indptr[r]; e < g.filter_bank.indptr[r + 1]; ++e) {
                        const int pos = spec + g.filter_bank.indices[e] - lo - u.col0;
                        const int h = pos / wp.width, j = pos % wp.width;
                        float4 &slot = coef[((size_t)dp.coef_base + wp.slot_base + j) * 16 + (size_t)h * 8 + ui];
                        (q == 0 ? slot.x : slot.z) = g.filter_bank.data[e].real();
                        (q == 0 ? slot.y : slot.w) = g.filter_bank.data[e].imag();
                    }
                    if (u.nlen > 0)
                        for (int e = g.negative_filter_bank.indptr[r]; e < g.negative_filter_bank.indptr[r + 1]; ++e) {
                            const int j = spec + g.negative_filter_bank.indices[e] - lo - u.ncol0;
                            float4 &slot = coef[((size_t)dp.coef_base + wp.nslot_base + j) * 16 + ui];
                            // conj(Kneg X) = conj(Kneg) conj(X): keep conj(Kneg); the h = 1 lanes keep zeros
                            (q == 0 ? slot.x : slot.z) = g.negative_filter_bank.data[e].real();
                            (q == 0 ? slot.y : slot.w) = -g.negative_filter_bank.data[e].imag();
                        }
                }
            }
        }
        dparts.push_back(dp);
    }

    ClusterParams &P = v->cluster;
    cudaError_t e;
    if ((e = upload(v, dparts, &P.parts)) != cudaSuccess || (e = upload(v, dwarps, &P.warps)) != cudaSuccess ||
        (e = upload(v, dlanes, &P.lanes)) != cudaSuccess || (e = upload(v, coef, &P.coef)) != cudaSuccess)
        return cuda_fail(e, "upload cluster plan");
    P.cluster_size = cs;
    P.n_buckets = (int32_t)v->kernel.n_buckets;
    P.spec_stride = F.spec_stride;
    P.ref_db = v->ref_db;
    int max_active = 0;
    e = configure_cluster(P.coef_bytes, P.max_rows, cs, &max_active);
    if (e != cudaSuccess || max_active <= 0) { cudaGetLastError(); return PVQT_OK; }  // keep the other kernels
    v->cluster_max_active = max_active;
    v->cluster_capable = true;
    v->cluster_ok = false;  // measured slower than the one-CTA-per-tile form on B200: opt-in, mode 2
    if (const char *m = std::getenv("PVQT_SPMM_MODE")) {
        v->cluster_ok = std::atoi(m) == 2;
        v->fused_ok = v->fused_capable && std::atoi(m) >= 1;
        v->pipe_ok = v->pipe_ok && v->fused_ok && std::atoi(m) != 1;
    }
    return PVQT_OK;
}

void prof_begin(pvqt *v, int kind, cudaStream_t stream)
{
    if (!v->profiling) return;
    pvqt::Timed t{};
    t.kind = kind;
    if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return;
    cudaEventRecord(t.a, stream);
    v->timed.push_back(t);
}

void prof_end(pvqt *v, cudaStream_t stream)
{
    if (!v->profiling || v->timed.empty()) return;
    cudaEventRecord(v->timed.back().b, stream);
}

// Spectrum / power scratch of one launch lane for up to `frames` frames.
int reserve_scratch(pvqt *v, pvqt::Lane &L, size_t frames, bool need_power, cudaStream_t stream)
{
    frames = std::min<size_t>(frames, v->chunk_frames);
    // a tile in the record layout, or in the plane layout of the pipeline form of K-spmm-db (whichever is larger)
    const size_t tile_bytes = std::max((size_t)v->fft.spec_stride * kTileFrames * 2 * sizeof(float),
                                       v->pipe_capable ? (size_t)v->fused.plane_stride * kTileFrames * 2 * sizeof(float) : 0);
    const size_t want = ((frames + kTileFrames - 1) / kTileFrames) * tile_bytes;
    if (want > L.spec.bytes) {
        cudaError_t e = L.spec.reserve(want);
        if (e != cudaSuccess) return cuda_fail(e, "allocate spectrum scratch");
        // frames past the end of the last tile are never written: keep them finite
        if ((e = cudaMemsetAsync(L.spec.ptr, 0, L.spec.bytes, stream)) != cudaSuccess)
            return cuda_fail(e, "clear spectrum scratch");
    }
    if (need_power) {
        cudaError_t e = L.power.reserve(frames * v->kernel.n_buckets * sizeof(float));
        if (e != cudaSuccess) return cuda_fail(e, "allocate power scratch");
    }
    return PVQT_OK;
}

// Launch the kernels for the frames `(n_streams, frames_per_stream, hop)` describe:
//   [K-sdft partial] -> K-fft (remaining groups) -> [K-sdft combine] -> K-spmm-db   (or K-spmm + K-db)
// every arrow a programmatic dependent launch.  Work is cut into launches of at most chunk_frames frames:
// whole streams when a stream is shorter than that, else frame ranges of one stream.
int run_device(pvqt *v, const float *d_audio, size_t n_streams, size_t stream_stride, size_t hop,
               size_t frames_per_stream, float *d_out, float *d_power, float *d_spec_out, cudaStream_t stream,
               int scratch_lane = 0)
{
    size_t total = 0;
    if (frames_per_stream > 0x7fffffffull || n_streams > 0xffffffffull || !mul_ok(n_streams, frames_per_stream, &total))
        return fail(PVQT_INVALID_ARGUMENT, "too many frames per stream or streams");
    if (total == 0) return PVQT_OK;
    if (hop == 0 && frames_per_stream > 1) return fail(PVQT_INVALID_ARGUMENT, "hop must be positive");
    {   // every sample offset a kernel forms must fit 64 bits with room to spare
        size_t span = 0, all = 0;
        if (!mul_ok(frames_per_stream - 1, hop, &span) || span > (SIZE_MAX >> 4) ||
            (n_streams > 1 && (!mul_ok(n_streams, stream_stride, &all) || all > (SIZE_MAX >> 4))))
            return fail(PVQT_INVALID_ARGUMENT, "hop / stream_stride too large");
    }
    const size_t nb = v->kernel.n_buckets;
    const size_t chunk = v->chunk_frames;  // multiple of kTileFrames
    const size_t tile_elems = (size_t)v->fft.spec_stride * kTileFrames * 2;  // floats per tile
    const bool need_power = d_power == nullptr && !v->fused_ok && !v->cluster_ok;

    // window groups on the sliding partial-DFT path for this call
    const pvqt::SdftPlan *plan = nullptr;
    std::vector<int> sdft_groups;  // indices into plan->groups
    const size_t job_frames = std::max(frames_per_stream, v->sdft_job_frames);
    if (v->sdft_enabled && job_frames > 1) {
        int st = PVQT_OK;
        plan = sdft_plan_for(v, hop, &st);
        if (st != PVQT_OK) return st;
        if (plan)
            for (size_t i = 0; i < plan->groups.size(); ++i)
                if ((int)sdft_groups.size() < kMaxSdft &&
                    (sdft_worthwhile((size_t)plan->groups[i].n_window, (size_t)plan->groups[i].nk, hop, job_frames) ||
                     (v->sdft_tc >= 2 && job_frames >= 256 && plan->tc[i].n_groups > 0 &&
                      sdft_tc_worthwhile((size_t)plan->groups[i].n_window, (size_t)plan->groups[i].nk, hop))))
                    sdft_groups.push_back((int)i);
    }
    v->last_sdft_mask = 0;
    for (int i : sdft_groups) v->last_sdft_mask |= 1u << plan->group_index[(size_t)i];
    auto on_sdft = [&](int gi) {
        for (int i : sdft_groups)
            if (plan->group_index[(size_t)i] == gi) return true;
        return false;
    };

    // Launch ranges: whole streams when a stream is shorter than a chunk, else frame ranges of one stream.  A
    // single range is cut in n_lanes pieces so that the lanes have something to overlap.
    struct Range { size_t s0, ns, t0, nf; };
    std::vector<Range> ranges;
    const bool by_stream = frames_per_stream <= chunk;
    const int lanes = d_spec_out ? 1 : std::max(1, std::min(v->n_lanes, (int)pvqt::kLanes));
    if (by_stream && n_streams > 1) {
        size_t spl = std::max<size_t>(1, chunk / frames_per_stream);
        if (n_streams <= spl && lanes > 1 && total >= (size_t)lanes * 1024) spl = (n_streams + lanes - 1) / lanes;
        for (size_t s0 = 0; s0 < n_streams; s0 += spl) ranges.push_back({s0, std::min(spl, n_streams - s0), 0, frames_per_stream});
    } else {
        for (size_t s0 = 0; s0 < n_streams; ++s0) {
            size_t step = chunk;
            if (frames_per_stream <= chunk && lanes > 1 && frames_per_stream >= (size_t)lanes * 1024)
                step = ((frames_per_stream + lanes - 1) / lanes + 2 * kTileFrames - 1) / (2 * kTileFrames) * (2 * kTileFrames);
            for (size_t t0 = 0; t0 < frames_per_stream; t0 += step)
                ranges.push_back({s0, 1, t0, std::min(step, frames_per_stream - t0)});
        }
    }
    const bool use_lanes = lanes > 1 && ranges.size() > 1;
    if (use_lanes) {
        PVQT_CUDA(cudaEventRecord(v->lane_fork, stream));
        for (int l = 0; l < lanes; ++l) PVQT_CUDA(cudaStreamWaitEvent(v->lane[l].stream, v->lane_fork, 0));
    }
    cudaStream_t caller_stream = stream;
    for (size_t ri = 0; ri < ranges.size(); ++ri) {
        {
            const size_t s0 = ranges[ri].s0, ns = ranges[ri].ns, t0 = ranges[ri].t0, nf = ranges[ri].nf;
            pvqt::Lane &L = v->lane[use_lanes ? ri % (size_t)lanes : (size_t)scratch_lane];
            stream = use_lanes ? L.stream : caller_stream;
            if (L.used && L.last_stream != stream && !v->use_graphs && !v->capturing)
                PVQT_CUDA(cudaStreamWaitEvent(stream, L.scratch_free, 0));
            struct ScratchRelease {   // records scratch_free behind the chain's last kernel, on every exit path
                pvqt *v; pvqt::Lane &L; cudaStream_t s;
                ~ScratchRelease()
                {
                    if (v->use_graphs || v->capturing) return;
                    if (cudaEventRecord(L.scratch_free, s) == cudaSuccess) { L.last_stream = s; L.used = true; }
                }
            } scratch_release{v, L, stream};
            if (!d_spec_out) {
                int rc = reserve_scratch(v, L, ns * nf, need_power, stream);
                if (rc) return rc;
            }
            const size_t f0 = s0 * frames_per_stream + t0;               // flat index of the launch's first frame
            const uint32_t n = (uint32_t)(ns * nf);
            float *spec = d_spec_out ? d_spec_out + (f0 / kTileFrames) * tile_elems : static_cast<float *>(L.spec.ptr);

            // ---- K-sdft partial sums ----
            bool counted = true;   // every partial-sum launch of this range reports to the lane's completion counter
            std::vector<SdftParams> sd;
            std::vector<int> sd_plan;   // index into plan->groups / plan->tc
            for (int i : sdft_groups) {
                sd_plan.push_back(i);
                SdftParams sp{};
                sp.g = plan->groups[(size_t)i];
                sp.audio = d_audio;
                sp.stream_stride = stream_stride;
                sp.valid_samples = (frames_per_stream - 1) * hop + (size_t)v->params.n_fft;
                sp.first_stream = (uint32_t)s0;
                sp.n_streams = (uint32_t)ns;
                sp.first_frame = (uint32_t)t0;
                sp.frames = (uint32_t)nf;
                sp.rows_per_stream = (uint32_t)(nf + sp.g.q);
                sp.spec_stride = v->fft.spec_stride;
                sp.spec = spec;
                sd.push_back(sp);
            }
            if (!sd.empty()) {
                size_t need = 0;
                for (const auto &sp : sd) need += (size_t)sp.n_streams * sp.rows_per_stream * sp.g.nk * sizeof(float2);
                need += 64;   // the pipeline form's bulk copies round their ends to 16 bytes
                if (L.sdft_c.reserve(need) != cudaSuccess || L.sdft_r.reserve(need) != cudaSuccess)
                    return cuda_fail(cudaGetLastError(), "allocate K-sdft scratch");
                size_t off = 0;
                for (size_t si = 0; si < sd.size(); ++si) {
                    SdftParams &sp = sd[si];
                    const SdftTcPlan &tc = plan->tc[(size_t)sd_plan[si]];
                    sp.partial_c = reinterpret_cast<float2 *>(static_cast<char *>(L.sdft_c.ptr) + off);
                    sp.partial_r = reinterpret_cast<float2 *>(static_cast<char *>(L.sdft_r.ptr) + off);
                    off += (size_t)sp.n_streams * sp.rows_per_stream * sp.g.nk * sizeof(float2);
                    const bool on_tc = v->sdft_tc >= 1 && tc.n_groups > 0;
                    // completion counter for K-spmm-db's early combine; not under graph capture (the expected value
                    // would be baked into the graph) and not for the tcgen05 form
                    sp.done_counter = nullptr;
                    if (v->tile_flags && !v->use_graphs && !v->capturing && !on_tc && L.sdft_done != nullptr) {
                        sp.done_counter = L.sdft_done;
                        L.sdft_expected += sdft_partial_ctas(sp, v->sdft_tensor_cores);
                    } else {
                        counted = false;
                    }
                    prof_begin(v, 4, stream);
                    cudaError_t e = on_tc ? launch_sdft_partial_tc(sp, tc, stream)
                                          : launch_sdft_partial(sp, v->sdft_tensor_cores, stream);
                    if (e != cudaSuccess) return cuda_fail(e, "launch sdft_partial_kernel");
                    prof_end(v, stream);
                    v->launches.fetch_add(1);
                }
            }

            // ---- K-spmm-db form of this launch: the pipeline form needs the combine staging of this hop's K-sdft groups
            // to fit beside its planes (else the one-CTA-per-tile form), and K-fft then writes the plane layout
            int pipe_sdft = 0;
            bool use_pipe = v->pipe_ok && v->fused_ok && !v->cluster_ok && !d_spec_out && !(v->tile_flags && !v->use_graphs && !v->capturing);
            if (use_pipe) {
                for (const auto &sp : sd) pipe_sdft = std::max<int>(pipe_sdft, (int)pipe_sdft_floats2(sp.g.q, sp.g.nk));
                const int total = pipe_sdft * (int)sd.size();   // the kernel's smem limit is raised to the largest total seen
                if (pipe_smem_bytes(v->fused.cols_touched, v->fused.n_buckets, v->fused.n_warps, pipe_sdft, (int)sd.size()) > 227 * 1024) {
                    use_pipe = false;
                } else if (total > v->pipe_sdft_configured) {
                    if (configure_pipe(v->fused.n_warps, v->fused.cols_touched, v->fused.n_buckets, total, 1) != cudaSuccess) {
                        cudaGetLastError();
                        use_pipe = false;
                    } else {
                        v->pipe_sdft_configured = total;
                    }
                }
            }

            // ---- K-fft for the other groups ----
            FftParams fp = v->fft;
            fp.plane_stride = use_pipe ? v->fused.plane_stride : 0;
            fp.frames.audio = d_audio;
            fp.frames.stream_stride = stream_stride;
            fp.frames.hop = hop;
            fp.frames.frames_per_stream = (uint32_t)frames_per_stream;
            fp.frames.n_frames = n;
            fp.frames.first_frame = f0;
            fp.frames.first_stream = (uint32_t)s0;
            fp.frames.first_t = (uint32_t)t0;
            fp.spec = spec;
            fp.wait_prior = sd.empty() ? 0 : 1;
            int ctas = 0, kept = 0;
            // One wave: the launch's work items (frames_per_cta frames each, about the same work whatever the group) are
            // spread over at most fft_wave CTAs, each group a share proportional to its items, and every CTA loops over
            // its items -- no last, mostly empty wave (5.19 waves of one-item CTAs at 3507 frames before).
            size_t items_total = 0;
            for (int g = 0; g < v->fft.n_groups; ++g)
                if (!on_sdft(g)) items_total += (n + v->fft.group[g].frames_per_cta - 1) / v->fft.group[g].frames_per_cta;
            const size_t wave = (size_t)v->fft_wave_ctas;
            for (int g = 0; g < v->fft.n_groups; ++g) {
                if (on_sdft(g)) continue;
                fp.group[kept] = v->fft.group[g];
                fp.group[kept].cta_begin = ctas;
                const size_t items = (n + fp.group[kept].frames_per_cta - 1) / fp.group[kept].frames_per_cta;
                size_t share = (wave == 0 || items_total <= wave) ? items : std::max<size_t>(1, (items * wave + items_total / 2) / items_total);
                share = std::min(share, items);
                fp.group[kept].n_ctas = (int32_t)share;
                ctas += (int)share;
                ++kept;
            }
            fp.n_groups = kept;
            fp.n_sdft = 0;
            // per-tile completion counts for K-spmm-db (one-CTA form only; not under graph capture; K-sdft must report too)
            const bool flags = !use_pipe && v->tile_flags && !v->use_graphs && !v->capturing && v->fused_ok && !v->cluster_ok && !d_spec_out &&
                               kept > 0 && counted;
            fp.tile_ready = nullptr;
            if (flags) {
                const size_t want = ((size_t)n / kTileFrames + 2) * sizeof(unsigned);
                if (want > L.tile_ready.bytes) {
                    if (L.tile_ready.reserve(std::max<size_t>(want, 4096)) != cudaSuccess)
                        return cuda_fail(cudaGetLastError(), "allocate tile counters");
                    PVQT_CUDA(cudaMemsetAsync(L.tile_ready.ptr, 0, L.tile_ready.bytes, stream));  // K-spmm-db leaves them at zero
                }
                fp.tile_ready = static_cast<unsigned *>(L.tile_ready.ptr);
            }
            // the combine step runs inside K-spmm-db (one-CTA form); K-fft's last CTAs take it over for the other
            // SpMM forms and for the spectra test hook
            const bool combine_in_spmm = v->fused_ok && !v->cluster_ok && !d_spec_out;
            if (kept > 0) {
                if (!combine_in_spmm)
                    for (const auto &sp : sd) fp.sdft[fp.n_sdft++] = sp;
                fp.combine_group = kept - 1;  // the last group: its CTAs start when the partial sums are long complete
                prof_begin(v, 0, stream);
                cudaError_t e = launch_fft(fp, ctas, v->fft_block_threads, stream);
                if (e != cudaSuccess) return cuda_fail(e, "launch fft_groups_kernel");
                prof_end(v, stream);
                v->launches.fetch_add(1);
            }

            // ---- K-sdft combine: run by K-fft's last CTAs; its own launch only when no FFT group is left ----
            if (kept == 0 && !combine_in_spmm)
                for (const auto &sp : sd) {
                    prof_begin(v, 5, stream);
                    cudaError_t e = launch_sdft_combine(sp, stream);
                    if (e != cudaSuccess) return cuda_fail(e, "launch sdft_combine_kernel");
                    prof_end(v, stream);
                    v->launches.fetch_add(1);
                }
            if (d_spec_out) continue;

            cudaError_t e;
            if (v->cluster_ok) {
                ClusterParams cp = v->cluster;
                cp.n_frames = n;
                cp.n_tiles = (n + kTileFrames - 1) / kTileFrames;
                cp.spec = spec;
                cp.out_db = d_out + f0 * nb;
                cp.power = d_power ? d_power + f0 * nb : nullptr;
                const int rounds = (int)((cp.n_tiles + 1) / 2);
                prof_begin(v, 6, stream);
                e = launch_spmm_db_cluster(cp, std::min(rounds, v->cluster_max_active), stream);
                if (e != cudaSuccess) return cuda_fail(e, "launch spmm_db_cluster_kernel");
                prof_end(v, stream);
                v->launches.fetch_add(1);
                continue;
            }
            if (v->fused_ok) {
                FusedParams up = v->fused;
                up.n_frames = n;
                up.n_tiles = (n + kTileFrames - 1) / kTileFrames;
                up.spec = spec;
                up.out_db = d_out + f0 * nb;
                up.power = d_power ? d_power + f0 * nb : nullptr;
                up.n_sdft = 0;
                for (const auto &sp : sd) up.sdft[up.n_sdft++] = sp;
                up.sdft_done = (!sd.empty() && counted) ? L.sdft_done : nullptr;
                up.sdft_expected = L.sdft_expected;
                up.tile_ready = fp.tile_ready;
                up.n_ready_groups = fp.n_groups;
                for (int g = 0; g < fp.n_groups; ++g) up.ready_fpc[g] = fp.group[g].frames_per_cta;
                prof_begin(v, 3, stream);
                e = use_pipe ? launch_spmm_db_pipe(up, v->sm_count, pipe_sdft, stream) : launch_spmm_db_fused(up, stream);
                if (e != cudaSuccess) return cuda_fail(e, "launch spmm_db_fused_kernel");
                prof_end(v, stream);
                v->launches.fetch_add(1);
                continue;
            }

            SpmmParams sp = v->spmm;
            sp.n_frames = n;
            sp.n_tiles = (n + kTileFrames - 1) / kTileFrames;
            sp.spec = spec;
            sp.power = d_power ? d_power + f0 * nb : static_cast<float *>(L.power.ptr);
            prof_begin(v, 1, stream);
            e = launch_spmm(sp, stream);
            if (e != cudaSuccess) return cuda_fail(e, "launch spmm_kernel");
            prof_end(v, stream);
            v->launches.fetch_add(1);

            DbParams dp{};
            dp.power = sp.power;
            dp.out_db = d_out + f0 * nb;
            dp.n_frames = n;
            dp.n_buckets = (int32_t)nb;
            dp.ref_db = v->ref_db;
            prof_begin(v, 2, stream);
            e = launch_power_to_db(dp, stream);
            if (e != cudaSuccess) return cuda_fail(e, "launch power_to_db_kernel");
            prof_end(v, stream);
            v->launches.fetch_add(1);
        }
    }
    if (use_lanes)
        for (int l = 0; l < lanes; ++l) {
            PVQT_CUDA(cudaEventRecord(v->lane[l].done, v->lane[l].stream));
            PVQT_CUDA(cudaStreamWaitEvent(caller_stream, v->lane[l].done, 0));
        }
    return PVQT_OK;
}

// Host buffers in/out, pipelined: the work is cut into segments (stream ranges, or frame ranges of one
// long recording); segment i+1's host->device copy and segment i-1's device->host copy run on their
// own streams while segment i computes.  Every segment owns a distinct region of the device audio /
// output buffers, so the three streams only need the event edges H2D(i) -> compute(i) -> D2H(i).
// Work larger than the staging budget is processed in consecutive super-batches.
struct HostJob {
    const float *audio; float *out;
    size_t n_streams, stream_stride, n_samples, hop, frames_per_stream;
    size_t n_fft, nb, span;
};

struct EventPool {
    pvqt *v;
    size_t used = 0;
    cudaError_t next(cudaEvent_t *e)
    {
        if (used == v->events.size()) {
            cudaEvent_t ne;
            cudaError_t rc = cudaEventCreateWithFlags(&ne, cudaEventDisableTiming);
            if (rc != cudaSuccess) return rc;
            v->events.push_back(ne);
        }
        *e = v->events[used++];
        return cudaSuccess;
    }
};

size_t segment_frames(const pvqt *v, size_t frames_in_batch)
{
    // a few segments per batch (copies overlap compute), at least 256 frames each; at most segment_cap frames, so
    // that the first H2D copy and the last D2H copy -- which nothing overlaps -- stay short in a long batch
    // (2048 streams x 511 frames, pinned buffers: 17.7 / 18.5 / 18.8 / 18.5 M frames/s at 4096 / 8192 / 16384 / 32768,
    // 16.5 M at 131072)
    const size_t n = v->segments_per_batch;
    return std::min<size_t>(std::max<size_t>((frames_in_batch + n - 1) / n, 256),
                            std::min<size_t>(v->segment_cap, v->chunk_frames));
}

// the copy streams join the compute stream's timeline (needed for graph capture, harmless otherwise)
int fork_streams(pvqt *v, EventPool &ev)
{
    cudaEvent_t fork;
    PVQT_CUDA(ev.next(&fork));
    PVQT_CUDA(cudaEventRecord(fork, v->stream));
    PVQT_CUDA(cudaStreamWaitEvent(v->s_in, fork, 0));
    PVQT_CUDA(cudaStreamWaitEvent(v->s_out, fork, 0));
    if (v->host_lanes > 1)
        for (int l = 0; l < v->host_lanes; ++l) PVQT_CUDA(cudaStreamWaitEvent(v->lane[l].stream, fork, 0));
    return PVQT_OK;
}

int join_streams(pvqt *v, EventPool &ev)
{
    for (cudaStream_t s : {v->s_in, v->s_out}) {
        cudaEvent_t j;
        PVQT_CUDA(ev.next(&j));
        PVQT_CUDA(cudaEventRecord(j, s));
        PVQT_CUDA(cudaStreamWaitEvent(v->stream, j, 0));
    }
    if (v->host_lanes > 1)
        for (int l = 0; l < v->host_lanes; ++l) {
            cudaEvent_t j;
            PVQT_CUDA(ev.next(&j));
            PVQT_CUDA(cudaEventRecord(j, v->lane[l].stream));
            PVQT_CUDA(cudaStreamWaitEvent(v->stream, j, 0));
        }
    return PVQT_OK;
}

int compute_and_copy_out(pvqt *v, EventPool &ev, const HostJob &J, cudaEvent_t copied, const float *d_audio, size_t ns,
                         size_t dstride, size_t nf, float *d_out, float *h_out, size_t segment)
{
    // consecutive segments compute on alternating lanes (stream + scratch): the kernels of a segment are
    // latency-bound at these sizes, so two segments in flight keep the device-to-host copies fed
    const int lane = v->host_lanes > 1 ? (int)(segment % (size_t)v->host_lanes) : 0;
    cudaStream_t cs = v->host_lanes > 1 ? v->lane[lane].stream : v->stream;
    PVQT_CUDA(cudaStreamWaitEvent(cs, copied, 0));
    int rc = run_device(v, d_audio, ns, dstride, J.hop, nf, d_out, nullptr, nullptr, cs, lane);
    if (rc) return rc;
    cudaEvent_t done;
    PVQT_CUDA(ev.next(&done));
    PVQT_CUDA(cudaEventRecord(done, cs));
    PVQT_CUDA(cudaStreamWaitEvent(v->s_out, done, 0));
    if (J.out != nullptr)   // nullptr: the dB spectra stay in HBM for the analysis epilogue (pvqt_calc_*_analysis)
        PVQT_CUDA(cudaMemcpyAsync(h_out, d_out, ns * nf * J.nb * sizeof(float), cudaMemcpyDeviceToHost, v->s_out));
    return PVQT_OK;
}

// stream-sharded super-batch: streams [b0, b0 + nbatch), whole streams per segment
int enqueue_stream_batch(pvqt *v, EventPool &ev, const HostJob &J, size_t b0, size_t nbatch, size_t dstride)
{
    float *d_audio = static_cast<float *>(v->d_audio.ptr), *d_out = static_cast<float *>(v->d_out.ptr);
    const size_t per_seg = std::max<size_t>(1, segment_frames(v, nbatch * J.frames_per_stream) / J.frames_per_stream);
    for (size_t s0 = 0; s0 < nbatch; s0 += per_seg) {
        const size_t ns = std::min(per_seg, nbatch - s0);
        // samples before `skip` are never read by any kernel (the union of the windows starts later): not uploaded
        const size_t skip = std::min(v->upload_skip, J.span);
        const size_t max_pitch = (size_t)0x7fffffff;   // cudaDevAttrMaxPitch
        if (ns > 1 && J.stream_stride * sizeof(float) <= max_pitch && dstride * sizeof(float) <= max_pitch) {
            PVQT_CUDA(cudaMemcpy2DAsync(d_audio + s0 * dstride + skip, dstride * sizeof(float),
                                        J.audio + (b0 + s0) * J.stream_stride + skip, J.stream_stride * sizeof(float),
                                        (J.span - skip) * sizeof(float), ns, cudaMemcpyHostToDevice, v->s_in));
        } else {
            for (size_t s = 0; s < ns; ++s)
                PVQT_CUDA(cudaMemcpyAsync(d_audio + (s0 + s) * dstride + skip, J.audio + (b0 + s0 + s) * J.stream_stride + skip,
                                          (J.span - skip) * sizeof(float), cudaMemcpyHostToDevice, v->s_in));
        }
        cudaEvent_t copied;
        PVQT_CUDA(ev.next(&copied));
        PVQT_CUDA(cudaEventRecord(copied, v->s_in));
        int rc = compute_and_copy_out(v, ev, J, copied, d_audio + s0 * dstride, ns, dstride, J.frames_per_stream,
                                      d_out + s0 * J.frames_per_stream * J.nb,
                                      J.out + (b0 + s0) * J.frames_per_stream * J.nb, s0 / per_seg);
        if (rc) return rc;
    }
    return PVQT_OK;
}

// frame-range super-batch of stream s: frames [b0, b0 + nbatch); every sample is copied once
int enqueue_frame_batch(pvqt *v, EventPool &ev, const HostJob &J, size_t s, size_t b0, size_t nbatch)
{
    float *d_audio = static_cast<float *>(v->d_audio.ptr), *d_out = static_cast<float *>(v->d_out.ptr);
    const float *h_audio = J.audio + s * J.stream_stride + b0 * J.hop;
    const size_t per_seg = segment_frames(v, nbatch);
    size_t copied_samples = std::min(v->upload_skip, (nbatch - 1) * J.hop + J.n_fft);   // never read: not uploaded
    for (size_t f0 = 0; f0 < nbatch; f0 += per_seg) {
        const size_t nf = std::min(per_seg, nbatch - f0);
        const size_t need = (f0 + nf - 1) * J.hop + J.n_fft;  // batch-relative end of this segment's samples
        if (need > copied_samples) {
            PVQT_CUDA(cudaMemcpyAsync(d_audio + copied_samples, h_audio + copied_samples,
                                      (need - copied_samples) * sizeof(float), cudaMemcpyHostToDevice, v->s_in));
            copied_samples = need;
        }
        cudaEvent_t copied;
        PVQT_CUDA(ev.next(&copied));
        PVQT_CUDA(cudaEventRecord(copied, v->s_in));
        int rc = compute_and_copy_out(v, ev, J, copied, d_audio + f0 * J.hop, 1, 0, nf, d_out + f0 * J.nb,
                                      J.out + (s * J.frames_per_stream + b0 + f0) * J.nb, f0 / per_seg);
        if (rc) return rc;
    }
    return PVQT_OK;
}

// keep_on_device: `out` may be NULL (no D2H); the spectra of the call are left in v->d_out ([n_streams][frames][nb]),
// complete once v->stream has drained -- the call then returns WITHOUT synchronising -- and the job must fit one
// staging batch (PVQT_UNSUPPORTED otherwise; the caller splits by stream).
int run_host(pvqt *v, const float *audio, size_t n_streams, size_t stream_stride, size_t n_samples, size_t hop,
             size_t frames_per_stream, float *out, bool keep_on_device = false)
{
    HostJob J{audio, out, n_streams, stream_stride, n_samples, hop, frames_per_stream,
              (size_t)v->params.n_fft, v->kernel.n_buckets, 0};
    if (n_streams == 0 || frames_per_stream == 0) return PVQT_OK;
    if (!audio || (!out && !keep_on_device)) return fail(PVQT_INVALID_ARGUMENT, "null buffer");
    if (hop == 0 && frames_per_stream > 1) return fail(PVQT_INVALID_ARGUMENT, "hop must be positive");
    if (!frames_fit(frames_per_stream, hop, J.n_fft, n_samples))
        return fail(PVQT_BAD_LENGTH, "each stream must hold (frames_per_stream - 1) * hop + n_fft samples");
    {
        size_t t = 0;
        if (frames_per_stream > 0x7fffffffull || n_streams > 0xffffffffull || !mul_ok(n_streams, frames_per_stream, &t) ||
            !mul_ok(t, J.nb * sizeof(float), &t))
            return fail(PVQT_INVALID_ARGUMENT, "too many frames per stream or streams");
    }
    if (n_streams > 1 && stream_stride < n_samples)
        return fail(PVQT_INVALID_ARGUMENT, "stream_stride must be >= n_samples");
    PVQT_CUDA(cudaSetDevice(v->device));
    J.span = (frames_per_stream - 1) * hop + J.n_fft;  // samples one stream needs
    struct JobFrames {   // every segment of this call decides its K-sdft groups from the whole call (or the enclosing job)
        pvqt *v; size_t saved;
        JobFrames(pvqt *v_, size_t f) : v(v_), saved(v_->sdft_job_frames) { v->sdft_job_frames = std::max(saved, f); }
        ~JobFrames() { v->sdft_job_frames = saved; }
    } job_frames(v, frames_per_stream);

    const size_t budget = v->staging_budget_samples;
    const bool by_stream = J.span <= budget && (n_streams > 1 || frames_per_stream <= 256);
    const size_t dstride = (J.span + 3) & ~(size_t)3;
    const size_t streams_per_batch = std::max<size_t>(1, budget / dstride);
    const size_t frames_per_batch =
        budget > J.n_fft ? std::max<size_t>(1, (budget - J.n_fft) / std::max<size_t>(hop, 1) + 1) : 1;
    const bool single_batch = by_stream ? n_streams <= streams_per_batch
                                        : (n_streams == 1 && frames_per_stream <= frames_per_batch);
    EventPool ev{v};
    if (keep_on_device && !single_batch) return fail(PVQT_UNSUPPORTED, "job exceeds one staging batch");

    if (single_batch) {
        const size_t audio_samples = by_stream ? n_streams * dstride : J.span;
        const size_t frames = n_streams * frames_per_stream;
        PVQT_CUDA(v->d_audio.reserve(audio_samples * sizeof(float)));
        PVQT_CUDA(v->d_out.reserve(frames * J.nb * sizeof(float)));
        int rc = PVQT_OK;
        for (auto &L : v->lane)
            if ((rc = reserve_scratch(v, L, frames, !v->fused_ok && !v->cluster_ok, v->stream)) != PVQT_OK) return rc;
        auto enqueue = [&]() -> int {
            int r = fork_streams(v, ev);
            if (r) return r;
            r = by_stream ? enqueue_stream_batch(v, ev, J, 0, n_streams, dstride)
                          : enqueue_frame_batch(v, ev, J, 0, 0, frames_per_stream);
            if (r) return r;
            return join_streams(v, ev);
        };
        pvqt::HostCallKey key;
        key.audio = audio; key.out = out; key.n_streams = n_streams; key.stream_stride = stream_stride;
        key.n_samples = n_samples; key.hop = hop; key.frames_per_stream = frames_per_stream;
        key.generations = v->d_audio.generation * 1000003u + v->d_out.generation * 10007u;
        for (const auto &L : v->lane)
            key.generations += L.spec.generation * 101u + L.power.generation + L.sdft_c.generation * 7u + L.sdft_r.generation * 13u;
        // A call shape is captured only after it has run eagerly with the same buffers: the first run builds the K-sdft
        // plan of this hop and reserves every scratch buffer (cudaMalloc / cudaMemcpy are illegal under capture).
        const bool replay = v->graph_exec && key == v->graph_key;
        const bool warm = v->graph_warm && key == v->graph_warm_key;
        if (v->use_graphs && !v->profiling && !keep_on_device && !replay && !warm) {
            v->graph_warm_key = key;
            v->graph_warm = true;
        }
        if (v->use_graphs && !v->profiling && !keep_on_device && (replay || warm)) {
            if (!replay) {
                if (v->graph_exec) { cudaGraphExecDestroy(v->graph_exec); v->graph_exec = nullptr; }
                cudaGraph_t graph = nullptr;
                const uint64_t launches_before = v->launches.load();
                PVQT_CUDA(cudaStreamBeginCapture(v->stream, cudaStreamCaptureModeThreadLocal));
                rc = enqueue();
                v->graph_launches = v->launches.load() - launches_before;
                v->launches.store(launches_before);  // counted per replay below
                cudaError_t e = cudaStreamEndCapture(v->stream, &graph);
                if (rc == PVQT_OK && e == cudaSuccess && graph) {
                    e = cudaGraphInstantiate(&v->graph_exec, graph, 0);
                    if (e != cudaSuccess) v->graph_exec = nullptr;
                }
                if (graph) cudaGraphDestroy(graph);
                if (rc != PVQT_OK) { cudaGetLastError(); return rc; }
                if (!v->graph_exec) { cudaGetLastError(); v->use_graphs = false; }  // capture unsupported: run eagerly
                v->graph_key = key;
            }
            if (v->graph_exec) {
                PVQT_CUDA(cudaGraphLaunch(v->graph_exec, v->stream));
                v->launches.fetch_add(v->graph_launches);
                PVQT_CUDA(cudaStreamSynchronize(v->stream));
                return PVQT_OK;
            }
        }
        ev.used = 0;
        rc = enqueue();
        if (rc) return rc;
        if (!keep_on_device) PVQT_CUDA(cudaStreamSynchronize(v->stream));
        return PVQT_OK;
    }

    // several super-batches: eager, one drain per batch
    if (by_stream) {
        for (size_t b0 = 0; b0 < n_streams; b0 += streams_per_batch) {
            const size_t nbatch = std::min(streams_per_batch, n_streams - b0);
            PVQT_CUDA(v->d_audio.reserve(nbatch * dstride * sizeof(float)));
            PVQT_CUDA(v->d_out.reserve(nbatch * frames_per_stream * J.nb * sizeof(float)));
            ev.used = 0;
            int rc = fork_streams(v, ev);
            if (!rc) rc = enqueue_stream_batch(v, ev, J, b0, nbatch, dstride);
            if (!rc) rc = join_streams(v, ev);
            if (rc) return rc;
            PVQT_CUDA(cudaStreamSynchronize(v->stream));
        }
    } else {
        for (size_t s = 0; s < n_streams; ++s)
            for (size_t b0 = 0; b0 < frames_per_stream; b0 += frames_per_batch) {
                const size_t nbatch = std::min(frames_per_batch, frames_per_stream - b0);
                PVQT_CUDA(v->d_audio.reserve(((nbatch - 1) * hop + J.n_fft) * sizeof(float)));
                PVQT_CUDA(v->d_out.reserve(nbatch * J.nb * sizeof(float)));
                ev.used = 0;
                int rc = fork_streams(v, ev);
                if (!rc) rc = enqueue_frame_batch(v, ev, J, s, b0, nbatch);
                if (!rc) rc = join_streams(v, ev);
                if (rc) return rc;
                PVQT_CUDA(cudaStreamSynchronize(v->stream));
            }
    }
    return PVQT_OK;
}

// ---- per-frame entry: one captured graph -------------------------------------------------------
void release_instant(pvqt *v)
{
    pvqt::Instant &I = v->instant;
    if (I.exec) cudaGraphExecDestroy(I.exec);
    if (I.h_in) cudaFreeHost(I.h_in);
    if (I.h_out) cudaFreeHost(I.h_out);
    if (I.d_in) cudaFree(I.d_in);
    if (I.d_out) cudaFree(I.d_out);
    I = pvqt::Instant{};
}

void release_frame_pipe(pvqt *v)
{
    pvqt::FramePipe &F = v->frame_pipe;
    if (F.exec) cudaGraphExecDestroy(F.exec);
    if (F.h_in) cudaFreeHost(F.h_in);
    if (F.h_res) cudaFreeHost(F.h_res);
    if (F.d_res) cudaFree(F.d_res);
    if (F.d_in) cudaFree(F.d_in);
    if (F.d_db) cudaFree(F.d_db);
    F = pvqt::FramePipe{};
}

// One frame through VQT + AnalysisState (see pvqt::FramePipe).  The first call of a shape runs eagerly (it reserves the
// scratch and the result mirrors: nothing may be allocated under capture), the second is captured, the later ones replay.
int run_frame_pipeline(pvqt *v, pvqt_analysis *a, const float *x, uint64_t frame_time_ns, const pvqt_analysis_outputs *out,
                       uint64_t *d2h_bytes)
{
    pvqt::FramePipe &F = v->frame_pipe;
    const size_t n_fft = (size_t)v->params.n_fft, nb = v->kernel.n_buckets;
    const size_t skip = std::min(v->upload_skip, n_fft), n_in = n_fft - skip;
    PVQT_CUDA(cudaSetDevice(v->device));
    pvqt_analysis_outputs host = *out;
    uint32_t wanted = 0;
    size_t res_bytes = 0, off[pvqt_detail::kAnalysisOutputs] = {};
    for (int i = 0; i < pvqt_detail::kAnalysisOutputs; ++i) {
        if (!pvqt_detail::analysis_output_member(host, i)) continue;
        wanted |= 1u << i;
        off[i] = res_bytes;
        res_bytes += (pvqt_detail::analysis_output_bytes_per_frame(a, i, out->max_peaks) + 15) & ~(size_t)15;
    }
    if (!F.h_in) {
        PVQT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&F.h_in), std::max<size_t>(n_in, 1) * sizeof(float), cudaHostAllocDefault));
        PVQT_CUDA(cudaMalloc(reinterpret_cast<void **>(&F.d_in), n_fft * sizeof(float)));
        PVQT_CUDA(cudaMalloc(reinterpret_cast<void **>(&F.d_db), nb * sizeof(float)));
        PVQT_CUDA(cudaMemset(F.d_in, 0, n_fft * sizeof(float)));
    }
    if (res_bytes > F.h_res_bytes) {
        if (F.exec) { cudaGraphExecDestroy(F.exec); F.exec = nullptr; F.calls = 0; }
        if (F.h_res) cudaFreeHost(F.h_res);
        if (F.d_res) cudaFree(F.d_res);
        F.h_res = F.d_res = nullptr;
        F.h_res_bytes = 0;
        PVQT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&F.h_res), res_bytes, cudaHostAllocDefault));
        PVQT_CUDA(cudaMalloc(reinterpret_cast<void **>(&F.d_res), res_bytes));
        F.h_res_bytes = res_bytes;
    }
    std::memcpy(F.h_in, x + skip, n_in * sizeof(float));
    auto enqueue = [&]() -> int {
        PVQT_CUDA(cudaMemcpyAsync(F.d_in + skip, F.h_in, n_in * sizeof(float), cudaMemcpyHostToDevice, v->stream));
        int rc = run_device(v, F.d_in, 1, 0, n_fft, 1, F.d_db, nullptr, nullptr, v->stream, 0);
        if (rc) return rc;
        // K-analysis writes the frame's results into one device buffer laid out like the pinned one: one clear (slots past
        // a frame's peak count stay 0, as in the batch calls), one copy back
        pvqt_analysis_outputs dev{};
        dev.max_peaks = out->max_peaks;
        for (int i = 0; i < pvqt_detail::kAnalysisOutputs; ++i)
            if (wanted >> i & 1u) pvqt_detail::analysis_output_member(dev, i) = F.d_res + off[i];
        if (res_bytes) PVQT_CUDA(cudaMemsetAsync(F.d_res, 0, res_bytes, v->stream));
        rc = pvqt_detail::analysis_run_device(a, F.d_db, 0, 1, 1, frame_time_ns, &dev, 0, v->stream);
        if (rc) return rc;
        v->launches.fetch_add(1);
        if (res_bytes) PVQT_CUDA(cudaMemcpyAsync(F.h_res, F.d_res, res_bytes, cudaMemcpyDeviceToHost, v->stream));
        return PVQT_OK;
    };
    const int config = (v->fused_ok ? 1 : 0) | (v->cluster_ok ? 2 : 0) | (v->pipe_ok ? 4 : 0);
    const bool same = F.analysis == a && F.analysis_generation == pvqt_detail::analysis_generation(a) &&
                      F.frame_time_ns == frame_time_ns && F.wanted == wanted && F.max_peaks == out->max_peaks &&
                      F.scratch_generation == v->lane[0].spec.generation && F.config == config;
    if (!same) {
        if (F.exec) { cudaGraphExecDestroy(F.exec); F.exec = nullptr; }
        F.calls = 0;
    }
    if (!F.exec && F.calls >= 1 && !v->profiling) {
        cudaGraph_t graph = nullptr;
        const uint64_t before = v->launches.load();
        PVQT_CUDA(cudaStreamBeginCapture(v->stream, cudaStreamCaptureModeThreadLocal));
        v->capturing = true;
        int rc = enqueue();
        v->capturing = false;
        F.graph_launches = v->launches.load() - before;
        v->launches.store(before);
        cudaError_t e = cudaStreamEndCapture(v->stream, &graph);
        if (rc == PVQT_OK && e == cudaSuccess && graph) {
            e = cudaGraphInstantiate(&F.exec, graph, 0);
            if (e != cudaSuccess) F.exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
        if (rc != PVQT_OK) { cudaGetLastError(); return rc; }
        if (!F.exec) cudaGetLastError();   // capture unsupported here: stay eager
    }
    if (F.exec && !v->profiling) {
        PVQT_CUDA(cudaGraphLaunch(F.exec, v->stream));
        v->launches.fetch_add(F.graph_launches);
    } else {
        int rc = enqueue();
        if (rc) return rc;
    }
    // (the keys are taken after the eager call: it may have grown the scratch or the mirrors)
    F.analysis = a;
    F.analysis_generation = pvqt_detail::analysis_generation(a);
    F.frame_time_ns = frame_time_ns;
    F.wanted = wanted;
    F.max_peaks = out->max_peaks;
    F.scratch_generation = v->lane[0].spec.generation;
    F.config = config;
    ++F.calls;
    PVQT_CUDA(cudaStreamSynchronize(v->stream));
    for (int i = 0; i < pvqt_detail::kAnalysisOutputs; ++i)
        if (wanted >> i & 1u)
            std::memcpy(pvqt_detail::analysis_output_member(host, i), F.h_res + off[i],
                        pvqt_detail::analysis_output_bytes_per_frame(a, i, out->max_peaks));
    if (d2h_bytes) *d2h_bytes = res_bytes;
    return PVQT_OK;
}

int run_instant(pvqt *v, const float *x, float *out)
{
    pvqt::Instant &I = v->instant;
    const size_t n_fft = (size_t)v->params.n_fft, nb = v->kernel.n_buckets;
    const size_t skip = std::min(v->upload_skip, n_fft), n_in = n_fft - skip;
    PVQT_CUDA(cudaSetDevice(v->device));
    if (!I.h_in) {
        PVQT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&I.h_in), n_in * sizeof(float), cudaHostAllocDefault));
        PVQT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&I.h_out), nb * sizeof(float), cudaHostAllocDefault));
        PVQT_CUDA(cudaMalloc(reinterpret_cast<void **>(&I.d_in), n_fft * sizeof(float)));
        PVQT_CUDA(cudaMalloc(reinterpret_cast<void **>(&I.d_out), nb * sizeof(float)));
        PVQT_CUDA(cudaMemset(I.d_in, 0, n_fft * sizeof(float)));
    }
    std::memcpy(I.h_in, x + skip, n_in * sizeof(float));
    auto enqueue = [&]() -> int {
        PVQT_CUDA(cudaMemcpyAsync(I.d_in + skip, I.h_in, n_in * sizeof(float), cudaMemcpyHostToDevice, v->stream));
        int rc = run_device(v, I.d_in, 1, 0, n_fft, 1, I.d_out, nullptr, nullptr, v->stream, 0);
        if (rc) return rc;
        PVQT_CUDA(cudaMemcpyAsync(I.h_out, I.d_out, nb * sizeof(float), cudaMemcpyDeviceToHost, v->stream));
        return PVQT_OK;
    };
    const int config = (v->fused_ok ? 1 : 0) | (v->cluster_ok ? 2 : 0) | (v->pipe_ok ? 4 : 0);
    if (I.exec && (I.scratch_generation != v->lane[0].spec.generation || I.config != config)) {
        cudaGraphExecDestroy(I.exec);
        I.exec = nullptr;
        I.calls = 0;
    }
    // the first call runs eagerly (it reserves the lane's scratch: no allocation may happen under capture),
    // the second is captured, every later one replays
    if (!I.exec && I.calls >= 1 && !v->profiling) {
        cudaGraph_t graph = nullptr;
        const uint64_t before = v->launches.load();
        PVQT_CUDA(cudaStreamBeginCapture(v->stream, cudaStreamCaptureModeThreadLocal));
        v->capturing = true;
        int rc = enqueue();
        v->capturing = false;
        I.graph_launches = v->launches.load() - before;
        v->launches.store(before);
        cudaError_t e = cudaStreamEndCapture(v->stream, &graph);
        if (rc == PVQT_OK && e == cudaSuccess && graph) {
            e = cudaGraphInstantiate(&I.exec, graph, 0);
            if (e != cudaSuccess) I.exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
        if (rc != PVQT_OK) { cudaGetLastError(); return rc; }
        if (!I.exec) cudaGetLastError();   // capture unsupported here: stay eager
        I.scratch_generation = v->lane[0].spec.generation;
        I.config = config;
    }
    if (I.exec && !v->profiling) {
        PVQT_CUDA(cudaGraphLaunch(I.exec, v->stream));
        v->launches.fetch_add(I.graph_launches);
    } else {
        int rc = enqueue();
        if (rc) return rc;
    }
    ++I.calls;
    PVQT_CUDA(cudaStreamSynchronize(v->stream));
    std::memcpy(out, I.h_out, nb * sizeof(float));
    return PVQT_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int pvqt_abi_version(void) { return PVQT_ABI_VERSION; }

const char *pvqt_last_error_string(void) { return g_last_error.c_str(); }

int pvqt_device_count(int *count)
{
    if (!count) return fail(PVQT_INVALID_ARGUMENT, "count is null");
    *count = 0;
    PVQT_CUDA(cudaGetDeviceCount(count));
    return PVQT_OK;
}

int pvqt_device_attributes(int device, int32_t *sm_count, int32_t *sm_clock_khz, int32_t *l2_bytes, int32_t *numa_node)
{
    PVQT_CUDA(cudaSetDevice(device));
    int v = 0;
    if (sm_count) { PVQT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); *sm_count = v; }
    if (sm_clock_khz) { PVQT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device)); *sm_clock_khz = v; }
    if (l2_bytes) { PVQT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device)); *l2_bytes = v; }
    if (numa_node) {
        v = -1;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrHostNumaId, device) != cudaSuccess) { cudaGetLastError(); v = -1; }
        *numa_node = v;
    }
    return PVQT_OK;
}

int pvqt_default_params(pvqt_params *out)
{
    if (!out) return fail(PVQT_INVALID_ARGUMENT, "out is null");
    // vqt.rs:180-214 / Default impl vqt.rs:333-348
    const float upscale = 1.0f;                // DEFAULT_UPSCALE_FACTOR
    out->sr = 22050.0f;                        // DEFAULT_SR
    out->n_fft = 2 * 16384;                    // DEFAULT_N_FFT
    out->min_freq = 55.0f;                     // DEFAULT_MIN_FREQ
    out->octaves = 7;                          // DEFAULT_OCTAVES
    out->buckets_per_octave = 12 * 7 * 1;      // DEFAULT_BUCKETS_PER_OCTAVE
    out->sparsity_quantile = 0.999f;           // DEFAULT_SPARSITY_QUANTILE
    out->quality = 1.6f / upscale;             // DEFAULT_Q
    out->gamma = 4.8f * out->quality;          // DEFAULT_GAMMA
    return PVQT_OK;
}

size_t pvqt_params_n_buckets(const pvqt_params *p)
{
    return p ? (size_t)p->buckets_per_octave * (size_t)p->octaves : 0;
}

static void fill_build_error(pvqt_error *err, const pvqt_host::BuildError &be)
{
    err->status = be.status;
    err->highest_frequency = be.highest_frequency;
    err->nyquist_frequency = be.nyquist_frequency;
    err->window_length = be.window_length;
    err->n_fft = be.n_fft;
}

int pvqt_filter_bank_params(const pvqt_params *params, pvqt_filter_params *out, size_t n, pvqt_error *err)
{
    pvqt_error local{};
    if (!err) err = &local;
    std::memset(err, 0, sizeof(*err));
    if (!params || !out) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    std::vector<pvqt_host::FilterParams> fp;
    pvqt_host::BuildError be;
    if (!pvqt_host::filter_bank_params(*params, fp, be)) { fill_build_error(err, be); return fail(be.status, be.message); }
    if (n != fp.size()) return fail(PVQT_BAD_LENGTH, "out must hold n_buckets entries");
    for (size_t i = 0; i < n; ++i) {
        out[i].freq = fp[i].freq;
        out[i].window_length = fp[i].window_length;
        out[i].sr_downscaling_factor = fp[i].sr_downscaling_factor;
        out[i].minimum_needed_window_size = fp[i].minimum_needed_window_size;
    }
    return PVQT_OK;
}

int pvqt_kernel_create(const pvqt_params *params, pvqt_kernel **out, pvqt_error *err)
{
    pvqt_error local{};
    if (!err) err = &local;
    std::memset(err, 0, sizeof(*err));
    if (!params || !out) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    std::unique_ptr<pvqt_kernel> k(new pvqt_kernel());
    pvqt_host::BuildError be;
    if (!pvqt_host::build_kernel(*params, k->kernel, be)) { fill_build_error(err, be); return fail(be.status, be.message); }
    *out = k.release();
    return PVQT_OK;
}

void pvqt_kernel_destroy(pvqt_kernel *k) { delete k; }
size_t pvqt_kernel_n_buckets(const pvqt_kernel *k) { return k ? k->kernel.n_buckets : 0; }
double pvqt_kernel_delay_seconds(const pvqt_kernel *k) { return k ? k->kernel.delay_seconds : 0.0; }
size_t pvqt_kernel_num_window_groups(const pvqt_kernel *k) { return k ? k->kernel.window_groups.size() : 0; }

static int group_window_of(const pvqt_host::Kernel &kernel, size_t group, uint64_t *begin, uint64_t *end)
{
    if (group >= kernel.window_groups.size()) return fail(PVQT_INVALID_ARGUMENT, "bad group index");
    if (begin) *begin = kernel.window_groups[group].window_begin;
    if (end) *end = kernel.window_groups[group].window_end;
    return PVQT_OK;
}

static int group_csr_of(const pvqt_host::Kernel &kernel, size_t group, int negative, pvqt_csr_view *out)
{
    if (!out || group >= kernel.window_groups.size()) return fail(PVQT_INVALID_ARGUMENT, "bad group index");
    const auto &g = kernel.window_groups[group];
    const pvqt_host::Csr &m = negative ? g.negative_filter_bank : g.filter_bank;
    out->rows = m.rows;
    out->cols = m.cols;
    out->nnz = m.nnz();
    out->indptr = m.indptr.data();
    out->indices = m.indices.data();
    out->data = reinterpret_cast<const float *>(m.data.data());
    return PVQT_OK;
}

int pvqt_kernel_group_window(const pvqt_kernel *k, size_t group, uint64_t *begin, uint64_t *end)
{
    if (!k) return fail(PVQT_INVALID_ARGUMENT, "null kernel");
    return group_window_of(k->kernel, group, begin, end);
}

int pvqt_kernel_group_csr(const pvqt_kernel *k, size_t group, int negative, pvqt_csr_view *out)
{
    if (!k) return fail(PVQT_INVALID_ARGUMENT, "null kernel");
    return group_csr_of(k->kernel, group, negative, out);
}

int pvqt_kernel_filter_bandwidths(const pvqt_kernel *k, float *lo_hz, float *hi_hz, size_t n)
{
    if (!k) return fail(PVQT_INVALID_ARGUMENT, "null kernel");
    if (n != k->kernel.n_buckets) return fail(PVQT_BAD_LENGTH, "lo / hi must hold n_buckets entries");
    if (lo_hz) std::copy(k->kernel.band_lo_hz.begin(), k->kernel.band_lo_hz.end(), lo_hz);
    if (hi_hz) std::copy(k->kernel.band_hi_hz.begin(), k->kernel.band_hi_hz.end(), hi_hz);
    return PVQT_OK;
}

int pvqt_kernel_coverage_gaps(const pvqt_kernel *k, uint32_t *out, size_t capacity, size_t *count)
{
    if (!k || !count) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    *count = k->kernel.coverage_gaps.size();
    if (out)
        for (size_t i = 0; i < std::min(capacity, k->kernel.coverage_gaps.size()); ++i) out[i] = k->kernel.coverage_gaps[i];
    return PVQT_OK;
}

int pvqt_create(const pvqt_params *params, int device, pvqt **out, pvqt_error *err)
{
    pvqt_error local{};
    if (!err) err = &local;
    std::memset(err, 0, sizeof(*err));
    if (!params || !out) { err->status = PVQT_INVALID_ARGUMENT; return fail(PVQT_INVALID_ARGUMENT, "null argument"); }
    *out = nullptr;

    std::unique_ptr<pvqt> v(new pvqt());
    v->params = *params;
    pvqt_host::BuildError be;
    if (!pvqt_host::build_kernel(*params, v->kernel, be)) { fill_build_error(err, be); return fail(be.status, be.message); }

    auto cuda_error = [&](cudaError_t e, const char *what) {
        err->status = PVQT_CUDA_ERROR;
        err->cuda_error = (int32_t)e;
        return cuda_fail(e, what);
    };
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess) return cuda_error(e, "cudaGetDeviceCount (no CPU fallback exists)");
    if (device < 0 || device >= n_dev) {
        err->status = PVQT_INVALID_ARGUMENT;
        return fail(PVQT_INVALID_ARGUMENT, "device " + std::to_string(device) + " out of range (" +
                                               std::to_string(n_dev) + " CUDA devices)");
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_error(e, "cudaSetDevice");
    v->device = device;
    if ((e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&v->s_in, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&v->s_out, cudaStreamNonBlocking)) != cudaSuccess)
        return cuda_error(e, "cudaStreamCreate");

    for (auto &L : v->lane) {
        if ((e = cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&L.scratch_free, cudaEventDisableTiming)) != cudaSuccess)
            return cuda_error(e, "create launch lanes");
        if ((e = cudaMalloc(&L.sdft_done, sizeof(unsigned))) != cudaSuccess ||
            (e = cudaMemset(L.sdft_done, 0, sizeof(unsigned))) != cudaSuccess)
            return cuda_error(e, "allocate K-sdft completion counter");
    }
    if (const char *s = std::getenv("PVQT_TILE_FLAGS")) v->tile_flags = std::atoi(s) != 0;
    if (const char *s = std::getenv("PVQT_FRAME_PIPE")) v->frame_pipe_enabled = std::atoi(s) != 0;
    if ((e = cudaEventCreateWithFlags(&v->lane_fork, cudaEventDisableTiming)) != cudaSuccess)
        return cuda_error(e, "create launch lanes");
    if (const char *s = std::getenv("PVQT_LANES")) v->n_lanes = std::max(1, std::min(std::atoi(s), (int)pvqt::kLanes));
    if (const char *s = std::getenv("PVQT_HOST_LANES")) v->host_lanes = std::max(1, std::min(std::atoi(s), (int)pvqt::kLanes));
    if (const char *s = std::getenv("PVQT_SEGMENTS")) v->segments_per_batch = std::max(1, std::atoi(s));
    if (const char *s = std::getenv("PVQT_GRAPHS")) v->use_graphs = std::atoi(s) != 0;
    if (const char *s = std::getenv("PVQT_SEGMENT_FRAMES")) v->segment_cap = (uint32_t)std::max(256, std::atoi(s));
    if (const char *s = std::getenv("PVQT_CHUNK_FRAMES"))   // frames per launch (tuning); kept a multiple of two tiles
        v->chunk_frames = (uint32_t)std::max(2 * kTileFrames, std::min(std::atoi(s), 1 << 20) / (2 * kTileFrames) * (2 * kTileFrames));
    if (const char *s = std::getenv("PVQT_SDFT_TC")) v->sdft_tc = std::max(0, std::min(std::atoi(s), 2));
    if (cudaDeviceGetAttribute(&v->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) v->sm_count = 148;
    pvqt *raw = v.release();
    int rc = build_device_plan(raw);
    if (rc == PVQT_OK) {
        // 0 (default): one work item per CTA, the hardware's CTA scheduler balances the SMs.  Measured against one wave of
        // persistent CTAs looping over their items (PVQT_FFT_WAVE=592: 4 resident CTAs x 148 SMs): 45.5 us against 51.8 us
        // per 3507 frames -- the static split loses more to imbalance than the last, partly empty wave costs.
        raw->fft_wave_ctas = 0;
        if (const char *s = std::getenv("PVQT_FFT_WAVE")) raw->fft_wave_ctas = std::max(0, std::atoi(s));
    }
    if (rc != PVQT_OK) {
        err->status = rc;
        std::string keep = g_last_error;
        pvqt_destroy(raw);
        g_last_error = keep;
        return rc;
    }
    *out = raw;
    return PVQT_OK;
}

void pvqt_destroy(pvqt *v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    for (cudaStream_t s : {v->s_in, v->stream, v->s_out})
        if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (cudaEvent_t e : v->events) cudaEventDestroy(e);
    if (v->graph_exec) cudaGraphExecDestroy(v->graph_exec);
    for (const auto &t : v->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (void *p : v->owned) cudaFree(p);
    for (auto &L : v->lane) {
        if (L.stream) { cudaStreamSynchronize(L.stream); cudaStreamDestroy(L.stream); }
        if (L.done) cudaEventDestroy(L.done);
        if (L.sdft_done) cudaFree(L.sdft_done);
        L.spec.release(); L.power.release(); L.sdft_c.release(); L.sdft_r.release(); L.tile_ready.release();
        if (L.scratch_free) cudaEventDestroy(L.scratch_free);
    }
    if (v->lane_fork) cudaEventDestroy(v->lane_fork);
    release_instant(v);
    release_frame_pipe(v);
    v->d_audio.release();
    v->d_out.release();
    delete v;
}

int pvqt_get_params(const pvqt *v, pvqt_params *out)
{
    if (!v || !out) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    *out = v->params;
    return PVQT_OK;
}
size_t pvqt_n_buckets(const pvqt *v) { return v ? v->kernel.n_buckets : 0; }
size_t pvqt_n_fft(const pvqt *v) { return v ? (size_t)v->params.n_fft : 0; }
double pvqt_delay_seconds(const pvqt *v) { return v ? v->kernel.delay_seconds : 0.0; }
size_t pvqt_num_window_groups(const pvqt *v) { return v ? v->kernel.window_groups.size() : 0; }
int pvqt_device(const pvqt *v) { return v ? v->device : -1; }
size_t pvqt_first_sample_used(const pvqt *v) { return v ? v->first_sample_used : 0; }

int pvqt_group_window(const pvqt *v, size_t group, uint64_t *begin, uint64_t *end)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    return group_window_of(v->kernel, group, begin, end);
}

int pvqt_group_csr(const pvqt *v, size_t group, int negative, pvqt_csr_view *out)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    return group_csr_of(v->kernel, group, negative, out);
}

size_t pvqt_spec_stride(const pvqt *v) { return v ? (size_t)v->fft.spec_stride : 0; }

int pvqt_group_columns(const pvqt *v, size_t group, uint32_t *first_col, uint32_t *n_cols, uint32_t *spec_offset)
{
    if (!v || group >= v->col_lo.size()) return fail(PVQT_INVALID_ARGUMENT, "bad group index");
    if (first_col) *first_col = v->col_lo[group];
    if (n_cols) *n_cols = v->n_cols[group];
    if (spec_offset) *spec_offset = v->spec_off[group];
    return PVQT_OK;
}

size_t pvqt_frames_in(const pvqt *v, size_t n_samples, size_t hop)
{
    if (!v || hop == 0 || n_samples < v->params.n_fft) return 0;
    return (n_samples - (size_t)v->params.n_fft) / hop + 1;
}

// ---- compute entry points ------------------------------------------------------------------
int pvqt_calc_instant_db(pvqt *v, const float *x, size_t n, float *out)
{
    if (!v || !x || !out) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    if (n != v->params.n_fft) return fail(PVQT_BAD_LENGTH, "input must be exactly n_fft samples");  // vqt.rs:867-871
    return run_instant(v, x, out);
}

int pvqt_calc_batch_db(pvqt *v, const float *audio, size_t n_samples, size_t hop, size_t n_frames, float *out)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    return run_host(v, audio, 1, 0, n_samples, hop, n_frames, out);
}

int pvqt_calc_frames_db(pvqt *v, const float *frames, size_t n_frames, float *out)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    const size_t n_fft = (size_t)v->params.n_fft;
    size_t n_samples = 0;
    if (!mul_ok(n_frames, n_fft, &n_samples)) return fail(PVQT_INVALID_ARGUMENT, "too many frames");
    return run_host(v, frames, 1, 0, n_samples, n_fft, n_frames, out);
}

int pvqt_calc_streams_db(pvqt *v, const float *audio, size_t n_streams, size_t stream_stride, size_t n_samples,
                         size_t hop, size_t frames_per_stream, float *out)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    return run_host(v, audio, n_streams, stream_stride, n_samples, hop, frames_per_stream, out);
}

int pvqt_calc_db_device(pvqt *v, const float *d_audio, size_t n_streams, size_t stream_stride, size_t hop,
                        size_t frames_per_stream, float *d_out, float *d_power, void *cuda_stream)
{
    if (!v || !d_audio || !d_out) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    PVQT_CUDA(cudaSetDevice(v->device));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : v->stream;
    return run_device(v, d_audio, n_streams, stream_stride, hop, frames_per_stream, d_out, d_power, nullptr, st);
}

int pvqt_fft_device(pvqt *v, const float *d_audio, size_t n_streams, size_t stream_stride, size_t hop,
                    size_t frames_per_stream, float *d_spec, void *cuda_stream)
{
    if (!v || !d_audio || !d_spec) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    PVQT_CUDA(cudaSetDevice(v->device));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : v->stream;
    return run_device(v, d_audio, n_streams, stream_stride, hop, frames_per_stream, nullptr, nullptr, d_spec, st);
}

// ---- VQT + AnalysisState in one call (BASELINE configs[4]) ------------------------------------
// What every caller of the reference does per frame -- calculate_vqt_instant_in_db then preprocess
// (pitchvis_viewer/src/vqt_system.rs:40-68 + analysis_system.rs:10-20, pitchvis_serial/src/main.rs:206-230) -- for
// whole recordings: the dB spectra go from K-spmm-db to K-analysis through HBM and never visit the host unless
// out_db asks for them.
int pvqt_calc_streams_analysis(pvqt *v, pvqt_analysis *a, const float *audio, size_t n_streams, size_t stream_stride,
                               size_t n_samples, size_t hop, size_t frames_per_stream, uint64_t frame_time_ns,
                               const pvqt_analysis_outputs *out, float *out_db, uint64_t *d2h_bytes)
{
    if (!v || !a || !audio) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    if (pvqt_detail::analysis_device(a) != v->device)
        return fail(PVQT_INVALID_ARGUMENT, "the Vqt and the AnalysisState must live on the same device");
    const size_t nb = v->kernel.n_buckets, n_fft = (size_t)v->params.n_fft;
    if (pvqt_analysis_n_buckets(a) != nb)
        return fail(PVQT_BAD_LENGTH, "x_vqt.len() must equal range.n_buckets()");   // analysis.rs:289
    if (pvqt_analysis_n_streams(a) != n_streams)
        return fail(PVQT_INVALID_ARGUMENT, "the AnalysisState holds another number of streams");
    if (d2h_bytes) *d2h_bytes = 0;
    if (n_streams == 0 || frames_per_stream == 0) return PVQT_OK;
    if (hop == 0 && frames_per_stream > 1) return fail(PVQT_INVALID_ARGUMENT, "hop must be positive");
    if (!frames_fit(frames_per_stream, hop, n_fft, n_samples))
        return fail(PVQT_BAD_LENGTH, "each stream must hold (frames_per_stream - 1) * hop + n_fft samples");
    size_t frames = 0;
    if (!mul_ok(n_streams, frames_per_stream, &frames)) return fail(PVQT_INVALID_ARGUMENT, "too many frames");
    if (n_streams == 1 && frames_per_stream == 1 && out && !out_db && v->frame_pipe_enabled)
        return run_frame_pipeline(v, a, audio, frame_time_ns, out, d2h_bytes);
    PVQT_CUDA(cudaSetDevice(v->device));
    pvqt_analysis_outputs dev{};
    int rc = pvqt_detail::analysis_outputs_reserve(a, out, frames, &dev, v->stream);
    if (rc) return rc;
    // stream groups that fit one staging batch (one long recording must fit it as a whole)
    const size_t span = (frames_per_stream - 1) * hop + n_fft, dstride = (span + 3) & ~(size_t)3;
    const size_t group = std::max<size_t>(1, v->staging_budget_samples / dstride);
    size_t total_d2h = 0;
    const bool timing = std::getenv("PVQT_DEBUG_TIMING") != nullptr;   // diagnostic: drains the stream between the stages
    auto stamp = [&](const char *what, size_t g0) {
        if (!timing) return;
        cudaStreamSynchronize(v->stream);
        static auto t_prev = std::chrono::steady_clock::now();
        const auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "pvqt_calc_streams_analysis: %-22s group at stream %zu: +%.2f ms\n", what, g0,
                     std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    stamp("start", 0);
    for (size_t g0 = 0; g0 < n_streams; g0 += group) {
        const size_t ng = std::min(group, n_streams - g0);
        rc = run_host(v, audio + g0 * stream_stride, ng, stream_stride, n_samples, hop, frames_per_stream, nullptr, true);
        if (rc) return rc;
        stamp("H2D + VQT", g0);
        const float *d_db = static_cast<const float *>(v->d_out.ptr);
        rc = pvqt_detail::analysis_run_device(a, d_db, g0, ng, frames_per_stream, frame_time_ns, out ? &dev : nullptr,
                                              g0 * frames_per_stream, v->stream);
        if (rc) return rc;
        v->launches.fetch_add(1);
        stamp("K-analysis", g0);
        if (out_db) {
            const size_t bytes = ng * frames_per_stream * nb * sizeof(float);
            PVQT_CUDA(cudaMemcpyAsync(out_db + g0 * frames_per_stream * nb, d_db, bytes, cudaMemcpyDeviceToHost, v->stream));
            total_d2h += bytes;
        }
        if (g0 + group < n_streams) PVQT_CUDA(cudaStreamSynchronize(v->stream));   // d_audio / d_out are reused
    }
    size_t res_bytes = 0;
    rc = pvqt_detail::analysis_outputs_download(a, out, &dev, frames, v->stream, &res_bytes);
    if (rc) return rc;
    PVQT_CUDA(cudaStreamSynchronize(v->stream));
    stamp("download of the results", 0);
    if (d2h_bytes) *d2h_bytes = total_d2h + res_bytes;
    return PVQT_OK;
}

int pvqt_calc_batch_analysis(pvqt *v, pvqt_analysis *a, const float *audio, size_t n_samples, size_t hop, size_t n_frames,
                             uint64_t frame_time_ns, const pvqt_analysis_outputs *out, float *out_db, uint64_t *d2h_bytes)
{
    return pvqt_calc_streams_analysis(v, a, audio, 1, 0, n_samples, hop, n_frames, frame_time_ns, out, out_db, d2h_bytes);
}

// ---- memory / timing helpers ---------------------------------------------------------------
int pvqt_dev_alloc(int device, size_t bytes, void **out)
{
    if (!out) return fail(PVQT_INVALID_ARGUMENT, "out is null");
    PVQT_CUDA(cudaSetDevice(device));
    PVQT_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return PVQT_OK;
}
int pvqt_dev_free(int device, void *p)
{
    PVQT_CUDA(cudaSetDevice(device));
    PVQT_CUDA(cudaFree(p));
    return PVQT_OK;
}
int pvqt_host_alloc_pinned(size_t bytes, void **out)
{
    if (!out) return fail(PVQT_INVALID_ARGUMENT, "out is null");
    PVQT_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return PVQT_OK;
}
int pvqt_host_free_pinned(void *p)
{
    PVQT_CUDA(cudaFreeHost(p));
    return PVQT_OK;
}
int pvqt_memcpy_h2d(pvqt *v, void *dst, const void *src, size_t bytes, int async)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    PVQT_CUDA(cudaSetDevice(v->device));
    PVQT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, v->stream));
    if (!async) PVQT_CUDA(cudaStreamSynchronize(v->stream));
    return PVQT_OK;
}
int pvqt_memcpy_d2h(pvqt *v, void *dst, const void *src, size_t bytes, int async)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    PVQT_CUDA(cudaSetDevice(v->device));
    PVQT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, v->stream));
    if (!async) PVQT_CUDA(cudaStreamSynchronize(v->stream));
    return PVQT_OK;
}
int pvqt_dev_memset(pvqt *v, void *dst, int value, size_t bytes)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    PVQT_CUDA(cudaSetDevice(v->device));
    PVQT_CUDA(cudaMemsetAsync(dst, value, bytes, v->stream));
    return PVQT_OK;
}
int pvqt_dev_flush_l2(pvqt *v, void *scratch, size_t bytes)
{
    if (!v || !scratch || bytes < 16) return fail(PVQT_INVALID_ARGUMENT, "null handle or scratch");
    PVQT_CUDA(cudaSetDevice(v->device));
    PVQT_CUDA(cudaMemsetAsync(scratch, 0, bytes, v->stream));
    PVQT_CUDA(launch_read_sweep(scratch, bytes, static_cast<unsigned *>(scratch), v->stream));
    return PVQT_OK;
}

int pvqt_synchronize(pvqt *v)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    PVQT_CUDA(cudaSetDevice(v->device));
    PVQT_CUDA(cudaStreamSynchronize(v->stream));
    return PVQT_OK;
}
int pvqt_event_create(pvqt *v, void **out_event)
{
    if (!v || !out_event) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    PVQT_CUDA(cudaSetDevice(v->device));
    cudaEvent_t ev;
    PVQT_CUDA(cudaEventCreate(&ev));
    *out_event = ev;
    return PVQT_OK;
}
int pvqt_event_destroy(pvqt *v, void *event)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    PVQT_CUDA(cudaSetDevice(v->device));
    PVQT_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
    return PVQT_OK;
}
int pvqt_event_record(pvqt *v, void *event)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    PVQT_CUDA(cudaSetDevice(v->device));
    PVQT_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), v->stream));
    return PVQT_OK;
}
int pvqt_event_elapsed_ms(pvqt *v, void *start, void *stop, float *ms)
{
    if (!v || !ms) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    PVQT_CUDA(cudaSetDevice(v->device));
    PVQT_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(stop)));
    PVQT_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
    return PVQT_OK;
}
uint64_t pvqt_launch_count(const pvqt *v) { return v ? v->launches.load() : 0; }

int pvqt_set_profiling(pvqt *v, int enabled)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    v->profiling = enabled != 0;
    return PVQT_OK;
}

int pvqt_set_fused_epilogue(pvqt *v, int mode)
{
    if (!v) return 0;
    v->cluster_ok = v->cluster_capable && mode == 2;
    v->fused_ok = v->fused_capable && mode >= 1;
    v->pipe_ok = v->pipe_capable && v->fused_ok && mode >= 3;
    if (v->graph_exec) { cudaGraphExecDestroy(v->graph_exec); v->graph_exec = nullptr; }
    return v->cluster_ok ? 2 : (v->pipe_ok ? 3 : (v->fused_ok ? 1 : 0));
}

int pvqt_plan_info(const pvqt *v, int32_t *out, size_t n)
{
    if (!v || !out) return fail(PVQT_INVALID_ARGUMENT, "null argument");
    int32_t walk = 0;   // band slots the warps of one K-spmm-db CTA walk per tile (padding included)
    if (v->fused_capable)
        for (int w = 0; w < v->fused.n_warps; ++w) walk += v->fused.warp[w].width + v->fused.warp[w].nwidth;
    const int32_t info[10] = {v->cluster_capable ? v->cluster.cluster_size : 0, v->cluster_max_active, v->cluster.coef_bytes,
                              v->cluster.max_rows, v->fused_capable ? v->fused.n_warps : 0, v->fft_block_threads,
                              v->fft.spec_stride, (int32_t)v->sdft_plans.size(), walk, (int32_t)v->last_sdft_mask};
    for (size_t i = 0; i < n && i < 10; ++i) out[i] = info[i];
    return PVQT_OK;
}


int pvqt_set_sliding_dft(pvqt *v, int mode)
{
    if (!v) return 0;
    v->sdft_enabled = mode != 0;
    v->sdft_tensor_cores = mode >= 2;
    v->sdft_tc = mode >= 3 ? 1 : 0;
    if (v->graph_exec) { cudaGraphExecDestroy(v->graph_exec); v->graph_exec = nullptr; }
    return v->sdft_enabled ? (v->sdft_tc ? 3 : v->sdft_tensor_cores ? 2 : 1) : 0;
}

int pvqt_get_profile(pvqt *v, int reset, double *kernel_ms, uint64_t *kernel_launches)
{
    if (!v) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    PVQT_CUDA(cudaSetDevice(v->device));
    double ms[PVQT_PROFILE_KINDS] = {};
    uint64_t n[PVQT_PROFILE_KINDS] = {};
    for (const auto &t : v->timed) {
        float e = 0.f;
        PVQT_CUDA(cudaEventSynchronize(t.b));
        PVQT_CUDA(cudaEventElapsedTime(&e, t.a, t.b));
        ms[t.kind] += e;
        n[t.kind] += 1;
    }
    if (reset) {
        for (const auto &t : v->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
        v->timed.clear();
    }
    for (int i = 0; i < PVQT_PROFILE_KINDS; ++i) {
        if (kernel_ms) kernel_ms[i] = ms[i];
        if (kernel_launches) kernel_launches[i] = n[i];
    }
    return PVQT_OK;
}

// ---- sharding ------------------------------------------------------------------------------
int pvqt_shard_range(size_t n_units, size_t n_parts, size_t part, size_t *begin, size_t *end)
{
    if (n_parts == 0 || part >= n_parts || !begin || !end) return fail(PVQT_INVALID_ARGUMENT, "bad shard request");
    const size_t base = n_units / n_parts, rem = n_units % n_parts;
    *begin = part * base + std::min(part, rem);
    *end = *begin + base + (part < rem ? 1 : 0);
    return PVQT_OK;
}

int pvqt_frame_range_samples(size_t n_fft, size_t hop, size_t f0, size_t f1, size_t *s0, size_t *s1)
{
    if (!s0 || !s1 || f1 < f0) return fail(PVQT_INVALID_ARGUMENT, "bad frame range");
    if (f1 == f0) { *s0 = 0; *s1 = 0; return PVQT_OK; }
    *s0 = f0 * hop;
    *s1 = (f1 - 1) * hop + n_fft;
    return PVQT_OK;
}

}  // extern "C"

// ---- single-process multi-GPU ----------------------------------------------------------------
struct pvqt_multi {
    std::vector<pvqt *> handles;
};

extern "C" {

int pvqt_multi_create(const pvqt_params *params, int n_devices, const int *device_ids, pvqt_multi **out,
                      pvqt_error *err)
{
    if (!out || n_devices <= 0) return fail(PVQT_INVALID_ARGUMENT, "bad multi-device request");
    *out = nullptr;
    std::unique_ptr<pvqt_multi> m(new pvqt_multi());
    for (int i = 0; i < n_devices; ++i) {
        pvqt *h = nullptr;
        int rc = pvqt_create(params, device_ids ? device_ids[i] : i, &h, err);
        if (rc != PVQT_OK) {
            std::string keep = g_last_error;
            for (pvqt *p : m->handles) pvqt_destroy(p);
            g_last_error = keep;
            return rc;
        }
        m->handles.push_back(h);
    }
    *out = m.release();
    return PVQT_OK;
}

void pvqt_multi_destroy(pvqt_multi *m)
{
    if (!m) return;
    for (pvqt *p : m->handles) pvqt_destroy(p);
    delete m;
}

int pvqt_multi_num_devices(const pvqt_multi *m) { return m ? (int)m->handles.size() : 0; }
pvqt *pvqt_multi_handle(pvqt_multi *m, int index)
{
    return (m && index >= 0 && index < (int)m->handles.size()) ? m->handles[index] : nullptr;
}

}  // extern "C"

namespace {

// run `fn(part)` on one host thread per device; the first non-OK status (and its message) wins
template <typename Fn>
int for_each_device(pvqt_multi *m, Fn fn)
{
    const size_t n = m->handles.size();
    std::vector<int> rc(n, PVQT_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> threads;
    for (size_t i = 0; i < n; ++i)
        threads.emplace_back([&, i] {
            rc[i] = fn(i);
            if (rc[i] != PVQT_OK) msg[i] = g_last_error;
        });
    for (auto &t : threads) t.join();
    for (size_t i = 0; i < n; ++i)
        if (rc[i] != PVQT_OK) { g_last_error = msg[i]; return rc[i]; }
    return PVQT_OK;
}

}  // namespace

extern "C" {

int pvqt_multi_calc_batch_db(pvqt_multi *m, const float *audio, size_t n_samples, size_t hop, size_t n_frames,
                             float *out)
{
    if (!m || m->handles.empty()) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    const size_t n_fft = pvqt_n_fft(m->handles[0]), nb = pvqt_n_buckets(m->handles[0]);
    if (n_frames == 0) return PVQT_OK;
    if (hop == 0 && n_frames > 1) return fail(PVQT_INVALID_ARGUMENT, "hop must be positive");
    if (!frames_fit(n_frames, hop, n_fft, n_samples))
        return fail(PVQT_BAD_LENGTH, "audio must hold (n_frames - 1) * hop + n_fft samples");
    return for_each_device(m, [&](size_t i) {
        size_t f0, f1, s0, s1;
        pvqt_shard_range(n_frames, m->handles.size(), i, &f0, &f1);
        if (f1 == f0) return (int)PVQT_OK;
        pvqt_frame_range_samples(n_fft, hop, f0, f1, &s0, &s1);
        pvqt *h = m->handles[i];
        h->sdft_job_frames = n_frames;   // the shard takes the code path of the whole recording: same bits as unsharded
        const int rc = pvqt_calc_batch_db(h, audio + s0, s1 - s0, hop, f1 - f0, out + f0 * nb);
        h->sdft_job_frames = 0;
        return rc;
    });
}

// What the host <-> device links of the box carry when every device of `m` copies at once: h2d_bytes in and d2h_bytes out
// per device and repetition, from / to pinned host memory, the two directions on their own streams -- the same traffic
// the host-buffer entries generate, without any kernel.  *seconds = wall time of `reps` repetitions on all devices.
int pvqt_multi_pcie_probe(pvqt_multi *m, size_t h2d_bytes, size_t d2h_bytes, int reps, double *seconds)
{
    if (!m || m->handles.empty() || !seconds || reps <= 0) return fail(PVQT_INVALID_ARGUMENT, "bad probe request");
    const size_t n = m->handles.size();
    std::vector<void *> h_in(n, nullptr), h_out(n, nullptr), d_in(n, nullptr), d_out(n, nullptr);
    int rc = for_each_device(m, [&](size_t i) {
        pvqt *v = m->handles[i];
        PVQT_CUDA(cudaSetDevice(v->device));
        PVQT_CUDA(cudaHostAlloc(&h_in[i], std::max<size_t>(h2d_bytes, 16), cudaHostAllocPortable));
        PVQT_CUDA(cudaHostAlloc(&h_out[i], std::max<size_t>(d2h_bytes, 16), cudaHostAllocPortable));
        PVQT_CUDA(cudaMalloc(&d_in[i], std::max<size_t>(h2d_bytes, 16)));
        PVQT_CUDA(cudaMalloc(&d_out[i], std::max<size_t>(d2h_bytes, 16)));
        std::memset(h_in[i], 0, std::max<size_t>(h2d_bytes, 16));
        PVQT_CUDA(cudaMemset(d_out[i], 0, std::max<size_t>(d2h_bytes, 16)));
        PVQT_CUDA(cudaMemcpyAsync(d_in[i], h_in[i], h2d_bytes, cudaMemcpyHostToDevice, v->s_in));   // warm-up
        PVQT_CUDA(cudaMemcpyAsync(h_out[i], d_out[i], d2h_bytes, cudaMemcpyDeviceToHost, v->s_out));
        PVQT_CUDA(cudaStreamSynchronize(v->s_in));
        PVQT_CUDA(cudaStreamSynchronize(v->s_out));
        return (int)PVQT_OK;
    });
    if (rc == PVQT_OK) {
        const auto t0 = std::chrono::steady_clock::now();
        rc = for_each_device(m, [&](size_t i) {
            pvqt *v = m->handles[i];
            PVQT_CUDA(cudaSetDevice(v->device));
            for (int r = 0; r < reps; ++r) {
                if (h2d_bytes) PVQT_CUDA(cudaMemcpyAsync(d_in[i], h_in[i], h2d_bytes, cudaMemcpyHostToDevice, v->s_in));
                if (d2h_bytes) PVQT_CUDA(cudaMemcpyAsync(h_out[i], d_out[i], d2h_bytes, cudaMemcpyDeviceToHost, v->s_out));
            }
            PVQT_CUDA(cudaStreamSynchronize(v->s_in));
            PVQT_CUDA(cudaStreamSynchronize(v->s_out));
            return (int)PVQT_OK;
        });
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    for (size_t i = 0; i < n; ++i) {
        cudaSetDevice(m->handles[i]->device);
        if (h_in[i]) cudaFreeHost(h_in[i]);
        if (h_out[i]) cudaFreeHost(h_out[i]);
        if (d_in[i]) cudaFree(d_in[i]);
        if (d_out[i]) cudaFree(d_out[i]);
    }
    return rc;
}

int pvqt_multi_calc_streams_db(pvqt_multi *m, const float *audio, size_t n_streams, size_t stream_stride,
                               size_t n_samples, size_t hop, size_t frames_per_stream, float *out)
{
    if (!m || m->handles.empty()) return fail(PVQT_INVALID_ARGUMENT, "null handle");
    const size_t nb = pvqt_n_buckets(m->handles[0]);
    return for_each_device(m, [&](size_t i) {
        size_t s0, s1;
        pvqt_shard_range(n_streams, m->handles.size(), i, &s0, &s1);
        if (s1 == s0) return (int)PVQT_OK;
        return pvqt_calc_streams_db(m->handles[i], audio + s0 * stream_stride, s1 - s0, stream_stride, n_samples, hop,
                                    frames_per_stream, out + s0 * frames_per_stream * nb);
    });
}

}  // extern "C"

#ifdef PVQT_PHASE_TIMERS
namespace pvqt_dev { cudaError_t read_phase_stamps(unsigned long long *out); cudaError_t read_sdft_stamps(unsigned long long *out); }
extern "C" int pvqt_debug_sdft_stamps(unsigned long long *out)
{
    cudaDeviceSynchronize();
    return pvqt_dev::read_sdft_stamps(out) == cudaSuccess ? 0 : 7;
}
// Diagnostic build only: out[2][8192][8] globaltimer stamps (kernel 0 = K-fft, 1 = K-spmm-db) of the last launches.
extern "C" int pvqt_debug_phase_stamps(unsigned long long *out)
{
    cudaDeviceSynchronize();
    return pvqt_dev::read_phase_stamps(out) == cudaSuccess ? 0 : 7;
}
#endif

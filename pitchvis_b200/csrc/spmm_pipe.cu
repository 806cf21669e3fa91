// spmm_pipe.cu -- K-spmm-db as a persistent, warp-specialised pipeline (one CTA per SM).
//
// Same arithmetic, in the same order, as spmm_db_fused_kernel (vqt_kernels.cu): the banded complex SpMM with its
// conjugate part (vqt.rs:889-910), |z|^2 and power_to_db (vqt.rs:922-954) for 8-frame tiles -- the results are
// bit-identical.  What changes is WHEN things happen.  The one-CTA-per-tile form runs its phases in lock-step on
// every SM (one wave of CTAs: all stage their tile from L2, all walk, all convert), so the L2 burst, the
// shared-memory-bound walk and the store phase never overlap: 35 us for 3507 frames of which 17.5 us are the walk
// (profiles/r01_zz_phase_timers_after.txt).  Here one CTA per SM loops over its tiles with two roles:
//
//   walkers  (n_warps warps, the lanes of the band walk)   wait FULL[b] -> walk planes[b] -> arrive EMPTY[b]
//                                                           -> log-spectrum of the tile into ls[b] -> arrive LSFULL[b]
//   helpers  (kPipeHelpers warps)                           one thread issues the tile's bulk copies (cp.async.bulk, TMA
//                                                           engine): the spectrum tile -- K-fft writes it as the very
//                                                           plane image the walk reads -- straight into planes[b^1], the
//                                                           K-sdft chunk rows of the tile into a staging buffer; the
//                                                           warps wait on the copies' mbarrier, run the K-sdft combine
//                                                           from shared memory, arrive FULL[b^1]; then power_to_db's
//                                                           frame-wise part for tile i-1 from ls[..]
//
// so the L2 traffic, the combine and the dB stores of neighbouring tiles hide behind the walk.  One CTA per SM leaves
// the walkers ~100 registers: the walk is software-pipelined (the spectrum records and the coefficient of slot j+1 are
// in registers before the FMAs of slot j issue), which is what lets 10 warps keep the FMA and shared-memory pipes
// busy at once (with 64 registers and the loads inside the loop body a lone CTA walked at half the rate of three
// co-resident ones: profiles/r02_pipe_stats.txt).
#include "device_helpers.cuh"
#include "sdft_combine.cuh"
#include "vqt_device.cuh"

#include <algorithm>

#ifndef PVQT_PIPE_HELPERS
#define PVQT_PIPE_HELPERS 6
#endif

namespace pvqt_dev {
#ifdef PVQT_PIPE_STATS
__device__ long long g_pipe_stats[256][16];
__device__ unsigned long long g_pipe_stamps[256][8];   // globaltimer: CTA start, after pdl_wait, first FULL, last walk done, end
#define PIPE_STAMP(slot)                                                                  \
    do {                                                                                  \
        if (lane == 0 && blockIdx.x < 256) {                                              \
            unsigned long long t_;                                                        \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                        \
            g_pipe_stamps[blockIdx.x][slot] = t_;                                         \
        }                                                                                 \
    } while (0)
#else
#define PIPE_STAMP(slot) do { } while (0)
#endif
namespace {

constexpr int kPipeHelpers = PVQT_PIPE_HELPERS;   // helper warps per CTA
constexpr int kPipeMaxThreads = 512;              // 16 warps: 128 registers per thread for the pipelined walk (ptxas sizes
                                                  // the register file for the thread count rounded up to 128: 576 -> 96)

__device__ __forceinline__ void bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// barrier ids (0 is __syncthreads)
enum : int { kBarFull = 1, kBarEmpty = 3, kBarLsFull = 5, kBarLsEmpty = 7, kBarHelpers = 9 };

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
// bulk copy global -> shared through the TMA engine; completes `bytes` on the mbarrier.  16-byte aligned, multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifdef PVQT_PIPE_STATS
// Diagnostic build only (scripts/pipe_stats.py): cycles lane 0 of walker warp 0 / helper warp 0 spends per phase.
// The stamp reads shared memory first: a plain clock read after BAR.SYNC.DEFER_BLOCKING issues before the barrier
// completes (the wait would be booked on the next phase); a memory access cannot.
#define PIPE_T(var)                                                                                            \
    long long var;                                                                                            \
    {                                                                                                         \
        unsigned d_;                                                                                          \
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(d_) : "r"(smem_addr(&mbar[0])) : "memory");   \
        var = clock64() + (long long)(d_ & 0u);                                                               \
    }
#define PIPE_ACC(slot, t0, t1) do { if (lane == 0 && (warp == 0 || warp == nw) && blockIdx.x < 256) g_pipe_stats[blockIdx.x][slot] += (t1) - (t0); } while (0)
#else
#define PIPE_T(var) do { } while (0)
#define PIPE_ACC(slot, t0, t1) do { } while (0)
#endif

// Staging of one K-sdft group for the combine (float2 entries): C rows | R rows | phase table, each part 16-byte aligned
__host__ __device__ inline int sd_c_cap(int q, int nk) { return ((q + kTileFrames) * nk + 2 + 1) & ~1; }
__host__ __device__ inline int sd_r_cap(int nk) { return (kTileFrames * nk + 2 + 1) & ~1; }

#ifdef PVQT_PIPE_NOK
#define PIPE_KLOAD(u, off) do { } while (0)
#else
#define PIPE_KLOAD(u, off) kq[u] = __ldg(kv + min(j + (off), last) * 32)
#endif

// One band slot in registers: the lane's coefficient pair and the four chunks of its spectrum record.
struct Slot {
    float4 k, xr03, xr47, xi03, xi47;
};

// y += k x (band slots) or y += k conj(x) (the conjugate-part band): the products and their order are those of
// mac8<false> / mac8<true> (device_helpers.cuh).  The two forms differ by the signs of two scalars, flipped with integer
// XORs on the sign bit: exact, and on the ALU pipe -- the FMA pipe, which bounds the walk, issues FFMA2s only.
__device__ __forceinline__ void mac_slot(float2 (&re)[2][4], float2 (&im)[2][4], const Slot &s, bool band)
{
    const float2 xr[4] = {make_float2(s.xr03.x, s.xr03.y), make_float2(s.xr03.z, s.xr03.w), make_float2(s.xr47.x, s.xr47.y),
                          make_float2(s.xr47.z, s.xr47.w)};
    const float2 xi[4] = {make_float2(s.xi03.x, s.xi03.y), make_float2(s.xi03.z, s.xi03.w), make_float2(s.xi47.x, s.xi47.y),
                          make_float2(s.xi47.z, s.xi47.w)};
    const float kk[2][2] = {{s.k.x, s.k.y}, {s.k.z, s.k.w}};
    const unsigned flip_b = band ? 0x80000000u : 0u, flip_c = band ? 0u : 0x80000000u;   // b = -kim | +kim, c = +kre | -kre
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const float kre = kk[r][0], kim = kk[r][1];
        const float bb = __uint_as_float(__float_as_uint(kim) ^ flip_b), cc = __uint_as_float(__float_as_uint(kre) ^ flip_c);
        const float2 a = make_float2(kre, kre), b = make_float2(bb, bb), c = make_float2(cc, cc), d = make_float2(kim, kim);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            re[r][p] = __ffma2_rn(a, xr[p], re[r][p]);
            re[r][p] = __ffma2_rn(b, xi[p], re[r][p]);
            im[r][p] = __ffma2_rn(c, xi[p], im[r][p]);
            im[r][p] = __ffma2_rn(d, xr[p], im[r][p]);
        }
    }
}

template <int PLANE>
__global__ void __launch_bounds__(kPipeMaxThreads, 1) spmm_db_pipe_kernel(const __grid_constant__ FusedParams P, const int sdft_floats2)
{
    extern __shared__ __align__(128) float4 pipe_smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = P.n_warps;                       // walker warps
    const int all = blockDim.x;                     // threads of both roles (barrier count of the hand-offs)
    const int nb = P.n_buckets;
    float4 *planes0 = pipe_smem;                    // [2][4 * PLANE]
    float *ls0 = reinterpret_cast<float *>(pipe_smem + 2 * 4 * PLANE);   // [2][8 * nb], 16-byte aligned halves
    const int ls_stride = (kTileFrames * nb + 3) & ~3;
    float2 *sd0 = reinterpret_cast<float2 *>(ls0 + 2 * ls_stride);       // combine staging: C rows | R rows | phase
    const uint32_t n_my = P.n_tiles > blockIdx.x ? (P.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();
    if (warp == 0) PIPE_STAMP(0);

    if (warp < nw) {
        // ------------------------------------------------------------------ walkers
        const FusedWarp W = P.warp[warp];
        const int4 meta = __ldg(P.lane_meta + warp * 32 + lane);
        const int2 rows = __ldg(P.lane_rows + warp * 32 + lane);
        const int width = W.width, n_slots = W.width + W.nwidth;
        // band slots, then the conjugate-part slots; CTA b streams copy b % copies of the (identical) coefficient arrays
        const float4 *kv = P.values + (size_t)(blockIdx.x % P.values_copies) * P.values_stride + (size_t)W.val_base * 32 + lane;
        const int last = n_slots - 1;
        // Does any lane of this warp read a column the helpers' K-sdft combine writes?  The other warps (8 of 10 at the
        // defaults) start a tile as soon as its bulk copy has landed and only signal the FULL barrier.
        bool reads_combined = false;
        for (int gi = 0; gi < P.n_sdft; ++gi) {
            const int c0 = P.sdft[gi].g.spec_offset, c1 = c0 + P.sdft[gi].g.nk;
            reads_combined |= (meta.x < c1 && meta.x + width > c0) || (W.nwidth > 0 && meta.z < c1 && meta.z + W.nwidth > c0);
        }
        reads_combined = __any_sync(0xffffffffu, reads_combined);
        for (uint32_t i = 0; i < n_my; ++i) {
            const int b = (int)(i & 1);
            const uint32_t tile = blockIdx.x + i * gridDim.x;
            const float4 *planes = planes0 + b * (4 * PLANE);
            float2 re[2][4], im[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int p = 0; p < 4; ++p) re[r][p] = im[r][p] = make_float2(0.f, 0.f);
            // The coefficient stream (198 KB per tile at the defaults, the same for every tile, L2-resident) goes through a
            // queue of kQ registers per lane, refilled kQ slots ahead with plain loads: through shared memory (the
            // cp.async ring of the one-CTA-per-tile form) every coefficient cost an LDS.128 and an LDGSTS -- 12 of the 28
            // cycles of the SM-wide shared-memory pipe a slot took, and that pipe, not the FMA pipe, bounded the walk
            // (profiles/r02_pipe_stats.txt).  The head of the queue is loaded before the tile's barrier: plan data only.
            // Also measured: a warp-private ring of three 4-slot chunks filled by cp.async.bulk (lane 0 refills a chunk
            // as soon as the warp has read it, 8-12 slots ahead): the per-chunk mbarrier wait, warp sync, proxy fence and
            // issue cost more than the deeper prefetch gave back (walk 46.1 k cycles against 37.9 k, step 87.0 against
            // 82.8 us).
            constexpr int kQ = 4;
            float4 kq[kQ];
#pragma unroll
            for (int u = 0; u < kQ; ++u) kq[u] = __ldg(kv + min(u, last) * 32);
            PIPE_T(w0);
            if (reads_combined) bar_sync(kBarFull + b, all);          // the helpers have staged and combined this tile
            else bar_arrive(kBarFull + b, all);                       // (counted, not waited for)
            mbar_wait(&mbar[b], (i >> 1) & 1);                        // the tile's bulk copy has landed (acquires its writes)
            PIPE_T(w1);
            PIPE_ACC(0, w0, w1);
            const float4 *xband = planes + meta.x, *xconj = planes + meta.z - width;   // record of slot j: x?[j]
            auto loadx = [&](Slot &s, int j) {
#ifdef PVQT_PIPE_NOX
                if (j > 0) return;
#endif
                const float4 *xp = (j < width ? xband : xconj) + j;
                s.xr03 = xp[0];
                s.xr47 = xp[PLANE];
                s.xi03 = xp[2 * PLANE];
                s.xi47 = xp[3 * PLANE];
            };
            auto mac = [&](Slot &s, const float4 &k, int j) {
                s.k = k;
#ifdef PVQT_PIPE_NOMAC
                re[0][0].x += s.k.x + s.xr03.x + s.xr47.y + s.xi03.z + s.xi47.w;
#else
                mac_slot(re, im, s, j < width);
#endif
            };
            // software pipeline, four slots per trip: the records of slot j+1 are in flight while slot j multiplies, the
            // coefficient of slot j+4 replaces the one just used
            Slot A, B;
            int j = 0;
            loadx(A, 0);
#pragma unroll 1
            for (; j + kQ <= n_slots; j += kQ) {
                loadx(B, j + 1);
                mac(A, kq[0], j);
                PIPE_KLOAD(0, 4);
                loadx(A, j + 2);
                mac(B, kq[1], j + 1);
                PIPE_KLOAD(1, 5);
                loadx(B, j + 3);
                mac(A, kq[2], j + 2);
                PIPE_KLOAD(2, 6);
                loadx(A, min(j + 4, last));
                mac(B, kq[3], j + 3);
                PIPE_KLOAD(3, 7);
            }
            const int rest = n_slots - j;                             // 0..3 slots left; A holds the records of slot j
            if (rest >= 1) {
                if (rest >= 2) loadx(B, j + 1);
                mac(A, kq[0], j);
            }
            if (rest >= 2) {
                if (rest >= 3) loadx(A, j + 2);
                mac(B, kq[1], j + 1);
            }
            if (rest >= 3) mac(A, kq[2], j + 2);
            PIPE_T(w2);
            PIPE_ACC(1, w1, w2);
            if (warp == 0 && i + 1 == n_my) PIPE_STAMP(3);
            if (i + 2 < n_my) bar_arrive(kBarEmpty + b, all);         // planes[b] may take tile i + 2
            if (i >= 2) bar_sync(kBarLsEmpty + b, all);               // the helpers are done with ls[b] of tile i - 2
            PIPE_T(w3);
            PIPE_ACC(2, w2, w3);
            // |z|^2 (norm_sqr) and log_spec (vqt.rs:930) per lane
            float *ls = ls0 + b * ls_stride;
            const uint32_t frame0 = tile * kTileFrames;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (rows.y > r) {
#pragma unroll
                    for (int f = 0; f < kTileFrames; ++f) {
                        const float zr = (f & 1) ? re[r][f >> 1].y : re[r][f >> 1].x;
                        const float zi = (f & 1) ? im[r][f >> 1].y : im[r][f >> 1].x;
                        const float p = zr * zr + zi * zi;
                        ls[f * nb + rows.x + r] = log_spec(p, P.ref_db);
                        if (P.power != nullptr && frame0 + f < P.n_frames) P.power[(size_t)(frame0 + f) * nb + rows.x + r] = p;
                    }
                }
            }
            __threadfence_block();
            bar_arrive(kBarLsFull + b, all);
            PIPE_T(w4);
            PIPE_ACC(3, w3, w4);
        }
        return;
    }

    // ---------------------------------------------------------------------- helpers
    const int ht = threadIdx.x - nw * 32, hn = all - nw * 32;         // helper thread index / count
    const int hw = ht >> 5, hwarps = hn >> 5;
    // phase tables of the K-sdft groups: constant, staged once (plan data: no dependency on the kernels before)
    for (int gi = 0; gi < P.n_sdft; ++gi) {
        const SdftGroup &G = P.sdft[gi].g;
        float2 *ph = sd0 + (size_t)gi * sdft_floats2 + sd_c_cap(G.q, G.nk) + sd_r_cap(G.nk);   // behind the C and R rows
        for (int e = ht; e < (G.q + 1) * G.nk; e += hn) ph[e] = __ldg(G.phase + e);
    }
    PIPE_T(h_start);
    pdl_wait();   // the spectra and the partial sums below come from K-fft / K-sdft
    PIPE_T(h_go);
    PIPE_ACC(4, h_start, h_go);
    if (warp == nw) PIPE_STAMP(1);

    for (uint32_t i = 0; i <= n_my; ++i) {
        if (i < n_my) {
            const int b = (int)(i & 1);
            const uint32_t tile = blockIdx.x + i * gridDim.x;
            const uint32_t lf0 = tile * kTileFrames;
            float4 *planes = planes0 + b * (4 * PLANE);
            PIPE_T(h0);
            bar_sync(kBarHelpers, hn);                                // every helper is done with the staging buffers of tile i - 1
            if (i >= 2) bar_sync(kBarEmpty + b, all);                 // the walkers have left planes[b] (tile i - 2)
            PIPE_T(h1);
            PIPE_ACC(5, h0, h1);
            // ---- bulk copies of the tile: the spectrum planes, and per K-sdft group the chunk rows of its frames.
            // A group takes the staged path when the tile's frames lie in one stream (always, except for the one tile in
            // ~64 that crosses a stream boundary): C rows row0 .. row0 + nv + q - 2 and R rows row0 + q .. + nv - 1 are
            // contiguous in the partial-sum arrays.
            bool staged[kMaxSdft];
            uint32_t c_delta[kMaxSdft], r_delta[kMaxSdft];            // the rows begin this many float2 into their buffers
            uint32_t nv_g[kMaxSdft];
            uint32_t tx = (uint32_t)(4 * PLANE * sizeof(float4));
            const float2 *c_src[kMaxSdft], *r_src[kMaxSdft];
            uint32_t c_bytes[kMaxSdft], r_bytes[kMaxSdft];
#pragma unroll
            for (int gi = 0; gi < kMaxSdft; ++gi) {
                staged[gi] = false;
                nv_g[gi] = c_delta[gi] = r_delta[gi] = c_bytes[gi] = r_bytes[gi] = 0;
                c_src[gi] = r_src[gi] = nullptr;
                if (gi >= P.n_sdft) continue;
                const SdftParams &D = P.sdft[gi];
                const SdftGroup &G = D.g;
                const uint32_t total = D.n_streams * D.frames;
                const uint32_t nv = lf0 < total ? min((uint32_t)kTileFrames, total - lf0) : 0u;
                nv_g[gi] = nv;
                if (nv == 0) continue;
                const uint32_t st = lf0 / D.frames, t0 = lf0 - st * D.frames;
                if (t0 + nv > D.frames) continue;                     // crosses into the next stream: direct loads below
                const size_t row0 = (size_t)st * D.rows_per_stream + t0;
                const float2 *c = D.partial_c + row0 * G.nk, *r = D.partial_r + (row0 + G.q) * G.nk;
                c_delta[gi] = (uint32_t)((reinterpret_cast<uintptr_t>(c) >> 3) & 1);      // float2 entries to the 16-byte line
                r_delta[gi] = (uint32_t)((reinterpret_cast<uintptr_t>(r) >> 3) & 1);
                c_src[gi] = c - c_delta[gi];
                r_src[gi] = r - r_delta[gi];
                c_bytes[gi] = ((c_delta[gi] + (nv + G.q - 1) * G.nk) * 8 + 15) & ~15u;
                r_bytes[gi] = G.rem != 0 ? ((r_delta[gi] + nv * G.nk) * 8 + 15) & ~15u : 0u;
                tx += c_bytes[gi] + r_bytes[gi];
                staged[gi] = true;
            }
            if (ht == 0) {
                fence_proxy_async();                                  // generic reads of planes[b] / staging before the async writes
                mbar_expect_tx(&mbar[b], tx);
                bulk_g2s(planes, P.spec + (size_t)tile * (4 * PLANE * 4), (uint32_t)(4 * PLANE * sizeof(float4)), &mbar[b]);
#pragma unroll
                for (int gi = 0; gi < kMaxSdft; ++gi) {
                    if (!staged[gi]) continue;
                    const SdftGroup &G = P.sdft[gi].g;
                    float2 *cb = sd0 + (size_t)gi * sdft_floats2, *rb = cb + sd_c_cap(G.q, G.nk);
                    bulk_g2s(cb, c_src[gi], c_bytes[gi], &mbar[b]);
                    if (r_bytes[gi]) bulk_g2s(rb, r_src[gi], r_bytes[gi], &mbar[b]);
                }
            }
            PIPE_T(h2);
            PIPE_ACC(6, h1, h2);
            mbar_wait(&mbar[b], (i >> 1) & 1);
            PIPE_T(h2b);
            PIPE_ACC(8, h2, h2b);
            // ---- K-sdft combine for this tile, into the planes:
            // X_t[k] = sum_i phase[i][k] C[row(t) + i][k] (+ phase[q][k] R[row(t) + q][k]) -- the sums of sdft_dot, in its order
#pragma unroll
            for (int gi = 0; gi < kMaxSdft; ++gi) {
                if (gi >= P.n_sdft) continue;
                const SdftParams &D = P.sdft[gi];
                const SdftGroup &G = D.g;
                const int nk = G.nk, q = G.q;
                if (staged[gi]) {
                    // one item = one bin, four consecutive frames (half a tile): the q + 3 chunk values are read once
                    const float2 *cb = sd0 + (size_t)gi * sdft_floats2 + c_delta[gi];
                    const float2 *rb = sd0 + (size_t)gi * sdft_floats2 + sd_c_cap(q, nk) + r_delta[gi];
                    const float2 *ph = sd0 + (size_t)gi * sdft_floats2 + sd_c_cap(q, nk) + sd_r_cap(nk);
                    const int nv = (int)nv_g[gi];
                    for (int item = ht; item < 2 * nk; item += hn) {
                        const int half = item / nk, k = item - half * nk;
                        const int f0 = 4 * half;                       // first frame of the item inside the tile
                        float2 a[4][4];
#pragma unroll
                        for (int f = 0; f < 4; ++f)
#pragma unroll
                            for (int u = 0; u < 4; ++u) a[f][u] = make_float2(0.f, 0.f);
                        // rows beyond the tile's valid frames are not in the buffer: those frames are not computed
                        const int nf = max(0, min(4, nv - f0));
                        auto cval = [&](int r) { return r < nv + q - 1 ? cb[(size_t)r * nk + k] : make_float2(0.f, 0.f); };
                        auto step = [&](float2 &acc, float2 w, float2 v) {
                            acc = __ffma2_rn(make_float2(w.x, w.x), v, acc);
                            acc = __ffma2_rn(make_float2(-w.y, w.y), make_float2(v.y, v.x), acc);
                        };
                        int i4 = 0;
                        float2 c0 = cval(f0), c1 = cval(f0 + 1), c2 = cval(f0 + 2);     // sliding window of chunk values
                        for (; i4 + 4 <= q; i4 += 4) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float2 w = ph[(size_t)(i4 + u) * nk + k];
                                const float2 c3 = cval(f0 + i4 + u + 3);
                                step(a[0][u], w, c0);
                                step(a[1][u], w, c1);
                                step(a[2][u], w, c2);
                                step(a[3][u], w, c3);
                                c0 = c1; c1 = c2; c2 = c3;
                            }
                        }
                        for (; i4 < q; ++i4) {                        // sdft_dot's tail goes to the first partial sum
                            const float2 w = ph[(size_t)i4 * nk + k];
                            const float2 c3 = cval(f0 + i4 + 3);
                            step(a[0][0], w, c0);
                            step(a[1][0], w, c1);
                            step(a[2][0], w, c2);
                            step(a[3][0], w, c3);
                            c0 = c1; c1 = c2; c2 = c3;
                        }
                        float xr[4], xi[4];
#pragma unroll
                        for (int f = 0; f < 4; ++f) {
                            float2 x = __fadd2_rn(__fadd2_rn(a[f][0], a[f][1]), __fadd2_rn(a[f][2], a[f][3]));
                            if (G.rem != 0 && f < nf) x = __fadd2_rn(x, cmul(rb[(size_t)(f0 + f) * nk + k], ph[(size_t)q * nk + k]));
                            if (f >= nf) x = make_float2(0.f, 0.f);
                            xr[f] = x.x;
                            xi[f] = x.y;
                        }
                        planes[half * PLANE + G.spec_offset + k] = make_float4(xr[0], xr[1], xr[2], xr[3]);
                        planes[(2 + half) * PLANE + G.spec_offset + k] = make_float4(xi[0], xi[1], xi[2], xi[3]);
                    }
                } else {
                    // the tile crosses a stream boundary (or is empty): one item per frame and bin, straight from global memory
                    const uint32_t total = D.n_streams * D.frames;
                    const int items = kTileFrames * nk;
                    for (int item = ht; item < items; item += hn) {
                        const int fi = item / nk, k = item - fi * nk;
                        const uint32_t lf = lf0 + fi;
                        float2 x = make_float2(0.f, 0.f);
                        if (lf < total) {
                            const uint32_t st = lf / D.frames, t = lf - st * D.frames;
                            const size_t row = (size_t)st * D.rows_per_stream + t;
                            x = sdft_dot<true>(D.partial_c + row * nk + k, G.phase + k, q, nk);
                            if (G.rem != 0)
                                x = __fadd2_rn(x, cmul(__ldcg(D.partial_r + (row + q) * nk + k), __ldg(G.phase + q * nk + k)));
                        }
                        float *re_plane = reinterpret_cast<float *>(planes + (fi >> 2) * PLANE + G.spec_offset + k);
                        float *im_plane = reinterpret_cast<float *>(planes + (2 + (fi >> 2)) * PLANE + G.spec_offset + k);
                        re_plane[fi & 3] = x.x;
                        im_plane[fi & 3] = x.y;
                    }
                }
            }
            PIPE_T(h3);
            PIPE_ACC(7, h2b, h3);
            __threadfence_block();
            bar_arrive(kBarFull + b, all);
            if (warp == nw && i == 0) PIPE_STAMP(2);
        }
        if (i >= 1) {
            // power_to_db's frame-wise part (vqt.rs:933-950) for tile i - 1: one warp per frame, coalesced stores
            const int pb = (int)((i - 1) & 1);
            const uint32_t tile = blockIdx.x + (i - 1) * gridDim.x;
            const uint32_t frame0 = tile * kTileFrames;
            const float *lsb = ls0 + pb * ls_stride;
            PIPE_T(h5);
            bar_sync(kBarLsFull + pb, all);
            PIPE_T(h6);
            PIPE_ACC(9, h5, h6);
            for (int f = hw; f < kTileFrames; f += hwarps) {
                if (frame0 + f >= P.n_frames) break;
                const float *l = lsb + f * nb;
                float *out = P.out_db + (size_t)(frame0 + f) * nb;
                float mx = -CUDART_INF_F, mn = CUDART_INF_F;
                const bool vec = (nb & 3) == 0 && (reinterpret_cast<uintptr_t>(P.out_db) & 15) == 0;
                if (vec) {
                    const float4 *l4 = reinterpret_cast<const float4 *>(l);
                    const int n4 = nb >> 2;
                    for (int e = lane; e < n4; e += 32) {
                        const float4 v = l4[e];
                        mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
                        mn = fminf(mn, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
                    }
                } else {
                    for (int r = lane; r < nb; r += 32) {
                        mx = fmaxf(mx, l[r]);
                        mn = fminf(mn, l[r]);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                }
                const float floor_db = mx - kTopDb, log_spec_min = fmaxf(mn, floor_db);  // vqt.rs:939-940
                if (vec) {
                    const float4 *l4 = reinterpret_cast<const float4 *>(l);
                    float4 *o4 = reinterpret_cast<float4 *>(out);
                    const int n4 = nb >> 2;
                    for (int e = lane; e < n4; e += 32) {
                        const float4 v = l4[e];
                        o4[e] = make_float4(db_out(v.x, floor_db, log_spec_min), db_out(v.y, floor_db, log_spec_min),
                                            db_out(v.z, floor_db, log_spec_min), db_out(v.w, floor_db, log_spec_min));
                    }
                } else {
                    for (int r = lane; r < nb; r += 32) out[r] = db_out(l[r], floor_db, log_spec_min);
                }
            }
            if (i + 1 < n_my) bar_arrive(kBarLsEmpty + pb, all);      // ls[pb] may take tile i + 1
            PIPE_T(h7);
            PIPE_ACC(10, h6, h7);
        }
    }
    if (warp == nw) PIPE_STAMP(4);
}

}  // namespace

int pipe_plane_stride(int cols_touched) { return cols_touched <= 832 ? 832 : 0; }

// float2 entries of one K-sdft group's staging: (q + 8) C rows, 8 R rows, (q + 1) phase rows, each part 16-byte aligned
size_t pipe_sdft_floats2(int q, int nk) { return ((size_t)sd_c_cap(q, nk) + sd_r_cap(nk) + (size_t)(q + 1) * nk + 1) & ~(size_t)1; }

size_t pipe_smem_bytes(int cols_touched, int n_buckets, int n_warps, int sdft_floats2, int n_sdft)
{
    const size_t planes = (size_t)2 * 4 * pipe_plane_stride(cols_touched) * sizeof(float4);
    const size_t ls = (size_t)2 * (((size_t)kTileFrames * n_buckets + 3) & ~(size_t)3) * sizeof(float);
    (void)n_warps;
    return planes + ls + (size_t)std::max(0, std::min(n_sdft, kMaxSdft)) * sdft_floats2 * sizeof(float2);
}

bool pipe_supported(int n_warps, int cols_touched, int n_buckets, int rows_per_lane, int min_slots)
{
    return rows_per_lane == 2 && n_warps >= 1 && (n_warps + kPipeHelpers) * 32 <= kPipeMaxThreads &&
           pipe_plane_stride(cols_touched) > 0 && min_slots >= 1 &&
           pipe_smem_bytes(cols_touched, n_buckets, n_warps, 0, 0) <= 200 * 1024;
}

cudaError_t configure_pipe(int n_warps, int cols_touched, int n_buckets, int sdft_floats2, int n_sdft)
{
    const size_t bytes = pipe_smem_bytes(cols_touched, n_buckets, n_warps, sdft_floats2, n_sdft);
    if (bytes > 227 * 1024) return cudaErrorInvalidConfiguration;
    return cudaFuncSetAttribute(spmm_db_pipe_kernel<832>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_spmm_db_pipe(const FusedParams &p, int n_ctas, int sdft_floats2, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)std::max(1, std::min<int>(n_ctas, (int)p.n_tiles)));
    cfg.blockDim = dim3((unsigned)(p.n_warps + kPipeHelpers) * 32);
    cfg.dynamicSmemBytes = pipe_smem_bytes(p.cols_touched, p.n_buckets, p.n_warps, sdft_floats2, p.n_sdft);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, spmm_db_pipe_kernel<832>, p, sdft_floats2);
}

}  // namespace pvqt_dev

#ifdef PVQT_PIPE_STATS
extern "C" int pvqt_debug_pipe_stamps(unsigned long long *out)
{
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, pvqt_dev::g_pipe_stamps, sizeof(pvqt_dev::g_pipe_stamps)) == cudaSuccess ? 0 : 7;
}

extern "C" int pvqt_debug_pipe_stats(long long *out, int reset)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, pvqt_dev::g_pipe_stats, sizeof(pvqt_dev::g_pipe_stats)) != cudaSuccess) return 7;
    if (reset) {
        static long long zeros[256][16] = {};
        cudaMemcpyToSymbol(pvqt_dev::g_pipe_stats, zeros, sizeof(zeros));
    }
    return 0;
}
#endif

// sdft_tc_kernel.cu -- K-sdft partial sums on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same result as sdft_partial_mma_kernel (sdft_kernels.cu): C[row][k] and R[row][k] of one window group, the chunk
// twiddle split in two levels e^{-2 pi i k (16 a + b) / N} = A[a][k] B[b][k].  The inner level is a dense real GEMM
//   S_a[row][col] = sum_{b < 16} x[row][16 a + b] * Bt[col][b],      col = 2 * bin + {0: re, 1: im}
// issued here as tcgen05.mma.kind::tf32, M = 128 chunk rows x N = 2 * BINS columns x K = 8 per instruction, with the
// 3xTF32 split (x = hi + lo, B = hi + lo; lo.hi + hi.lo + hi.hi, f32 accumulate in TMEM): six MMAs per 16-sample
// block.  The outer level, acc += A[a][k] S_a (one complex FMA per bin and block), stays on the FP32 pipe with the
// running sums in registers: one thread per (row, 32 bins), read back from TMEM with tcgen05.ld.32x32b.
//
// Shared-memory operands use the no-swizzle K-major canonical layout (8 rows x 16 bytes core matrices):
//   element (row r, sample b) of a 16-sample block at float index ((b / 4) * 128 + r) * 4 + b % 4
// so that one MMA's K = 8 slice is two 2048-byte planes (leading byte offset 2048, stride byte offset 128).
// Pipeline per CTA (all warps in lock step, the tensor core asynchronous behind an mbarrier):
//   cp.async raw block a+3 | split block a into hi / lo operand planes | one thread issues MMA(a) -> TMEM buffer a % 2
//   | every thread folds S_{a-1} (TMEM buffer (a-1) % 2) into its running sums while MMA(a) runs.
// Selected for groups whose remainder `rem` is a multiple of 16 (R is then a snapshot of the running sum).
#include "device_helpers.cuh"
#include "vqt_device.cuh"

namespace pvqt_dev {
__device__ unsigned long long g_tc_debug[40];
namespace {

constexpr int kTcRows = 128;
constexpr int kTcRawStride = 20;   // floats per row of a raw block: LDS.128 of 8 consecutive rows hit 32 banks
constexpr int kTcRawRing = 4;
constexpr int kTcRawBytes = kTcRows * kTcRawStride * 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async4_zfill_tc(void *smem_dst, const void *gmem_src, unsigned src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// Bounded: a lost completion ends the kernel with wrong numbers (caught by the parity tests) instead of a hang.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spin = 0; spin < (1u << 22); ++spin)
        if (mbar_try_wait(bar, parity)) break;
    __syncwarp();
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// Shared-memory matrix descriptor: K-major, 64-byte swizzle (one 16-sample block = one 64-byte row; the 16-byte chunks
// of row r XOR-ed with (r / 2) % 4), 8-row groups 512 bytes apart, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(512u >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
}
// byte offset of 16-byte chunk `kq` of row `r` inside an operand tile
__device__ __forceinline__ int sw64_offset(int r, int kq) { return r * 64 + ((kq ^ ((r >> 1) & 3)) << 4); }
// Instruction descriptor: D f32, A / B tf32, both K-major, M = 128, N = n.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive accumulator columns of the thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t tf32_rna(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kTcBins = 64;                 // bins per CTA tile: MMA N = 128 accumulator columns
constexpr int kTcCols = 2 * kTcBins;
constexpr int kTcSlots = 4;                 // A-operand ring (hi + lo planes per slot)
constexpr int kTcAhead = 2;                 // blocks of audio a loader thread keeps in flight
constexpr int kTcLoaderWarps = 4;           // thread = chunk row: loads, splits hi / lo, fills the ring
constexpr int kTcEpiWarps = 8;              // thread = (chunk row, 32 bins): folds S_a into the running sums
constexpr int kTcThreads = (kTcLoaderWarps + 1 + kTcEpiWarps) * 32;
constexpr int kTcPlaneA = 128 * 16;         // bytes of one k-quad plane of the A operand
constexpr int kTcPlaneB = kTcCols * 16;
constexpr int kTcOperandA = 4 * kTcPlaneA, kTcOperandB = 4 * kTcPlaneB;
constexpr int kTcBufs = 4;                  // TMEM accumulator buffers
constexpr uint32_t kTcTmemCols = kTcBufs * kTcCols;

struct TcBarriers {
    uint64_t full[kTcSlots];    // loaders -> MMA: slot holds block a            (one arrival per loader warp)
    uint64_t empty[kTcSlots];   // MMA -> loaders: the MMAs reading the slot are done (tcgen05.commit)
    uint64_t tfull[kTcBufs];          // MMA -> epilogue: accumulator buffer holds S_a  (tcgen05.commit)
    uint64_t tempty[kTcBufs];         // epilogue -> MMA: buffer read back              (one arrival per epilogue warp)
};

// Warp-specialised: warps 0-3 load + split, warp 4 issues the MMAs, warps 5-12 fold.  A wait that times out
// raises `abort_flag`; every later wait then falls through, so a protocol bug ends in wrong numbers, not a hang.
__global__ void __launch_bounds__(kTcThreads, 1) sdft_partial_tc_kernel(const __grid_constant__ SdftParams P)
{
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) TcBarriers bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int abort_flag;

    unsigned char *b_hi = tc_smem, *b_lo = b_hi + kTcOperandB;
    unsigned char *a_op = b_lo + kTcOperandB;                                    // [slot][hi, lo]
    float2 *tw_a_s = reinterpret_cast<float2 *>(a_op + kTcSlots * 2 * kTcOperandA);   // [n_blocks][kTcBins]

    const SdftGroup &G = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bin0 = blockIdx.y * kTcBins;
    const uint32_t total_rows = P.n_streams * P.rows_per_stream;
    const uint32_t row0 = blockIdx.x * kTcRows;
    const int nb = G.n_blocks;

    pdl_launch_dependents();   // K-fft of the other groups may run beside this kernel
    const long long t_start = clock64();
    const bool dbg_cta = blockIdx.x == 1 && blockIdx.y == 0;
    long long d_wait0 = 0, d_wait1 = 0, d_work = 0;

    auto wait = [&](uint64_t *bar, uint32_t parity) {
        if (abort_flag) return;
        for (uint32_t spin = 0; spin < (1u << 16); ++spin)
            if (mbar_try_wait(bar, parity)) return;
        abort_flag = 1;
    };

    if (warp == kTcLoaderWarps) tmem_alloc(&tmem_base_s, kTcTmemCols);
    if (tid == 0) {
        abort_flag = 0;
        for (int i = 0; i < kTcSlots; ++i) {
            mbar_init(&bars.full[i], kTcLoaderWarps);
            mbar_init(&bars.empty[i], 1);
        }
        for (int i = 0; i < kTcBufs; ++i) {
            mbar_init(&bars.tfull[i], 1);
            mbar_init(&bars.tempty[i], kTcEpiWarps);
        }
        fence_mbar_init();
    }
    // B operand (inner twiddles, split hi / lo) and the outer twiddles of this CTA's bins
    for (int idx = tid; idx < 16 * kTcCols; idx += kTcThreads) {
        const int b = idx / kTcCols, n = idx - b * kTcCols, bin = bin0 + (n >> 1);
        float v = 0.f;
        if (bin < G.nk) {
            const float2 w = __ldg(G.tw_b + b * G.nk + bin);
            v = (n & 1) ? w.y : w.x;
        }
        const uint32_t hi = tf32_rna(v), lo = tf32_rna(v - __uint_as_float(hi));
        const int off = sw64_offset(n, b >> 2) + (b & 3) * 4;
        *reinterpret_cast<uint32_t *>(b_hi + off) = hi;
        *reinterpret_cast<uint32_t *>(b_lo + off) = lo;
    }
    for (int idx = tid; idx < nb * kTcBins; idx += kTcThreads) {
        const int a = idx / kTcBins, j = idx - a * kTcBins;
        tw_a_s[idx] = bin0 + j < G.nk ? __ldg(G.tw_a + a * G.nk + bin0 + j) : make_float2(0.f, 0.f);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const long long t_pro = clock64();

    if (warp < kTcLoaderWarps) {
        // ---------------- loaders: thread = chunk row ----------------
        const int r = tid;
        const uint32_t row = row0 + r;
        const float *src = P.audio;
        int valid = 0;
        if (row < total_rows) {
            const uint32_t s = row / P.rows_per_stream;
            const uint32_t c = P.first_frame + (row - s * P.rows_per_stream);
            const uint64_t base = (uint64_t)c * G.hop + G.window_begin;
            src = P.audio + (uint64_t)(P.first_stream + s) * P.stream_stride + base;
            valid = base < P.valid_samples ? (int)min((uint64_t)G.hop, P.valid_samples - base) : 0;
        }
        // the row as aligned 16-byte vectors: vector v holds row elements 4 v - sh .. 4 v - sh + 3
        const int sh = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
        const float4 *p4 = reinterpret_cast<const float4 *>(src - sh);
        auto ldv = [&](int v) {
            const int e0 = 4 * v - sh;
            if (e0 >= 0 && e0 + 3 < valid) return __ldg(p4 + v);
            float4 o;
            o.x = (e0 >= 0 && e0 < valid) ? __ldg(src + e0) : 0.f;
            o.y = (e0 + 1 >= 0 && e0 + 1 < valid) ? __ldg(src + e0 + 1) : 0.f;
            o.z = (e0 + 2 >= 0 && e0 + 2 < valid) ? __ldg(src + e0 + 2) : 0.f;
            o.w = (e0 + 3 >= 0 && e0 + 3 < valid) ? __ldg(src + e0 + 3) : 0.f;
            return o;
        };
        // register queue: the vectors of the next kTcAhead blocks are in flight
        float4 carry = ldv(0), q[kTcAhead][4];
#pragma unroll
        for (int u = 0; u < kTcAhead; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) q[u][i] = u < nb ? ldv(4 * u + 1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        // the 4 row elements starting `sh` floats into the aligned pair (lo, hi): two levels of selects
        const bool s1 = sh & 1, s2 = sh & 2;
        auto shifted = [&](const float4 &lo, const float4 &hi) {
            const float a0 = s2 ? lo.z : lo.x, a1 = s2 ? lo.w : lo.y, a2 = s2 ? hi.x : lo.z, a3 = s2 ? hi.y : lo.w,
                        a4 = s2 ? hi.z : hi.x;
            return make_float4(s1 ? a1 : a0, s1 ? a2 : a1, s1 ? a3 : a2, s1 ? a4 : a3);
        };
        for (int a0 = 0; a0 < nb; a0 += kTcAhead) {
#pragma unroll
        for (int u = 0; u < kTcAhead; ++u) {
            const int a = a0 + u;
            if (a >= nb) break;
            float4 x[4];
            x[0] = shifted(carry, q[u][0]);
#pragma unroll
            for (int i = 1; i < 4; ++i) x[i] = shifted(q[u][i - 1], q[u][i]);
            carry = q[u][3];
            if (a + kTcAhead < nb) {
#pragma unroll
                for (int i = 0; i < 4; ++i) q[u][i] = ldv(4 * (a + kTcAhead) + 1 + i);
            }
            const int slot = a % kTcSlots;
            long long c0 = clock64();
            if (a >= kTcSlots) wait(&bars.empty[slot], (uint32_t)(a / kTcSlots - 1) & 1u);
            d_wait0 += clock64() - c0;
            unsigned char *hi_p = a_op + slot * 2 * kTcOperandA, *lo_p = hi_p + kTcOperandA;
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {
                uint4 h, l;
                h.x = tf32_rna(x[kq].x); l.x = tf32_rna(x[kq].x - __uint_as_float(h.x));
                h.y = tf32_rna(x[kq].y); l.y = tf32_rna(x[kq].y - __uint_as_float(h.y));
                h.z = tf32_rna(x[kq].z); l.z = tf32_rna(x[kq].z - __uint_as_float(h.z));
                h.w = tf32_rna(x[kq].w); l.w = tf32_rna(x[kq].w - __uint_as_float(h.w));
                *reinterpret_cast<uint4 *>(hi_p + sw64_offset(r, kq)) = h;
                *reinterpret_cast<uint4 *>(lo_p + sw64_offset(r, kq)) = l;
            }
            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.full[slot]);
        }
        }
        if (dbg_cta && tid == 0) { g_tc_debug[0] = d_wait0; g_tc_debug[1] = clock64() - t_pro; g_tc_debug[9] = t_pro - t_start; }
    } else if (warp == kTcLoaderWarps) {
        // ---------------- MMA issue: one thread ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_tf32(kTcCols);
            const uint32_t bh0 = smem_u32(b_hi), bl0 = smem_u32(b_lo);
            for (int a = 0; a < nb; ++a) {
                const int slot = a % kTcSlots, buf = a % kTcBufs;
                long long c0 = clock64();
                wait(&bars.full[slot], (uint32_t)(a / kTcSlots) & 1u);
                d_wait0 += clock64() - c0; c0 = clock64();
                if (a >= kTcBufs) wait(&bars.tempty[buf], (uint32_t)(a / kTcBufs - 1) & 1u);
                d_wait1 += clock64() - c0; c0 = clock64();
                tc_fence_after();
                const uint32_t d = tmem_base + buf * kTcCols;
                const uint32_t ah0 = smem_u32(a_op + slot * 2 * kTcOperandA), al0 = ah0 + kTcOperandA;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t ah = smem_desc(ah0 + h * 32), al = smem_desc(al0 + h * 32);   // K = 8 samples = 32 bytes
                    const uint64_t bh = smem_desc(bh0 + h * 32), bl = smem_desc(bl0 + h * 32);
                    mma_tf32_ss(d, al, bh, idesc, h);   // small terms first; h = 0 overwrites the buffer
                    mma_tf32_ss(d, ah, bl, idesc, 1);
                    mma_tf32_ss(d, ah, bh, idesc, 1);
                }
                mma_commit(&bars.empty[slot]);
                mma_commit(&bars.tfull[buf]);
                d_work += clock64() - c0;
            }
            if (dbg_cta) { g_tc_debug[2] = d_wait0; g_tc_debug[3] = d_wait1; g_tc_debug[4] = d_work; g_tc_debug[5] = clock64() - t_pro; }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: thread = (chunk row, 32 bins) ----------------
        const int e = warp - (kTcLoaderWarps + 1);
        const int quarter = warp & 3;                 // the TMEM lanes a warp may touch: 32 (warp % 4) .. + 31
        const int my_bin = 32 * (e >> 2);             // first of the thread's 32 bins inside the CTA tile
        const uint32_t my_row = row0 + 32 * quarter + lane;
        const uint32_t my_taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + 2 * my_bin;
        const int ra = G.rem >> 4;
        float2 acc[32];
        const unsigned long long xmode = g_tc_debug[15];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = make_float2(0.f, 0.f);
        auto store = [&](float2 *dst) {
            if (my_row < total_rows) {
                float2 *o = dst + (size_t)my_row * G.nk + bin0 + my_bin;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (bin0 + my_bin + j < G.nk) o[j] = acc[j];
            }
        };
        for (int a = 0; a < nb; ++a) {
            const int buf = a % kTcBufs;
            long long c0 = clock64();
            wait(&bars.tfull[buf], (uint32_t)(a / kTcBufs) & 1u);
            d_wait0 += clock64() - c0;
            __syncwarp();
            tc_fence_after();
            const float2 *A = tw_a_s + a * kTcBins + my_bin;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float s[32];
                if (!(xmode & 1)) tmem_ld32(my_taddr + buf * kTcCols + 32 * half, s);
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) s[j] = (float)a;
                }
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.tempty[buf]);   // both halves are in registers: the buffer may be overwritten
                }
                if (!(xmode & 2))
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 w = A[16 * half + j];
                    float2 t = __ffma2_rn(make_float2(w.x, w.x), make_float2(s[2 * j], s[2 * j + 1]), acc[16 * half + j]);
                    acc[16 * half + j] = __ffma2_rn(make_float2(-w.y, w.y), make_float2(s[2 * j + 1], s[2 * j]), t);
                }
            }
            if (a + 1 == ra && G.rem != 0) store(P.partial_r);   // rem = 16 ra: R is the running sum after ra blocks
        }
        const long long c1 = clock64();
        store(P.partial_c);
        if (dbg_cta && e == 0 && lane == 0) { g_tc_debug[6] = d_wait0; g_tc_debug[7] = c1 - t_pro; g_tc_debug[8] = clock64() - c1; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTcLoaderWarps) tmem_dealloc(tmem_base, kTcTmemCols);
}

size_t tc_smem_bytes(int n_blocks)
{
    return (size_t)2 * kTcOperandB + (size_t)kTcSlots * 2 * kTcOperandA + (size_t)n_blocks * kTcBins * sizeof(float2);
}

}  // namespace

bool sdft_tc_supported(const SdftGroup &g)
{
    return g.rem % 16 == 0 && g.hop_pad == g.n_blocks * 16 && tc_smem_bytes(g.n_blocks) <= 200 * 1024;
}

cudaError_t configure_sdft_tc(int n_blocks)
{
    static int configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const int want = (int)tc_smem_bytes(n_blocks);
    if (want > 200 * 1024 || dev < 0 || dev >= 64) return cudaErrorInvalidConfiguration;
    if (want <= configured[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(sdft_partial_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e == cudaSuccess) configured[dev] = want;
    return e;
}

void sdft_tc_debug_set(unsigned long long mode) { cudaMemcpyToSymbol(g_tc_debug, &mode, sizeof(mode), 15 * sizeof(unsigned long long)); }
void sdft_tc_debug(unsigned long long *out) { cudaMemcpyFromSymbol(out, g_tc_debug, sizeof(unsigned long long) * 40); }

cudaError_t launch_sdft_partial_tc(const SdftParams &p, cudaStream_t stream)
{
    const uint32_t rows = p.n_streams * p.rows_per_stream;
    const dim3 grid((rows + kTcRows - 1) / kTcRows, (p.g.nk + kTcBins - 1) / kTcBins);
    sdft_partial_tc_kernel<<<grid, kTcThreads, tc_smem_bytes(p.g.n_blocks), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace pvqt_dev

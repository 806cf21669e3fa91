// sdft_tc_kernel.cu -- K-sdft partial sums on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same result as sdft_partial_mma_kernel (sdft_kernels.cu): C[row][k] and R[row][k] of one window group, the chunk
// twiddle split in two levels: the chunk is cut into accumulation groups of up to G samples (G = 16, 32 or 64; a
// cut at `rem`), e^{-2 pi i k (s_g + b) / N} = A[g][k] B[b][k] for sample b of the group starting at s_g.  The inner
// level is a dense real GEMM
//   S_g[row][col] = sum_{b < len_g} x[row][s_g + b] * Bt[col][b]
// issued as tcgen05.mma.kind::tf32, M = 128 chunk rows x N = 128 columns x K = 8 per instruction, with the 3xTF32
// split (x = hi + lo, B = hi + lo; lo.hi + hi.lo + hi.hi, f32 accumulate in TMEM): six MMAs per 16 samples.  The outer
// level, acc += A[g][k] S_g (one complex FMA per bin and group), stays on the FP32 pipe with the running sums in
// registers: one thread per (row, 32 bins), S_g read back from TMEM with tcgen05.ld.32x32b.  Columns come in
// bin pairs, (re b0, re b1, im b0, im b1), so that the packed FFMA2s of the outer level need no register swaps.
//
// Shared-memory operands: K-major, 64-byte swizzle -- one 16-sample block of a row is one 64-byte line.
// Warp-specialised pipeline per CTA, every hand-over an mbarrier:
//   loaders (thread = chunk row): aligned 16-byte loads, hi / lo split, fill a 4-slot operand ring
//   one MMA thread: 6 MMAs per 16-sample block into one of 4 TMEM accumulators; tcgen05.commit frees the slot and,
//                   after the group's last block, publishes the accumulator
//   fold warps: tcgen05.ld, release the accumulator, acc += A[g] S_g
// Selected for groups whose remainder `rem` is a multiple of 16 (R is then a snapshot of the running sum).
#include "device_helpers.cuh"
#include "vqt_device.cuh"

namespace pvqt_dev {
namespace {

constexpr int kTcRows = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

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

// A 16-byte vector of a chunk row that reaches past the samples the stream holds: element-wise, zero filled.
__device__ __noinline__ float4 ldv_tail(const float *src, int e0, int valid)
{
    float4 o;
    o.x = (e0 >= 0 && e0 < valid) ? __ldg(src + e0) : 0.f;
    o.y = (e0 + 1 >= 0 && e0 + 1 < valid) ? __ldg(src + e0 + 1) : 0.f;
    o.z = (e0 + 2 >= 0 && e0 + 2 < valid) ? __ldg(src + e0 + 2) : 0.f;
    o.w = (e0 + 3 >= 0 && e0 + 3 < valid) ? __ldg(src + e0 + 3) : 0.f;
    return o;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kTcBins = 64;                 // bins per CTA tile: MMA N = 128 accumulator columns
constexpr int kTcCols = 2 * kTcBins;
constexpr int kTcSlots = 4;                 // A-operand ring (hi + lo tile per slot)
constexpr int kTcAhead = 2;                 // blocks of audio a loader thread keeps in flight
constexpr int kTcLoaderWarps = 4;           // thread = chunk row: loads, splits hi / lo, fills the ring
constexpr int kTcEpiWarps = 8;              // thread = (chunk row, 32 bins): folds S_g into the running sums
constexpr int kTcThreads = (kTcLoaderWarps + 1 + kTcEpiWarps) * 32;
constexpr int kTcOperandA = 128 * 64;       // bytes of one operand tile: 128 rows x 16 samples
constexpr int kTcOperandB = kTcCols * 64;
constexpr int kTcBufs = 4;                  // TMEM accumulator buffers
constexpr uint32_t kTcTmemCols = kTcBufs * kTcCols;

struct TcBarriers {
    uint64_t full[kTcSlots];    // loaders -> MMA: slot holds block a            (one arrival per loader warp)
    uint64_t empty[kTcSlots];   // MMA -> loaders: the MMAs reading the slot are done (tcgen05.commit)
    uint64_t tfull[kTcBufs];    // MMA -> fold: accumulator holds S_g             (tcgen05.commit)
    uint64_t tempty[kTcBufs];   // fold -> MMA: accumulator read back             (one arrival per fold warp)
};

// Every wait is bounded (2^26 polls, seconds: far beyond any preemption of a healthy device).  A wait that does run
// out raises `abort_flag` and every later wait falls through, so a protocol bug ends in wrong numbers -- which the
// parity tests catch -- instead of a hung device.
__global__ void __launch_bounds__(kTcThreads, 1) sdft_partial_tc_kernel(const __grid_constant__ SdftParams P,
                                                                         const __grid_constant__ SdftTcPlan T)
{
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) TcBarriers bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int abort_flag;

    const int g16 = T.group16;                                                   // 16-sample blocks per full group
    unsigned char *b_op = tc_smem;                                               // [sub-block][hi, lo]
    unsigned char *a_op = b_op + g16 * 2 * kTcOperandB;                          // [slot][hi, lo]
    float4 *tw_g_s = reinterpret_cast<float4 *>(a_op + kTcSlots * 2 * kTcOperandA);   // [group][bin pair]: (Ar0, Ar1, Ai0, Ai1)

    const SdftGroup &G = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bin0 = blockIdx.y * kTcBins;
    const uint32_t total_rows = P.n_streams * P.rows_per_stream;
    const uint32_t row0 = blockIdx.x * kTcRows;
    const int nb = G.n_blocks;

    pdl_launch_dependents();   // K-fft of the other groups may run beside this kernel

    auto wait = [&](uint64_t *bar, uint32_t parity) {
        if (abort_flag) return;
#pragma unroll 1
        for (uint32_t spin = 0; spin < (1u << 26); ++spin)
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
    // B operand: inner twiddles of a full group, split hi / lo.  Column n of bin pair p = n / 4: re / im of bins
    // 2p, 2p + 1 in the order (re0, re1, im0, im1).
    for (int idx = tid; idx < 16 * g16 * kTcCols; idx += kTcThreads) {
        const int b = idx / kTcCols, n = idx - b * kTcCols, bin = bin0 + 2 * (n >> 2) + (n & 1);
        float v = 0.f;
        if (bin < G.nk) {
            const float2 w = __ldg(T.tw_b + b * G.nk + bin);
            v = (n & 2) ? w.y : w.x;
        }
        const uint32_t hi = tf32_rna(v), lo = tf32_rna(v - __uint_as_float(hi));
        const int off = (b >> 4) * 2 * kTcOperandB + sw64_offset(n, (b >> 2) & 3) + (b & 3) * 4;
        *reinterpret_cast<uint32_t *>(b_op + off) = hi;
        *reinterpret_cast<uint32_t *>(b_op + off + kTcOperandB) = lo;
    }
    for (int idx = tid; idx < T.n_groups * (kTcBins / 2); idx += kTcThreads) {
        const int g = idx / (kTcBins / 2), p = idx - g * (kTcBins / 2), bin = bin0 + 2 * p;
        const float2 w0 = bin < G.nk ? __ldg(T.tw_g + g * G.nk + bin) : make_float2(0.f, 0.f);
        const float2 w1 = bin + 1 < G.nk ? __ldg(T.tw_g + g * G.nk + bin + 1) : make_float2(0.f, 0.f);
        tw_g_s[idx] = make_float4(w0.x, w1.x, w0.y, w1.y);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

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
        // the row as aligned 16-byte vectors: vector v holds row elements 4 v - sh .. 4 v - sh + 3.  Elements before
        // the row (v = 0, sh > 0) are read and dropped: they lie inside the same allocation.
        const int sh = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
        const float4 *p4 = reinterpret_cast<const float4 *>(src - sh);
        auto ldv = [&](int v) {
            const int e0 = 4 * v - sh;
            if (e0 + 3 < valid) return __ldg(p4 + v);
            return ldv_tail(src, e0, valid);
        };
        // the 4 row elements starting `sh` floats into the aligned pair (lo, hi): two levels of selects
        const bool s1 = sh & 1, s2 = sh & 2;
        auto shifted = [&](const float4 &lo, const float4 &hi) {
            const float a0 = s2 ? lo.z : lo.x, a1 = s2 ? lo.w : lo.y, a2 = s2 ? hi.x : lo.z, a3 = s2 ? hi.y : lo.w,
                        a4 = s2 ? hi.z : hi.x;
            return make_float4(s1 ? a1 : a0, s1 ? a2 : a1, s1 ? a3 : a2, s1 ? a4 : a3);
        };
        // register queue: the vectors of the next kTcAhead blocks are in flight
        float4 carry = ldv(0), q[kTcAhead][4];
#pragma unroll
        for (int u = 0; u < kTcAhead; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) q[u][i] = u < nb ? ldv(4 * u + 1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
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
                if (a >= kTcSlots) wait(&bars.empty[slot], (uint32_t)(a / kTcSlots - 1) & 1u);
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
    } else if (warp == kTcLoaderWarps) {
        // ---------------- MMA issue: one thread ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_tf32(kTcCols);
            const uint32_t b0 = smem_u32(b_op);
            for (int g = 0; g < T.n_groups; ++g) {
                const int buf = g % kTcBufs, len = T.len16[g];
                if (g >= kTcBufs) wait(&bars.tempty[buf], (uint32_t)(g / kTcBufs - 1) & 1u);
                const uint32_t d = tmem_base + buf * kTcCols;
                for (int j = 0; j < len; ++j) {
                    const int a = T.start16[g] + j, slot = a % kTcSlots;
                    wait(&bars.full[slot], (uint32_t)(a / kTcSlots) & 1u);
                    tc_fence_after();
                    const uint32_t ah0 = smem_u32(a_op + slot * 2 * kTcOperandA), al0 = ah0 + kTcOperandA;
                    const uint32_t bh0 = b0 + j * 2 * kTcOperandB, bl0 = bh0 + kTcOperandB;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {   // K = 8 samples = 32 bytes along the swizzled line
                        const uint64_t ah = smem_desc(ah0 + h * 32), al = smem_desc(al0 + h * 32);
                        const uint64_t bh = smem_desc(bh0 + h * 32), bl = smem_desc(bl0 + h * 32);
                        mma_tf32_ss(d, al, bh, idesc, (j | h) != 0);   // small terms first; the first MMA overwrites
                        mma_tf32_ss(d, ah, bl, idesc, 1);
                        mma_tf32_ss(d, ah, bh, idesc, 1);
                    }
                    mma_commit(&bars.empty[slot]);
                }
                mma_commit(&bars.tfull[buf]);
            }
        }
        __syncwarp();
    } else {
        // ---------------- fold: thread = (chunk row, 32 bins = 16 bin pairs) ----------------
        const int e = warp - (kTcLoaderWarps + 1);
        const int quarter = warp & 3;                 // the TMEM lanes a warp may touch: 32 (warp % 4) .. + 31
        const int my_pair = 16 * (e >> 2);            // first of the thread's bin pairs inside the CTA tile
        const uint32_t my_row = row0 + 32 * quarter + lane;
        const uint32_t my_taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + 4 * my_pair;
        float2 re[16], im[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) re[j] = im[j] = make_float2(0.f, 0.f);
        auto store = [&](float2 *dst) {
            if (my_row < total_rows) {
                float2 *o = dst + (size_t)my_row * G.nk + bin0 + 2 * my_pair;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (bin0 + 2 * (my_pair + j) < G.nk) o[2 * j] = make_float2(re[j].x, im[j].x);
                    if (bin0 + 2 * (my_pair + j) + 1 < G.nk) o[2 * j + 1] = make_float2(re[j].y, im[j].y);
                }
            }
        };
        for (int g = 0; g < T.n_groups; ++g) {
            const int buf = g % kTcBufs;
            wait(&bars.tfull[buf], (uint32_t)(g / kTcBufs) & 1u);
            __syncwarp();
            tc_fence_after();
            const float4 *A = tw_g_s + g * (kTcBins / 2) + my_pair;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float s[32];
                tmem_ld32(my_taddr + buf * kTcCols + 32 * half, s);
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.tempty[buf]);   // everything is in registers: the buffer may be overwritten
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w = A[8 * half + j];
                    const float2 ar = make_float2(w.x, w.y), ai = make_float2(w.z, w.w);
                    const float2 sr = make_float2(s[4 * j], s[4 * j + 1]), si = make_float2(s[4 * j + 2], s[4 * j + 3]);
                    float2 &cr = re[8 * half + j], &ci = im[8 * half + j];
                    cr = __ffma2_rn(ar, sr, cr);
                    cr = __ffma2_rn(make_float2(-ai.x, -ai.y), si, cr);
                    ci = __ffma2_rn(ar, si, ci);
                    ci = __ffma2_rn(ai, sr, ci);
                }
            }
            if (g + 1 == T.r_groups) store(P.partial_r);   // the groups so far cover exactly the first `rem` samples
        }
        store(P.partial_c);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTcLoaderWarps) tmem_dealloc(tmem_base, kTcTmemCols);
}

size_t tc_smem_bytes(int group16, int n_groups)
{
    return (size_t)group16 * 2 * kTcOperandB + (size_t)kTcSlots * 2 * kTcOperandA + (size_t)n_groups * (kTcBins / 2) * sizeof(float4);
}

}  // namespace

bool sdft_tc_supported(const SdftGroup &g)
{
    return g.rem % 16 == 0 && g.hop_pad == g.n_blocks * 16 && g.n_blocks <= kTcMaxGroups;
}

// Accumulation groups of up to 16 * group16 samples, cut at `rem`.
void sdft_tc_make_plan(const SdftGroup &g, int group16, SdftTcPlan *t)
{
    *t = SdftTcPlan{};
    t->group16 = group16;
    const int r16 = g.rem / 16;
    auto cut = [&](int a, int b) {
        for (int s = a; s < b; s += group16) {
            t->start16[t->n_groups] = (uint8_t)s;
            t->len16[t->n_groups] = (uint8_t)std::min(group16, b - s);
            ++t->n_groups;
        }
    };
    cut(0, r16);
    t->r_groups = t->n_groups;
    cut(r16, g.n_blocks);
}

cudaError_t configure_sdft_tc(const SdftTcPlan &t)
{
    static int configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const int want = (int)tc_smem_bytes(t.group16, t.n_groups);
    if (want > 200 * 1024 || dev < 0 || dev >= 64) return cudaErrorInvalidConfiguration;
    if (want <= configured[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(sdft_partial_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e == cudaSuccess) configured[dev] = want;
    return e;
}

cudaError_t launch_sdft_partial_tc(const SdftParams &p, const SdftTcPlan &t, cudaStream_t stream)
{
    const uint32_t rows = p.n_streams * p.rows_per_stream;
    const dim3 grid((rows + kTcRows - 1) / kTcRows, (p.g.nk + kTcBins - 1) / kTcBins);
    sdft_partial_tc_kernel<<<grid, kTcThreads, tc_smem_bytes(t.group16, t.n_groups), stream>>>(p, t);
    return cudaGetLastError();
}

}  // namespace pvqt_dev

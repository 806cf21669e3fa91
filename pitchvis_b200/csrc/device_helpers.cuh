// device_helpers.cuh -- small device-side helpers shared by the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace pvqt_dev {
namespace {

// Programmatic dependent launch: K-spmm and K-db are launched with the programmatic-stream-
// serialization attribute, so their CTAs may become resident while the producer kernel drains.
// pdl_wait() blocks until the producer grid has completed and its writes are visible;
// pdl_launch_dependents() lets the next kernel in the stream start its prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" :::); }

// ------------------------------------------------------------------------------------------
// packed complex helpers (float2 = one 64-bit register pair)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a * (-i): a swap + partial negation, folded by ptxas into the consumer's operand modifiers
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }
// a * w:  t = w.y * (a.y, a.x);  result = w.x * a + (-t.x, t.y)          (FMUL2 + FFMA2)
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    const float2 t = __fmul2_rn(make_float2(w.y, w.y), make_float2(a.y, a.x));
    return __ffma2_rn(make_float2(w.x, w.x), a, make_float2(-t.x, t.y));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
}


// ------------------------------------------------------------------------------------------
// SpMM multiply-accumulate and power_to_db pieces shared by the SpMM kernels
// ------------------------------------------------------------------------------------------
// One kernel coefficient k applied to the 8 frames of one spectrum record.  Accumulators are planar
// frame pairs: re[p] = (Re y_f2p, Re y_f2p+1), im[p] likewise; xr* / xi* are the record's chunks.
//   y += k x:        re += k.re xr - k.im xi,   im += k.re xi + k.im xr          (vqt.rs:889-894)
//   y += k conj(x):  re += k.re xr + k.im xi,   im += k.im xr - k.re xi          (vqt.rs:896-910)
template <bool kConj>
__device__ __forceinline__ void mac8(float2 (&re)[4], float2 (&im)[4], float kre, float kim, const float4 &xr03,
                                     const float4 &xr47, const float4 &xi03, const float4 &xi47)
{
    const float2 xr[4] = {make_float2(xr03.x, xr03.y), make_float2(xr03.z, xr03.w), make_float2(xr47.x, xr47.y),
                          make_float2(xr47.z, xr47.w)};
    const float2 xi[4] = {make_float2(xi03.x, xi03.y), make_float2(xi03.z, xi03.w), make_float2(xi47.x, xi47.y),
                          make_float2(xi47.z, xi47.w)};
    const float s_im_xi = kConj ? kim : -kim;   // coefficient of xi in re
    const float s_re_xi = kConj ? -kre : kre;   // coefficient of xi in im
    const float2 a = make_float2(kre, kre), b = make_float2(s_im_xi, s_im_xi), c = make_float2(s_re_xi, s_re_xi),
                 d = make_float2(kim, kim);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        re[p] = __ffma2_rn(a, xr[p], re[p]);
        re[p] = __ffma2_rn(b, xi[p], re[p]);
        im[p] = __ffma2_rn(c, xi[p], im[p]);
        im[p] = __ffma2_rn(d, xr[p], im[p]);
    }
}

constexpr float kAMin = 1e-6f * 1e-6f;  // vqt.rs:924
constexpr float kTopDb = 60.0f;         // vqt.rs:925

// 10 log10(max(p, amin)) - ref_db (vqt.rs:930) as 10 log10(2) * lg2(p): one MUFU.LG2 and an FMA instead of the
// ~30 instructions of log10f, 16 times per thread at the end of K-spmm-db's critical path.  Error: lg2.approx
// is within 2^-22 absolute on [0.5, 2] and 2 ulp elsewhere; |lg2 p| <= 40 here (p >= 1e-12), so at most
// 2 * 2^-18 * 3.0103 = 2.3e-5 dB, plus the rounding of the product (4e-6 dB): 3 % of the 1e-3 dB tolerance.
__device__ __forceinline__ float log_spec(float p, float ref_db)
{
    return fmaf(3.01029995663981195f, __log2f(fmaxf(p, kAMin)), -ref_db);
}

__device__ __forceinline__ float db_out(float l, float floor_db, float log_spec_min)
{
    const float clamped = fmaxf(l, floor_db);          // vqt.rs:945-950
    return log_spec_min > 0.0f ? clamped - log_spec_min : fmaxf(clamped, 0.0f);
}


}  // namespace
}  // namespace pvqt_dev

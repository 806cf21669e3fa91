// device_helpers.cuh -- small device-side helpers shared by the kernel translation units.
#pragma once

#include <cuda_runtime.h>
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


}  // namespace
}  // namespace pvqt_dev

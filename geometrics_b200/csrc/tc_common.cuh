// tc_common.cuh -- tcgen05 / TMEM / TMA / mbarrier helpers shared by the tensor-core filtered searches
// (chamfer_tc.cu, tri_tc.cu).  sm_100a only.
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace geom {
namespace tc {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor: K-major, no swizzle ("interleave"): 8-row x 16-byte core matrices;
// LBO = byte distance between the two 16-byte k-chunks, SBO = byte distance between 8-row groups.
__device__ __forceinline__ u64 make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    u64 d = 0;
    d |= (u64)((saddr & 0x3FFFFu) >> 4);
    d |= (u64)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (u64)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (u64)1 << 46;  // descriptor version 1 (sm_100)
    return d;
}

__device__ __forceinline__ unsigned short bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16_val(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) { return (unsigned)bf16_bits(lo) | ((unsigned)bf16_bits(hi) << 16); }

// 64 consecutive columns of this warp's 32 TMEM lanes: load and wait in ONE statement, so that no use of
// the registers can be scheduled between the load and its wait.
__device__ __forceinline__ void tmem_ld64_wait(unsigned (&v)[64], unsigned taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr)
                 : "memory");
}


__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)  // suspend-time hint (ns): park instead of spinning
                     : "memory");
    }
}
// The same wait with a non-blocking probe first: when the phase is already complete (the common case for a consumer that
// is behind its producer) test_wait answers in one shared-memory round trip, while try_wait costs ~140 cycles even then
// (profiles/r01_tc_probes.txt).
__device__ __forceinline__ void mbar_wait_fast(unsigned long long *bar, unsigned parity)
{
    unsigned done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (!done) mbar_wait(bar, parity);
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// The same, called by a whole (converged) warp: one elected lane announces and issues the copy.  With warp-uniform
// operands this compiles to straight-line code (UBLKCP on uniform registers) instead of a per-lane R2UR loop.
__device__ __forceinline__ void tma_load_elect(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
        "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// the same in two steps, for several copies that complete one barrier phase: announce the bytes once, then copy
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_copy(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// minimum of 32 approximate distances v[o..o+31]: 16 three-input minima as a depth-4 tree
__device__ __forceinline__ float tree_min32(const unsigned (&v)[64], int o)
{
    float l1[12];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        l1[i] = min3(__uint_as_float(v[o + 3 * i]), __uint_as_float(v[o + 3 * i + 1]), __uint_as_float(v[o + 3 * i + 2]));
    l1[10] = __uint_as_float(v[o + 30]);
    l1[11] = __uint_as_float(v[o + 31]);
    const float a0 = min3(l1[0], l1[1], l1[2]), a1 = min3(l1[3], l1[4], l1[5]);
    const float a2 = min3(l1[6], l1[7], l1[8]), a3 = min3(l1[9], l1[10], l1[11]);
    return fminf(min3(a0, a1, a2), a3);
}

}  // namespace tc
}  // namespace geom

// common.cuh -- shared device helpers for libgeom_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/geom_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgeom_b200 is written for sm_100a (B200) only"
#endif

namespace geom {

typedef unsigned long long u64;

// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2), IEEE round-to-nearest per half ----
__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b)
{
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// Same operations as `asm volatile`: the compiler keeps volatile asm statements in program order, which
// lets a kernel hand-schedule its packed-FP32 stream (dependent ops spaced >= the pipe latency apart).
__device__ __forceinline__ u64 add2v(u64 a, u64 b)
{
    u64 r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2v(u64 a, u64 b)
{
    u64 r;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2v(u64 a, u64 b, u64 c)
{
    u64 r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// 4-byte asynchronous global -> shared copy (LDGSTS) and its group fence.
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 3-input minimum (FMNMX3).  NaN operands are ignored unless all are NaN.
__device__ __forceinline__ float min3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// Squared distance with the reference's rounding sequence (chamfer_distance.cu:35-38 as compiled):
// d = fma(dz,dz, fma(dx,dx, rn(dy*dy))), d* = candidate - query.
__device__ __forceinline__ float sqdist_ref(float qx, float qy, float qz, float cx, float cy, float cz)
{
    const float dx = __fsub_rn(cx, qx);
    const float dy = __fsub_rn(cy, qy);
    const float dz = __fsub_rn(cz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// One query against candidates c[0, nc_full) with the reference kernel's own control flow (chamfer_distance.cu:18-54):
// 512-candidate tiles, the first candidate of a tile taken unconditionally (:39), strict '<' inside a tile, strict '>'
// across tiles (:47), the last (tile length mod 4) candidates never examined (:31-33) -- the slow path that gives
// NaN / Inf inputs exactly what the reference gives them.  exact (GEOM_FLAG_EXACT_TAIL): one flat tile, nothing skipped.
__device__ inline void chamfer_ref_row(float qx, float qy, float qz, const float *__restrict__ c, int nc_full, bool exact,
                                       float &res, int &res_i)
{
    res = 0.f;
    res_i = 0;
    const int step = exact ? (nc_full > 0 ? nc_full : 1) : 512;
    for (int k2 = 0; k2 < nc_full; k2 += step) {
        const int end_k = (nc_full < k2 + step ? nc_full : k2 + step) - k2;
        const int end_ka = exact ? end_k : end_k - (end_k & 3);
        float best = 0.f;
        int best_i = 0;
        for (int k = 0; k < end_ka; ++k) {
            const float *p = c + (size_t)(k2 + k) * 3;
            const float d = sqdist_ref(qx, qy, qz, __ldg(p), __ldg(p + 1), __ldg(p + 2));
            if (k == 0 || d < best) {
                best = d;
                best_i = k + k2;
            }
        }
        if (k2 == 0 || res > best) {
            res = best;
            res_i = best_i;
        }
    }
}

// (distance bits << 32) | tag.  Distances are >= +0, so their bit patterns order like the values and
// an unsigned 64-bit min picks the smallest distance, then the smallest tag.
__device__ __forceinline__ u64 make_key(float d, unsigned tag)
{
    return ((u64)__float_as_uint(d) << 32) | (u64)tag;
}

// Filtered atomic min: the plain read may be stale, but keys only ever decrease, so a stale value
// can only cause a redundant atomic, never a missed update.
__device__ __forceinline__ void key_min(u64 *addr, u64 key)
{
    const u64 cur = *(volatile u64 *)addr;
    if (key < cur) atomicMin(addr, key);
}

// ---- optional per-kernel timing (prof.cu); ids index geom_prof_end()'s arrays ----
namespace prof {
enum Kernel {
    kChamferTwoWay = 0, kChamferRefine, kChamferOneWay, kTri, kChamferLoss, kChamferLossBwd,
    kSampleFwd, kSampleBwd, kSampleArea, kGcnGemm, kGcnSpmm, kGcnBwd, kTriClosest, kTriClosestBwd,
    kMisc, kMisc2, kChamferTcScan, kChamferTcPack, kPoolFwd, kPoolBwd, kTriTcScan, kTriTcPrep, kGcnBwdGemm, kBnFwd,
    kBnBwd, kSpare, kKernelCount
};
static_assert(kKernelCount == GEOM_PROF_KERNELS, "geom_b200.h GEOM_PROF_KERNELS out of sync");
struct Scope {
    Scope(cudaStream_t st, int id);
    ~Scope();
    cudaStream_t st_;
    int rec_;
};
}  // namespace prof

// tensor-core filtered Chamfer search (chamfer_tc.cu)
namespace chamfer_tc {
size_t workspace_bytes(int b, int n, int m);
int ctas_per_item(int n, int m);  // thread blocks one batch item contributes to the search grid (one CTA per SM)
cudaError_t launch(int b, int n, const float *xyz1, int m, const float *xyz2, int n_scan, int m_scan, int zero1,
                   int zero2, int exact, float *dist1, int *idx1, float *dist2, int *idx2, void *workspace, cudaStream_t st);
}  // namespace chamfer_tc

// tensor-core filtered point-to-triangle search (tri_tc.cu)
namespace tri_tc {
size_t workspace_bytes(int b, int n, int m);
cudaError_t launch(int b, int n, const float *xyz, int m, const float *t1, const float *t2, const float *t3, int m_scan_end,
                   int exact, int empty_tail, float *dist, int *point, int *index, void *workspace, float eps_factor,
                   cudaStream_t st);
}  // namespace tri_tc

// process-wide debug knobs (misc.cu, geom_debug_set): 0 = not set
namespace debug {
float knob(int id);
}

// tensor-core (3xTF32 tcgen05) feature GEMM of the 0N-GCN layer (gcn_tc.cu)
namespace gcn_tc {
bool supported(int N);
size_t workspace_bytes(int K, int N);
cudaError_t launch(int M, int K, int N, const float *X, const float *W, const float *bias, float *Y, float *T,
                   int ksplit, int ldt, int relu, void *wsplit, cudaStream_t st);
}  // namespace gcn_tc

// Opt `kernel` in to `bytes` (> 48 KB) of dynamic shared memory, once per device.  `done` is the caller's function-local
// static bit set (one bit per device ordinal < 64); atomic, so concurrent callers are safe -- at worst the idempotent
// attribute call runs twice.
template <class K>
inline cudaError_t ensure_dynamic_smem(K kernel, int bytes, std::atomic<unsigned long long> &done)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if (tracked) done.fetch_or(1ull << dev, std::memory_order_release);
    return cudaSuccess;
}

inline int status_from_cuda(cudaError_t e) { return e == cudaSuccess ? GEOM_OK : (int)e; }

}  // namespace geom

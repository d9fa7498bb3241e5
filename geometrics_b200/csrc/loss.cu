// loss.cu -- Chamfer loss on top of the nearest-neighbour indices, forward sums and backward scatter.
//
// Replaces the chain of torch kernels the reference runs after chamfer_dist() (reference:
// utils.py:402-417 in batch_point_to_point, utils.py:451-462 in batch_point_to_surface: offset the
// indices, two index_select gathers, subtract, square, sum, mean) and its autograd backward
// (index_select -> index_add).  The same arithmetic exists as CPU C in the legacy tree
// (old_GEOMetrics/chamfer_distance/src/my_lib.c:67-108 nnd_backward), which is what the backward
// kernel is checked against.
//
// Both kernels are pure HBM streams: 12 B point + 4 B index + 12 B gathered neighbour per point for
// the forward (28 B/point), plus 2 x 12 B of gradient read-modify-write for the backward.
#include "common.cuh"

namespace geom {
namespace loss {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// partial[blockIdx.x] = { sum_j ||xyz2[idx1[j]] - xyz1[j]||^2 , sum_k ||xyz1[idx2[k]] - xyz2[k]||^2 }
// over this block's grid-stride share; optionally the per-point squared distances too, and (F1, utils.py:424-436) the
// per-item counts of points whose distance on f1_scale-scaled coordinates is <= f1_thr, evaluated like the reference
// expression sqrt(sum((s*q - s*p)**2)) <= thr: every product and sum rounded separately, IEEE sqrt.
__global__ void __launch_bounds__(kThreads)
loss_partial_kernel(int b, int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2,
                    const int *__restrict__ idx1, const int *__restrict__ idx2,
                    float *__restrict__ d1_out, float *__restrict__ d2_out, double2 *__restrict__ partial,
                    float f1_scale, float f1_thr, int *__restrict__ cnt1, int *__restrict__ cnt2,
                    unsigned *__restrict__ arrive, float *__restrict__ sums)
{
    const long long rows1 = (long long)b * n, rows2 = (long long)b * m;
    const int lane = threadIdx.x & 31;
    float s1 = 0.f, s2 = 0.f;
    // warp-uniform trip count: the F1 counts are aggregated per warp before they touch memory
    for (long long base = (long long)blockIdx.x * kThreads + (threadIdx.x & ~31); base < rows1 + rows2;
         base += (long long)gridDim.x * kThreads) {
        const long long t = base + lane;
        const bool active = t < rows1 + rows2;
        bool first = true, hit = false;
        int bi = 0;
        if (active) {
            first = t < rows1;
            const long long r = first ? t : t - rows1;
            const int cnt = first ? n : m, ocnt = first ? m : n;
            bi = (int)(r / cnt);
            const float *p = (first ? xyz1 : xyz2) + r * 3;
            const int k = (first ? idx1 : idx2)[r];
            const float *q = (first ? xyz2 : xyz1) + ((long long)bi * ocnt + k) * 3;
            const float px = p[0], py = p[1], pz = p[2], qx = q[0], qy = q[1], qz = q[2];
            const float d = sqdist_ref(px, py, pz, qx, qy, qz);
            if (first) {
                s1 += d;
                if (d1_out) d1_out[r] = d;
            } else {
                s2 += d;
                if (d2_out) d2_out[r] = d;
            }
            if (cnt1) {
                const float tx = __fsub_rn(__fmul_rn(f1_scale, qx), __fmul_rn(f1_scale, px));
                const float ty = __fsub_rn(__fmul_rn(f1_scale, qy), __fmul_rn(f1_scale, py));
                const float tz = __fsub_rn(__fmul_rn(f1_scale, qz), __fmul_rn(f1_scale, pz));
                const float ss = __fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), __fmul_rn(tz, tz));
                hit = __fsqrt_rn(ss) <= f1_thr;
            }
        }
        if (cnt1) {
            const int key = active ? (bi * 2 + (first ? 0 : 1)) : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const unsigned hits = __ballot_sync(0xffffffffu, hit) & peers;
            if (active && hits && lane == __ffs((int)peers) - 1) atomicAdd((first ? cnt1 : cnt2) + bi, __popc(hits));
        }
    }
    __shared__ float sh[2][kThreads / 32];
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = s1;
        sh[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, c = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            a += (double)sh[0][w];
            c += (double)sh[1][w];
        }
        partial[blockIdx.x] = make_double2(a, c);
    }
    // the last block to arrive adds the partials in a fixed order (deterministic) -- no second launch
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(arrive, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    __shared__ double sa[kThreads], sc[kThreads];
    double a = 0, c = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) {
        a += __ldcg(&partial[i].x);  // L2 reads: the partials were written by other blocks
        c += __ldcg(&partial[i].y);
    }
    sa[threadIdx.x] = a;
    sc[threadIdx.x] = c;
    __syncthreads();
    for (int st = kThreads / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) {
            sa[threadIdx.x] += sa[threadIdx.x + st];
            sc[threadIdx.x] += sc[threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sums[0] = (float)sa[0];
        sums[1] = (float)sc[0];
        *arrive = 0u;  // ready for the next call on this workspace
    }
}

// d/dxyz of  L = w[0] * s1 + w[1] * s2  with the indices held constant:
//   row j of xyz1 with q = xyz2[idx1[j]]:  g = 2 w0 (p - q);  grad1[j] += g;  grad2[idx1[j]] -= g
//   row k of xyz2 with q = xyz1[idx2[k]]:  g = 2 w1 (p - q);  grad2[k] += g;  grad1[idx2[k]] -= g
// (my_lib.c:67-108).  Gradient buffers must be zeroed by the caller (the entry point does it).
__global__ void __launch_bounds__(kThreads)
loss_bwd_kernel(int b, int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2,
                const int *__restrict__ idx1, const int *__restrict__ idx2, const float *__restrict__ w,
                float *__restrict__ grad1, float *__restrict__ grad2)
{
    const long long rows1 = (long long)b * n, rows2 = (long long)b * m;
    const float w0 = 2.0f * w[0], w1 = 2.0f * w[1];
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < rows1 + rows2;
         t += (long long)gridDim.x * kThreads) {
        const bool first = t < rows1;
        const long long r = first ? t : t - rows1;
        const int cnt = first ? n : m, ocnt = first ? m : n;
        const int bi = (int)(r / cnt);
        const long long qo = ((long long)bi * ocnt + (first ? idx1 : idx2)[r]) * 3;
        const float *p = (first ? xyz1 : xyz2) + r * 3;
        const float *q = (first ? xyz2 : xyz1) + qo;
        float *gp = (first ? grad1 : grad2);
        float *gq = (first ? grad2 : grad1);
        const float ww = first ? w0 : w1;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float g = ww * (p[c] - q[c]);
            if (gp) atomicAdd(gp + r * 3 + c, g);
            if (gq) atomicAdd(gq + qo + c, -g);
        }
    }
}

static int blocks_for(long long items)
{
    long long bl = (items + kThreads - 1) / kThreads;
    if (bl > kMaxBlocks) bl = kMaxBlocks;
    if (bl < 1) bl = 1;
    return (int)bl;
}

}  // namespace loss
}  // namespace geom

using namespace geom;

extern "C" size_t geom_chamfer_loss_workspace_bytes(int b, int n, int m)
{
    if (b < 0 || n < 0 || m < 0) return 0;
    return sizeof(double2) * (size_t)loss::blocks_for((long long)b * ((long long)n + m)) + 16;  // partials + arrival counter
}

static int chamfer_loss_fwd_impl(int b, int n, const float *xyz1, int m, const float *xyz2, const int *idx1, const int *idx2,
                                 float *dist1, float *dist2, float *sums, float f1_scale, float f1_thr, int *count1, int *count2,
                                 void *workspace, size_t workspace_bytes, geom_stream_t stream)
{
    using namespace geom::loss;
    if (b < 0 || n < 0 || m < 0 || !sums) return GEOM_ERR_INVALID_ARG;
    if ((count1 == nullptr) != (count2 == nullptr)) return GEOM_ERR_INVALID_ARG;
    if ((long long)b * n * 3 > 0x7fffffffLL || (long long)b * m * 3 > 0x7fffffffLL) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (count1 && b > 0) {
        if ((e = cudaMemsetAsync(count1, 0, sizeof(int) * (size_t)b, st)) != cudaSuccess) return status_from_cuda(e);
        if ((e = cudaMemsetAsync(count2, 0, sizeof(int) * (size_t)b, st)) != cudaSuccess) return status_from_cuda(e);
    }
    const long long items = (long long)b * ((long long)n + m);
    // a direction with no neighbours to gather from contributes nothing (reference: empty index_select)
    if (items == 0 || n == 0 || m == 0) return status_from_cuda(cudaMemsetAsync(sums, 0, 2 * sizeof(float), st));
    if (!xyz1 || !xyz2 || !idx1 || !idx2) return GEOM_ERR_INVALID_ARG;
    const int blocks = blocks_for(items);
    if (!workspace || workspace_bytes < sizeof(double2) * (size_t)blocks + 16 || ((uintptr_t)workspace & 15u))
        return GEOM_ERR_WORKSPACE;
    unsigned *arrive = (unsigned *)((double2 *)workspace + blocks);
    if ((e = cudaMemsetAsync(arrive, 0, 4, st)) != cudaSuccess) return status_from_cuda(e);
    {
        prof::Scope ps(st, prof::kChamferLoss);
        loss_partial_kernel<<<blocks, kThreads, 0, st>>>(b, n, xyz1, m, xyz2, idx1, idx2, dist1, dist2, (double2 *)workspace,
                                                         f1_scale, f1_thr, count1, count2, arrive, sums);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_chamfer_loss_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                                     const int *idx1, const int *idx2, float *dist1, float *dist2,
                                     float *sums, void *workspace, size_t workspace_bytes,
                                     geom_stream_t stream)
{
    return chamfer_loss_fwd_impl(b, n, xyz1, m, xyz2, idx1, idx2, dist1, dist2, sums, 0.f, 0.f, nullptr, nullptr, workspace,
                                 workspace_bytes, stream);
}

extern "C" int geom_chamfer_loss_f1_fwd(int b, int n, const float *xyz1, int m, const float *xyz2, const int *idx1,
                                        const int *idx2, float *dist1, float *dist2, float *sums, float f1_scale,
                                        float f1_threshold, int *count1, int *count2, void *workspace, size_t workspace_bytes,
                                        geom_stream_t stream)
{
    if (!count1 || !count2) return GEOM_ERR_INVALID_ARG;
    return chamfer_loss_fwd_impl(b, n, xyz1, m, xyz2, idx1, idx2, dist1, dist2, sums, f1_scale, f1_threshold, count1, count2,
                                 workspace, workspace_bytes, stream);
}

extern "C" int geom_chamfer_loss_bwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                                     const int *idx1, const int *idx2, const float *weights,
                                     float *grad1, float *grad2, geom_stream_t stream)
{
    using namespace geom::loss;
    if (b < 0 || n < 0 || m < 0 || !weights) return GEOM_ERR_INVALID_ARG;
    if ((long long)b * n * 3 > 0x7fffffffLL || (long long)b * m * 3 > 0x7fffffffLL) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (grad1 && (size_t)b * n) e = cudaMemsetAsync(grad1, 0, sizeof(float) * 3 * (size_t)b * n, st);
    if (e == cudaSuccess && grad2 && (size_t)b * m) e = cudaMemsetAsync(grad2, 0, sizeof(float) * 3 * (size_t)b * m, st);
    if (e != cudaSuccess) return status_from_cuda(e);
    const long long items = (long long)b * ((long long)n + m);
    if (items == 0 || n == 0 || m == 0 || (!grad1 && !grad2)) return GEOM_OK;
    if (!xyz1 || !xyz2 || !idx1 || !idx2) return GEOM_ERR_INVALID_ARG;
    {
        prof::Scope ps(st, prof::kChamferLossBwd);
        loss_bwd_kernel<<<blocks_for(items), kThreads, 0, st>>>(b, n, xyz1, m, xyz2, idx1, idx2, weights, grad1, grad2);
    }
    return status_from_cuda(cudaGetLastError());
}

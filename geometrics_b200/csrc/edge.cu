// edge.cu -- edge-length regulariser (reference: batch_calc_edge, utils.py:636-651; called for the three predicted
// meshes of every step, GEOMetrics.py:151-153).  SURVEY.md 8f-2.
//
//   loss = ( mean|p2-p1|^2 + mean|p3-p1|^2 + mean|p2-p3|^2 ) / 3   over all B*F faces (p1,p2,p3 = face vertices)
//
// The reference runs 3 index_selects, 3 subtractions, 3 squares, 3 row sums, 3 means and 3 adds (15 launches, and as
// many again in autograd's backward); here one launch sums the 3*B*F squared edge lengths (fp64 accumulation, one
// atomic per CTA) and one launch scatters the gradient.  Gather-bound: 24 B of indices + 36 B of vertices per face.
#include "common.cuh"

namespace geom {
namespace edge {

constexpr int kThreads = 256;

__device__ __forceinline__ void load_face(const float *__restrict__ vb, const long long *__restrict__ f, float (&p)[3][3], int (&id)[3])
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        id[k] = (int)f[k];
#pragma unroll
        for (int c = 0; c < 3; ++c) p[k][c] = vb[(size_t)id[k] * 3 + c];
    }
}

__global__ void __launch_bounds__(kThreads)
fwd_kernel(long long total, int V, int F, const float *__restrict__ verts, const long long *__restrict__ faces, double *__restrict__ acc)
{
    __shared__ double part[kThreads / 32];
    double s = 0.0;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        const int b = (int)(t / F), f = (int)(t - (long long)b * F);
        float p[3][3];
        int id[3];
        load_face(verts + (size_t)b * V * 3, faces + (size_t)f * 3, p, id);
        float e = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float e1 = p[1][c] - p[0][c], e2 = p[2][c] - p[0][c], e3 = p[1][c] - p[2][c];
            e += e1 * e1 + e2 * e2 + e3 * e3;
        }
        s += (double)e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) tsum += part[w];
        atomicAdd(acc, tsum);
    }
}

// grad_verts += g * dL/dp with L = |p2-p1|^2 + |p3-p1|^2 + |p2-p3|^2 per face, g = *grad_scalar * scale
__global__ void __launch_bounds__(kThreads)
bwd_kernel(long long total, int V, int F, const float *__restrict__ verts, const long long *__restrict__ faces,
           const float *__restrict__ grad_scalar, float scale, float *__restrict__ grad_verts)
{
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const float g = 2.f * scale * grad_scalar[0];
    const int b = (int)(t / F), f = (int)(t - (long long)b * F);
    float p[3][3];
    int id[3];
    load_face(verts + (size_t)b * V * 3, faces + (size_t)f * 3, p, id);
    float *gb = grad_verts + (size_t)b * V * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float e1 = p[1][c] - p[0][c], e2 = p[2][c] - p[0][c], e3 = p[1][c] - p[2][c];
        atomicAdd(gb + (size_t)id[0] * 3 + c, g * (-e1 - e2));
        atomicAdd(gb + (size_t)id[1] * 3 + c, g * (e1 + e3));
        atomicAdd(gb + (size_t)id[2] * 3 + c, g * (e2 - e3));
    }
}

}  // namespace edge
}  // namespace geom

using namespace geom;

extern "C" int geom_edge_loss_fwd(int b, int V, const float *verts, int F, const int64_t *faces, double *sum, geom_stream_t stream)
{
    if (b < 0 || V < 0 || F < 0 || !sum) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sum, 0, sizeof(double), st);
    if (e != cudaSuccess) return status_from_cuda(e);
    const long long total = (long long)b * F;
    if (total == 0) return GEOM_OK;
    if (!verts || !faces || V == 0) return GEOM_ERR_INVALID_ARG;
    {
        prof::Scope ps(st, prof::kMisc);
        const long long blocks = (total + edge::kThreads - 1) / edge::kThreads;
        edge::fwd_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), edge::kThreads, 0, st>>>(total, V, F, verts, (const long long *)faces, sum);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_edge_loss_bwd(int b, int V, const float *verts, int F, const int64_t *faces, const float *grad_scalar,
                                  float scale, float *grad_verts, geom_stream_t stream)
{
    if (b < 0 || V < 0 || F < 0 || !grad_verts) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)b * V == 0) return GEOM_OK;
    cudaError_t e = cudaMemsetAsync(grad_verts, 0, sizeof(float) * 3 * (size_t)b * V, st);
    if (e != cudaSuccess) return status_from_cuda(e);
    const long long total = (long long)b * F;
    if (total == 0) return GEOM_OK;
    if (!verts || !faces || !grad_scalar) return GEOM_ERR_INVALID_ARG;
    {
        prof::Scope ps(st, prof::kMisc);
        edge::bwd_kernel<<<(unsigned)((total + edge::kThreads - 1) / edge::kThreads), edge::kThreads, 0, st>>>(
            total, V, F, verts, (const long long *)faces, grad_scalar, scale, grad_verts);
    }
    return status_from_cuda(cudaGetLastError());
}

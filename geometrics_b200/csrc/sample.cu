// sample.cu -- differentiable area-weighted face sampler, fused.
//
// Replaces the ~25 small torch kernels of batch_sample (reference: utils.py:590-633):
//   face areas (utils.py:596-602)            -> area_kernel   (one CTA per mesh, normalised in place)
//   faces[choices] + 3 index_select + the barycentric combine (utils.py:614-631) -> sample_fwd_kernel
//   autograd of the three index_selects (index_add scatter)                      -> sample_bwd_kernel
// The random draws themselves (torch.multinomial over the areas, two uniform vectors) stay in torch
// so that a seed gives the same (choices, r1, r2) stream the reference would consume.
//
// All three kernels are gather/scatter streams bounded by HBM/L2 bandwidth; per sample the forward
// reads 8 B choice + 8 B randoms + 24 B face + 36 B vertices and writes 12 B.
#include "common.cuh"

namespace geom {
namespace sample {

constexpr int kThreads = 256;

// Area = sqrt(a + b + c) / 2 with x = v0 - v1, y = v1 - v2 and
//   a = (x2*y3 - x3*y2)^2, b = (x3*y1 - x1*y3)^2, c = (x1*y2 - x2*y1)^2      (utils.py:596-601)
// each operation rounded separately, as the chain of torch element-wise kernels does.
__device__ __forceinline__ float face_area(const float *__restrict__ v, const long long *__restrict__ f)
{
    const float *p0 = v + f[0] * 3, *p1 = v + f[1] * 3, *p2 = v + f[2] * 3;
    const float x1 = __fsub_rn(p0[0], p1[0]), x2 = __fsub_rn(p0[1], p1[1]), x3 = __fsub_rn(p0[2], p1[2]);
    const float y1 = __fsub_rn(p1[0], p2[0]), y2 = __fsub_rn(p1[1], p2[1]), y3 = __fsub_rn(p1[2], p2[2]);
    const float ta = __fsub_rn(__fmul_rn(x2, y3), __fmul_rn(x3, y2));
    const float tb = __fsub_rn(__fmul_rn(x3, y1), __fmul_rn(x1, y3));
    const float tc = __fsub_rn(__fmul_rn(x1, y2), __fmul_rn(x2, y1));
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(ta, ta), __fmul_rn(tb, tb)), __fmul_rn(tc, tc));
    return __fdiv_rn(__fsqrt_rn(s), 2.0f);
}

// One CTA per mesh: areas -> block sum (fixed order) -> areas / sum  (utils.py:602).
__global__ void __launch_bounds__(kThreads)
area_kernel(int V, const float *__restrict__ verts, int F, const long long *__restrict__ faces,
            float *__restrict__ areas, int normalise)
{
    const int bi = blockIdx.x;
    const float *v = verts + (size_t)bi * V * 3;
    float *out = areas + (size_t)bi * F;
    float acc = 0.f;
    for (int f = threadIdx.x; f < F; f += kThreads) {
        const float a = face_area(v, faces + (size_t)f * 3);
        out[f] = a;
        acc += a;
    }
    if (!normalise) return;
    __shared__ float sh[kThreads];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    const float total = sh[0];
    for (int f = threadIdx.x; f < F; f += kThreads) out[f] = __fdiv_rn(out[f], total);  // own writes only
}

// Barycentric weights from the two uniforms: u = sqrt(r1), v = r2 -> (1-u, u(1-v), uv)  (utils.py:627-630)
__device__ __forceinline__ void bary(float r1, float r2, float &w0, float &w1, float &w2)
{
    const float u = __fsqrt_rn(r1);
    w0 = __fsub_rn(1.0f, u);
    w1 = __fmul_rn(u, __fsub_rn(1.0f, r2));
    w2 = __fmul_rn(u, r2);
}

__global__ void __launch_bounds__(kThreads)
sample_fwd_kernel(long long total, int num, int V, const float *__restrict__ verts, int F,
                  const long long *__restrict__ faces, const long long *__restrict__ choices,
                  const float *__restrict__ r1, const float *__restrict__ r2, float *__restrict__ points)
{
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
         t += (long long)gridDim.x * kThreads) {
        const int bi = (int)(t / num);
        const long long *f = faces + choices[t] * 3;
        const float *v = verts + (size_t)bi * V * 3;
        const float *x = v + f[0] * 3, *y = v + f[1] * 3, *z = v + f[2] * 3;
        float w0, w1, w2;
        bary(r1[t], r2[t], w0, w1, w2);
        float *o = points + t * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)  // ((1-u)*xs + (u*(1-v))*ys) + u*v*zs, each product and sum rounded
            o[c] = __fadd_rn(__fadd_rn(__fmul_rn(w0, x[c]), __fmul_rn(w1, y[c])), __fmul_rn(w2, z[c]));
    }
}

// grad_verts[b, f_k] += w_k * grad_points[b, s]   (the index_add that autograd runs for utils.py:624-626)
__global__ void __launch_bounds__(kThreads)
sample_bwd_kernel(long long total, int num, int V, int F, const long long *__restrict__ faces,
                  const long long *__restrict__ choices, const float *__restrict__ r1,
                  const float *__restrict__ r2, const float *__restrict__ gpoints, float *__restrict__ gverts)
{
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
         t += (long long)gridDim.x * kThreads) {
        const int bi = (int)(t / num);
        const long long *f = faces + choices[t] * 3;
        float *gv = gverts + (size_t)bi * V * 3;
        float w0, w1, w2;
        bary(r1[t], r2[t], w0, w1, w2);
        const float *g = gpoints + t * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float gc = g[c];
            atomicAdd(gv + f[0] * 3 + c, w0 * gc);
            atomicAdd(gv + f[1] * 3 + c, w1 * gc);
            atomicAdd(gv + f[2] * 3 + c, w2 * gc);
        }
    }
}

static int blocks_for(long long items)
{
    long long bl = (items + kThreads - 1) / kThreads;
    if (bl > 148 * 16) bl = 148 * 16;
    return (int)(bl < 1 ? 1 : bl);
}

}  // namespace sample
}  // namespace geom

using namespace geom;

extern "C" int geom_face_areas(int b, int V, const float *verts, int F, const int64_t *faces,
                               float *areas, int normalise, geom_stream_t stream)
{
    if (b < 0 || V < 0 || F < 0) return GEOM_ERR_INVALID_ARG;
    if (b == 0 || F == 0) return GEOM_OK;
    if (!verts || !faces || !areas || V == 0) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        prof::Scope ps(st, prof::kSampleArea);
        sample::area_kernel<<<b, sample::kThreads, 0, st>>>(V, verts, F, (const long long *)faces, areas, normalise);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_sample_fwd(int b, int V, const float *verts, int F, const int64_t *faces, int num,
                               const int64_t *choices, const float *r1, const float *r2, float *points,
                               geom_stream_t stream)
{
    if (b < 0 || V < 0 || F < 0 || num < 0) return GEOM_ERR_INVALID_ARG;
    const long long total = (long long)b * num;
    if (total == 0) return GEOM_OK;
    if (!verts || !faces || !choices || !r1 || !r2 || !points || V == 0 || F == 0) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        prof::Scope ps(st, prof::kSampleFwd);
        sample::sample_fwd_kernel<<<sample::blocks_for(total), sample::kThreads, 0, st>>>(
            total, num, V, verts, F, (const long long *)faces, (const long long *)choices, r1, r2, points);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_sample_bwd(int b, int V, int F, const int64_t *faces, int num, const int64_t *choices,
                               const float *r1, const float *r2, const float *grad_points,
                               float *grad_verts, geom_stream_t stream)
{
    if (b < 0 || V < 0 || F < 0 || num < 0 || !grad_verts) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)b * V == 0) return GEOM_OK;
    cudaError_t e = cudaMemsetAsync(grad_verts, 0, sizeof(float) * 3 * (size_t)b * V, st);
    if (e != cudaSuccess) return status_from_cuda(e);
    const long long total = (long long)b * num;
    if (total == 0) return GEOM_OK;
    if (!faces || !choices || !r1 || !r2 || !grad_points || F == 0) return GEOM_ERR_INVALID_ARG;
    {
        prof::Scope ps(st, prof::kSampleBwd);
        sample::sample_bwd_kernel<<<sample::blocks_for(total), sample::kThreads, 0, st>>>(
            total, num, V, F, (const long long *)faces, (const long long *)choices, r1, r2, grad_points, grad_verts);
    }
    return status_from_cuda(cudaGetLastError());
}

// triloss.cu -- point-to-surface loss: differentiable closest point on the winning triangle.
//
// Replaces, after tri_dist() has returned (option, face index) per ground-truth point, the torch chain
// of batch_point_to_surface (reference: utils.py:467-481: three index_selects of the predicted
// vertices by face, three more by winning face) and calc_point_to_line (utils.py:506-550: all seven
// candidate closest points for every row, seven boolean-mask scatters = seven host syncs) plus the
// helper classes edge (utils.py:553-570) and Plane (utils.py:572-587), and their autograd backward.
//
// Per point only the candidate its option selects is evaluated, straight from (verts, faces):
//   1/2/3: vertex A/B/C;   4/5/6: S + t (T - S), t = (p-S).(T-S)/|T-S|^2 on AB / BC / CA -- the CA
//   edge uses its own direction here, as the reference's Python does (utils.py:515,543), unlike the
//   search kernel's ranking quirk;   0: p - ((p-A).n) n with n = cross(A-B, A-C) normalised.
// forward: sum over all points of |c - p|^2 (double accumulation, deterministic two-stage reduction).
// backward: d/dverts of  weight * sum, scattered with atomicAdd into (B,V,3).
#include "common.cuh"

namespace geom {
namespace triloss {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;

struct F3 {
    float x, y, z;
};
__device__ __forceinline__ F3 ld3(const float *p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator*(float s, F3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ void atomic_add3(float *dst, F3 g)
{
    atomicAdd(dst + 0, g.x);
    atomicAdd(dst + 1, g.y);
    atomicAdd(dst + 2, g.z);
}

struct Item {
    F3 p, A, B, C;
    long long ia, ib, ic;  // float offsets of the three vertices inside this mesh's (V,3) block
    int opt;
};

__device__ __forceinline__ Item load_item(long long t, int n, int V, const float *__restrict__ xyz,
                                          const float *__restrict__ verts, int F,
                                          const long long *__restrict__ faces,
                                          const int *__restrict__ option, const int *__restrict__ index)
{
    Item it;
    const int bi = (int)(t / n);
    const float *v = verts + (size_t)bi * V * 3;
    int fi = index[t];
    fi = min(max(fi, 0), F - 1);
    const long long *f = faces + (size_t)fi * 3;
    it.ia = f[0] * 3;
    it.ib = f[1] * 3;
    it.ic = f[2] * 3;
    it.p = ld3(xyz + t * 3);
    it.A = ld3(v + it.ia);
    it.B = ld3(v + it.ib);
    it.C = ld3(v + it.ic);
    it.opt = option[t];
    return it;
}

// closest point for the item's option
__device__ __forceinline__ F3 closest(const Item &it)
{
    switch (it.opt) {
        case 1: return it.A;
        case 2: return it.B;
        case 3: return it.C;
        case 4:
        case 5:
        case 6: {
            const F3 S = it.opt == 4 ? it.A : it.opt == 5 ? it.B : it.C;
            const F3 T = it.opt == 4 ? it.B : it.opt == 5 ? it.C : it.A;
            const F3 D = T - S;
            const float t = dot(it.p - S, D) / dot(D, D);  // edge.Project, utils.py:564-569
            return S + t * D;                              // edge.PointAt, utils.py:559-560
        }
        default: {
            const F3 N = cross(it.A - it.B, it.A - it.C);  // utils.py:518
            const F3 nh = (1.0f / sqrtf(dot(N, N))) * N;   // Plane.__init__, utils.py:573-575
            const float s = dot(it.p - it.A, nh);          // Plane.Project, utils.py:580-586
            return it.p - s * nh;
        }
    }
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kThreads)
fwd_kernel(long long total, int n, int V, const float *__restrict__ xyz, const float *__restrict__ verts,
           int F, const long long *__restrict__ faces, const int *__restrict__ option,
           const int *__restrict__ index, float *__restrict__ dist_out, float *__restrict__ closest_out,
           double *__restrict__ partial)
{
    float acc = 0.f;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
         t += (long long)gridDim.x * kThreads) {
        const Item it = load_item(t, n, V, xyz, verts, F, faces, option, index);
        const F3 c = closest(it);
        const F3 e = c - it.p;
        const float d = dot(e, e);
        acc += d;
        if (dist_out) dist_out[t] = d;
        if (closest_out) {
            closest_out[t * 3 + 0] = c.x;
            closest_out[t * 3 + 1] = c.y;
            closest_out[t * 3 + 2] = c.z;
        }
    }
    __shared__ float sh[kThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) a += (double)sh[w];
        partial[blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(256) final_kernel(const double *__restrict__ partial, int nblocks,
                                                     float *__restrict__ sum)
{
    __shared__ double sa[256];
    double a = 0;
    for (int i = threadIdx.x; i < nblocks; i += 256) a += partial[i];
    sa[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sa[threadIdx.x] += sa[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) sum[0] = (float)sa[0];
}

// Gradient of weight * |c(A,B,C) - p|^2 with respect to the three vertices (p constant).
// Derivation in DESIGN.md "Point-to-surface backward"; checked against torch autograd of the
// reference expressions in tests/test_gpu_losses.py.
__global__ void __launch_bounds__(kThreads)
bwd_kernel(long long total, int n, int V, const float *__restrict__ xyz, const float *__restrict__ verts,
           int F, const long long *__restrict__ faces, const int *__restrict__ option,
           const int *__restrict__ index, const float *__restrict__ weight, float *__restrict__ gverts)
{
    const float w2 = 2.0f * weight[0];
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
         t += (long long)gridDim.x * kThreads) {
        const Item it = load_item(t, n, V, xyz, verts, F, faces, option, index);
        float *gv = gverts + (size_t)(t / n) * V * 3;
        if (it.opt >= 1 && it.opt <= 3) {
            const F3 Vx = it.opt == 1 ? it.A : it.opt == 2 ? it.B : it.C;
            atomic_add3(gv + (it.opt == 1 ? it.ia : it.opt == 2 ? it.ib : it.ic), w2 * (Vx - it.p));
        } else if (it.opt >= 4 && it.opt <= 6) {
            const F3 S = it.opt == 4 ? it.A : it.opt == 5 ? it.B : it.C;
            const F3 T = it.opt == 4 ? it.B : it.opt == 5 ? it.C : it.A;
            const long long is = it.opt == 4 ? it.ia : it.opt == 5 ? it.ib : it.ic;
            const long long itx = it.opt == 4 ? it.ib : it.opt == 5 ? it.ic : it.ia;
            const F3 D = T - S, pS = it.p - S;
            const float DD = dot(D, D);
            const float tt = dot(pS, D) / DD;
            const F3 g = w2 * ((S + tt * D) - it.p);
            const float k = dot(g, D) / DD;
            const F3 gD = tt * g + k * (pS - (2.0f * tt) * D);
            atomic_add3(gv + is, (g - k * D) - gD);
            atomic_add3(gv + itx, gD);
        } else {
            const F3 U = it.A - it.B, W = it.A - it.C;
            const F3 N = cross(U, W);
            const float inv = 1.0f / sqrtf(dot(N, N));
            const F3 nh = inv * N;
            const F3 pA = it.p - it.A;
            const float s = dot(pA, nh);
            const F3 g = w2 * ((-s) * nh);  // c - p = -s n
            const float gn = dot(g, nh);
            const F3 h = (-1.0f) * (gn * pA + s * g);
            const F3 G = inv * (h - dot(h, nh) * nh);
            const F3 gU = cross(W, G), gW = cross(G, U);
            atomic_add3(gv + it.ia, (gn * nh + gU) + gW);
            atomic_add3(gv + it.ib, (-1.0f) * gU);
            atomic_add3(gv + it.ic, (-1.0f) * gW);
        }
    }
}

static int blocks_for(long long items)
{
    long long bl = (items + kThreads - 1) / kThreads;
    if (bl > kMaxBlocks) bl = kMaxBlocks;
    return (int)(bl < 1 ? 1 : bl);
}

}  // namespace triloss
}  // namespace geom

using namespace geom;

extern "C" size_t geom_tri_loss_workspace_bytes(int b, int n)
{
    if (b < 0 || n < 0) return 0;
    return sizeof(double) * (size_t)triloss::blocks_for((long long)b * n);
}

extern "C" int geom_tri_loss_fwd(int b, int n, const float *xyz, int V, const float *verts, int F,
                                 const int64_t *faces, const int *option, const int *index,
                                 float *dist, float *closest, float *sum, void *workspace,
                                 size_t workspace_bytes, geom_stream_t stream)
{
    using namespace geom::triloss;
    if (b < 0 || n < 0 || V < 0 || F < 0 || !sum) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)b * n;
    if (total == 0) return status_from_cuda(cudaMemsetAsync(sum, 0, sizeof(float), st));
    if (!xyz || !verts || !faces || !option || !index || V == 0 || F == 0) return GEOM_ERR_INVALID_ARG;
    const int blocks = blocks_for(total);
    if (!workspace || workspace_bytes < sizeof(double) * (size_t)blocks || ((uintptr_t)workspace & 7u))
        return GEOM_ERR_WORKSPACE;
    {
        prof::Scope ps(st, prof::kTriClosest);
        fwd_kernel<<<blocks, kThreads, 0, st>>>(total, n, V, xyz, verts, F, (const long long *)faces, option,
                                                index, dist, closest, (double *)workspace);
        final_kernel<<<1, 256, 0, st>>>((const double *)workspace, blocks, sum);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_tri_loss_bwd(int b, int n, const float *xyz, int V, const float *verts, int F,
                                 const int64_t *faces, const int *option, const int *index,
                                 const float *weight, float *grad_verts, geom_stream_t stream)
{
    using namespace geom::triloss;
    if (b < 0 || n < 0 || V < 0 || F < 0 || !grad_verts || !weight) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)b * V == 0) return GEOM_OK;
    cudaError_t e = cudaMemsetAsync(grad_verts, 0, sizeof(float) * 3 * (size_t)b * V, st);
    if (e != cudaSuccess) return status_from_cuda(e);
    const long long total = (long long)b * n;
    if (total == 0) return GEOM_OK;
    if (!xyz || !verts || !faces || !option || !index || F == 0) return GEOM_ERR_INVALID_ARG;
    {
        prof::Scope ps(st, prof::kTriClosestBwd);
        bwd_kernel<<<blocks_for(total), kThreads, 0, st>>>(total, n, V, xyz, verts, F, (const long long *)faces,
                                                           option, index, weight, grad_verts);
    }
    return status_from_cuda(cudaGetLastError());
}

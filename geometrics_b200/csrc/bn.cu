// bn.cu -- BatchNorm1d(V) + ReLU + averaged residual of the mesh deformation block, one launch each way.
//
// Replaces, per stage of BatchMeshDeformationBlock.forward (reference models.py:241-294),
//     x = F.relu(self.bnK(x));  features = (features + x) / 2            (the second line only after even layers)
// i.e. torch's batch_norm + relu + add + div kernels (and their four backward kernels).  nn.BatchNorm1d(verts) applied
// to a (B, V, C) tensor treats V as the channel axis: statistics per VERTEX over the B*C values of that vertex
// (models.py:222-235, checkpoint shapes bnK.weight (482,)).
//   training:  mean_v, var_v (biased) over (b, c);  y = (x - mean_v) / sqrt(var_v + eps) * gamma_v + beta_v;
//              running_mean/var updated with momentum (unbiased variance), as torch does
//   eval:      the running statistics instead
//   out = relu(y)                       (res == NULL)
//   out = (res + relu(y)) / 2           (res != NULL; res may be a column slice of a wider tensor: row stride ld_res)
// One CTA per vertex: the vertex' B rows of C floats (B*C*4 = 12 KB at the reference's B=16, C=192) are read three times
// from L2 (sum, centred sum of squares, normalise) -- two-pass statistics like torch's, no E[x^2]-E[x]^2 cancellation.
// Backward (one CTA per vertex): dy = dout * [y > 0] (x 1/2 with a residual), dgamma = sum dy xhat, dbeta = sum dy,
//   dx = gamma invstd (dy - dbeta/N - xhat dgamma/N)     (training);   dx = gamma invstd dy     (eval)
// and dres = dout / 2 written in the same pass.
#include "common.cuh"

namespace geom {
namespace bn {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_sum(float v, float *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // sh may still be read from the previous reduction
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    return t;
}

struct FwdArgs {
    const float *x;       // (B, V, C)
    const float *res;     // (B, V, >= C) with row stride ld_res, or null
    const float *gamma, *beta;  // (V) or null (affine off)
    float *running_mean, *running_var;  // (V) or null
    float *save_mean, *save_invstd;     // (V): what the backward needs
    float *out;           // (B, V, C)
    long long *batches_tracked;  // num_batches_tracked (one int64) or null; incremented in training
    int B, V, C, ld_res, training;
    float eps, momentum;
};

__global__ void __launch_bounds__(kThreads) fwd_kernel(const FwdArgs a)
{
    __shared__ float sh[kThreads / 32];
    const int v = blockIdx.x;
    const int n = a.B * a.C;
    const size_t vstride = (size_t)a.V * a.C;
    const float *xv = a.x + (size_t)v * a.C;
    float mean, invstd;
    if (a.training) {
        float s = 0.f;
        for (int i = threadIdx.x; i < n; i += kThreads) s += xv[(size_t)(i / a.C) * vstride + (i % a.C)];
        mean = block_sum(s, sh) / (float)n;
        float q = 0.f;
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const float d = xv[(size_t)(i / a.C) * vstride + (i % a.C)] - mean;
            q = fmaf(d, d, q);
        }
        const float var = block_sum(q, sh) / (float)n;
        invstd = rsqrtf(var + a.eps);
        if (threadIdx.x == 0) {
            if (a.running_mean) a.running_mean[v] = (1.f - a.momentum) * a.running_mean[v] + a.momentum * mean;
            if (a.running_var) a.running_var[v] = (1.f - a.momentum) * a.running_var[v] + a.momentum * (var * (float)n / (float)max(n - 1, 1));
            if (v == 0 && a.batches_tracked) *a.batches_tracked += 1;
        }
    } else {
        mean = a.running_mean[v];
        invstd = rsqrtf(a.running_var[v] + a.eps);
    }
    if (threadIdx.x == 0) {
        a.save_mean[v] = mean;
        a.save_invstd[v] = invstd;
    }
    const float g = a.gamma ? a.gamma[v] : 1.f, bt = a.beta ? a.beta[v] : 0.f;
    const float scale = g * invstd, shift = fmaf(-mean, scale, bt);
    float *ov = a.out + (size_t)v * a.C;
    const float *rv = a.res ? a.res + (size_t)v * a.ld_res : nullptr;
    const size_t rstride = (size_t)a.V * a.ld_res;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int b = i / a.C, c = i % a.C;
        float y = fmaxf(fmaf(xv[(size_t)b * vstride + c], scale, shift), 0.f);
        if (rv) y = (rv[(size_t)b * rstride + c] + y) * 0.5f;
        ov[(size_t)b * vstride + c] = y;
    }
}

struct BwdArgs {
    const float *x, *dout;              // (B, V, C)
    const float *gamma, *beta;          // (V) or null
    const float *save_mean, *save_invstd;
    float *dx;                          // (B, V, C)
    float *dres;                        // (B, V, C) or null (no residual / residual needs no gradient)
    float *dgamma, *dbeta;              // (V) or null
    int B, V, C, training, has_res;
};

__global__ void __launch_bounds__(kThreads) bwd_kernel(const BwdArgs a)
{
    __shared__ float sh[kThreads / 32];
    const int v = blockIdx.x;
    const int n = a.B * a.C;
    const size_t vstride = (size_t)a.V * a.C;
    const size_t base = (size_t)v * a.C;
    const float mean = a.save_mean[v], invstd = a.save_invstd[v];
    const float g = a.gamma ? a.gamma[v] : 1.f, bt = a.beta ? a.beta[v] : 0.f;
    const float half = a.has_res ? 0.5f : 1.0f;
    const float scale = g * invstd, shift = fmaf(-mean, scale, bt);  // the forward's own expression decides the ReLU mask
    float sdy = 0.f, sdyx = 0.f;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const size_t o = base + (size_t)(i / a.C) * vstride + (i % a.C);
        const float xv = a.x[o];
        const float xh = (xv - mean) * invstd;
        const float dy = fmaf(xv, scale, shift) > 0.f ? a.dout[o] * half : 0.f;
        sdy += dy;
        sdyx = fmaf(dy, xh, sdyx);
    }
    sdy = block_sum(sdy, sh);
    sdyx = block_sum(sdyx, sh);
    if (threadIdx.x == 0) {
        if (a.dgamma) a.dgamma[v] = sdyx;
        if (a.dbeta) a.dbeta[v] = sdy;
    }
    const float inv_n = 1.f / (float)n;
    const float k1 = a.training ? sdy * inv_n : 0.f, k2 = a.training ? sdyx * inv_n : 0.f;
    const float gi = g * invstd;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const size_t o = base + (size_t)(i / a.C) * vstride + (i % a.C);
        const float xv = a.x[o];
        const float xh = (xv - mean) * invstd;
        const float d = a.dout[o];
        const float dy = fmaf(xv, scale, shift) > 0.f ? d * half : 0.f;
        a.dx[o] = gi * (dy - k1 - xh * k2);
        if (a.dres) a.dres[o] = d * 0.5f;
    }
}

}  // namespace bn
}  // namespace geom

extern "C" int geom_bn_relu_res_fwd(int b, int V, int C, const float *x, const float *res, int ld_res, const float *gamma,
                                    const float *beta, float *running_mean, float *running_var, long long *batches_tracked,
                                    int training, float momentum, float eps, float *save_mean, float *save_invstd, float *out,
                                    geom_stream_t stream)
{
    using namespace geom;
    if (b < 0 || V < 0 || C < 0 || !save_mean || !save_invstd) return GEOM_ERR_INVALID_ARG;
    if ((long long)b * V * C == 0) return GEOM_OK;
    if (!x || !out || (res && ld_res < C) || (!training && (!running_mean || !running_var))) return GEOM_ERR_INVALID_ARG;
    if ((long long)b * C > 0x7fffffffLL) return GEOM_ERR_INVALID_ARG;
    bn::FwdArgs a;
    a.x = x, a.res = res, a.gamma = gamma, a.beta = beta, a.running_mean = running_mean, a.running_var = running_var;
    a.save_mean = save_mean, a.save_invstd = save_invstd, a.out = out, a.batches_tracked = batches_tracked;
    a.B = b, a.V = V, a.C = C, a.ld_res = ld_res, a.training = training, a.eps = eps, a.momentum = momentum;
    cudaStream_t st = (cudaStream_t)stream;
    {
        prof::Scope ps(st, prof::kBnFwd);
        bn::fwd_kernel<<<V, bn::kThreads, 0, st>>>(a);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_bn_relu_res_bwd(int b, int V, int C, const float *x, const float *dout, const float *gamma, const float *beta,
                                    const float *save_mean, const float *save_invstd, int training, int has_res, float *dx,
                                    float *dres, float *dgamma, float *dbeta, geom_stream_t stream)
{
    using namespace geom;
    if (b < 0 || V < 0 || C < 0) return GEOM_ERR_INVALID_ARG;
    if ((long long)b * V * C == 0) return GEOM_OK;
    if (!x || !dout || !save_mean || !save_invstd || !dx) return GEOM_ERR_INVALID_ARG;
    bn::BwdArgs a;
    a.x = x, a.dout = dout, a.gamma = gamma, a.beta = beta, a.save_mean = save_mean, a.save_invstd = save_invstd;
    a.dx = dx, a.dres = dres, a.dgamma = dgamma, a.dbeta = dbeta;
    a.B = b, a.V = V, a.C = C, a.training = training, a.has_res = has_res;
    cudaStream_t st = (cudaStream_t)stream;
    {
        prof::Scope ps(st, prof::kBnBwd);
        bn::bwd_kernel<<<V, bn::kThreads, 0, st>>>(a);
    }
    return status_from_cuda(cudaGetLastError());
}

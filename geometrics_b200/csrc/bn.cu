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
// One CTA per vertex: the vertex' B rows of C floats (B*C*4 = 12 KB at the reference's B=16, C=192) are read once into
// registers (up to 8192 values per vertex; beyond that three passes over L2): sum, centred sum of squares, normalise --
// two-pass statistics like torch's, no E[x^2]-E[x]^2 cancellation.
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

// EPT > 0: the vertex' B*C values (<= 256 * EPT) are read ONCE into registers; EPT == 0: streaming fallback for any size
// (three passes over L2).  Element i of the vertex is (b, c) = (i / C, i % C); a thread walks i = tid, tid + 256, ... by
// stepping (b, c) with the precomputed (256 / C, 256 % C) -- no division in the loops.
template <int EPT>
__global__ void __launch_bounds__(kThreads) fwd_kernel(const FwdArgs a)
{
    __shared__ float sh[kThreads / 32];
    const int v = blockIdx.x;
    const int n = a.B * a.C;
    const size_t vstride = (size_t)a.V * a.C;
    const float *xv = a.x + (size_t)v * a.C;
    const int sb = kThreads / a.C, sc = kThreads % a.C;
    const int b0 = threadIdx.x / a.C, c0 = threadIdx.x % a.C;
    constexpr int R = EPT > 0 ? EPT : 1;
    float xr[R];
    float mean, invstd;
    if (EPT > 0) {
        int bb = b0, cc = c0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            xr[k] = (threadIdx.x + k * kThreads < n) ? xv[(size_t)bb * vstride + cc] : 0.f;
            bb += sb, cc += sc;
            if (cc >= a.C) cc -= a.C, ++bb;
        }
    }
    if (a.training) {
        float s = 0.f;
        if (EPT > 0) {
#pragma unroll
            for (int k = 0; k < R; ++k) s += xr[k];
        } else {
            int bb = b0, cc = c0;
            for (int i = threadIdx.x; i < n; i += kThreads) {
                s += xv[(size_t)bb * vstride + cc];
                bb += sb, cc += sc;
                if (cc >= a.C) cc -= a.C, ++bb;
            }
        }
        mean = block_sum(s, sh) / (float)n;
        float q = 0.f;
        if (EPT > 0) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const float d = (threadIdx.x + k * kThreads < n) ? xr[k] - mean : 0.f;
                q = fmaf(d, d, q);
            }
        } else {
            int bb = b0, cc = c0;
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const float d = xv[(size_t)bb * vstride + cc] - mean;
                q = fmaf(d, d, q);
                bb += sb, cc += sc;
                if (cc >= a.C) cc -= a.C, ++bb;
            }
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
    int bb = b0, cc = c0;
    if (EPT > 0) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (threadIdx.x + k * kThreads < n) {
                float y = fmaxf(fmaf(xr[k], scale, shift), 0.f);
                if (rv) y = (rv[(size_t)bb * rstride + cc] + y) * 0.5f;
                ov[(size_t)bb * vstride + cc] = y;
            }
            bb += sb, cc += sc;
            if (cc >= a.C) cc -= a.C, ++bb;
        }
    } else {
        for (int i = threadIdx.x; i < n; i += kThreads) {
            float y = fmaxf(fmaf(xv[(size_t)bb * vstride + cc], scale, shift), 0.f);
            if (rv) y = (rv[(size_t)bb * rstride + cc] + y) * 0.5f;
            ov[(size_t)bb * vstride + cc] = y;
            bb += sb, cc += sc;
            if (cc >= a.C) cc -= a.C, ++bb;
        }
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

template <int EPT>
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
    const int sb = kThreads / a.C, sc = kThreads % a.C;
    const int b0 = threadIdx.x / a.C, c0 = threadIdx.x % a.C;
    constexpr int R = EPT > 0 ? EPT : 1;
    float xr[R], dr[R];
    float sdy = 0.f, sdyx = 0.f;
    {
        int bb = b0, cc = c0;
        if (EPT > 0) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const bool in = threadIdx.x + k * kThreads < n;
                const size_t o = base + (size_t)bb * vstride + cc;
                xr[k] = in ? a.x[o] : 0.f;
                dr[k] = in ? a.dout[o] : 0.f;
                const float xh = (xr[k] - mean) * invstd;
                const float dy = (in && fmaf(xr[k], scale, shift) > 0.f) ? dr[k] * half : 0.f;
                sdy += dy;
                sdyx = fmaf(dy, xh, sdyx);
                bb += sb, cc += sc;
                if (cc >= a.C) cc -= a.C, ++bb;
            }
        } else {
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const size_t o = base + (size_t)bb * vstride + cc;
                const float xv = a.x[o];
                const float xh = (xv - mean) * invstd;
                const float dy = fmaf(xv, scale, shift) > 0.f ? a.dout[o] * half : 0.f;
                sdy += dy;
                sdyx = fmaf(dy, xh, sdyx);
                bb += sb, cc += sc;
                if (cc >= a.C) cc -= a.C, ++bb;
            }
        }
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
    int bb = b0, cc = c0;
    if (EPT > 0) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (threadIdx.x + k * kThreads < n) {
                const size_t o = base + (size_t)bb * vstride + cc;
                const float xh = (xr[k] - mean) * invstd;
                const float dy = fmaf(xr[k], scale, shift) > 0.f ? dr[k] * half : 0.f;
                a.dx[o] = gi * (dy - k1 - xh * k2);
                if (a.dres) a.dres[o] = dr[k] * 0.5f;
            }
            bb += sb, cc += sc;
            if (cc >= a.C) cc -= a.C, ++bb;
        }
    } else {
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const size_t o = base + (size_t)bb * vstride + cc;
            const float xv = a.x[o];
            const float xh = (xv - mean) * invstd;
            const float d = a.dout[o];
            const float dy = fmaf(xv, scale, shift) > 0.f ? d * half : 0.f;
            a.dx[o] = gi * (dy - k1 - xh * k2);
            if (a.dres) a.dres[o] = d * 0.5f;
            bb += sb, cc += sc;
            if (cc >= a.C) cc -= a.C, ++bb;
        }
    }
}

template <class Args, class F4, class F8, class F16, class F32, class F0>
static void dispatch_ept(long long n, F4 f4, F8 f8, F16 f16, F32 f32, F0 f0)
{
    if (n <= 4 * kThreads) f4();
    else if (n <= 8 * kThreads) f8();
    else if (n <= 16 * kThreads) f16();
    else if (n <= 32 * kThreads) f32();
    else f0();
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
        bn::dispatch_ept<bn::FwdArgs>((long long)b * C, [&] { bn::fwd_kernel<4><<<V, bn::kThreads, 0, st>>>(a); },
                                      [&] { bn::fwd_kernel<8><<<V, bn::kThreads, 0, st>>>(a); },
                                      [&] { bn::fwd_kernel<16><<<V, bn::kThreads, 0, st>>>(a); },
                                      [&] { bn::fwd_kernel<32><<<V, bn::kThreads, 0, st>>>(a); },
                                      [&] { bn::fwd_kernel<0><<<V, bn::kThreads, 0, st>>>(a); });
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
        bn::dispatch_ept<bn::BwdArgs>((long long)b * C, [&] { bn::bwd_kernel<4><<<V, bn::kThreads, 0, st>>>(a); },
                                      [&] { bn::bwd_kernel<8><<<V, bn::kThreads, 0, st>>>(a); },
                                      [&] { bn::bwd_kernel<16><<<V, bn::kThreads, 0, st>>>(a); },
                                      [&] { bn::bwd_kernel<32><<<V, bn::kThreads, 0, st>>>(a); },
                                      [&] { bn::bwd_kernel<0><<<V, bn::kThreads, 0, st>>>(a); });
    }
    return status_from_cuda(cudaGetLastError());
}

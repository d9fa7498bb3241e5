// gcn_bwd.cu -- backward of the 0N-GCN layer  Y = [adj . (XW)[:, :k] | (XW)[:, k:]] + bias  (autograd of reference
// layers.py:107-116 / 143-152 / 34-41), every GEMM of it written here (round 1 called cuBLAS through torch.matmul):
//     dS = [adj^T . dY[:, :k] | dY[:, k:]]          spmm_bwd_kernel (gcn.cu), CSR of adj^T
//     db = column sums of dY                        colsum_kernel
//     dW = X^T dS        (in x M) . (M x out)       gemm_tn_kernel: the long M = B*V reduction is split over CTAs
//     dX = dS W^T        (M x out) . (out x in)     the forward's own feature GEMM (tensor cores when in <= 256) on the
//                                                   transposed weight
#include <algorithm>

#include "common.cuh"

namespace geom {
namespace gcn_bwd {

// db[n] = sum_m dY[m, n]
__global__ void __launch_bounds__(256) colsum_kernel(int M, int N, const float *__restrict__ dY, float *__restrict__ db)
{
    // block = 8 row lanes x 32 columns; grid.x over column groups of 32, grid.y over row slices
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r0 = threadIdx.x >> 5;
    float s = 0.f;
    if (col < N)
        for (int m = blockIdx.y * 8 + r0; m < M; m += gridDim.y * 8) s += dY[(size_t)m * N + col];
    __shared__ float sh[8][33];
    sh[r0][threadIdx.x & 31] = s;
    __syncthreads();
    if (r0 == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x & 31];
        atomicAdd(db + col, t);
    }
}

// out[c][r] = in[r][c]
__global__ void __launch_bounds__(256) transpose_kernel(int R, int C, const float *__restrict__ in, float *__restrict__ out)
{
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.x >> 5; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + (threadIdx.x & 31);
        t[i][threadIdx.x & 31] = (r < R && c < C) ? in[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x >> 5; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + (threadIdx.x & 31);
        if (r < R && c < C) out[(size_t)c * R + r] = t[threadIdx.x & 31][i];
    }
}

// dW partial: P[slice][k][n] = sum over the slice's rows m of X[m][k] * dS[m][n].  128 (k) x 64 (n) output tile per CTA,
// 8 x 4 per thread (256 threads), 8 rows of m per shared-memory step; a second kernel adds the slices in a fixed order
// (deterministic, no atomics).
constexpr int TK = 128, TN = 64, TM = 8;
__global__ void __launch_bounds__(256) gemm_tn_kernel(int M, int K, int N, const float *__restrict__ X, const float *__restrict__ dS,
                                                      float *__restrict__ P, int rows_per_slice)
{
    __shared__ __align__(16) float sx[2][TM][TK], sd[2][TM][TN];
    const int k0 = blockIdx.x * TK, n0 = blockIdx.y * TN;
    const int m_begin = blockIdx.z * rows_per_slice, m_end = min(M, m_begin + rows_per_slice);
    const int tk = (threadIdx.x >> 4) * 8, tn = (threadIdx.x & 15) * 4;
    // staging: thread t copies X[m0 + t/32][k0 + (t%32)*4 .. +3] and dS[m0 + t/16 (first 128 threads)][n0 + (t%16)*4 .. +3]
    const int xr = threadIdx.x >> 5, xc = (threadIdx.x & 31) * 4;
    const int dr = (threadIdx.x & 127) >> 4, dc = (threadIdx.x & 15) * 4;
    const bool vec = (K % 4 == 0) && (N % 4 == 0);
    auto load_x = [&](int m0) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int m = m0 + xr;
        if (m < m_end) {
            const float *p = X + (size_t)m * K + k0 + xc;
            if (vec && k0 + xc + 3 < K) v = *reinterpret_cast<const float4 *>(p);
            else {
                if (k0 + xc < K) v.x = p[0];
                if (k0 + xc + 1 < K) v.y = p[1];
                if (k0 + xc + 2 < K) v.z = p[2];
                if (k0 + xc + 3 < K) v.w = p[3];
            }
        }
        return v;
    };
    auto load_d = [&](int m0) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int m = m0 + dr;
        if (threadIdx.x < 128 && m < m_end) {
            const float *p = dS + (size_t)m * N + n0 + dc;
            if (vec && n0 + dc + 3 < N) v = *reinterpret_cast<const float4 *>(p);
            else {
                if (n0 + dc < N) v.x = p[0];
                if (n0 + dc + 1 < N) v.y = p[1];
                if (n0 + dc + 2 < N) v.z = p[2];
                if (n0 + dc + 3 < N) v.w = p[3];
            }
        }
        return v;
    };
    float acc[8][4] = {};
    float4 vx = load_x(m_begin), vd = load_d(m_begin);
    int buf = 0;
    for (int m0 = m_begin; m0 < m_end; m0 += TM) {
        *reinterpret_cast<float4 *>(&sx[buf][xr][xc]) = vx;
        if (threadIdx.x < 128) *reinterpret_cast<float4 *>(&sd[buf][dr][dc]) = vd;
        __syncthreads();
        if (m0 + TM < m_end) {  // next step's operands travel while this one is multiplied
            vx = load_x(m0 + TM);
            vd = load_d(m0 + TM);
        }
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            const float4 xa = *reinterpret_cast<const float4 *>(&sx[buf][r][tk]);
            const float4 xb = *reinterpret_cast<const float4 *>(&sx[buf][r][tk + 4]);
            const float4 da = *reinterpret_cast<const float4 *>(&sd[buf][r][tn]);
            const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w}, ds[4] = {da.x, da.y, da.z, da.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xs[i], ds[j], acc[i][j]);
        }
        buf ^= 1;  // two buffers: the next store cannot overtake this step's reads (one barrier per step)
    }
    float *out = P + (size_t)blockIdx.z * K * N;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + tk + i < K && n0 + tn + j < N) out[(size_t)(k0 + tk + i) * N + n0 + tn + j] = acc[i][j];
}

// dW[e] = sum over slices of P[slice][e], slices in ascending order
__global__ void __launch_bounds__(256) slice_sum_kernel(int slices, long long elems, const float *__restrict__ P, float *__restrict__ dW)
{
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= elems) return;
    float s = 0.f;
    for (int z = 0; z < slices; ++z) s += P[(size_t)z * elems + e];
    dW[e] = s;
}

}  // namespace gcn_bwd
}  // namespace geom

using namespace geom;

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
constexpr int kMaxSlices = 32;  // M slices of the dW GEMM

extern "C" size_t geom_zeron_gcn_bwd_workspace_bytes(int b, int V, int in_features, int out_features)
{
    if (b < 0 || V < 0 || in_features < 0 || out_features < 0) return 0;
    const size_t M = (size_t)b * V;
    // dS | W^T | dW slice partials | workspace of the dX GEMM (out -> in features, nothing smoothed)
    return align256(M * out_features * 4) + align256((size_t)in_features * out_features * 4) +
           align256((size_t)kMaxSlices * in_features * out_features * 4) +
           geom_zeron_gcn_tensor_workspace_bytes(b, V, out_features, in_features, 0) + 256;
}

extern "C" int geom_zeron_gcn_bwd(int b, int V, int in_features, int out_features, int ksplit, const float *X, const float *W,
                                  const float *dY, const int *t_row_ptr, const int *t_col_idx, const float *t_val, float *dX,
                                  float *dW, float *db, int flags, void *workspace, size_t workspace_bytes, geom_stream_t stream)
{
    if (b < 0 || V < 0 || in_features < 0 || out_features < 0 || ksplit < 0 || ksplit > out_features) return GEOM_ERR_INVALID_ARG;
    const long long M = (long long)b * V;
    if (M == 0 || out_features == 0 || in_features == 0) return GEOM_OK;
    if (M > 0x7fffffffLL || !dY || (dW && !X) || (dX && !W)) return GEOM_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < geom_zeron_gcn_bwd_workspace_bytes(b, V, in_features, out_features) ||
        ((uintptr_t)workspace & 15u))
        return GEOM_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char *w = (char *)workspace;
    float *dS = (float *)w;
    float *Wt = (float *)(w + align256((size_t)M * out_features * 4));
    float *P = (float *)((char *)Wt + align256((size_t)in_features * out_features * 4));
    void *gemm_ws = (char *)P + align256((size_t)kMaxSlices * in_features * out_features * 4);
    const size_t gemm_ws_bytes = workspace_bytes - (size_t)((char *)gemm_ws - w);
    int rc = GEOM_OK;
    cudaError_t e = cudaSuccess;
    if (db) {
        if ((e = cudaMemsetAsync(db, 0, sizeof(float) * out_features, st)) != cudaSuccess) return status_from_cuda(e);
        prof::Scope ps(st, prof::kGcnBwd);
        const int slices = (int)std::min<long long>(64, (M + 63) / 64);
        gcn_bwd::colsum_kernel<<<dim3((out_features + 31) / 32, slices), 256, 0, st>>>((int)M, out_features, dY, db);
    }
    if (!dX && !dW) return status_from_cuda(cudaGetLastError());
    const float *dSp = dY;  // nothing smoothed: dS == dY
    if (ksplit > 0) {
        if ((rc = geom_zeron_gcn_bwd_smooth(b, V, out_features, ksplit, t_row_ptr, t_col_idx, t_val, dY, dS, stream)) != GEOM_OK) return rc;
        dSp = dS;
    }
    if (dW) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int tiles = ((in_features + gcn_bwd::TK - 1) / gcn_bwd::TK) * ((out_features + gcn_bwd::TN - 1) / gcn_bwd::TN);
        int slices = std::max(1, (2 * sms + tiles - 1) / tiles);                                  // ~2 CTAs per SM
        slices = (int)std::min<long long>(std::min(slices, kMaxSlices), (M + 63) / 64);           // >= 64 rows per slice
        const int rows = (int)((((M + slices - 1) / slices) + gcn_bwd::TM - 1) / gcn_bwd::TM * gcn_bwd::TM);
        slices = (int)((M + rows - 1) / rows);
        const long long elems = (long long)in_features * out_features;
        prof::Scope ps(st, prof::kGcnBwdGemm);
        gcn_bwd::gemm_tn_kernel<<<dim3((in_features + gcn_bwd::TK - 1) / gcn_bwd::TK, (out_features + gcn_bwd::TN - 1) / gcn_bwd::TN, slices),
                                  256, 0, st>>>((int)M, in_features, out_features, X, dSp, P, rows);
        gcn_bwd::slice_sum_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(slices, elems, P, dW);
    }
    if (dX) {
        gcn_bwd::transpose_kernel<<<dim3((out_features + 31) / 32, (in_features + 31) / 32), 256, 0, st>>>(in_features, out_features, W, Wt);
        if ((e = cudaGetLastError()) != cudaSuccess) return status_from_cuda(e);
        // dX (M x in) = dS (M x out) . W^T (out x in): the forward GEMM with nothing smoothed and no bias
        rc = geom_zeron_gcn_fwd(b, V, out_features, in_features, 0, dSp, Wt, nullptr, nullptr, nullptr, nullptr, dX, flags & GEOM_GCN_TENSOR,
                                gemm_ws, gemm_ws_bytes, stream);
        if (rc != GEOM_OK) return rc;
    }
    return status_from_cuda(cudaGetLastError());
}

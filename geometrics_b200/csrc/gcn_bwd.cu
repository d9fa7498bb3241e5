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

// dW[k][n] += sum over this CTA's slice of m of X[m][k] * dS[m][n]: 64 x 64 output tile, 4 x 4 per thread, 16 rows of m per
// shared-memory step; the partial tiles of the M slices meet in fp32 atomics (dW zeroed by the entry point).
constexpr int TK = 64, TN = 64, TM = 16;
__global__ void __launch_bounds__(256) gemm_tn_kernel(int M, int K, int N, const float *__restrict__ X, const float *__restrict__ dS,
                                                      float *__restrict__ dW, int rows_per_slice)
{
    __shared__ __align__(16) float sx[TM][TK + 4], sd[TM][TN + 4];
    const int k0 = blockIdx.x * TK, n0 = blockIdx.y * TN;
    const int m_begin = blockIdx.z * rows_per_slice, m_end = min(M, m_begin + rows_per_slice);
    const int tk = (threadIdx.x >> 4) * 4, tn = (threadIdx.x & 15) * 4;
    float acc[4][4] = {};
    for (int m0 = m_begin; m0 < m_end; m0 += TM) {
        // 16 x 64 of each operand: 1024 floats, 4 per thread, coalesced along the row
        for (int e = threadIdx.x; e < TM * TK; e += 256) {
            const int r = e / TK, c = e % TK;
            const int m = m0 + r;
            sx[r][c] = (m < m_end && k0 + c < K) ? X[(size_t)m * K + k0 + c] : 0.f;
            sd[r][c] = (m < m_end && n0 + c < N) ? dS[(size_t)m * N + n0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            const float4 xa = *reinterpret_cast<const float4 *>(&sx[r][tk]);
            const float4 da = *reinterpret_cast<const float4 *>(&sd[r][tn]);
            const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ds[4] = {da.x, da.y, da.z, da.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xs[i], ds[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + tk + i < K && n0 + tn + j < N) atomicAdd(dW + (size_t)(k0 + tk + i) * N + n0 + tn + j, acc[i][j]);
}

}  // namespace gcn_bwd
}  // namespace geom

using namespace geom;

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t geom_zeron_gcn_bwd_workspace_bytes(int b, int V, int in_features, int out_features)
{
    if (b < 0 || V < 0 || in_features < 0 || out_features < 0) return 0;
    const size_t M = (size_t)b * V;
    // dS | W^T | workspace of the dX GEMM (out -> in features, nothing smoothed)
    return align256(M * out_features * 4) + align256((size_t)in_features * out_features * 4) +
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
    void *gemm_ws = (char *)Wt + align256((size_t)in_features * out_features * 4);
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
        if ((e = cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)in_features * out_features, st)) != cudaSuccess) return status_from_cuda(e);
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int tiles = ((in_features + gcn_bwd::TK - 1) / gcn_bwd::TK) * ((out_features + gcn_bwd::TN - 1) / gcn_bwd::TN);
        int slices = std::max(1, (3 * sms + tiles - 1) / tiles);                        // ~3 CTAs per SM
        slices = (int)std::min<long long>(slices, (M + 4 * gcn_bwd::TM - 1) / (4 * gcn_bwd::TM));  // >= 64 rows per slice
        const int rows = (int)((((M + slices - 1) / slices) + gcn_bwd::TM - 1) / gcn_bwd::TM * gcn_bwd::TM);
        prof::Scope ps(st, prof::kGcnBwdGemm);
        gcn_bwd::gemm_tn_kernel<<<dim3((in_features + gcn_bwd::TK - 1) / gcn_bwd::TK, (out_features + gcn_bwd::TN - 1) / gcn_bwd::TN,
                                       (unsigned)((M + rows - 1) / rows)),
                                  256, 0, st>>>((int)M, in_features, out_features, X, dSp, dW, rows);
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

// gcn.cu -- 0N-GCN layer: feature GEMM + sparse-adjacency smoothing of the first k output columns.
//
// Replaces the torch chain of the reference's graph layers (layers.py:34-41 ZERON_GCN,
// :107-116 Batch_Image_ZERON_GCNGCN, :143-152 BatchZERON_GCN, :176-187 BatchGCNMax):
//     support = X W ;  Y = [ adj . support[:, :k]  |  support[:, k:] ] + bias
// where the reference multiplies by a DENSE (V,V) adjacency (2 B V^2 k flops, 35.7 GFLOP per layer at
// BASELINE config 4) although a mesh row has ~7 non-zeros.  Here the adjacency is converted once to
// CSR (csr_* kernels), the GEMM writes columns >= k straight into Y with the bias fused and columns
// < k into a compact scratch T, and one SpMM pass produces Y[:, :k] = adj . T + bias.  No torch.cat,
// no strided-slice copy.
//
// The GEMM is fp32 (parity target 1e-5 relative against the fp32 reference rules out plain TF32)
// on the CUDA cores with packed FFMA2: 128x64 tile per 128-thread CTA, 8x8 outputs per thread, A
// staged k-major and duplicated {v,v} in shared memory so that one 64-bit operand feeds a packed FMA
// against two adjacent B columns (half the issue slots of scalar FFMA).
#include "common.cuh"

namespace geom {
namespace gcn {

// ------------------------------------------------------------------------------------------
// dense (V,V) -> CSR.  Element (i,j) is read at adj[i*rs + j*cs], so the same kernels build the
// CSR of the transpose (rs=1, cs=V) for the backward pass.  Column indices ascend within a row.
// ------------------------------------------------------------------------------------------
__global__ void csr_count_kernel(int V, const float *__restrict__ adj, long long rs, long long cs,
                                 int *__restrict__ row_cnt)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= V) return;
    int cnt = 0;
    for (int j = lane; j < V; j += 32) cnt += adj[row * rs + j * cs] != 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) row_cnt[row] = cnt;
}

// exclusive scan of row_cnt[0..V) into row_ptr[0..V]; single block.
__global__ void csr_scan_kernel(int V, const int *__restrict__ row_cnt, int *__restrict__ row_ptr)
{
    __shared__ int carry;
    __shared__ int buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < V; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < V ? row_cnt[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < V) row_ptr[i] = carry + buf[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += buf[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) row_ptr[V] = carry;
}

__global__ void csr_fill_kernel(int V, const float *__restrict__ adj, long long rs, long long cs,
                                const int *__restrict__ row_ptr, int *__restrict__ col_idx,
                                float *__restrict__ val)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= V) return;
    int out = row_ptr[row];
    for (int j0 = 0; j0 < V; j0 += 32) {
        const int j = j0 + lane;
        const float a = j < V ? adj[row * rs + j * cs] : 0.0f;
        const unsigned mask = __ballot_sync(0xffffffffu, a != 0.0f);
        if (a != 0.0f) {
            const int pos = out + __popc(mask & ((1u << lane) - 1u));
            col_idx[pos] = j;
            val[pos] = a;
        }
        out += __popc(mask);
    }
}

// ------------------------------------------------------------------------------------------
// feature GEMM  S = X (M,K) . W (K,N), columns < ksplit -> T (M, ldt), columns >= ksplit -> Y + bias
// ------------------------------------------------------------------------------------------
constexpr int BN = 64, BK = 16, GT = 128, APAD = 4;

struct GemmArgs {
    const float *X;     // (M, K) row-major
    const float *W;     // (K, N) row-major
    const float *bias;  // (N) or null
    float *Y;           // (M, N)
    float *T;           // (M, ldt) scratch for the smoothed columns, or null when ksplit == 0
    int M, K, N, ksplit, ldt, relu;
};

// BMT x 64 output tile per 128-thread CTA (BMT = 128: 8x8 outputs per thread; 64: 4x8, used when the
// 128-row grid would not fill the GPU).  Thread (ty, tx) owns rows {ty*4..+3} (+64) and columns
// {tx*4..+3} and {32+tx*4..+3}: every LDS.128 of a warp covers one contiguous 64/128-byte span, so
// shared-memory reads are conflict-free (4 wavefronts per k-step per warp for 32 FFMA2).
template <int BMT, bool VEC>
__global__ void __launch_bounds__(GT, 3) gemm_kernel(const GemmArgs g)
{
    constexpr int RH = BMT / 64;  // row halves per thread (1 or 2)
    constexpr int RT = 4 * RH;    // rows per thread
    constexpr int AL = BMT / 32;  // A-tile float4 loads per thread
    __shared__ __align__(16) float As[2][BK][BMT + APAD];  // k-major
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int ty = tid >> 3, tx = tid & 7;
    const int m0 = blockIdx.x * BMT, n0 = blockIdx.y * BN;

    // global -> register staging: A rows (tid>>2) + 32*i with k-offset (tid&3)*4; B k = (tid>>4) + 8*i, cols (tid&15)*4
    const int a_row = tid >> 2, a_k = (tid & 3) * 4;
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;
    float4 ra[AL], rb[2];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < AL; ++i) {
            const int r = m0 + a_row + 32 * i, k = k0 + a_k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < g.M) {
                const float *src = g.X + (size_t)r * g.K + k;
                if (VEC && k + 3 < g.K) {
                    v = *reinterpret_cast<const float4 *>(src);
                } else {
                    if (k + 0 < g.K) v.x = src[0];
                    if (k + 1 < g.K) v.y = src[1];
                    if (k + 2 < g.K) v.z = src[2];
                    if (k + 3 < g.K) v.w = src[3];
                }
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int k = k0 + b_k + 8 * i, n = n0 + b_n;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < g.K) {
                const float *src = g.W + (size_t)k * g.N + n;
                if (VEC && n + 3 < g.N) {
                    v = *reinterpret_cast<const float4 *>(src);
                } else {
                    if (n + 0 < g.N) v.x = src[0];
                    if (n + 1 < g.N) v.y = src[1];
                    if (n + 2 < g.N) v.z = src[2];
                    if (n + 3 < g.N) v.w = src[3];
                }
            }
            rb[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < AL; ++i) {
            const int r = a_row + 32 * i;
            As[buf][a_k + 0][r] = ra[i].x;
            As[buf][a_k + 1][r] = ra[i].y;
            As[buf][a_k + 2][r] = ra[i].z;
            As[buf][a_k + 3][r] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) *reinterpret_cast<float4 *>(&Bs[buf][b_k + 8 * i][b_n]) = rb[i];
    };

    u64 acc[RT][4];  // [row][column pair]: pairs (0,1),(2,3) of the first column group, then of the second
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;

    const int nk = (g.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[RT];
#pragma unroll
            for (int h = 0; h < RH; ++h) {
                const float4 a4 = *reinterpret_cast<const float4 *>(&As[buf][k][h * 64 + ty * 4]);
                av[h * 4 + 0] = a4.x, av[h * 4 + 1] = a4.y, av[h * 4 + 2] = a4.z, av[h * 4 + 3] = a4.w;
            }
            const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][k][tx * 4]);
            const ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][k][32 + tx * 4]);
            const u64 b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const u64 a2 = pack2(av[i], av[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma2(a2, b[j], acc[i][j]);
            }
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int r = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (r >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[2];
            unpack2(acc[i][j], v[0], v[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = n0 + (j >> 1) * 32 + tx * 4 + (j & 1) * 2 + h;
                if (c >= g.N) continue;
                if (c < g.ksplit) {
                    g.T[(size_t)r * g.ldt + c] = v[h];
                } else {
                    float o = g.bias ? v[h] + g.bias[c] : v[h];
                    if (g.relu) o = fmaxf(o, 0.f);
                    g.Y[(size_t)r * g.N + c] = o;
                }
            }
        }
    }
}

// skinny outputs (N <= 16, e.g. the 192 -> 3 coordinate head): one warp per row.
constexpr int kSkinnyMaxN = 16;
__global__ void __launch_bounds__(256) skinny_kernel(const GemmArgs g)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= g.M) return;
    float acc[kSkinnyMaxN];
#pragma unroll
    for (int c = 0; c < kSkinnyMaxN; ++c) acc[c] = 0.f;
    const float *x = g.X + (size_t)row * g.K;
    for (int k = lane; k < g.K; k += 32) {
        const float xv = x[k];
        const float *w = g.W + (size_t)k * g.N;
#pragma unroll
        for (int c = 0; c < kSkinnyMaxN; ++c)
            if (c < g.N) acc[c] = fmaf(xv, w[c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < kSkinnyMaxN; ++c) {
        if (c >= g.N) break;
        float v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            if (c < g.ksplit) {
                g.T[(size_t)row * g.ldt + c] = v;
            } else {
                float o = g.bias ? v + g.bias[c] : v;
                if (g.relu) o = fmaxf(o, 0.f);
                g.Y[(size_t)row * g.N + c] = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// SpMM over the smoothed columns:  Y[b,i,c] = bias[c] + sum_j val(i,j) T[b,j,c],  c < ksplit.
// One warp per (b,i); lanes stride the columns (coalesced T rows).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spmm_kernel(int rows, int V, int N, int ksplit, int ldt, const int *__restrict__ row_ptr,
            const int *__restrict__ col_idx, const float *__restrict__ val, const float *__restrict__ T,
            const float *__restrict__ bias, float *__restrict__ Y, int relu)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int bi = row / V, i = row - bi * V;
    const int p0 = row_ptr[i], p1 = row_ptr[i + 1];
    const float *Tb = T + (size_t)bi * V * ldt;
    // Four columns per lane: the T rows are ldt = round4(ksplit) floats, 16-byte aligned, so a neighbour's row is
    // read with 128-bit loads (22 lanes cover k = 85).  Each column keeps the scalar kernel's fmaf chain over the
    // neighbours in CSR order, so the result is bit-identical to a one-column-per-lane sweep.  Columns
    // ksplit..ldt-1 of T are never written by the GEMM: they are accumulated but not stored.
    const bool vec_store = (N & 3) == 0 && (((uintptr_t)Y & 15u) == 0);
    for (int q = lane; q < (ldt >> 2); q += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = p0; p < p1; ++p) {
            const float w = val[p];
            const float4 t = *reinterpret_cast<const float4 *>(Tb + (size_t)col_idx[p] * ldt + 4 * q);
            acc.x = fmaf(w, t.x, acc.x), acc.y = fmaf(w, t.y, acc.y), acc.z = fmaf(w, t.z, acc.z), acc.w = fmaf(w, t.w, acc.w);
        }
        const int c = 4 * q;
        float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (bias && c + j < ksplit) o[j] += bias[c + j];
            if (relu) o[j] = fmaxf(o[j], 0.f);
        }
        float *dst = Y + (size_t)row * N + c;
        if (vec_store && c + 3 < ksplit) {
            *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + j < ksplit) dst[j] = o[j];
        }
    }
}

// Backward of the concat/smoothing: dS[:, :k] = adj^T . dY[:, :k] (CSR of the transpose),
// dS[:, k:] = dY[:, k:].  One warp per (b,i).
__global__ void __launch_bounds__(256)
spmm_bwd_kernel(int rows, int V, int N, int ksplit, const int *__restrict__ t_row_ptr,
                const int *__restrict__ t_col_idx, const float *__restrict__ t_val,
                const float *__restrict__ dY, float *__restrict__ dS)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int bi = row / V, i = row - bi * V;
    const int p0 = t_row_ptr[i], p1 = t_row_ptr[i + 1];
    const float *dYb = dY + (size_t)bi * V * N;
    for (int c = lane; c < N; c += 32) {
        float acc;
        if (c < ksplit) {
            acc = 0.f;
            for (int p = p0; p < p1; ++p) acc = fmaf(t_val[p], dYb[(size_t)t_col_idx[p] * N + c], acc);
        } else {
            acc = dY[(size_t)row * N + c];
        }
        dS[(size_t)row * N + c] = acc;
    }
}

// Plain CSR SpMM over every feature column: Y[b,i,c] = sum_j val(i,j) X[b,j,c]  (Laplacian
// regulariser, utils.py:654-662).  One thread per output element.
__global__ void __launch_bounds__(256)
spmm_plain_kernel(long long total, int V, int C, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                  const float *__restrict__ val, const float *__restrict__ X, float *__restrict__ Y)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C);
    const long long row = t / C;
    const int i = (int)(row % V);
    const float *Xb = X + (row - i) * C;
    float acc = 0.f;
    for (int p = row_ptr[i]; p < row_ptr[i + 1]; ++p) acc = fmaf(val[p], Xb[(size_t)col_idx[p] * C + c], acc);
    Y[t] = acc;
}

}  // namespace gcn
}  // namespace geom

using namespace geom;

extern "C" int geom_csr_spmm(int b, int V, int C, const int *row_ptr, const int *col_idx, const float *val,
                             const float *X, float *Y, geom_stream_t stream)
{
    if (b < 0 || V < 0 || C < 0) return GEOM_ERR_INVALID_ARG;
    const long long total = (long long)b * V * C;
    if (total == 0) return GEOM_OK;
    if (!row_ptr || !col_idx || !val || !X || !Y) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        prof::Scope ps(st, prof::kMisc2);
        gcn::spmm_plain_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, V, C, row_ptr, col_idx, val, X, Y);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_csr_count(int V, const float *adj, int transpose, int *row_cnt, int *row_ptr,
                              geom_stream_t stream)
{
    if (V < 0) return GEOM_ERR_INVALID_ARG;
    if (!row_ptr || (V > 0 && (!adj || !row_cnt))) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rs = transpose ? 1 : V, cs = transpose ? V : 1;
    if (V > 0) gcn::csr_count_kernel<<<(V + 7) / 8, 256, 0, st>>>(V, adj, rs, cs, row_cnt);
    gcn::csr_scan_kernel<<<1, 1024, 0, st>>>(V, row_cnt, row_ptr);
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_csr_fill(int V, const float *adj, int transpose, const int *row_ptr, int *col_idx,
                             float *val, geom_stream_t stream)
{
    if (V < 0) return GEOM_ERR_INVALID_ARG;
    if (V == 0) return GEOM_OK;
    if (!adj || !row_ptr || !col_idx || !val) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rs = transpose ? 1 : V, cs = transpose ? V : 1;
    gcn::csr_fill_kernel<<<(V + 7) / 8, 256, 0, st>>>(V, adj, rs, cs, row_ptr, col_idx, val);
    return status_from_cuda(cudaGetLastError());
}

extern "C" size_t geom_zeron_gcn_workspace_bytes(int b, int V, int ksplit)
{
    if (b < 0 || V < 0 || ksplit < 0) return 0;
    const size_t ldt = ((size_t)ksplit + 3) & ~(size_t)3;
    return sizeof(float) * (size_t)b * V * ldt;
}

extern "C" size_t geom_zeron_gcn_tensor_workspace_bytes(int b, int V, int in_features, int out_features, int ksplit)
{
    if (b < 0 || V < 0 || in_features < 0 || out_features < 0 || ksplit < 0) return 0;
    const size_t t = (geom_zeron_gcn_workspace_bytes(b, V, ksplit) + 255) & ~(size_t)255;
    return t + gcn_tc::workspace_bytes(in_features, out_features);
}

extern "C" int geom_zeron_gcn_fwd(int b, int V, int in_features, int out_features, int ksplit,
                                  const float *X, const float *W, const float *bias,
                                  const int *row_ptr, const int *col_idx, const float *val, float *Y,
                                  int flags, void *workspace, size_t workspace_bytes, geom_stream_t stream)
{
    using namespace geom::gcn;
    const int relu = flags & 1;
    if (b < 0 || V < 0 || in_features < 0 || out_features < 0 || ksplit < 0 || ksplit > out_features)
        return GEOM_ERR_INVALID_ARG;
    const long long M = (long long)b * V;
    if (M == 0 || out_features == 0) return GEOM_OK;
    if (M > 0x7fffffffLL || M * (long long)max(in_features, out_features) > (1LL << 40)) return GEOM_ERR_INVALID_ARG;
    if (!X || !W || !Y) return GEOM_ERR_INVALID_ARG;
    if (ksplit > 0 && (!row_ptr || !col_idx || !val)) return GEOM_ERR_INVALID_ARG;
    const size_t need = geom_zeron_gcn_workspace_bytes(b, V, ksplit);
    if (ksplit > 0 && (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 15u))) return GEOM_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;

    GemmArgs g;
    g.X = X;
    g.W = W;
    g.bias = bias;
    g.Y = Y;
    g.T = (float *)workspace;
    g.M = (int)M;
    g.K = in_features;
    g.N = out_features;
    g.ksplit = ksplit;
    g.ldt = (ksplit + 3) & ~3;
    g.relu = relu;
    const bool tensor = (flags & 2) && gcn_tc::supported(out_features) && in_features >= 8;
    if (tensor) {
        const size_t tneed = geom_zeron_gcn_tensor_workspace_bytes(b, V, in_features, out_features, ksplit);
        if (!workspace || workspace_bytes < tneed || ((uintptr_t)workspace & 15u)) return GEOM_ERR_WORKSPACE;
        void *wsplit = (char *)workspace + ((need + 255) & ~(size_t)255);
        prof::Scope ps(st, prof::kGcnGemm);
        cudaError_t e = gcn_tc::launch((int)M, in_features, out_features, X, W, bias, Y, g.T, ksplit, g.ldt, relu, wsplit, st);
        if (e != cudaSuccess) return status_from_cuda(e);
    } else {
        prof::Scope ps(st, prof::kGcnGemm);
        if (out_features <= kSkinnyMaxN) {
            skinny_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(g);
        } else {
            const bool vec = (in_features % 4 == 0) && (out_features % 4 == 0) && (((uintptr_t)X | (uintptr_t)W) & 15u) == 0;
            const unsigned ny = (unsigned)((out_features + BN - 1) / BN);
            int dev = 0, sms = 148;
            if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            // 128-row tiles when they give the GPU at least ~2 waves of CTAs (3 resident per SM), else 64-row tiles
            const bool big = ((M + 127) / 128) * (long long)ny >= 6LL * sms;
            if (big) {
                dim3 grid((unsigned)((M + 127) / 128), ny);
                if (vec)
                    gemm_kernel<128, true><<<grid, GT, 0, st>>>(g);
                else
                    gemm_kernel<128, false><<<grid, GT, 0, st>>>(g);
            } else {
                dim3 grid((unsigned)((M + 63) / 64), ny);
                if (vec)
                    gemm_kernel<64, true><<<grid, GT, 0, st>>>(g);
                else
                    gemm_kernel<64, false><<<grid, GT, 0, st>>>(g);
            }
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return status_from_cuda(e);
    if (ksplit > 0) {
        prof::Scope ps(st, prof::kGcnSpmm);
        spmm_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>((int)M, V, out_features, ksplit, g.ldt, row_ptr, col_idx,
                                                             val, g.T, bias, Y, relu);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_zeron_gcn_bwd_smooth(int b, int V, int out_features, int ksplit, const int *t_row_ptr,
                                         const int *t_col_idx, const float *t_val, const float *dY,
                                         float *dS, geom_stream_t stream)
{
    using namespace geom::gcn;
    if (b < 0 || V < 0 || out_features < 0 || ksplit < 0 || ksplit > out_features) return GEOM_ERR_INVALID_ARG;
    const long long M = (long long)b * V;
    if (M == 0 || out_features == 0) return GEOM_OK;
    if (M > 0x7fffffffLL) return GEOM_ERR_INVALID_ARG;
    if (!dY || !dS || (ksplit > 0 && (!t_row_ptr || !t_col_idx || !t_val))) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        prof::Scope ps(st, prof::kGcnBwd);
        spmm_bwd_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>((int)M, V, out_features, ksplit, t_row_ptr, t_col_idx,
                                                                 t_val, dY, dS);
    }
    return status_from_cuda(cudaGetLastError());
}

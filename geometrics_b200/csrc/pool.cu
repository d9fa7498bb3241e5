// pool.cu -- vertex feature pooling: bilinear gather of CNN feature maps at projected vertex positions
// (reference: batched_pooling, utils.py:316-389, called before every deformation block,
// GEOMetrics.py:119,124,129).  SURVEY.md 8f-1.
//
// The reference runs, per feature map, 4 index_select gathers over a (C, B*dim*dim) view, 4 permutes,
// 8 broadcast multiplies, 3 adds and a cat: ~60 launches and ~10 full-size temporaries per call.  Here one
// launch gathers every feature map and writes straight into the caller's (B, V, ld) feature buffer at a column
// offset (so the torch.cat with the vertex positions, models.py:239, can be skipped too).
//
// Semantics kept bit for bit (given the same projected coordinates xs, ys):
//   cx = clamp(xs*dim, 0, dim-1), cy likewise;  x1 = floor(cx), x2 = ceil(cx), y1, y2 likewise;
//   A = x2-cx, B = cx-x1, G = y2-cy, H = cy-y1;
//   f = ((A*C)*G + (H*D)*A + (G*E)*B) + (B*F)*H      with C,D,E,F = map[x1,y1], [x1,y2], [x2,y1], [x2,y2]
// including the reference's quirk that an integral cx (e.g. a vertex clamped to the image border) has
// x1 == x2, hence A == B == 0 and a ZERO feature (utils.py:341-346).  Every product / sum is rounded separately
// (no FMA contraction), in the reference's association order.
//
// Layout: maps are NCHW fp32 (B, C, dim, dim); a CTA owns 32 consecutive vertices of one mesh and 128 channels of
// one map (grid.z walks the (map, channel chunk) pairs); each warp takes
// channels round-robin with lanes = vertices (the 4 corner reads of neighbouring vertices fall into the same
// few rows of one 12 KB channel plane: L1/L2 hits), results go through a 32x33 shared-memory tile so that the
// output rows are written as 128-byte segments.
// Backward: d(map) by atomicAdd on the 4 corners (the reference's index_select backward is the same
// scatter-add), d(xs), d(ys) through the bilinear weights (floor / ceil / clamp carry torch's gradients:
// zero, zero, pass-through inside [0, dim-1]).
#include "common.cuh"

namespace geom {
namespace pool {

constexpr int kMaxMaps = 8;
constexpr int kVertsPerCta = 32;
constexpr int kWarps = 8;
constexpr int kChunk = 128;     // channels per CTA: blockIdx.z walks (map, channel chunk) pairs, so that the real
constexpr int kMaxChunks = 64;  // shapes (16 meshes x 482 vertices x 960 channels) give ~2000 CTAs instead of 256

struct Maps {
    const float *ptr[kMaxMaps];  // forward: feature maps; backward: unused
    float *grad[kMaxMaps];       // backward: gradient of the maps (accumulated into)
    int channels[kMaxMaps];
    int dim[kMaxMaps];
    int col[kMaxMaps];                 // first output column of the map
    int n;
    int n_chunks;
    unsigned char chunk_map[kMaxChunks];   // chunk -> map
    unsigned short chunk_c0[kMaxChunks];   // chunk -> first channel / kChunk
};

struct Corner {  // per (vertex, map): flat indices of the 4 corners inside one channel plane and the weights
    int i11, i12, i21, i22;
    float A, B, G, H;
};

__device__ __forceinline__ Corner make_corner(float xs, float ys, int dim, bool *pass_x, bool *pass_y)
{
    const float fd = (float)dim, hi = (float)(dim - 1);
    const float px = __fmul_rn(xs, fd), py = __fmul_rn(ys, fd);
    const float cx = fminf(fmaxf(px, 0.f), hi), cy = fminf(fmaxf(py, 0.f), hi);
    if (pass_x) *pass_x = (px >= 0.f) && (px <= hi);  // torch.clamp passes the gradient inside the closed range
    if (pass_y) *pass_y = (py >= 0.f) && (py <= hi);
    const float x1 = floorf(cx), x2 = ceilf(cx), y1 = floorf(cy), y2 = ceilf(cy);
    Corner c;
    c.A = __fsub_rn(x2, cx), c.B = __fsub_rn(cx, x1), c.G = __fsub_rn(y2, cy), c.H = __fsub_rn(cy, y1);
    const int ix1 = (int)x1, ix2 = (int)x2, iy1 = (int)y1, iy2 = (int)y2;
    c.i11 = ix1 * dim + iy1, c.i12 = ix1 * dim + iy2, c.i21 = ix2 * dim + iy1, c.i22 = ix2 * dim + iy2;
    return c;
}

__global__ void __launch_bounds__(kWarps * 32)
fwd_kernel(int V, const Maps m, const float *__restrict__ xs, const float *__restrict__ ys,
           float *__restrict__ out, int ld, int col0)
{
    __shared__ Corner corners[kVertsPerCta];
    __shared__ float tile[kVertsPerCta][33];
    const int b = blockIdx.y, v0 = blockIdx.x * kVertsPerCta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = m.chunk_map[blockIdx.z];
    const int c_lo = (int)m.chunk_c0[blockIdx.z] * kChunk;
    if (threadIdx.x < kVertsPerCta) {
        const int v = min(v0 + (int)threadIdx.x, V - 1);
        corners[threadIdx.x] = make_corner(xs[(size_t)b * V + v], ys[(size_t)b * V + v], m.dim[k], nullptr, nullptr);
    }
    __syncthreads();
    {
        const int C = m.channels[k], plane = m.dim[k] * m.dim[k], col = m.col[k] + col0;
        const int c_hi = min(C, c_lo + kChunk);
        const float *base = m.ptr[k] + (size_t)b * C * plane;
        const Corner c = corners[lane];
        for (int cb = c_lo; cb < c_hi; cb += 32) {  // 32 channels per round: 4 per warp, then one coalesced flush
#pragma unroll
            for (int i = 0; i < 32 / kWarps; ++i) {
                const int ch = cb + warp + kWarps * i;
                float f = 0.f;
                if (ch < C) {
                    const float *p = base + (size_t)ch * plane;
                    const float vC = __ldg(p + c.i11), vD = __ldg(p + c.i12), vE = __ldg(p + c.i21), vF = __ldg(p + c.i22);
                    const float s1 = __fmul_rn(__fmul_rn(c.A, vC), c.G);
                    const float s2 = __fmul_rn(__fmul_rn(c.H, vD), c.A);
                    const float s3 = __fmul_rn(__fmul_rn(c.G, vE), c.B);
                    const float s4 = __fmul_rn(__fmul_rn(c.B, vF), c.H);
                    f = __fadd_rn(__fadd_rn(__fadd_rn(s1, s2), s3), s4);
                }
                tile[lane][warp + kWarps * i] = f;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kVertsPerCta / kWarps; ++i) {
                const int j = warp + kWarps * i, v = v0 + j;
                if (v < V && cb + lane < C) out[((size_t)b * V + v) * ld + col + cb + lane] = tile[j][lane];
            }
            __syncthreads();
        }
    }
}

// d(maps) += scatter of the weighted output gradient; d(xs), d(ys) = sum over channels and maps.
__global__ void __launch_bounds__(kWarps * 32)
bwd_kernel(int V, const Maps m, const float *__restrict__ xs, const float *__restrict__ ys,
           const float *__restrict__ dout, int ld, int col0, float *__restrict__ dxs, float *__restrict__ dys)
{
    __shared__ Corner corners[kVertsPerCta];
    __shared__ float tile[kVertsPerCta][33];
    __shared__ float gx[kWarps][kVertsPerCta], gy[kWarps][kVertsPerCta];
    __shared__ unsigned char passx[kVertsPerCta], passy[kVertsPerCta];
    const int b = blockIdx.y, v0 = blockIdx.x * kVertsPerCta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = m.chunk_map[blockIdx.z];
    const int c_lo = (int)m.chunk_c0[blockIdx.z] * kChunk;
    if (threadIdx.x < kVertsPerCta) {
        const int v = min(v0 + (int)threadIdx.x, V - 1);
        bool px, py;
        corners[threadIdx.x] = make_corner(xs[(size_t)b * V + v], ys[(size_t)b * V + v], m.dim[k], &px, &py);
        passx[threadIdx.x] = px, passy[threadIdx.x] = py;
    }
    __syncthreads();
    const bool live = v0 + lane < V;
    float ax = 0.f, ay = 0.f;  // this lane's vertex, this warp's channels
    {
        const int C = m.channels[k], plane = m.dim[k] * m.dim[k], col = m.col[k] + col0;
        const int c_hi = min(C, c_lo + kChunk);
        const size_t boff = (size_t)b * C * plane;
        const float *base = m.ptr[k] ? m.ptr[k] + boff : nullptr;
        float *gbase = m.grad[k] ? m.grad[k] + boff : nullptr;
        const Corner c = corners[lane];
        const float fd = (float)m.dim[k];
        const float sx = passx[lane] ? fd : 0.f, sy = passy[lane] ? fd : 0.f;  // d(cx)/d(xs), d(cy)/d(ys)
        for (int cb = c_lo; cb < c_hi; cb += 32) {
            // coalesced read of the 32 x 32 gradient tile
#pragma unroll
            for (int i = 0; i < kVertsPerCta / kWarps; ++i) {
                const int j = warp + kWarps * i, v = v0 + j;
                tile[j][lane] = (v < V && cb + lane < C) ? dout[((size_t)b * V + v) * ld + col + cb + lane] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 32 / kWarps; ++i) {
                const int ch = cb + warp + kWarps * i;
                if (ch < C && live) {
                    const float g = tile[lane][warp + kWarps * i];
                    if (gbase) {
                        float *p = gbase + (size_t)ch * plane;
                        atomicAdd(p + c.i11, g * c.A * c.G);
                        atomicAdd(p + c.i12, g * c.H * c.A);
                        atomicAdd(p + c.i21, g * c.G * c.B);
                        atomicAdd(p + c.i22, g * c.B * c.H);
                    }
                    if (base && (dxs || dys)) {
                        const float *p = base + (size_t)ch * plane;
                        const float vC = __ldg(p + c.i11), vD = __ldg(p + c.i12), vE = __ldg(p + c.i21), vF = __ldg(p + c.i22);
                        // f = A C G + H D A + G E B + B F H;  dA/dcx = -1, dB/dcx = +1, dG/dcy = -1, dH/dcy = +1
                        ax += g * (-(vC * c.G) - vD * c.H + vE * c.G + vF * c.H) * sx;
                        ay += g * (-(vC * c.A) + vD * c.A - vE * c.B + vF * c.B) * sy;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (dxs || dys) {  // partial sums of this channel chunk: the outputs were zeroed by the launcher
        gx[warp][lane] = ax, gy[warp][lane] = ay;
        __syncthreads();
        if (warp == 0 && live) {
            float tx = 0.f, ty = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) tx += gx[w][lane], ty += gy[w][lane];
            if (dxs) atomicAdd(&dxs[(size_t)b * V + v0 + lane], tx);
            if (dys) atomicAdd(&dys[(size_t)b * V + v0 + lane], ty);
        }
    }
}

static int fill(Maps &m, int n_maps, const float *const *maps, float *const *grads, const int *channels, const int *dims,
                long long *total_channels)
{
    if (n_maps < 0 || n_maps > kMaxMaps || (n_maps > 0 && (!channels || !dims))) return GEOM_ERR_INVALID_ARG;
    m.n = n_maps;
    long long tc = 0;
    for (int k = 0; k < kMaxMaps; ++k) {
        m.ptr[k] = (k < n_maps && maps) ? maps[k] : nullptr;
        m.grad[k] = (k < n_maps && grads) ? grads[k] : nullptr;
        m.channels[k] = k < n_maps ? channels[k] : 0;
        m.dim[k] = k < n_maps ? dims[k] : 1;
        if (k < n_maps) {
            if (channels[k] < 0 || dims[k] < 1 || dims[k] > 32768) return GEOM_ERR_INVALID_ARG;
            tc += channels[k];
        }
    }
    *total_channels = tc;
    m.n_chunks = 0;
    int col = 0;
    for (int k = 0; k < n_maps; ++k) {
        m.col[k] = col;
        col += channels[k];
        for (int c0 = 0; c0 < channels[k]; c0 += kChunk) {
            if (m.n_chunks == kMaxChunks || c0 / kChunk > 65535) return GEOM_ERR_INVALID_ARG;  // > 8192 channels in total
            m.chunk_map[m.n_chunks] = (unsigned char)k;
            m.chunk_c0[m.n_chunks] = (unsigned short)(c0 / kChunk);
            ++m.n_chunks;
        }
    }
    return GEOM_OK;
}

}  // namespace pool
}  // namespace geom

using namespace geom;

extern "C" int geom_pool_fwd(int b, int V, int n_maps, const float *const *maps, const int *channels, const int *dims,
                             const float *xs, const float *ys, float *out, int ld, int col0, geom_stream_t stream)
{
    pool::Maps m;
    long long tc = 0;
    if (b < 0 || V < 0 || ld < 0 || col0 < 0) return GEOM_ERR_INVALID_ARG;
    const int rc = pool::fill(m, n_maps, maps, nullptr, channels, dims, &tc);
    if (rc != GEOM_OK) return rc;
    if (b == 0 || V == 0 || tc == 0) return GEOM_OK;
    if (!maps || !xs || !ys || !out || col0 + tc > ld || b > 65535) return GEOM_ERR_INVALID_ARG;
    for (int k = 0; k < n_maps; ++k)
        if (channels[k] > 0 && !maps[k]) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        prof::Scope ps(st, prof::kPoolFwd);
        pool::fwd_kernel<<<dim3((unsigned)((V + pool::kVertsPerCta - 1) / pool::kVertsPerCta), (unsigned)b, (unsigned)m.n_chunks), pool::kWarps * 32, 0, st>>>(
            V, m, xs, ys, out, ld, col0);
    }
    return status_from_cuda(cudaGetLastError());
}

extern "C" int geom_pool_bwd(int b, int V, int n_maps, const float *const *maps, float *const *grad_maps,
                             const int *channels, const int *dims, const float *xs, const float *ys,
                             const float *grad_out, int ld, int col0, float *grad_xs, float *grad_ys, geom_stream_t stream)
{
    pool::Maps m;
    long long tc = 0;
    if (b < 0 || V < 0 || ld < 0 || col0 < 0) return GEOM_ERR_INVALID_ARG;
    const int rc = pool::fill(m, n_maps, maps, grad_maps, channels, dims, &tc);
    if (rc != GEOM_OK) return rc;
    if (b == 0 || V == 0) return GEOM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (tc == 0) {  // no features: the coordinate gradients are zero
        cudaError_t e = cudaSuccess;
        if (grad_xs) e = cudaMemsetAsync(grad_xs, 0, sizeof(float) * (size_t)b * V, st);
        if (e == cudaSuccess && grad_ys) e = cudaMemsetAsync(grad_ys, 0, sizeof(float) * (size_t)b * V, st);
        return status_from_cuda(e);
    }
    if (!xs || !ys || !grad_out || col0 + tc > ld || b > 65535) return GEOM_ERR_INVALID_ARG;
    if ((grad_xs || grad_ys) && !maps) return GEOM_ERR_INVALID_ARG;
    {
        cudaError_t e = cudaSuccess;
        if (grad_xs) e = cudaMemsetAsync(grad_xs, 0, sizeof(float) * (size_t)b * V, st);
        if (e == cudaSuccess && grad_ys) e = cudaMemsetAsync(grad_ys, 0, sizeof(float) * (size_t)b * V, st);
        if (e != cudaSuccess) return status_from_cuda(e);
        prof::Scope ps(st, prof::kPoolBwd);
        pool::bwd_kernel<<<dim3((unsigned)((V + pool::kVertsPerCta - 1) / pool::kVertsPerCta), (unsigned)b, (unsigned)m.n_chunks), pool::kWarps * 32, 0, st>>>(
            V, m, xs, ys, grad_out, ld, col0, grad_xs, grad_ys);
    }
    return status_from_cuda(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------
// Projection of the vertices into the image (batch_camera_info + utils.py:319-330) as ONE launch instead of
// ~30 tiny (B,3) / (B,V) torch kernels.  Same expressions in the same order as the reference; torch's own
// kernels may contract a*b - c*d differently, so this path matches the torch expression to fp32 rounding
// (~1e-6 relative), not bit for bit.  img_info rows are (azimuth deg, elevation deg, distance).
// ------------------------------------------------------------------------------------------------------
namespace geom {
namespace pool {

struct Camera {
    float c[3];      // position
    float ax[3], ay[3], az[3];  // rows of cam_mat
};

__device__ __forceinline__ float rem360(float x)  // torch's % on floats: result takes the sign of the divisor
{
    float r = fmodf(x, 360.f);
    if (r != 0.f && r < 0.f) r += 360.f;
    return r;
}

__device__ __forceinline__ Camera make_camera(const float *__restrict__ p)
{
    const float kPi = 3.14159265358979323846f;
    const float theta = rem360(p[0] * kPi / 180.f), phi = rem360(p[1] * kPi / 180.f);
    Camera k;
    const float t = p[2] * cosf(phi);
    k.c[0] = t * cosf(theta), k.c[1] = p[2] * sinf(phi), k.c[2] = t * sinf(theta);
    // axisX = cross((0,1,0), c) = (c.z, 0, -c.x);  axisY = cross(c, axisX);  axisZ = c;  each normalised
    float x[3] = {k.c[2], 0.f, -k.c[0]};
    float y[3] = {k.c[1] * x[2] - k.c[2] * x[1], k.c[2] * x[0] - k.c[0] * x[2], k.c[0] * x[1] - k.c[1] * x[0]};
    const float nx = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const float ny = sqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    const float nz = sqrtf(k.c[0] * k.c[0] + k.c[1] * k.c[1] + k.c[2] * k.c[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) k.ax[i] = x[i] / nx, k.ay[i] = y[i] / ny, k.az[i] = k.c[i] / nz;
    return k;
}

__global__ void __launch_bounds__(256)
project_fwd_kernel(int V, const float *__restrict__ verts, const float *__restrict__ info, float *__restrict__ xs,
                   float *__restrict__ ys)
{
    const int b = blockIdx.y, v = blockIdx.x * 256 + threadIdx.x;
    if (v >= V) return;
    const Camera k = make_camera(info + 3 * b);
    const float *p = verts + ((size_t)b * V + v) * 3;
    const float q0 = p[0] * .57f - k.c[0], q1 = p[1] * .57f - k.c[1], q2 = p[2] * .57f - k.c[2];
    const float X = q0 * k.ax[0] + q1 * k.ax[1] + q2 * k.ax[2];
    const float Y = q0 * k.ay[0] + q1 * k.ay[1] + q2 * k.ay[2];
    const float Z = q0 * k.az[0] + q1 * k.az[1] + q2 * k.az[2];
    const float h = (-Y) / (-Z) * 248.f + 112.f;
    const float w = X / (-Z) * 248.f + 112.f;
    xs[(size_t)b * V + v] = h / 223.f;
    ys[(size_t)b * V + v] = w / 223.f;
}

// d(verts) from d(xs), d(ys):  xs = (248 Y/Z + 112)/223,  ys = (-248 X/Z + 112)/223,  (X,Y,Z) = M (0.57 p - c)
__global__ void __launch_bounds__(256)
project_bwd_kernel(int V, const float *__restrict__ verts, const float *__restrict__ info, const float *__restrict__ gxs,
                   const float *__restrict__ gys, float *__restrict__ gverts)
{
    const int b = blockIdx.y, v = blockIdx.x * 256 + threadIdx.x;
    if (v >= V) return;
    const Camera k = make_camera(info + 3 * b);
    const float *p = verts + ((size_t)b * V + v) * 3;
    const float q0 = p[0] * .57f - k.c[0], q1 = p[1] * .57f - k.c[1], q2 = p[2] * .57f - k.c[2];
    const float X = q0 * k.ax[0] + q1 * k.ax[1] + q2 * k.ax[2];
    const float Y = q0 * k.ay[0] + q1 * k.ay[1] + q2 * k.ay[2];
    const float Z = q0 * k.az[0] + q1 * k.az[1] + q2 * k.az[2];
    const float s = 248.f / 223.f;
    const float gx = gxs ? gxs[(size_t)b * V + v] : 0.f, gy = gys ? gys[(size_t)b * V + v] : 0.f;
    const float dX = -gy * s / Z;
    const float dY = gx * s / Z;
    const float dZ = (-gx * Y + gy * X) * s / (Z * Z);
    float *g = gverts + ((size_t)b * V + v) * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = .57f * (dX * k.ax[i] + dY * k.ay[i] + dZ * k.az[i]);
}

}  // namespace pool
}  // namespace geom

extern "C" int geom_project_fwd(int b, int V, const float *verts, const float *img_info, float *xs, float *ys,
                                geom_stream_t stream)
{
    if (b < 0 || V < 0) return GEOM_ERR_INVALID_ARG;
    if (b == 0 || V == 0) return GEOM_OK;
    if (!verts || !img_info || !xs || !ys || b > 65535) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        geom::prof::Scope ps(st, geom::prof::kMisc);
        geom::pool::project_fwd_kernel<<<dim3((unsigned)((V + 255) / 256), (unsigned)b), 256, 0, st>>>(V, verts, img_info, xs, ys);
    }
    return geom::status_from_cuda(cudaGetLastError());
}

extern "C" int geom_project_bwd(int b, int V, const float *verts, const float *img_info, const float *grad_xs,
                                const float *grad_ys, float *grad_verts, geom_stream_t stream)
{
    if (b < 0 || V < 0) return GEOM_ERR_INVALID_ARG;
    if (b == 0 || V == 0) return GEOM_OK;
    if (!verts || !img_info || !grad_verts || b > 65535) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    {
        geom::prof::Scope ps(st, geom::prof::kMisc);
        geom::pool::project_bwd_kernel<<<dim3((unsigned)((V + 255) / 256), (unsigned)b), 256, 0, st>>>(V, verts, img_info, grad_xs, grad_ys,
                                                                                                    grad_verts);
    }
    return geom::status_from_cuda(cudaGetLastError());
}

// host.cu -- host-buffer entry points: the caller hands HOST memory (pinned for full overlap), the
// library pipelines  H2D copy -> search -> D2H copy  over batch chunks on three streams and returns
// when the results are in host memory.  This is the end-to-end path a caller without device-resident
// data takes (the reference's wrapper also starts from host allocations: chamfer_distance.py:16-26
// builds its outputs on the CPU and copies them over, synchronously and pageable).
//
// A geom_host_ctx owns the streams, events and a device arena that grows on demand and is reused
// across calls; it is not thread-safe (one context per calling thread).
#include <new>

#include "common.cuh"

struct geom_host_ctx {
    cudaStream_t h2d = nullptr, comp = nullptr, d2h = nullptr;
    static constexpr int kMaxChunks = 8;
    cudaEvent_t in_done[kMaxChunks] = {}, out_ready[kMaxChunks] = {};
    void *arena = nullptr;
    size_t arena_bytes = 0;
    int device = 0;
};

static cudaError_t ensure_arena(geom_host_ctx *c, size_t bytes)
{
    if (bytes <= c->arena_bytes) return cudaSuccess;
    if (c->arena) {
        cudaError_t e = cudaFree(c->arena);  // synchronises the device: nothing of ours is in flight here
        c->arena = nullptr;
        c->arena_bytes = 0;
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaMalloc(&c->arena, bytes);
    if (e == cudaSuccess) c->arena_bytes = bytes;
    return e;
}

extern "C" int geom_host_ctx_create(geom_host_ctx **out)
{
    if (!out) return GEOM_ERR_INVALID_ARG;
    geom_host_ctx *c = new (std::nothrow) geom_host_ctx();
    if (!c) return GEOM_ERR_INVALID_ARG;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->comp, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking);
    for (int i = 0; e == cudaSuccess && i < geom_host_ctx::kMaxChunks; ++i) {
        e = cudaEventCreateWithFlags(&c->in_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->out_ready[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        geom_host_ctx_destroy(c);
        return geom::status_from_cuda(e);
    }
    *out = c;
    return GEOM_OK;
}

extern "C" int geom_host_ctx_destroy(geom_host_ctx *c)
{
    if (!c) return GEOM_OK;
    for (int i = 0; i < geom_host_ctx::kMaxChunks; ++i) {
        if (c->in_done[i]) cudaEventDestroy(c->in_done[i]);
        if (c->out_ready[i]) cudaEventDestroy(c->out_ready[i]);
    }
    if (c->h2d) cudaStreamDestroy(c->h2d);
    if (c->comp) cudaStreamDestroy(c->comp);
    if (c->d2h) cudaStreamDestroy(c->d2h);
    if (c->arena) cudaFree(c->arena);
    delete c;
    return GEOM_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Slice plan of geom_chamfer_fwd_host: lo_of[k]..lo_of[k+1] is slice k; returns the number of slices (<= 8).
extern "C" int geom_chamfer_host_plan(int b, int n, int m, int chunks, unsigned flags, int sm_count, int *lo_of)
{
    if (b < 1 || n < 0 || m < 0 || !lo_of) return 0;
    if (chunks < 1) chunks = 4;
    if (chunks > geom_host_ctx::kMaxChunks) chunks = geom_host_ctx::kMaxChunks;
    if (chunks > b) chunks = b;
    const bool cuda_cores = (flags & (GEOM_FLAG_FP32 | GEOM_FLAG_ONEWAY)) != 0;
    const int sms = sm_count;
    const int per_item = cuda_cores ? 0 : geom::chamfer_tc::ctas_per_item(n, m);
    if (per_item > 0 && 2 * per_item <= sms && chunks >= 3) {
        // Tensor-core search: one CTA per SM, so a launch costs whole waves of the grid.  Slices of 1, 2, 3, 3, ...
        // waves: the first copy-in is short, every later copy-in hides behind the previous slice's search, and no
        // slice pays for a nearly empty final wave.  `chunks` is the maximum number of slices.
        int k = 0, pos = 0, w = 1;
        lo_of[0] = 0;
        const int one_wave = sms / per_item;
        while (pos < b) {
            int take = (w * sms) / per_item;
            if (k == chunks - 1 || b - pos - take < one_wave) take = b - pos;  // the last slice takes the remainder
            lo_of[++k] = (pos += take);
            if (w < 3) ++w;
        }
        return k;
    }
    // CUDA-core search (ticket-scheduled, no wave quantisation): small first and last slices (only the first copy-in
    // and the last copy-out are exposed), the rest split evenly.
    const int edge = (chunks >= 3) ? ((b / 8 > 0) ? b / 8 : 1) : 0;
    const int inner = chunks - (edge ? 2 : 0), rest = b - 2 * edge;
    int pos = 0, k = 0;
    lo_of[k++] = 0;
    if (edge) lo_of[k++] = (pos += edge);
    for (int i = 0; i < inner; ++i) lo_of[k++] = (pos += rest / inner + (i < rest % inner ? 1 : 0));
    if (edge) lo_of[k++] = (pos += edge);
    return k - 1;
}

extern "C" int geom_chamfer_fwd_host(geom_host_ctx *c, int b, int n, const float *h_xyz1, int m,
                                     const float *h_xyz2, float *h_dist1, int *h_idx1, float *h_dist2,
                                     int *h_idx2, int chunks, unsigned flags)
{
    if (!c || b < 0 || n < 0 || m < 0) return GEOM_ERR_INVALID_ARG;
    {
        // the context's streams, events and arena live on the device it was created on: calling from another current
        // device would launch there with pointers from here
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != c->device) return GEOM_ERR_INVALID_ARG;
    }
    if (b == 0 || (n == 0 && m == 0)) return GEOM_OK;
    if ((n > 0 && (!h_xyz1 || !h_idx1)) || (m > 0 && (!h_xyz2 || !h_idx2))) return GEOM_ERR_INVALID_ARG;
    int lo_of[geom_host_ctx::kMaxChunks + 1];
    const bool cuda_cores = (flags & (GEOM_FLAG_FP32 | GEOM_FLAG_ONEWAY)) != 0;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess) sms = 0;
    chunks = geom_chamfer_host_plan(b, n, m, chunks, flags, sms, lo_of);
    if (chunks < 1) return GEOM_ERR_INVALID_ARG;
    int per = 0;
    for (int k = 0; k < chunks; ++k) per = (lo_of[k + 1] - lo_of[k] > per) ? lo_of[k + 1] - lo_of[k] : per;

    // arena layout: xyz1 | xyz2 | idx1 | idx2 | dist1 | dist2 | one search workspace per chunk
    const size_t s_x1 = align_up(sizeof(float) * 3 * (size_t)b * n, 256), s_x2 = align_up(sizeof(float) * 3 * (size_t)b * m, 256);
    const size_t s_i1 = align_up(sizeof(int) * (size_t)b * n, 256), s_i2 = align_up(sizeof(int) * (size_t)b * m, 256);
    const size_t s_d1 = h_dist1 ? s_i1 : 0, s_d2 = h_dist2 ? s_i2 : 0;
    const size_t ws_one = align_up(cuda_cores ? geom_chamfer_workspace_bytes(per, n, m) : geom_chamfer_tensor_workspace_bytes(per, n, m), 256);
    cudaError_t e = ensure_arena(c, s_x1 + s_x2 + s_i1 + s_i2 + s_d1 + s_d2 + ws_one * chunks);
    if (e != cudaSuccess) return geom::status_from_cuda(e);
    char *base = (char *)c->arena;
    float *d_x1 = (float *)base, *d_x2 = (float *)(base + s_x1);
    int *d_i1 = (int *)(base + s_x1 + s_x2), *d_i2 = (int *)(base + s_x1 + s_x2 + s_i1);
    float *d_d1 = h_dist1 ? (float *)(base + s_x1 + s_x2 + s_i1 + s_i2) : nullptr;
    float *d_d2 = h_dist2 ? (float *)(base + s_x1 + s_x2 + s_i1 + s_i2 + s_d1) : nullptr;
    char *d_ws = base + s_x1 + s_x2 + s_i1 + s_i2 + s_d1 + s_d2;

    for (int k = 0; k < chunks; ++k) {
        const int lo = lo_of[k], hi = lo_of[k + 1];
        if (lo >= hi) continue;
        const size_t o1 = (size_t)lo * n, o2 = (size_t)lo * m, c1 = (size_t)(hi - lo) * n, c2 = (size_t)(hi - lo) * m;
        if (c1) e = cudaMemcpyAsync(d_x1 + o1 * 3, h_xyz1 + o1 * 3, sizeof(float) * 3 * c1, cudaMemcpyHostToDevice, c->h2d);
        if (e == cudaSuccess && c2)
            e = cudaMemcpyAsync(d_x2 + o2 * 3, h_xyz2 + o2 * 3, sizeof(float) * 3 * c2, cudaMemcpyHostToDevice, c->h2d);
        if (e == cudaSuccess) e = cudaEventRecord(c->in_done[k], c->h2d);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->comp, c->in_done[k], 0);
        if (e != cudaSuccess) break;
        const int rc = geom_chamfer_fwd(hi - lo, n, d_x1 + o1 * 3, m, d_x2 + o2 * 3, d_d1 ? d_d1 + o1 : nullptr, d_i1 + o1,
                                        d_d2 ? d_d2 + o2 : nullptr, d_i2 + o2, d_ws + ws_one * k, ws_one, flags, c->comp);
        if (rc != GEOM_OK) {
            cudaDeviceSynchronize();
            return rc;
        }
        e = cudaEventRecord(c->out_ready[k], c->comp);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->d2h, c->out_ready[k], 0);
        if (e == cudaSuccess && c1) e = cudaMemcpyAsync(h_idx1 + o1, d_i1 + o1, sizeof(int) * c1, cudaMemcpyDeviceToHost, c->d2h);
        if (e == cudaSuccess && c2) e = cudaMemcpyAsync(h_idx2 + o2, d_i2 + o2, sizeof(int) * c2, cudaMemcpyDeviceToHost, c->d2h);
        if (e == cudaSuccess && d_d1 && c1) e = cudaMemcpyAsync(h_dist1 + o1, d_d1 + o1, sizeof(float) * c1, cudaMemcpyDeviceToHost, c->d2h);
        if (e == cudaSuccess && d_d2 && c2) e = cudaMemcpyAsync(h_dist2 + o2, d_d2 + o2, sizeof(float) * c2, cudaMemcpyDeviceToHost, c->d2h);
        if (e != cudaSuccess) break;
    }
    const cudaError_t es = cudaStreamSynchronize(c->d2h);
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) cudaDeviceSynchronize();
    return geom::status_from_cuda(e);
}

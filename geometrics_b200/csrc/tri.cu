// tri.cu -- point-to-triangle nearest search for sm_100a.
//
// Replaces TriDistanceKernel / TriDistanceKernelLauncher (reference:
// tri_distance/tri_distance.cu:94-228).  Outputs (dist, region option, face index) are bit-identical
// to the reference kernel, including on degenerate triangles (NaN handling per 512-tile) and the
// reference's option-6 direction-vector quirk (tri_distance.cu:180), because both decide which
// face index wins.
//
// Design (not a translation):
//   * Everything that depends only on the triangle (edge vectors, squared edge lengths and their
//     correctly rounded reciprocals, the face normal, the three in-plane edge normals cross(N, edge)
//     used by is_above, the unit normal of the plane projection) is computed ONCE per triangle when
//     a warp stages 32 triangles into its shared-memory slab, not once per (point, triangle) pair as
//     the reference does (tri_distance.cu:140-148).  Each value is produced by the same rounding
//     sequence the reference binary uses (explicit _rn intrinsics; DESIGN.md "TriDistance").
//   * The three IEEE divisions per pair become  q = a*y; r = fma(-b,q,a); q = fma(r,y,q)  (twice)
//     with y = rn(1/b) precomputed per triangle: Markstein's correction, correctly rounded whenever
//     no intermediate can under/overflow; operands outside that comfortable range take __fdiv_rn.
//   * The inner loop is branch-free (the reference's if/else region cascade becomes predicate logic
//     and selects; a warp holds 32 different points, so the cascade diverged on every triangle) and
//     tracks only (best distance, face index).  The region option of the winner is recomputed in the
//     epilogue by the exact cascade.
//   * Every warp is an independent worker; work units (32*PR points) x (chunk of a 512-triangle
//     reference tile) are pulled from a global ticket, so small batches still fill the GPU and there
//     is no wave quantisation.  Chunks meet in 64-bit keys (distance bits << 32 | face index) merged by
//     filtered atomicMin = "strictly smaller wins, lowest index on ties", the reference's rule.
#include "tri_common.cuh"

namespace geom {
namespace tri {

constexpr int kWarps = 4;
constexpr int kThreads = 32 * kWarps;

struct TriArgs {
    const float *xyz, *t1, *t2, *t3;
    u64 *keys;              // [b*n]  (distance bits << 32 | face index), 0xff.. = nothing yet
    unsigned char *poison;  // [b*n]  0xff = clean, 0 = tile 0's first candidate evaluated to NaN
    unsigned *ticket;       // starts at 0xffffffff
    int n, m;
    int ref_tile;           // 512, or m rounded up in exact (flat scan) mode
    int full_tiles;         // number of complete reference tiles
    int last_len;           // scanned length of the trailing partial tile (0 = none)
    int chunk;              // triangles per work unit, a multiple of 32 dividing ref_tile
    int chunks_per_tile;    // ref_tile / chunk (full tiles)
    int n_chunks;           // all tiles
    int n_pblk;             // point blocks per batch item
    unsigned total_units;
};

template <int PR>
__global__ void __launch_bounds__(kThreads, 4) search_kernel(const TriArgs a)
{
    __shared__ float4 s_rec[kWarps][32 * kRecVec];
    const int lane = threadIdx.x & 31;
    float4 *slab = s_rec[threadIdx.x >> 5];

    for (;;) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(a.ticket, 1u) + 1u;
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= a.total_units) break;
        const int ch = (int)(u % (unsigned)a.n_chunks);
        const unsigned pu = u / (unsigned)a.n_chunks;
        const int bi = (int)(pu / (unsigned)a.n_pblk);
        const int pbase = (int)(pu % (unsigned)a.n_pblk) * (32 * PR);
        // chunk -> (reference tile start k2, offset inside the tile, length)
        int k2, off, len;
        if (ch < a.full_tiles * a.chunks_per_tile) {
            k2 = (ch / a.chunks_per_tile) * a.ref_tile;
            off = (ch % a.chunks_per_tile) * a.chunk;
            len = a.chunk;
        } else {
            k2 = a.full_tiles * a.ref_tile;
            off = (ch - a.full_tiles * a.chunks_per_tile) * a.chunk;
            len = min(a.chunk, a.last_len - off);
        }
        const size_t tbase = (size_t)bi * a.m * 3;
        const float *pp = a.xyz + (size_t)bi * a.n * 3;

        V3 p[PR];
        float best[PR];
        int bidx[PR];
        bool live[PR];  // false: this tile's first candidate is NaN for the point -> the tile is void
#pragma unroll
        for (int r = 0; r < PR; ++r) {
            const int row = min(pbase + r * 32 + lane, a.n - 1);
            p[r] = v3(pp[row * 3 + 0], pp[row * 3 + 1], pp[row * 3 + 2]);
        }
        // The reference takes the first candidate of every 512-tile unconditionally
        // (tri_distance.cu:194, k == 0): if it evaluates to NaN nothing in that tile can replace it.
        {
            const size_t g = tbase + (size_t)k2 * 3;
            if (lane == 0)
                build_record(slab, v3(a.t1[g], a.t1[g + 1], a.t1[g + 2]), v3(a.t2[g], a.t2[g + 1], a.t2[g + 2]),
                             v3(a.t3[g], a.t3[g + 1], a.t3[g + 2]));
            __syncwarp();
#pragma unroll
            for (int r = 0; r < PR; ++r) {
                float opt;
                const float d0 = tri_pair_exact(p[r], slab, opt);
                live[r] = !(d0 != d0);
                best[r] = (off == 0) ? d0 : __int_as_float(0x7f800000);  // candidate k2 belongs to chunk 0
                bidx[r] = k2;
            }
            __syncwarp();
        }

        for (int s0 = 0; s0 < len; s0 += 32) {
            const int cnt = min(32, len - s0);
            if (lane < cnt) {
                const size_t g = tbase + (size_t)(k2 + off + s0 + lane) * 3;
                build_record(slab + lane * kRecVec, v3(a.t1[g], a.t1[g + 1], a.t1[g + 2]),
                             v3(a.t2[g], a.t2[g + 1], a.t2[g + 2]), v3(a.t3[g], a.t3[g + 1], a.t3[g + 2]));
            }
            __syncwarp();
#pragma unroll 1
            for (int j = 0; j < cnt; ++j) {
                const TriRegs t = load_record(slab + j * kRecVec);
                const int fi = k2 + off + s0 + j;
#pragma unroll
                for (int r = 0; r < PR; ++r) {
                    const float d = tri_pair_fast(p[r], t);
                    if (d < best[r]) {  // NaN never wins; ascending scan + strict '<' = lowest index on ties
                        best[r] = d;
                        bidx[r] = fi;
                    }
                }
            }
            __syncwarp();
        }

#pragma unroll
        for (int r = 0; r < PR; ++r) {
            const int row = pbase + r * 32 + lane;
            if (row >= a.n) continue;
            const size_t o = (size_t)bi * a.n + row;
            if (live[r]) {
                if (best[r] < __int_as_float(0x7f800000) || off == 0)
                    key_min(&a.keys[o], make_key(best[r], (unsigned)bidx[r]));
            } else if (k2 == 0 && off == 0) {
                a.poison[o] = 0;
            }
        }
    }
}

struct FinishArgs {
    const float *xyz, *t1, *t2, *t3;
    const u64 *keys;
    const unsigned char *poison;
    float *dist;
    int *point, *index;
    int b, n, m;
    int empty_tail;  // the trailing reference tile holds 1-3 triangles: its (10000, 0, 0) wins over d > 10000
};

// Decode the winner per point and recompute its region option with the exact cascade.
__global__ void __launch_bounds__(256) finish_kernel(const FinishArgs a)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)a.b * a.n) return;
    const int bi = (int)(t / a.n);
    const u64 key = a.keys[t];
    float d = 10000.0f, opt = 0.0f;  // tri_distance.cu:127-128 with nothing scanned
    int idx = 0;
    const bool poisoned = a.poison[t] == 0;
    if (poisoned || key != ~0ull) {
        idx = poisoned ? 0 : (int)(unsigned)key;
        const size_t g = ((size_t)bi * a.m + idx) * 3;
        float4 rec[kRecVec];
        build_record(rec, v3(a.t1[g], a.t1[g + 1], a.t1[g + 2]), v3(a.t2[g], a.t2[g + 1], a.t2[g + 2]),
                     v3(a.t3[g], a.t3[g + 1], a.t3[g + 2]));
        d = tri_pair_exact(v3(a.xyz[t * 3], a.xyz[t * 3 + 1], a.xyz[t * 3 + 2]), rec, opt);
        if (a.empty_tail && d > 10000.0f) {  // merge rule `dist > best` against the empty tile, :202
            d = 10000.0f;
            opt = 0.0f;
            idx = 0;
        }
    }
    if (a.dist) a.dist[t] = d;
    a.point[t] = (int)opt;
    a.index[t] = idx;
}

}  // namespace tri
}  // namespace geom

extern "C" size_t geom_tri_workspace_bytes(int b, int n)
{
    if (b < 0 || n < 0) return 0;
    const size_t pts = (size_t)b * n;
    return sizeof(geom::u64) * pts + ((pts + 15) & ~(size_t)15) + 16;  // keys + poison bytes + ticket
}

extern "C" size_t geom_tri_tensor_workspace_bytes(int b, int n, int m)
{
    if (b < 0 || n < 0 || m < 0) return 0;
    const size_t plain = geom_tri_workspace_bytes(b, n);  // small or degenerate sizes fall back to the plain kernel
    const size_t tc = geom::tri_tc::workspace_bytes(b, n, m);
    return plain > tc ? plain : tc;
}

extern "C" int geom_tri_fwd(int b, int n, const float *xyz, int m, const float *tri1, const float *tri2,
                            const float *tri3, float *dist, int *point, int *index, void *workspace,
                            size_t workspace_bytes, unsigned flags, geom_stream_t stream)
{
    using namespace geom;
    using namespace geom::tri;
    if (b < 0 || n < 0 || m < 0) return GEOM_ERR_INVALID_ARG;
    if ((long long)b * n * 3 > 0x7fffffffLL || (long long)b * m * 3 > 0x7fffffffLL) return GEOM_ERR_INVALID_ARG;
    if (b == 0 || n == 0) return GEOM_OK;
    if (!xyz || !point || !index) return GEOM_ERR_INVALID_ARG;
    if (m > 0 && (!tri1 || !tri2 || !tri3)) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (m == 0) {
        // the reference kernel's tile loop never runs: outputs keep the caller's zeros
        if (dist) e = cudaMemsetAsync(dist, 0, sizeof(float) * (size_t)b * n, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(point, 0, sizeof(int) * (size_t)b * n, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(index, 0, sizeof(int) * (size_t)b * n, st);
        return status_from_cuda(e);
    }
    const bool exact = (flags & GEOM_FLAG_EXACT_TAIL) != 0;
    {
        // Path choice, as for Chamfer: GEOM_FLAG_FP32 forces the CUDA-core brute force, GEOM_FLAG_TENSOR the tensor-core
        // filtered search; with neither, the filtered search runs when the workspace is large enough for it and the
        // problem is big enough to amortise its two extra launches.
        const int ft = exact ? 0 : m / kRefTile;
        const int rem_ = m - ft * kRefTile;
        const int scan_end = exact ? m : ft * kRefTile + (rem_ - (rem_ & 3));  // one past the last scanned triangle
        const bool tc_possible = !(flags & GEOM_FLAG_FP32) && scan_end > 0 && b <= 65535 && m <= (1 << 22);
        const size_t tc_need = tc_possible ? geom_tri_tensor_workspace_bytes(b, n, m) : 0;
        const bool tc_fits = workspace && workspace_bytes >= tc_need && !((uintptr_t)workspace & 15u);
        const bool worth = (long long)b * n * m >= (1ll << 24);
        if (tc_possible && ((flags & GEOM_FLAG_TENSOR) || (tc_fits && worth))) {
            if (!tc_fits) return GEOM_ERR_WORKSPACE;
            const int empty_tail_tc = (!exact && ft > 0 && rem_ > 0 && (rem_ - (rem_ & 3)) == 0) ? 1 : 0;
            const float ef = debug::knob(GEOM_DEBUG_TRI_TC_EPS_FACTOR);
            return status_from_cuda(tri_tc::launch(b, n, xyz, m, tri1, tri2, tri3, scan_end, exact ? 1 : 0, empty_tail_tc, dist,
                                                   point, index, workspace, ef > 0.f ? ef : 1.0f, st));
        }
    }
    const size_t need = geom_tri_workspace_bytes(b, n);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 7u)) return GEOM_ERR_WORKSPACE;
    const size_t pts = (size_t)b * n;

    TriArgs a;
    a.xyz = xyz;
    a.t1 = tri1;
    a.t2 = tri2;
    a.t3 = tri3;
    a.keys = (u64 *)workspace;
    a.poison = (unsigned char *)(a.keys + pts);
    a.ticket = (unsigned *)(a.poison + ((pts + 15) & ~(size_t)15));
    a.n = n;
    a.m = m;
    // exact mode is a flat scan: one "tile" spanning every face, no tail skip, no per-tile restart
    a.ref_tile = exact ? ((m + 511) / 512) * 512 : kRefTile;
    a.full_tiles = exact ? 0 : m / kRefTile;
    const int rem = m - a.full_tiles * a.ref_tile;
    a.last_len = exact ? m : rem - (rem & 3);  // tri_distance.cu:129
    const int empty_tail = (!exact && a.full_tiles > 0 && rem > 0 && a.last_len == 0) ? 1 : 0;

    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return status_from_cuda(e);
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return status_from_cuda(e);
    const long long slots = (long long)sms * 4 * kWarps;
    // Largest points-per-lane / chunk that still gives every warp ~3 units to pull.
    int pr = 4, chunk = 512;
    auto units_for = [&](int pr_, int chunk_) {
        const long long nch = (long long)a.full_tiles * (a.ref_tile / chunk_) + (a.last_len + chunk_ - 1) / chunk_;
        return (long long)b * ((n + 32 * pr_ - 1) / (32 * pr_)) * nch;
    };
    while (units_for(pr, chunk) < 3 * slots) {
        if (pr > 1)
            pr >>= 1;
        else if (chunk > 64)
            chunk >>= 1;
        else
            break;
    }
    {  // tests / tuning: force a variant through geom_debug_set (both knobs must be set)
        const int epr = (int)debug::knob(GEOM_DEBUG_TRI_PR), ech = (int)debug::knob(GEOM_DEBUG_TRI_CHUNK);
        if ((epr == 1 || epr == 2 || epr == 4) && (ech == 64 || ech == 128 || ech == 256 || ech == 512)) {
            pr = epr;
            chunk = ech;
        }
    }
    a.chunk = chunk;
    a.chunks_per_tile = a.ref_tile / chunk;
    a.n_chunks = a.full_tiles * a.chunks_per_tile + (a.last_len + chunk - 1) / chunk;
    a.n_pblk = (n + 32 * pr - 1) / (32 * pr);
    const long long units = units_for(pr, chunk);
    if (units > 0x7fffffffLL) return GEOM_ERR_INVALID_ARG;
    a.total_units = (unsigned)units;

    if ((e = cudaMemsetAsync(workspace, 0xff, need, st)) != cudaSuccess) return status_from_cuda(e);
    if (units > 0) {
        const int ctas = (int)(((units < slots ? units : slots) + kWarps - 1) / kWarps);
        prof::Scope ps(st, prof::kTri);
        if (pr == 4)
            search_kernel<4><<<ctas, kThreads, 0, st>>>(a);
        else if (pr == 2)
            search_kernel<2><<<ctas, kThreads, 0, st>>>(a);
        else
            search_kernel<1><<<ctas, kThreads, 0, st>>>(a);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return status_from_cuda(e);

    FinishArgs f;
    f.xyz = xyz;
    f.t1 = tri1;
    f.t2 = tri2;
    f.t3 = tri3;
    f.keys = a.keys;
    f.poison = a.poison;
    f.dist = dist;
    f.point = point;
    f.index = index;
    f.b = b;
    f.n = n;
    f.m = m;
    f.empty_tail = empty_tail;
    {
        prof::Scope ps(st, prof::kMisc);
        finish_kernel<<<(unsigned)((pts + 255) / 256), 256, 0, st>>>(f);
    }
    return status_from_cuda(cudaGetLastError());
}

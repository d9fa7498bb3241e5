// tri_tc.cu -- tensor-core FILTERED exact point-to-triangle search (tcgen05 + TMEM + TMA), sm_100a.
//
// Same results as the brute-force kernel of tri.cu -- (dist, region option, face index) bit-identical to the
// reference's TriDistanceKernel (tri_distance/tri_distance.cu:94-211), quirks included -- at ~1 % of its FP32
// work.  The reference evaluates ~100 FP32 instructions for every (point, triangle) pair; here a bf16
// tensor-core contraction bounds every pair from below and only the pairs that can still win are evaluated with
// the exact arithmetic (tri_common.cuh, the same functions tri.cu uses).
//
//   Bound.  Each triangle j gets a bounding sphere (c_j, r_j) around its centroid.  The reference's COMPUTED
//   distance d_j(p) = |p - p2|^2 uses a closest point p2 that lies in the triangle (options 1-5), is the plane
//   projection when p projects inside it (option 0), or -- the reference's option-6 quirk, tri_distance.cu:180 --
//   C + u(B-A), which is farther from p than the true closest point.  In every case
//        d_j(p) >= max(0, |p - c_j| - r_j)^2      (up to fp32 rounding, covered by explicit margins below)
//   (checked over 6.1 M pairs of four shape families in profiles/r01_tri_filter_study.txt: never violated).
//   Triangles for which no such argument holds (zero / denormal edges, slivers with sin^2 < 1e-10, non-finite
//   vertices: "suspect") are never filtered: they are always evaluated exactly.
//
//   1. prep_kernel (per triangle): compact 64-byte record for the exact arithmetic, sphere (c, r), and the packed
//      bf16 B operand such that ONE K=16 MMA row-times-row product against a packed point is
//          h(i,j) = |p_i - c_j|^2 - r_j^2          (fp32 accumulate, |error| <= eps = kEpsScale * R^2)
//      exactly the operand construction of chamfer_tc.cu with |c|^2 - r^2 in place of the squared norm.
//      Suspect, unscanned (the reference's tail skip, tri_distance.cu:129) and padding triangles get h = +1e30.
//   2. scan_kernel: a CTA owns 4 x 128 points and sweeps the triangle blocks of its mesh TWICE through the same
//      TMA ring / tcgen05.mma / TMEM pipeline as chamfer_tc.cu.
//      Pass 1 keeps only the running minimum of h and the 128-triangle tile that holds it; the thread then finds
//      the triangle with the smallest fp32 sphere distance in that tile and evaluates it exactly: U, a tight
//      upper bound on the point's final distance, available before anything is listed.
//      Pass 2 lists (tile, 16-bit mask of its 8-triangle groups) for the groups whose minimum h satisfies
//          h <= s^2 + 2 s r_max + eps,   s = sqrt(U) (1 + 1e-3) + 1e-4 R,   r_max = largest radius in the tile
//      (h > that  =>  |p - c_j| - r_j > s  =>  d_j > U, so the triangle cannot win or tie).  Listed groups are
//      re-tested per triangle with the fp32 sphere (its own r_j, no eps) and the survivors are evaluated with the
//      reference arithmetic; U tightens as better triangles are found.  Selection rule: smaller distance wins,
//      equal distances go to the lower face index = the reference's strict '<' in ascending order with
//      earlier-tile-wins merges (tri_distance.cu:194,202).
//   Reference quirks kept: the first triangle of every 512-tile is taken unconditionally (:194, k == 0), so a
//   NaN there voids that tile for the point (and tile 0 poisons the point's result): NaN can only come from a
//   suspect triangle, so tiles whose first triangle is suspect are flagged by prep_kernel and the scanner
//   evaluates that triangle exactly for its point before trusting anything in the tile.  Points with non-finite
//   coordinates, and inputs whose magnitude defeats the bf16 split (R >= 1e15 or R <= 1e-15), are handled by a
//   per-thread transcription of the reference loop (brute()).
#include "tc_common.cuh"
#include "tri_common.cuh"

namespace geom {
namespace tri_tc {

using namespace geom::tc;
using namespace geom::tri;

constexpr int kGroups = 4;                      // 128-point blocks per CTA == scanner warp groups
constexpr int kRowsPerCta = 128 * kGroups;
constexpr int kTileCols = 128;                  // triangles per ring stage (one packed block)
constexpr int kHalfCols = 64;                   // triangles per MMA / accumulator
constexpr int kScanThreads = kGroups * 128;     // warps 0-15: scanners
constexpr int kThreads = kScanThreads + kGroups * 32;  // warps 16-19: one producer per group
constexpr int kStages = 8;                      // triangle-block ring (TMA)
constexpr int kPrefetch = kStages - 2;
constexpr int kStageRing = 256;                 // per-warp ring of survivors (>= 31 + one batch of 128)
constexpr int kPoolCap = 16384;                 // (point, triangle) pairs per CTA awaiting exact evaluation (64 KB)
constexpr int kTmemCols = kGroups * 2 * kHalfCols;
constexpr unsigned kBlockBytes = 2 * 128 * 16;  // one packed 128-row block
constexpr int kRecVec4 = 4;                     // float4 per compact triangle record
constexpr float kEpsScale = 2.5e-4f;            // |h_mma - h| <= kEpsScale * R^2 (same operand split as chamfer_tc.cu)
constexpr float kBigNorm = 1.0e30f;
constexpr float kRelMargin = 1.0e-3f;           // on sqrt(U)
constexpr float kAbsMargin = 1.0e-4f;           // times R, on sqrt(U)
constexpr float kSphereShrink = 1.0f - 3.0e-5f; // fp32 rounding of the sphere test
constexpr float kDiscShrink = 1.0f - 1.0e-3f;   // fp32 rounding of the disc test
constexpr float kInf = __builtin_huge_valf();

// ---------------------------------------------------------------------------------------------------------------
// compact triangle record (4 x float4):  A | |AB|^2 ;  B | |BC|^2 ;  C | |CA|^2 ;  1/|AB|^2, 1/|BC|^2, 1/|CA|^2, 1/|N|
// Everything else the pair evaluation needs is re-derived with the operations build_record() uses.
// 128-bit read-only load that does not allocate in L1: the records stream through, the sphere array stays cached
__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float sqrt_approx(float v)
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
struct Rec4 {
    float4 r0, r1, r2, r3;
};
__device__ __forceinline__ Rec4 load_rec(const float4 *__restrict__ r)
{
    Rec4 q;
    q.r0 = ldg_stream(r), q.r1 = ldg_stream(r + 1), q.r2 = ldg_stream(r + 2), q.r3 = ldg_stream(r + 3);
    return q;
}
__device__ __forceinline__ TriRegs derive(const Rec4 &q)
{
    TriRegs t;
    t.A = xyz(q.r0), t.B = xyz(q.r1), t.C = xyz(q.r2);
    t.lab = q.r0.w, t.lbc = q.r1.w, t.lca = q.r2.w;
    t.rab = q.r3.x, t.rbc = q.r3.y, t.rca = q.r3.z;
    t.AB = vsub(t.B, t.A), t.BC = vsub(t.C, t.B), t.CA = vsub(t.A, t.C);
    const V3 N = cross_s2(vsub(t.A, t.B), t.CA);
    t.xab = cross_s2(N, t.AB), t.xbc = cross_s2(N, t.BC), t.xca = cross_s2(N, t.CA);
    t.nn = v3(__fmul_rn(N.x, q.r3.w), __fmul_rn(N.y, q.r3.w), __fmul_rn(N.z, q.r3.w));
    t.safe = div_safe_den(t.lab) && div_safe_den(t.lbc) && div_safe_den(t.lca);
    return t;
}
__device__ __forceinline__ TriRegs load_compact(const float4 *__restrict__ r) { return derive(load_rec(r)); }

// The reference cascade on one triangle (distance and region option), from the compact record.
__device__ __forceinline__ float pair_exact_compact(V3 p, const float4 *__restrict__ r, float &option)
{
    const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
    float4 rec[kRecVec];
    build_record(rec, xyz(r0), xyz(r1), xyz(r2));
    return tri_pair_exact(p, rec, option);
}

struct PrepArgs {
    const float *t1, *t2, *t3;
    uint4 *formB;
    float4 *sph, *nrm, *rec;
    unsigned *tinfo, *rmax_bits, *ritem_bits;
    int b, m, m_pad, ntiles128;
    int exact;  // flat scan: no tail skip, one reference tile
};

__global__ void __launch_bounds__(256) prep_kernel(const PrepArgs a)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    float amax = 0.f;
    if (t < (long long)a.b * a.m_pad) {
        const int bi = (int)(t / a.m_pad), j = (int)(t % a.m_pad);
        uint4 b0 = make_uint4(0, 0, 0, 0), b1 = make_uint4(0, 0, pack_bf16x2(kBigNorm, 0.f), 0);
        float4 s = make_float4(0.f, 0.f, 0.f, -1.0f);  // radius < 0: not a candidate
        float rwarp = 0.f;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, r3 = r0, nv = r0;
        if (j < a.m) {
            const size_t g = ((size_t)bi * a.m + j) * 3;
            const V3 A = v3(a.t1[g], a.t1[g + 1], a.t1[g + 2]), B = v3(a.t2[g], a.t2[g + 1], a.t2[g + 2]),
                     C = v3(a.t3[g], a.t3[g + 1], a.t3[g + 2]);
            const V3 AB = vsub(B, A), BC = vsub(C, B), CA = vsub(A, C);
            const V3 N = cross_s2(vsub(A, B), CA);
            const float nn2 = dot_s1(N, N);
            const float inv = __frcp_rn(__fsqrt_rn(nn2));
            const float lab = dot_s1(AB, AB), lbc = dot_s1(BC, BC), lca = dot_s1(CA, CA);
            r0 = make_float4(A.x, A.y, A.z, lab);
            r1 = make_float4(B.x, B.y, B.z, lbc);
            r2 = make_float4(C.x, C.y, C.z, lca);
            r3 = make_float4(__frcp_rn(lab), __frcp_rn(lbc), __frcp_rn(lca), inv);
            // bounding sphere around (an fp32 value near) the centroid; the radius is measured from that value
            const float third = 1.0f / 3.0f;
            const float cx = (A.x + B.x + C.x) * third, cy = (A.y + B.y + C.y) * third, cz = (A.z + B.z + C.z) * third;
            auto d2 = [&](V3 P) {
                const float x = P.x - cx, y = P.y - cy, z = P.z - cz;
                return fmaf(z, z, fmaf(x, x, y * y));
            };
            const float rr = sqrtf(fmaxf(fmaxf(d2(A), d2(B)), d2(C))) * (1.0f + 1.0e-5f);
            bool scanned = true;
            if (!a.exact) {  // tri_distance.cu:129: the last (tile length mod 4) triangles of a tile are never examined
                const int k2 = (j / kRefTile) * kRefTile;
                const int end_k = min(a.m, k2 + kRefTile) - k2;
                scanned = (j - k2) < end_k - (end_k & 3);
            }
            const float big = fmaxf(fmaxf(fmaxf(fabsf(cx), fabsf(cy)), fabsf(cz)), rr);
            // "regular": the geometric lower bound provably holds and the pair evaluation is finite for finite points
            const bool regular = lab >= 1.0e-30f && lbc >= 1.0e-30f && lca >= 1.0e-30f && nn2 >= 1.0e-30f &&
                                 nn2 > 1.0e-10f * lab * lca && inv < kInf && big < 1.0e15f;  // false on NaN
            if (scanned) {
                if (regular) {
                    s = make_float4(cx, cy, cz, rr);
                    nv = make_float4(__fmul_rn(N.x, inv), __fmul_rn(N.y, inv), __fmul_rn(N.z, inv), 0.f);
                    amax = fmaxf(fmaxf(fabsf(cx), fabsf(cy)), fabsf(cz));
                    rwarp = rr;
                    const float xh = bf16_val(cx), yh = bf16_val(cy), zh = bf16_val(cz);
                    const float xl = cx - xh, yl = cy - yh, zl = cz - zh;
                    const float nn = fmaf(cz, cz, fmaf(cx, cx, cy * cy)) - rr * rr;
                    const float n1 = bf16_val(nn), n2 = bf16_val(nn - n1), n3 = nn - n1 - bf16_val(nn - n1);
                    b0 = make_uint4(pack_bf16x2(xh, yh), pack_bf16x2(zh, xh), pack_bf16x2(yh, zh), pack_bf16x2(xl, yl));
                    b1 = make_uint4(pack_bf16x2(zl, 1.f), pack_bf16x2(1.f, 1.f), pack_bf16x2(n1, n2), pack_bf16x2(n3, 0.f));
                } else {
                    s = make_float4(0.f, 0.f, 0.f, kInf);  // infinite radius: passes every sphere test
                    unsigned bits = 1u << ((j & 127) >> 3);
                    if (j % kRefTile == 0 && (!a.exact || j == 0)) bits |= 1u << 16;  // first triangle of a reference tile
                    atomicOr(a.tinfo + (size_t)bi * a.ntiles128 + (j >> 7), bits);
                }
            }
        }
        const size_t o = ((size_t)bi * a.m_pad + (size_t)(j & ~127)) * 2 + (size_t)(j & 127);
        a.formB[o] = b0;
        a.formB[o + 128] = b1;
        a.sph[t] = s;
        a.nrm[t] = nv;
        float4 *rp = a.rec + (size_t)t * kRecVec4;
        rp[0] = r0, rp[1] = r1, rp[2] = r2, rp[3] = r3;
        // m_pad is a multiple of 128, so a warp never straddles two meshes (nor two tiles)
        rwarp = fmaxf(rwarp, __shfl_xor_sync(0xffffffffu, rwarp, 16));
        rwarp = fmaxf(rwarp, __shfl_xor_sync(0xffffffffu, rwarp, 8));
        rwarp = fmaxf(rwarp, __shfl_xor_sync(0xffffffffu, rwarp, 4));
        rwarp = fmaxf(rwarp, __shfl_xor_sync(0xffffffffu, rwarp, 2));
        rwarp = fmaxf(rwarp, __shfl_xor_sync(0xffffffffu, rwarp, 1));
        // one atomic per warp for the largest radius of the 128-triangle TILE (a warp's 32 triangles share it)
        if ((threadIdx.x & 31) == 0 && rwarp > 0.f) atomicMax(a.ritem_bits + (size_t)bi * a.ntiles128 + (j >> 7), __float_as_uint(rwarp));
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 16));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 8));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(a.rmax_bits, __float_as_uint(amax));
}

// Points as the A operand:  -2ph(3) -2pl(3) -2ph(3)  n1 n2 n3  1 1 1  0   (chamfer_tc.cu's query form).
// Non-finite coordinates raise R to inf / NaN bits, which switches the whole call to brute().
__global__ void __launch_bounds__(256)
pack_points_kernel(int b, int n, int n_pad, const float *__restrict__ xyz, uint4 *__restrict__ formA, unsigned *__restrict__ rmax_bits)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    float amax = 0.f;
    if (t < (long long)b * n_pad) {
        const int bi = (int)(t / n_pad), r = (int)(t % n_pad);
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
        if (r < n) {
            const float *p = xyz + ((size_t)bi * n + r) * 3;
            const float x = p[0], y = p[1], z = p[2];
            amax = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
            if (!(amax < kInf)) amax = 0.f;  // non-finite points are handled per thread (brute()), not through R
            const float xh = bf16_val(x), yh = bf16_val(y), zh = bf16_val(z);
            const float xl = x - xh, yl = y - yh, zl = z - zh;
            const float nn = fmaf(z, z, fmaf(x, x, y * y));
            const float n1 = bf16_val(nn), n2 = bf16_val(nn - n1), n3 = nn - n1 - bf16_val(nn - n1);
            a0 = make_uint4(pack_bf16x2(-2.f * xh, -2.f * yh), pack_bf16x2(-2.f * zh, -2.f * xl),
                            pack_bf16x2(-2.f * yl, -2.f * zl), pack_bf16x2(-2.f * xh, -2.f * yh));
            a1 = make_uint4(pack_bf16x2(-2.f * zh, n1), pack_bf16x2(n2, n3), pack_bf16x2(1.f, 1.f), pack_bf16x2(1.f, 0.f));
        }
        const size_t o = ((size_t)bi * n_pad + (size_t)(r & ~127)) * 2 + (size_t)(r & 127);
        formA[o] = a0;
        formA[o + 128] = a1;
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 16));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 8));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(rmax_bits, __float_as_uint(amax));
}

struct ScanArgs {
    const float *xyz;       // (b, n, 3)
    const uint4 *formA;     // packed points     (b, n_pad / 128 blocks)
    const uint4 *formB;     // packed triangles  (b, m_pad / 128 blocks)
    const float4 *sph;      // (b, m_pad)  centre | radius (inf: always evaluate, < 0: not a candidate)
    const float4 *nrm;      // (b, m_pad)  unit normal (regular triangles)
    const float4 *rec;      // (b, m_pad, 4) compact records
    const unsigned *tinfo;  // (b, ntiles128)  bits 0-15: groups of 8 that hold a suspect triangle; bit 16: the tile starts a
                            //                 reference tile whose first triangle is suspect
    const unsigned *rmax_bits, *ritem_bits;
    float *dist;            // (b, n) or null
    int *point, *index;     // (b, n)
    int n, n_pad, m, m_pad;
    int ntiles;             // 128-triangle tiles that hold scanned triangles (swept twice)
    int ntiles128;          // row stride of tinfo (m_pad / 128)
    int ref_tiles128;       // 128-tiles per reference tile: 4, or a number > ntiles for the flat (exact) scan
    int exact;
    int empty_tail;         // the trailing reference tile holds 1-3 triangles: its (10000, 0, 0) beats d > 10000 (:202)
    float eps_scale;
};

#ifdef TRI_TC_TIMING
// development build only (tools/tri_tc_timing.py): per scanner warp phase time stamps and work counters
__device__ unsigned long long g_timing[1024 * 16 * 16];
#define TT_STAMP(slot) do { if (lane == 0) tt_[slot] = clock64(); } while (0)
#define TT_COUNT(slot, v) do { tt_[slot] += (unsigned long long)(v); } while (0)
#else
#define TT_STAMP(slot) do { } while (0)
#define TT_COUNT(slot, v) do { } while (0)
#endif

struct Stage {                 // one ring stage = one 128-triangle tile
    uint4 b[2][kTileCols];     // packed B operand [k-chunk][row]
    float4 sph[kTileCols];     // centre | radius   (second sweep only)
    float4 nrm[kTileCols];     // unit normal       (second sweep only)
};
struct ScanSmem {
    uint4 sA[kGroups][2][128];             // [row block][k-chunk][row]  canonical no-swizzle K-major
    Stage st[kStages];
    u64 key[kScanThreads];                 // per point: (exact distance bits << 32 | face index) of the best triangle so far
    float px[kScanThreads], py[kScanThreads], pz[kScanThreads];  // the CTA's points
    float b2[kScanThreads];                // per point: squared conservative bound s^2 the disc test compares with
    unsigned pool[kPoolCap];               // (point << 22 | face) pairs that passed the disc test, awaiting exact evaluation
    unsigned short items[kScanThreads / 32][32 * 16];  // per warp: (lane << 4 | group) work items of the tile being swept
    unsigned stage[kScanThreads / 32][kStageRing];     // per warp: survivors on their way to the pool (ring buffer)
    unsigned long long full_bar[kGroups][2], empty_bar[kGroups][2];
    unsigned long long b_full[kStages], b_free[kStages];
    unsigned long long a_full;
    unsigned tmem_slot;
    int pool_n;                            // pairs appended (may exceed kPoolCap: the excess was evaluated on the spot)
    int next_chunk;                        // work-stealing cursor of the final evaluation
};

// minimum of the 8 values v[o..o+7]
__device__ __forceinline__ float min8(const unsigned (&v)[64], int o)
{
    const float m1 = min3(__uint_as_float(v[o]), __uint_as_float(v[o + 1]), __uint_as_float(v[o + 2]));
    const float m2 = min3(__uint_as_float(v[o + 3]), __uint_as_float(v[o + 4]), __uint_as_float(v[o + 5]));
    return fminf(m2, min3(__uint_as_float(v[o + 6]), __uint_as_float(v[o + 7]), m1));
}

__global__ void __launch_bounds__(kThreads, 1) scan_kernel(const ScanArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScanSmem &S = *reinterpret_cast<ScanSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bi = blockIdx.y;
    const int row0 = blockIdx.x * kRowsPerCta;
    const int ntiles = a.ntiles;
    const int T = 2 * ntiles;  // two sweeps
    const uint4 *gB = a.formB + (size_t)bi * a.m_pad * 2;
    const float4 *sphb = a.sph + (size_t)bi * a.m_pad;
    const float4 *nrmb = a.nrm + (size_t)bi * a.m_pad;

    // ring refill: step tt of the double sweep reads tile tt % ntiles; the second sweep also needs its spheres and normals
    auto issue_stage = [&](int tt) {
        Stage &st = S.st[tt % kStages];
        unsigned long long *bar = &S.b_full[tt % kStages];
        const int t = tt % ntiles;
        if (tt < ntiles) {
            mbar_expect_tx(bar, kBlockBytes);
            tma_copy(&st.b[0][0], gB + (size_t)t * 256, kBlockBytes, bar);
        } else {
            mbar_expect_tx(bar, kBlockBytes + 2 * kTileCols * 16);
            tma_copy(&st.b[0][0], gB + (size_t)t * 256, kBlockBytes, bar);
            tma_copy(&st.sph[0], sphb + (size_t)t * kTileCols, kTileCols * 16, bar);
            tma_copy(&st.nrm[0], nrmb + (size_t)t * kTileCols, kTileCols * 16, bar);
        }
    };

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_slot)), "r"((unsigned)kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == kScanThreads) {
        for (int g = 0; g < kGroups; ++g)
            for (int h = 0; h < 2; ++h) {
                mbar_init(&S.full_bar[g][h], 1);
                mbar_init(&S.empty_bar[g][h], 4);
            }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&S.b_full[s], 1);
            mbar_init(&S.b_free[s], kScanThreads / 32);  // every scanner warp releases the stage when it is done with the tile
        }
        mbar_init(&S.a_full, 1);
        S.pool_n = 0;
        S.next_chunk = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tma_load(&S.sA[0][0][0], a.formA + ((size_t)bi * a.n_pad + row0) * 2, kGroups * kBlockBytes, &S.a_full);
        for (int t = 0; t < kPrefetch && t < T; ++t) issue_stage(t);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = S.tmem_slot;

    if (warp >= kGroups * 4) {
        // ================================ producers (one warp per point block) ================================
        // warp-uniform, visibly to ptxas: the MMA operands live in uniform registers and the issue is one elected
        // instruction instead of a per-lane R2UR loop
        const int g = __shfl_sync(0xffffffffu, warp, 0) - kGroups * 4;
        const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kHalfCols >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const u64 da = make_desc(smem_u32(&S.sA[g][0][0]), 128 * 16, 128);
        mbar_wait(&S.a_full, 0);
        for (int t = 0; t < T; ++t) {
            const int s = t % kStages;
            if (g == 0) {
                const int tl = t + kPrefetch;
                if (tl < T) {
                    if (t >= 2) mbar_wait(&S.b_free[tl % kStages], (unsigned)(((t - 2) / kStages) & 1));
                    if (lane == 0) issue_stage(tl);
                }
            }
            mbar_wait(&S.b_full[s], (unsigned)((t / kStages) & 1));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (t >= 1) mbar_wait(&S.empty_bar[g][h], (unsigned)((t - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
                    const u64 db = make_desc(smem_u32(&S.st[s].b[0][h * kHalfCols]), kTileCols * 16, 128);
                    asm volatile(
                        "{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}" ::"r"(tmem_u + (unsigned)((g * 2 + h) * kHalfCols)),
                        "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(smem_u32(&S.full_bar[g][h]))
                        : "memory");
                }
            }
        }
    } else {
        // ================================ scanners ================================
        const int g = warp >> 2;
        const int rloc = (warp & 3) * 32 + lane;
        const int qrow = row0 + g * 128 + rloc;
        const bool valid = qrow < a.n;
        V3 p0;
        {
            const float *q = a.xyz + ((size_t)bi * a.n + min(qrow, a.n - 1)) * 3;
            p0 = v3(q[0], q[1], q[2]);
        }
        const float R = __uint_as_float(*a.rmax_bits);
        const unsigned *rtile = a.ritem_bits + (size_t)bi * a.ntiles128;  // largest sphere radius per 128-triangle tile
        const float pmax = fabsf(p0.x) + fabsf(p0.y) + fabsf(p0.z);  // NaN / inf propagate
        // the filter needs finite operands of a magnitude the bf16 split handles; anything else: the reference loop below
        const bool filter_ok = (R < 1.0e15f) && (R > 1.0e-15f) && (pmax < kInf);  // false on NaN
        const bool active = valid && filter_ok;
        const float eps = a.eps_scale * R * R;
        const float4 *recb = a.rec + (size_t)bi * a.m_pad * kRecVec4;
        const unsigned *tinfo = a.tinfo + (size_t)bi * a.ntiles128;
        unsigned short *items = &S.items[warp][0];
        const int wbase = tid & ~31;  // first point of this warp

        // The best triangle so far lives in S.key[point] (distance bits << 32 | face index; ~0 = none): an unsigned 64-bit
        // minimum is "smaller distance wins, equal distances go to the lower face index" = the reference's rule
        // (distances are >= +0, so their bits order like their values; NaN never enters).
        S.key[tid] = ~0ull;
        S.px[tid] = p0.x, S.py[tid] = p0.y, S.pz[tid] = p0.z;
        // outside the sweeps the point is re-read from shared memory: three registers less through the TMEM loops
        auto P = [&]() { return v3(S.px[tid], S.py[tid], S.pz[tid]); };
#ifdef TRI_TC_TIMING
        unsigned long long tt_[16] = {};
#endif
        float thr0 = active ? kInf : -kInf, thr1 = 0.f;  // listing threshold on h (see below); -inf: this thread lists nothing
        bool poisoned = false;   // the first triangle of reference tile 0 evaluates to NaN for this point (:194,:202)
        bool dead_now = false;   // the same for the reference tile being swept: nothing in it can be selected

        // first triangle of a reference tile that prep flagged: evaluate it for this point (uniform branch)
        int ref_pos = 0;  // position of the tile inside its reference tile (a runtime modulo costs ~25 instructions)
        auto check_first = [&](int t, unsigned ti) {
            const bool first = ref_pos == 0;
            ref_pos = (ref_pos + 1 == a.ref_tiles128) ? 0 : ref_pos + 1;
            if (first) {
                dead_now = false;
                if (ti & (1u << 16)) {
                    float opt;
                    const float d0 = pair_exact_compact(P(), recb + (size_t)t * 128 * kRecVec4, opt);
                    dead_now = d0 != d0;
                    if (t == 0) poisoned = dead_now;
                }
            }
        };
        // exact evaluation of one pooled pair
        auto eval_code = [&](unsigned code) {
            const int pt = (int)(code >> 22), j = (int)(code & 0x3fffffu);
            const float d = tri_pair_fast(v3(S.px[pt], S.py[pt], S.pz[pt]), load_compact(recb + (size_t)j * kRecVec4));
            if (d == d) atomicMin(&S.key[pt], make_key(d, (unsigned)j));  // NaN never wins (:194)
        };

        const unsigned tbase = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(g * 2 * kHalfCols);
        // one half tile: wait for its MMA, pull the 64 values, hand the accumulator back
        auto pull = [&](unsigned (&v)[64], int tt, int h) {
            // ONE wait per tile: the MMAs complete in order, so half 1 done implies half 0 done
            if (h == 0) mbar_wait(&S.full_bar[g][1], (unsigned)(tt & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem_ld64_wait(v, tbase + (unsigned)(h * kHalfCols));
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty_bar[g][h]);
        };

        TT_STAMP(0);
        // ---------------- sweep 1: which group of 8 triangles holds the smallest h? ----------------
        float mt = kInf;
        int bg = -1;
        for (int t = 0; t < ntiles; ++t) {
            check_first(t, __ldg(tinfo + t));
            float hm = kInf;
            int hg = 0;
            {
                unsigned v[64];
                pull(v, t, 0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float m = min8(v, 8 * i);
                    if (m < hm) hm = m, hg = i;
                }
            }
            {
                unsigned v[64];
                pull(v, t, 1);
                if (lane == 0) mbar_arrive(&S.b_free[t % kStages]);  // both MMAs of the tile are complete for this warp's group
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float m = min8(v, 8 * i);
                    if (m < hm) hm = m, hg = 8 + i;
                }
            }
            if (!dead_now && hm < mt) mt = hm, bg = t * 16 + hg;
        }
        TT_STAMP(1);
        ref_pos = 0;
        // the triangle of that group with the smallest fp32 sphere value gives the first exact distance U0
        if (active && bg >= 0) {
            const V3 p = P();
            float4 s8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s8[i] = __ldg(sphb + bg * 8 + i);
            float hb = kInf;
            int jb = -1;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float x = p.x - s8[i].x, y = p.y - s8[i].y, z = p.z - s8[i].z;
                const float hh = fmaf(-s8[i].w, s8[i].w, fmaf(z, z, fmaf(x, x, y * y)));
                if (s8[i].w >= 0.0f && s8[i].w < kInf && hh < hb) hb = hh, jb = bg * 8 + i;
            }
            if (jb >= 0) {
                const float d = tri_pair_fast(P(), load_compact(recb + (size_t)jb * kRecVec4));
                if (d == d) S.key[tid] = make_key(d, (unsigned)jb);
            }
        }
        {
            // bounds for the second sweep: s = sqrt(U0) (1 + rel) + abs R;  list h <= s^2 + 2 s r_max + eps;  disc test vs s^2
            const u64 k0 = S.key[tid];
            const float u0 = (k0 == ~0ull) ? kInf : __uint_as_float((unsigned)(k0 >> 32));
            const float sUp = fmaf(sqrtf(u0), 1.0f + kRelMargin, kAbsMargin * R);
            S.b2[tid] = sUp * sUp;
            // per tile t: thr = s^2 + 2 s r_tile(t) + eps = thr0 + thr1 * r_tile(t)
            thr0 = active ? fmaf(sUp, sUp, eps) : -kInf;
            thr1 = active ? 2.0f * sUp : 0.0f;
        }
        __syncwarp();
        TT_STAMP(2);

        // survivors go through a per-warp ring and reach the CTA pool 32 at a time (one shared-memory atomic per 32 pairs)
        unsigned *stg = &S.stage[warp][0];
        int st_head = 0, st_tail = 0;
        auto push32 = [&](int n) {  // n <= 32 entries from the ring head to the pool
            int slot = 0;
            if (lane == 0) slot = atomicAdd(&S.pool_n, n);
            slot = __shfl_sync(0xffffffffu, slot, 0) + lane;
            if (lane < n) {
                const unsigned code = stg[(st_head + lane) & (kStageRing - 1)];
                if (slot < kPoolCap)
                    S.pool[slot] = code;
                else
                    eval_code(code);  // pool full (pathological inputs): evaluate on the spot
            }
            st_head += n;
            __syncwarp();
        };
        // ---------------- sweep 2: list groups, disc-test their triangles while the tile is on chip, pool the survivors ----------------
        for (int t = 0; t < ntiles; ++t) {
            const int tt = ntiles + t;
            const unsigned ti = __ldg(tinfo + t);
            check_first(t, ti);
            const float thr = fmaf(thr1, __uint_as_float(__ldg(rtile + t)), thr0);  // inf stays inf, -inf stays -inf (thr1 = 0)
            unsigned mask = 0;
            {
                unsigned v[64];
                pull(v, tt, 0);
#pragma unroll
                for (int i = 0; i < 8; ++i) mask |= !(min8(v, 8 * i) > thr) ? (1u << i) : 0u;
            }
            {
                unsigned v[64];
                pull(v, tt, 1);
#pragma unroll
                for (int i = 0; i < 8; ++i) mask |= !(min8(v, 8 * i) > thr) ? (1u << (8 + i)) : 0u;
            }
            if (active) mask |= ti & 0xffffu;  // groups with suspect triangles are always listed
            if (dead_now) mask = 0;
            if (__any_sync(0xffffffffu, mask != 0)) {
                // (lane, group) work items of the whole warp, then 4 items x 8 triangles per step: every lane tests one triangle
                int incl = __popc(mask);
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += o;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                {
                    int o = incl - __popc(mask);
                    for (unsigned mm = mask; mm; mm &= mm - 1) items[o++] = (unsigned short)((lane << 4) | (__ffs((int)mm) - 1));
                }
                mbar_wait(&S.b_full[tt % kStages], (unsigned)((tt / kStages) & 1));  // the tile's spheres and normals have landed
                __syncwarp();
                const Stage &st = S.st[tt % kStages];
                TT_COUNT(9, total);
                for (int base = 0; base < total; base += 16) {
                    // a batch of up to 16 items = 4 independent steps of 32 triangle tests (one latency chain per batch)
                    bool ok[4];
                    unsigned code[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int it = base + 4 * k + (lane >> 3);
                        ok[k] = false;
                        code[k] = 0;
                        if (it < total) {
                            const unsigned item = items[it];
                            const int pt = wbase + (int)(item >> 4);
                            const int c = (int)(item & 15u) * 8 + (lane & 7);
                            const float4 s4 = st.sph[c];
                            const float x = S.px[pt] - s4.x, y = S.py[pt] - s4.y, z = S.pz[pt] - s4.z;
                            bool o = s4.w >= 0.0f;  // < 0: not a candidate (tail skip, padding)
                            if (s4.w < kInf) {  // regular triangle: the disc (centre, radius, normal) that contains it bounds the distance
                                const float4 nv = st.nrm[c];
                                const float dc2 = fmaf(z, z, fmaf(x, x, y * y));
                                const float h = fmaf(z, nv.z, fmaf(x, nv.x, y * nv.y));
                                const float h2 = h * h;
                                const float s2 = fmaxf(fmaf(-4.0e-6f, dc2, dc2 - h2), 0.0f);  // conservative in-plane distance^2
                                const float lat = fmaxf(sqrt_approx(s2) - s4.w, 0.0f);  // 2^-22 relative: inside kDiscShrink
                                o = o && !(fmaf(lat, lat, h2) * kDiscShrink > S.b2[pt]);
                            }
                            ok[k] = o;
                            code[k] = ((unsigned)pt << 22) | (unsigned)(t * kTileCols + c);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned okb = __ballot_sync(0xffffffffu, ok[k]);
                        if (ok[k]) stg[(st_tail + __popc(okb & ((1u << lane) - 1u))) & (kStageRing - 1)] = code[k];
                        st_tail += __popc(okb);
                    }
                    __syncwarp();
                    while (st_tail - st_head >= 32) push32(32);
                }
            }
            if (lane == 0) mbar_arrive(&S.b_free[tt % kStages]);  // this warp no longer reads the stage
        }
        if (st_tail > st_head) push32(st_tail - st_head);
        TT_STAMP(3);
        // ---------------- exact evaluation of the pooled pairs: 32 at a time, any warp, any point ----------------
        asm volatile("bar.sync 1, %0;" ::"n"(kScanThreads) : "memory");
        TT_STAMP(4);
        {
            const int npool = min(S.pool_n, kPoolCap);
            auto grab = [&]() {
                int c0 = 0;
                if (lane == 0) c0 = atomicAdd(&S.next_chunk, 32);
                return __shfl_sync(0xffffffffu, c0, 0);
            };
            // the next chunk's records are requested before the current chunk is evaluated (L2 latency off the chain)
            int c0 = grab();
            unsigned code = c0 + lane < npool ? S.pool[c0 + lane] : 0xffffffffu;
            Rec4 rq;
            if (code != 0xffffffffu) rq = load_rec(recb + (size_t)(code & 0x3fffffu) * kRecVec4);
            while (c0 < npool) {
                const int c1 = grab();
                const unsigned code1 = c1 + lane < npool ? S.pool[c1 + lane] : 0xffffffffu;
                Rec4 rn;
                if (code1 != 0xffffffffu) rn = load_rec(recb + (size_t)(code1 & 0x3fffffu) * kRecVec4);
                TT_COUNT(8, 1);
                if (code != 0xffffffffu) {
                    const int pt = (int)(code >> 22), j = (int)(code & 0x3fffffu);
                    const float d = tri_pair_fast(v3(S.px[pt], S.py[pt], S.pz[pt]), derive(rq));
                    if (d == d) atomicMin(&S.key[pt], make_key(d, (unsigned)j));
                }
                c0 = c1, code = code1, rq = rn;
            }
        }
        TT_STAMP(5);
        asm volatile("bar.sync 1, %0;" ::"n"(kScanThreads) : "memory");
        TT_STAMP(6);
#ifdef TRI_TC_TIMING
        if (lane == 0 && (blockIdx.y * gridDim.x + blockIdx.x) < 1024) {
            unsigned long long *o = g_timing + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + warp) * 16;
            for (int i = 0; i < 16; ++i) o[i] = tt_[i];
        }
#endif

        if (valid) {
            float d = 10000.0f, opt = 0.0f;  // tri_distance.cu:127-128 with nothing scanned
            int idx = 0;
            if (!filter_ok) {
                // the reference loop itself (tri_distance.cu:111-206) for this one point
                float res = 0.f;
                int res_i = 0;
                bool res_empty = true;
                const int step = a.exact ? max(a.m, 1) : kRefTile;
                for (int k2 = 0; k2 < a.m; k2 += step) {
                    const int end_k = min(a.m, k2 + step) - k2;
                    const int end_ka = a.exact ? end_k : end_k - (end_k & 3);
                    float tb = 10000.0f;
                    int tb_i = 0;
                    for (int k = 0; k < end_ka; ++k) {
                        const float dk = tri_pair_fast(P(), load_compact(recb + (size_t)(k2 + k) * kRecVec4));
                        if (k == 0 || dk < tb) {
                            tb = dk;
                            tb_i = k2 + k;
                        }
                    }
                    if (k2 == 0 || res > tb) {
                        res = tb;
                        res_i = tb_i;
                        res_empty = end_ka == 0;
                    }
                }
                if (!res_empty) {
                    idx = res_i;
                    d = pair_exact_compact(P(), recb + (size_t)idx * kRecVec4, opt);
                }
            } else {
                if (poisoned) {
                    idx = 0;
                    d = pair_exact_compact(P(), recb, opt);
                } else if (S.key[tid] != ~0ull) {
                    idx = (int)(unsigned)S.key[tid];
                    d = pair_exact_compact(P(), recb + (size_t)idx * kRecVec4, opt);
                }
                if (a.empty_tail && d > 10000.0f) {
                    d = 10000.0f;
                    opt = 0.0f;
                    idx = 0;
                }
            }
            const size_t o = (size_t)bi * a.n + qrow;
            if (a.dist) a.dist[o] = d;
            a.point[o] = (int)opt;
            a.index[o] = idx;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)kTmemCols));
}

static size_t pad_to(size_t v, size_t q) { return (v + q - 1) / q * q; }

struct Layout {
    size_t n_pad, m_pad, off_ritem, off_tinfo, off_formA, off_formB, off_sph, off_nrm, off_rec, total;
};
static Layout layout(int b, int n, int m)
{
    Layout L;
    L.n_pad = pad_to((size_t)n, kRowsPerCta);
    L.m_pad = pad_to((size_t)m, kTileCols);
    size_t o = 256;  // [0]: R bits
    L.off_ritem = o;
    o += pad_to((size_t)b * (L.m_pad / kTileCols) * 4, 256);  // largest radius per tile
    L.off_tinfo = o;
    o += pad_to((size_t)b * (L.m_pad / kTileCols) * 4, 256);
    L.off_formA = o;
    o += (size_t)b * L.n_pad * 32;
    L.off_formB = o;
    o += (size_t)b * L.m_pad * 32;
    L.off_sph = o;
    o += (size_t)b * L.m_pad * 16;
    L.off_nrm = o;
    o += (size_t)b * L.m_pad * 16;
    L.off_rec = o;
    o += (size_t)b * L.m_pad * 16 * kRecVec4;
    L.total = o;
    return L;
}

size_t workspace_bytes(int b, int n, int m) { return layout(b, n, m).total; }

// m_scan_end: one past the last scanned triangle (> 0).  exact: flat scan (GEOM_FLAG_EXACT_TAIL).
cudaError_t launch(int b, int n, const float *xyz, int m, const float *t1, const float *t2, const float *t3, int m_scan_end,
                   int exact, int empty_tail, float *dist, int *point, int *index, void *workspace, float eps_factor,
                   cudaStream_t st)
{
    const Layout L = layout(b, n, m);
    char *w = (char *)workspace;
    const int ntiles128 = (int)(L.m_pad / kTileCols);
    cudaError_t e = cudaMemsetAsync(w, 0, L.off_formA, st);  // R, per-mesh radius, tile flags
    if (e != cudaSuccess) return e;
    PrepArgs pa;
    pa.t1 = t1, pa.t2 = t2, pa.t3 = t3;
    pa.formB = (uint4 *)(w + L.off_formB);
    pa.sph = (float4 *)(w + L.off_sph);
    pa.nrm = (float4 *)(w + L.off_nrm);
    pa.rec = (float4 *)(w + L.off_rec);
    pa.tinfo = (unsigned *)(w + L.off_tinfo);
    pa.rmax_bits = (unsigned *)w;
    pa.ritem_bits = (unsigned *)(w + L.off_ritem);
    pa.b = b, pa.m = m, pa.m_pad = (int)L.m_pad, pa.ntiles128 = ntiles128, pa.exact = exact;
    {
        prof::Scope ps(st, prof::kTriTcPrep);
        prep_kernel<<<(unsigned)(((size_t)b * L.m_pad + 255) / 256), 256, 0, st>>>(pa);
        pack_points_kernel<<<(unsigned)(((size_t)b * L.n_pad + 255) / 256), 256, 0, st>>>(b, n, (int)L.n_pad, xyz, (uint4 *)(w + L.off_formA),
                                                                                      (unsigned *)w);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ScanArgs s;
    s.xyz = xyz;
    s.formA = (const uint4 *)(w + L.off_formA);
    s.formB = pa.formB;
    s.sph = pa.sph;
    s.nrm = pa.nrm;
    s.rec = pa.rec;
    s.tinfo = pa.tinfo;
    s.rmax_bits = pa.rmax_bits;
    s.ritem_bits = pa.ritem_bits;
    s.dist = dist, s.point = point, s.index = index;
    s.n = n, s.n_pad = (int)L.n_pad, s.m = m, s.m_pad = (int)L.m_pad;
    s.ntiles = (m_scan_end + kTileCols - 1) / kTileCols;
    s.ntiles128 = ntiles128;
    s.ref_tiles128 = exact ? 0x40000000 : kRefTile / kTileCols;
    s.exact = exact;
    s.empty_tail = empty_tail;
    s.eps_scale = kEpsScale * eps_factor;
    {
        prof::Scope ps(st, prof::kTriTcScan);
        static std::atomic<unsigned long long> smem_done{0};
        if ((e = ensure_dynamic_smem(scan_kernel, (int)sizeof(ScanSmem), smem_done)) != cudaSuccess) return e;
        scan_kernel<<<dim3((unsigned)(L.n_pad / kRowsPerCta), b, 1), kThreads, sizeof(ScanSmem), st>>>(s);
    }
    return cudaGetLastError();
}

}  // namespace tri_tc
}  // namespace geom

#ifdef TRI_TC_TIMING
extern "C" __attribute__((visibility("default"))) int geom_debug_read_tri_timing(unsigned long long *host, size_t count)
{
    return (int)cudaMemcpyFromSymbol(host, geom::tri_tc::g_timing, count * sizeof(unsigned long long));
}
#endif

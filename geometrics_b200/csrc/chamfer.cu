// chamfer.cu -- Chamfer nearest-neighbour search for sm_100a.
//
// Replaces ChamferDistanceKernel / ChamferDistanceKernelLauncher
// (reference: chamfer_distance/chamfer_distance.cu:6-73).  Results are bit-identical to the
// reference kernel on finite inputs: same per-pair rounding sequence, lowest-index tie rule, and
// (by default) the same tail-skip behaviour.  See DESIGN.md "Chamfer" for the derivation.
//
// Design (not a translation of the reference):
//   * The reference evaluates every pair twice (one launch per direction).  d(p,q) is bitwise
//     symmetric, so `twoway_kernel` evaluates each pair ONCE and feeds both the row minimum
//     (nearest xyz2 point of every xyz1 point) and the column minimum (nearest xyz1 point of every
//     xyz2 point).
//   * Distances are computed two at a time with packed fp32x2 instructions (FADD2/FMUL2/FFMA2):
//     3 issue slots per pair instead of 6.
//   * The inner loop tracks only minimum VALUES (3-input FMNMX3) plus a coarse tag saying where the
//     minimum was reached (row side: group of 8 columns; column side: block of 256 rows, obtained
//     with one warp-wide CREDUX.MIN per column).  `refine_kernel` then re-evaluates the tagged
//     8 / 256 candidates to recover the exact lowest index.  This removes the per-pair
//     compare+select+select of the classic argmin loop.
//   * Every warp is an independent worker over an equal share of (batch, row-block, column-tile)
//     work items: no block-level barriers, no tail imbalance, 16 warps/SM.
//   * Partial results from different workers meet in 64-bit keys (distance bits << 32 | tag) merged
//     with filtered atomicMin, which also implements "lowest index wins ties".
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace geom {
namespace chamfer {

constexpr int kWarps = 4;
constexpr int kThreads = 32 * kWarps;
constexpr int kGroupCols = 8;           // tag granularity of the row keys (two 4-column steps)
constexpr int kRefTile = 512;           // the reference kernel's smem tile (chamfer_distance.cu:15)

struct TwoWayArgs {
    const float *xyz1;
    const float *xyz2;
    u64 *rowkey;  // [b][n_full]
    u64 *colkey;  // [b][m_full]
    unsigned *ticket;    // work counter, starts at 0xffffffff (part of the 0xff-filled workspace)
    int n, m;            // scanned extents (tail-trimmed in compatibility mode)
    int n_full, m_full;  // real point counts = array strides
    int n_rblk, n_ctile, tile_cols;
    int strip;           // column tiles per work unit
    int strips_per_row;  // ceil(n_ctile / strip)
    unsigned total_units;
};

// One warp = one worker.  Work units (row block of 32*R xyz1 rows) x (strip of column tiles) are
// handed out by a global ticket counter, so every warp stays busy until the whole search is done
// (a static split leaves the SM sub-partitions' lower-priority warps running alone at the end).
// The next column tile (of this unit, or the first tile of the next unit, whose ticket is drawn
// early) is prefetched with LDGSTS into a raw per-warp buffer while the current tile is evaluated.
//   ORD = 0: the compiler schedules the packed-FP32 stream; ORD = 1: phased `asm volatile` order
//   (8 independent chains per phase, dependent operations 8 issues apart).
template <int R, int MINB, int TILE, int ORD>
__global__ void __launch_bounds__(kThreads, MINB) twoway_kernel(const TwoWayArgs a)
{
    constexpr int kRowsPerWarp = 32 * R;
    // Per-warp private staging: candidate tile as duplicated SoA pairs {v,v} so one 64-bit register
    // pair feeds a packed op against two different rows; the tile's column minima; the raw prefetch.
    __shared__ __align__(16) float2 s_q[kWarps][3][TILE];
    __shared__ __align__(16) unsigned s_col[kWarps][TILE];
    __shared__ __align__(16) float s_raw[kWarps][3 * TILE];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float2(*sq)[TILE] = s_q[warp];
    unsigned *scol = s_col[warp];
    float *raw = s_raw[warp];

    u64 nx[R / 2], ny[R / 2], nz[R / 2];  // negated row coordinates, rows (2i, 2i+1) packed
    float best[R];
    unsigned grp[R];

    auto fetch = [&]() {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1u) + 1u;
        return __shfl_sync(0xffffffffu, t, 0);
    };
    // asynchronous copy of column tile `ct` of batch item `bi_` into the raw buffer (AoS, tc*3 floats)
    auto prefetch = [&](int bi_, int ct) {
        const int c0 = ct * a.tile_cols;
        const int tc = min(a.tile_cols, a.m - c0);
        const float *p2 = a.xyz2 + ((size_t)bi_ * a.m_full + c0) * 3;
        for (int e = lane; e < tc * 3; e += 32) cp_async4(&raw[e], &p2[e]);
        cp_async_commit();
    };
    auto first_tile = [&](unsigned u_, int &bi_, int &ct_) {
        const unsigned rid = u_ / (unsigned)a.strips_per_row;
        bi_ = (int)(rid / (unsigned)a.n_rblk);
        ct_ = (int)(u_ % (unsigned)a.strips_per_row) * a.strip;
    };

    unsigned u = fetch();
    if (u < a.total_units) {
        int b0, c0;
        first_tile(u, b0, c0);
        prefetch(b0, c0);
    }
    while (u < a.total_units) {
        const unsigned rid = u / (unsigned)a.strips_per_row;
        const int ct0 = (int)(u % (unsigned)a.strips_per_row) * a.strip;
        const int ct1 = min(ct0 + a.strip, a.n_ctile);
        const int bi = (int)(rid / (unsigned)a.n_rblk);
        const int rb = (int)(rid % (unsigned)a.n_rblk);
        const int rbase = rb * kRowsPerWarp;
        unsigned u_next = a.total_units;
        {
            const float *p1 = a.xyz1 + (size_t)bi * a.n_full * 3;
#pragma unroll
            for (int i = 0; i < R / 2; ++i) {
                // rows past the end are clamped duplicates: they add no new values to any minimum
                const int r0 = min(rbase + (2 * i) * 32 + lane, a.n - 1);
                const int r1 = min(rbase + (2 * i + 1) * 32 + lane, a.n - 1);
                nx[i] = pack2(-p1[r0 * 3 + 0], -p1[r1 * 3 + 0]);
                ny[i] = pack2(-p1[r0 * 3 + 1], -p1[r1 * 3 + 1]);
                nz[i] = pack2(-p1[r0 * 3 + 2], -p1[r1 * 3 + 2]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                best[r] = __int_as_float(0x7f800000);
                grp[r] = 0u;
            }
        }

        for (int ct = ct0; ct < ct1; ++ct) {
            // ---- raw tile (prefetched) -> duplicated SoA pairs ----
            const int c0 = ct * a.tile_cols;
            const int tc = min(a.tile_cols, a.m - c0);
            const int ngroups = (tc + kGroupCols - 1) / kGroupCols;
            cp_async_wait_all();
            __syncwarp();
            {
                const int nel = ngroups * kGroupCols * 3;
                for (int e = lane; e < nel; e += 32) {
                    const int pt = e / 3;
                    const int c = e - pt * 3;
                    const float v = raw[min(pt, tc - 1) * 3 + c];  // padding = duplicate of the last column
                    sq[c][pt] = make_float2(v, v);
                }
            }
            __syncwarp();
            // ---- prefetch what this warp evaluates next ----
            if (ct + 1 < ct1) {
                prefetch(bi, ct + 1);
            } else {
                u_next = fetch();
                if (u_next < a.total_units) {
                    int bn, cn;
                    first_tile(u_next, bn, cn);
                    prefetch(bn, cn);
                }
            }

            // ---- all pairs of (32*R rows) x (tile columns) ----
            const unsigned gid0 = (unsigned)(c0 / kGroupCols);
#pragma unroll 1
            for (int g = 0; g < ngroups; ++g) {
                float nb[R];
#pragma unroll
                for (int r = 0; r < R; ++r) nb[r] = best[r];
                if (ORD == 2) {
                    // two columns per step: half the live registers of the 4-column step (more warps/SM)
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int col = g * kGroupCols + s * 2;
                        const ulonglong2 xa = *reinterpret_cast<const ulonglong2 *>(&sq[0][col]);
                        const ulonglong2 ya = *reinterpret_cast<const ulonglong2 *>(&sq[1][col]);
                        const ulonglong2 za = *reinterpret_cast<const ulonglong2 *>(&sq[2][col]);
                        const u64 qx[2] = {xa.x, xa.y}, qy[2] = {ya.x, ya.y}, qz[2] = {za.x, za.y};
                        float d[2][R];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
#pragma unroll
                            for (int i = 0; i < R / 2; ++i) {
                                const u64 dx = add2(qx[c], nx[i]);
                                const u64 dy = add2(qy[c], ny[i]);
                                const u64 dz = add2(qz[c], nz[i]);
                                const u64 dd = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
                                unpack2(dd, d[c][2 * i], d[c][2 * i + 1]);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < R; ++r) nb[r] = min3(nb[r], d[0][r], d[1][r]);
                        unsigned cv[2];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            float cm = min3(d[c][0], d[c][1], d[c][2]);
                            if (R == 4) {
                                cm = fminf(cm, d[c][3]);
                            } else {
#pragma unroll
                                for (int r = 3; r + 1 < R; r += 2) cm = min3(cm, d[c][r], d[c][r + 1]);
                                cm = fminf(cm, d[c][R - 1]);
                            }
                            cv[c] = __reduce_min_sync(0xffffffffu, __float_as_uint(cm));
                        }
                        if (lane == 0) *reinterpret_cast<uint2 *>(&scol[col]) = make_uint2(cv[0], cv[1]);
                    }
                } else {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int col = g * kGroupCols + s * 4;
                    const ulonglong2 xa = *reinterpret_cast<const ulonglong2 *>(&sq[0][col]);
                    const ulonglong2 xb = *reinterpret_cast<const ulonglong2 *>(&sq[0][col + 2]);
                    const ulonglong2 ya = *reinterpret_cast<const ulonglong2 *>(&sq[1][col]);
                    const ulonglong2 yb = *reinterpret_cast<const ulonglong2 *>(&sq[1][col + 2]);
                    const ulonglong2 za = *reinterpret_cast<const ulonglong2 *>(&sq[2][col]);
                    const ulonglong2 zb = *reinterpret_cast<const ulonglong2 *>(&sq[2][col + 2]);
                    const u64 qx[4] = {xa.x, xa.y, xb.x, xb.y};
                    const u64 qy[4] = {ya.x, ya.y, yb.x, yb.y};
                    const u64 qz[4] = {za.x, za.y, zb.x, zb.y};
                    float d[4][R];  // [column][row]
                    if (ORD == 0) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
#pragma unroll
                            for (int i = 0; i < R / 2; ++i) {
                                const u64 dx = add2(qx[c], nx[i]);
                                const u64 dy = add2(qy[c], ny[i]);
                                const u64 dz = add2(qz[c], nz[i]);
                                const u64 dd = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
                                unpack2(dd, d[c][2 * i], d[c][2 * i + 1]);
                            }
                        }
                    } else {
                        // two columns at a time: R independent chains, each phase R issues long
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            u64 t[2][R / 2], w[2][R / 2];
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int i = 0; i < R / 2; ++i) w[c][i] = add2v(qy[ch * 2 + c], ny[i]);
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int i = 0; i < R / 2; ++i) t[c][i] = add2v(qx[ch * 2 + c], nx[i]);
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int i = 0; i < R / 2; ++i) w[c][i] = mul2v(w[c][i], w[c][i]);
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int i = 0; i < R / 2; ++i) w[c][i] = fma2v(t[c][i], t[c][i], w[c][i]);
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int i = 0; i < R / 2; ++i) t[c][i] = add2v(qz[ch * 2 + c], nz[i]);
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int i = 0; i < R / 2; ++i) {
                                    w[c][i] = fma2v(t[c][i], t[c][i], w[c][i]);
                                    unpack2(w[c][i], d[ch * 2 + c][2 * i], d[ch * 2 + c][2 * i + 1]);
                                }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        nb[r] = min3(nb[r], d[0][r], d[1][r]);
                        nb[r] = min3(nb[r], d[2][r], d[3][r]);
                    }
                    unsigned cv[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float cm = min3(d[c][0], d[c][1], d[c][2]);
                        if (R == 4) {
                            cm = fminf(cm, d[c][3]);
                        } else {
#pragma unroll
                            for (int r = 3; r + 1 < R; r += 2) cm = min3(cm, d[c][r], d[c][r + 1]);
                            cm = fminf(cm, d[c][R - 1]);
                        }
                        cv[c] = __reduce_min_sync(0xffffffffu, __float_as_uint(cm));
                    }
                    if (lane == 0) *reinterpret_cast<uint4 *>(&scol[col]) = make_uint4(cv[0], cv[1], cv[2], cv[3]);
                }
                }
                const unsigned gid = gid0 + (unsigned)g;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    grp[r] = (nb[r] < best[r]) ? gid : grp[r];
                    best[r] = nb[r];
                }
            }
            __syncwarp();

            // ---- publish the tile's column minima ----
            {
                u64 *ck = a.colkey + (size_t)bi * a.m_full + c0;
                for (int cc = lane; cc < tc; cc += 32) key_min(&ck[cc], ((u64)scol[cc] << 32) | (u64)(unsigned)rb);
            }
            __syncwarp();
        }

        // ---- publish this strip's row minima ----
        u64 *rk = a.rowkey + (size_t)bi * a.n_full;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = rbase + r * 32 + lane;
            if (row < a.n) key_min(&rk[row], make_key(best[r], grp[r]));
        }
        u = u_next;
    }
}

struct RefineArgs {
    const float *xyz1;
    const float *xyz2;
    const u64 *rowkey;
    const u64 *colkey;
    float *dist1;
    int *idx1;
    float *dist2;
    int *idx2;
    int b, n, m, n_full, m_full;
    int rows_per_warp;  // column-key tag granularity (32 * R of the search kernel that ran)
    int zero1, zero2;   // reference "empty final tile" behaviour, see geom_chamfer_fwd
};

// Recover exact indices from (minimum distance, coarse tag): re-evaluate the tagged candidates in
// ascending order and take the first whose distance equals the minimum bit for bit.
//   blocks [0, row_blocks): one thread per xyz1 row, 8 tagged columns (96 contiguous bytes);
//   the rest: one warp per xyz2 column, the 256 tagged rows 32 at a time (coalesced), first hit by ballot.
constexpr int kRefineThreads = 256;

__global__ void __launch_bounds__(kRefineThreads) refine_kernel(const RefineArgs a, int row_blocks)
{
    if ((int)blockIdx.x < row_blocks) {
        const long long gid = (long long)blockIdx.x * kRefineThreads + threadIdx.x;
        if (gid >= (long long)a.b * a.n) return;
        const int bi = (int)(gid / a.n), j = (int)(gid % a.n);
        const u64 key = a.rowkey[(size_t)bi * a.n_full + j];
        float dmin = __uint_as_float((unsigned)(key >> 32));
        const int cb = (int)(unsigned)key * kGroupCols;
        const int ce = min(cb + kGroupCols, a.m);
        const float *q = a.xyz1 + ((size_t)bi * a.n_full + j) * 3;
        const float *cand = a.xyz2 + (size_t)bi * a.m_full * 3;
        const float qx = q[0], qy = q[1], qz = q[2];
        int idx = cb;
        for (int c = cb; c < ce; ++c) {
            if (sqdist_ref(qx, qy, qz, cand[c * 3 + 0], cand[c * 3 + 1], cand[c * 3 + 2]) == dmin) {
                idx = c;
                break;
            }
        }
        if (a.zero1 && dmin > 0.0f) {
            dmin = 0.0f;
            idx = 0;
        }
        if (a.dist1) a.dist1[(size_t)bi * a.n_full + j] = dmin;
        a.idx1[(size_t)bi * a.n_full + j] = idx;
        return;
    }
    const int lane = threadIdx.x & 31;
    const long long g2 = ((long long)blockIdx.x - row_blocks) * (kRefineThreads / 32) + (threadIdx.x >> 5);
    if (g2 >= (long long)a.b * a.m) return;
    const int bi = (int)(g2 / a.m), k = (int)(g2 % a.m);
    const u64 key = a.colkey[(size_t)bi * a.m_full + k];
    float dmin = __uint_as_float((unsigned)(key >> 32));
    const int rb0 = (int)(unsigned)key * a.rows_per_warp;
    const float *q = a.xyz2 + ((size_t)bi * a.m_full + k) * 3;
    const float *cand = a.xyz1 + (size_t)bi * a.n_full * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    int idx = rb0;
#pragma unroll 2
    for (int r0 = rb0; r0 < min(rb0 + a.rows_per_warp, a.n); r0 += 32) {
        const int r = r0 + lane;
        bool hit = false;
        if (r < a.n) hit = sqdist_ref(qx, qy, qz, cand[r * 3 + 0], cand[r * 3 + 1], cand[r * 3 + 2]) == dmin;
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
            idx = r0 + __ffs(mask) - 1;
            break;
        }
    }
    if (lane == 0) {
        if (a.zero2 && dmin > 0.0f) {
            dmin = 0.0f;
            idx = 0;
        }
        if (a.dist2) a.dist2[(size_t)bi * a.m_full + k] = dmin;
        a.idx2[(size_t)bi * a.m_full + k] = idx;
    }
}

// Plain one-direction search (classic argmin loop, QR queries per thread).  Used for the <=3
// leftover queries the tail-trimmed two-way pass does not cover, for degenerate sizes, and when
// GEOM_FLAG_ONEWAY is set.  Candidates are scanned in ascending order with strict '<', so the
// lowest index wins ties exactly as in the reference.
constexpr int kOneWayQR = 4;
constexpr int kOneWayThreads = 128;
constexpr int kOneWayTile = 512;

__global__ void __launch_bounds__(kOneWayThreads)
oneway_kernel(const float *__restrict__ qpts, int q_full, int q0, int nq,
              const float *__restrict__ cpts, int c_full, int mc,
              float *__restrict__ dist, int *__restrict__ idx, int zero_if_pos, int flat)
{
    __shared__ __align__(16) float sx[kOneWayTile], sy[kOneWayTile], sz[kOneWayTile];
    const int bi = blockIdx.y;
    const float *q = qpts + (size_t)bi * q_full * 3;
    const float *c = cpts + (size_t)bi * c_full * 3;
    float qx[kOneWayQR], qy[kOneWayQR], qz[kOneWayQR], best[kOneWayQR];
    int besti[kOneWayQR], row[kOneWayQR];
#pragma unroll
    for (int r = 0; r < kOneWayQR; ++r) {
        row[r] = q0 + (blockIdx.x * kOneWayQR + r) * kOneWayThreads + threadIdx.x;
        const int rr = min(row[r], q0 + nq - 1);
        qx[r] = q[rr * 3 + 0];
        qy[r] = q[rr * 3 + 1];
        qz[r] = q[rr * 3 + 2];
        best[r] = 0.f;
        besti[r] = 0;
    }
    // The reference's control flow, so that NaN / Inf inputs get the reference's answers: the first candidate of every
    // 512-tile is taken unconditionally (chamfer_distance.cu:39), tiles merge with a strict '>' (:47).  flat
    // (GEOM_FLAG_EXACT_TAIL): one tile spanning every candidate.
    float tb[kOneWayQR];
    int tbi[kOneWayQR];
    for (int k0 = 0; k0 < mc; k0 += kOneWayTile) {
        const int tk = min(kOneWayTile, mc - k0);
        const bool tile_start = !flat || k0 == 0;
        __syncthreads();
        for (int e = threadIdx.x; e < tk * 3; e += kOneWayThreads) {
            const int pt = e / 3, cc = e - pt * 3;
            const float v = c[(size_t)k0 * 3 + e];
            (cc == 0 ? sx : cc == 1 ? sy : sz)[pt] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < tk; ++k) {
            const float cx = sx[k], cy = sy[k], cz = sz[k];
#pragma unroll
            for (int r = 0; r < kOneWayQR; ++r) {
                const float d = sqdist_ref(qx[r], qy[r], qz[r], cx, cy, cz);
                if ((k == 0 && tile_start) || d < tb[r]) {
                    tb[r] = d;
                    tbi[r] = k0 + k;
                }
            }
        }
        if (!flat || k0 + kOneWayTile >= mc) {
#pragma unroll
            for (int r = 0; r < kOneWayQR; ++r)
                if ((flat ? true : k0 == 0) || best[r] > tb[r]) {
                    best[r] = tb[r];
                    besti[r] = tbi[r];
                }
        }
    }
#pragma unroll
    for (int r = 0; r < kOneWayQR; ++r) {
        if (row[r] < q0 + nq) {
            float d = best[r];
            int bi_ = besti[r];
            if (zero_if_pos && d > 0.0f) {
                d = 0.0f;
                bi_ = 0;
            }
            if (dist) dist[(size_t)bi * q_full + row[r]] = d;
            idx[(size_t)bi * q_full + row[r]] = bi_;
        }
    }
}

// Two-way path only: its key merges assume ordered (non-NaN) distances.  One pass flags non-finite coordinates ...
__global__ void __launch_bounds__(256) finite_check_kernel(const float *__restrict__ a, long long na, const float *__restrict__ b_, long long nb,
                                                           unsigned *__restrict__ flag)
{
    const long long stride = (long long)gridDim.x * 256;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < na + nb; i += stride) acc += (i < na ? a[i] : b_[i - na]) * 0.0f;
    if (__syncthreads_or(acc != 0.0f || acc != acc)) {  // x * 0 is NaN exactly for NaN / Inf
        if (threadIdx.x == 0) atomicExch(flag, 1u);
    }
}
// ... and when the flag is set every row is redone with the reference's own control flow (a no-op launch otherwise).
__global__ void __launch_bounds__(256)
nonfinite_fixup_kernel(const unsigned *__restrict__ flag, int b, int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2,
                       float *dist1, int *idx1, float *dist2, int *idx2, int exact)
{
    if (*flag == 0u) return;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)b * (n + m)) return;
    const bool second = t >= (long long)b * n;
    const long long u = second ? t - (long long)b * n : t;
    const int nq = second ? m : n, nc = second ? n : m;
    const int bi = (int)(u / nq);
    const float *q = (second ? xyz2 : xyz1) + (size_t)u * 3;
    const float *c = (second ? xyz1 : xyz2) + (size_t)bi * nc * 3;
    float res;
    int res_i;
    chamfer_ref_row(q[0], q[1], q[2], c, nc, exact != 0, res, res_i);
    float *dist = second ? dist2 : dist1;
    if (dist) dist[u] = res;
    (second ? idx2 : idx1)[u] = res_i;
}

static cudaError_t launch_oneway(int b, const float *q, int q_full, int q0, int nq, const float *c,
                                 int c_full, int mc, float *dist, int *idx, int zero, int flat, cudaStream_t st)
{
    if (nq <= 0 || b <= 0) return cudaSuccess;
    const int per_cta = kOneWayThreads * kOneWayQR;
    dim3 grid((nq + per_cta - 1) / per_cta, b);
    {
        prof::Scope ps(st, prof::kChamferOneWay);
        oneway_kernel<<<grid, kOneWayThreads, 0, st>>>(q, q_full, q0, nq, c, c_full, mc, dist, idx, zero, flat);
    }
    return cudaGetLastError();
}

// Fill one direction's outputs with (0, 0): what the reference leaves behind when a direction has
// no candidate to scan (chamfer_distance.cu:29-30 with end_ka == 0 at k2 == 0).
static cudaError_t fill_zero(int b, int cnt, float *dist, int *idx, cudaStream_t st)
{
    if ((size_t)b * cnt == 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (dist) e = cudaMemsetAsync(dist, 0, sizeof(float) * (size_t)b * cnt, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(idx, 0, sizeof(int) * (size_t)b * cnt, st);
    return e;
}

struct Tune {
    int variant, strip, tile_cols;
};
static Tune read_tune()
{
    Tune t = {1, 2, 0};  // phased FP2 order, 256-column tiles, strips of 2 (profiles/r01_chamfer_tune_v2.jsonl)
    // tuning sweeps (tools/tune_chamfer.py) override through geom_debug_set; knob value v+1 selects variant v
    const int v = (int)debug::knob(GEOM_DEBUG_CHAMFER_VARIANT), s = (int)debug::knob(GEOM_DEBUG_CHAMFER_STRIP),
              tc = (int)debug::knob(GEOM_DEBUG_CHAMFER_TILE_COLS);
    if (v >= 1 && v <= 8) t.variant = v - 1;
    if (s >= 1) t.strip = s;
    if (tc >= 32) t.tile_cols = tc;
    return t;
}

}  // namespace chamfer
}  // namespace geom

using namespace geom;

extern "C" size_t geom_chamfer_workspace_bytes(int b, int n, int m)
{
    if (b < 0 || n < 0 || m < 0) return 0;
    return sizeof(u64) * (size_t)b * ((size_t)n + (size_t)m) + 16;  // keys + the work ticket
}

extern "C" size_t geom_chamfer_tensor_workspace_bytes(int b, int n, int m)
{
    if (b < 0 || n < 0 || m < 0) return 0;
    const size_t keys = geom_chamfer_workspace_bytes(b, n, m);  // degenerate sizes fall back to the plain kernels
    const size_t tc = chamfer_tc::workspace_bytes(b, n, m);     // packed bf16 operands
    return keys > tc ? keys : tc;
}

extern "C" int geom_chamfer_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                                float *dist1, int *idx1, float *dist2, int *idx2,
                                void *workspace, size_t workspace_bytes, unsigned flags,
                                geom_stream_t stream)
{
    using namespace geom::chamfer;
    if (b < 0 || n < 0 || m < 0) return GEOM_ERR_INVALID_ARG;
    if ((long long)b * n * 3 > 0x7fffffffLL || (long long)b * m * 3 > 0x7fffffffLL) return GEOM_ERR_INVALID_ARG;
    if (b == 0 || (n == 0 && m == 0)) return GEOM_OK;
    if ((n > 0 && (!xyz1 || !idx1)) || (m > 0 && (!xyz2 || !idx2))) return GEOM_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;

    // Reference tail semantics (chamfer_distance.cu:19-33,47): the scan of the final 512-tile is
    // truncated to a multiple of 4; if that leaves the final tile empty its (best=0, best_i=0)
    // overwrites every positive result.
    const bool exact = (flags & GEOM_FLAG_EXACT_TAIL) != 0;
    const int m_scan = exact ? m : m - ((m % kRefTile) & 3);  // candidates of direction 1
    const int n_scan = exact ? n : n - ((n % kRefTile) & 3);  // candidates of direction 2
    const int zero1 = (!exact && (m % kRefTile) >= 1 && (m % kRefTile) <= 3) ? 1 : 0;
    const int zero2 = (!exact && (n % kRefTile) >= 1 && (n % kRefTile) <= 3) ? 1 : 0;

    cudaError_t e = cudaSuccess;
    // Path choice: GEOM_FLAG_FP32 / GEOM_FLAG_ONEWAY force the CUDA-core kernels, GEOM_FLAG_TENSOR the tensor-core
    // filtered search; with neither, the tensor-core search runs whenever the caller's workspace is large enough.
    const bool tc_possible = !(flags & (GEOM_FLAG_ONEWAY | GEOM_FLAG_FP32)) && n_scan > 0 && m_scan > 0 && b <= 65535;
    const size_t tc_need = tc_possible ? geom_chamfer_tensor_workspace_bytes(b, n, m) : 0;
    const bool tc_fits = workspace && workspace_bytes >= tc_need && !((uintptr_t)workspace & 15u);
    if (tc_possible && ((flags & GEOM_FLAG_TENSOR) || tc_fits)) {
        if (!tc_fits) return GEOM_ERR_WORKSPACE;
        return status_from_cuda(chamfer_tc::launch(b, n, xyz1, m, xyz2, n_scan, m_scan, zero1, zero2, exact ? 1 : 0, dist1,
                                                   idx1, dist2, idx2, workspace, st));
    }
    const bool twoway = !(flags & GEOM_FLAG_ONEWAY) && n_scan > 0 && m_scan > 0;
    if (twoway) {
        const size_t need = geom_chamfer_workspace_bytes(b, n, m);
        if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 7u)) return GEOM_ERR_WORKSPACE;
        int dev = 0, sms = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return status_from_cuda(e);
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
            return status_from_cuda(e);

        // Kernel variant: rows per lane R, resident CTAs per SM, widest column tile.  The default was
        // picked on B200 (profiles/); GEOM_CHAMFER_TUNE="variant,strip,tile_cols" overrides for tuning.
        struct Variant {
            int rows_per_lane, ctas_per_sm, max_tile;
            void (*kernel)(const TwoWayArgs);
        };
        static const Variant kVariants[] = {
            {8, 4, 256, twoway_kernel<8, 4, 256, 0>},
            {8, 4, 256, twoway_kernel<8, 4, 256, 1>},
            {8, 4, 128, twoway_kernel<8, 4, 128, 0>},
            {8, 4, 128, twoway_kernel<8, 4, 128, 1>},
            {8, 3, 256, twoway_kernel<8, 3, 256, 1>},
            {8, 5, 256, twoway_kernel<8, 5, 256, 2>},
            {8, 6, 128, twoway_kernel<8, 6, 128, 2>},
            {8, 4, 256, twoway_kernel<8, 4, 256, 2>},
        };
        static const Tune tune = read_tune();
        const Variant &v = kVariants[tune.variant];
        const int rows_per_warp = 32 * v.rows_per_lane;

        TwoWayArgs a;
        a.xyz1 = xyz1;
        a.xyz2 = xyz2;
        a.rowkey = (u64 *)workspace;
        a.colkey = a.rowkey + (size_t)b * n;
        a.ticket = (unsigned *)(a.colkey + (size_t)b * m);
        a.n = n_scan;
        a.m = m_scan;
        a.n_full = n;
        a.m_full = m;
        a.n_rblk = (n_scan + rows_per_warp - 1) / rows_per_warp;
        int ctas = sms * v.ctas_per_sm;
        const long long slots = (long long)ctas * kWarps;
        // widest column tile that still gives every worker >= 8 tiles to pull
        int tile_cols = tune.tile_cols > 0 ? min(tune.tile_cols, v.max_tile) : v.max_tile;
        for (; tile_cols > 32; tile_cols >>= 1) {
            const long long tiles = (long long)b * a.n_rblk * ((m_scan + tile_cols - 1) / tile_cols);
            if (tiles >= 8 * slots) break;
        }
        a.tile_cols = tile_cols;
        a.n_ctile = (m_scan + tile_cols - 1) / tile_cols;
        a.strip = max(1, min(tune.strip, a.n_ctile));
        a.strips_per_row = (a.n_ctile + a.strip - 1) / a.strip;
        const long long units = (long long)b * a.n_rblk * a.strips_per_row;
        if (units > 0x7fffffffLL) return GEOM_ERR_INVALID_ARG;
        a.total_units = (unsigned)units;
        if (units < slots) ctas = (int)((units + kWarps - 1) / kWarps);

        if ((e = cudaMemsetAsync(workspace, 0xff, need, st)) != cudaSuccess) return status_from_cuda(e);
        {
            prof::Scope ps(st, prof::kChamferTwoWay);
            v.kernel<<<ctas, kThreads, 0, st>>>(a);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return status_from_cuda(e);

        RefineArgs r;
        r.xyz1 = xyz1;
        r.xyz2 = xyz2;
        r.rowkey = a.rowkey;
        r.colkey = a.colkey;
        r.dist1 = dist1;
        r.idx1 = idx1;
        r.dist2 = dist2;
        r.idx2 = idx2;
        r.b = b;
        r.n = n_scan;
        r.m = m_scan;
        r.n_full = n;
        r.m_full = m;
        r.rows_per_warp = rows_per_warp;
        r.zero1 = zero1;
        r.zero2 = zero2;
        const long long row_blocks = ((long long)b * n_scan + kRefineThreads - 1) / kRefineThreads;
        const long long col_blocks = ((long long)b * m_scan + kRefineThreads / 32 - 1) / (kRefineThreads / 32);
        {
            prof::Scope ps(st, prof::kChamferRefine);
            refine_kernel<<<(unsigned)(row_blocks + col_blocks), kRefineThreads, 0, st>>>(r, (int)row_blocks);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return status_from_cuda(e);

        // queries beyond the trimmed extents still need an answer over the trimmed candidates
        if ((e = launch_oneway(b, xyz1, n, n_scan, n - n_scan, xyz2, m, m_scan, dist1, idx1, zero1, exact ? 1 : 0, st)) != cudaSuccess)
            return status_from_cuda(e);
        if ((e = launch_oneway(b, xyz2, m, m_scan, m - m_scan, xyz1, n, n_scan, dist2, idx2, zero2, exact ? 1 : 0, st)) != cudaSuccess)
            return status_from_cuda(e);
        {
            // NaN / Inf inputs: detect, then redo with the reference's control flow (both launches are ~free when finite)
            unsigned *flag = a.ticket + 2;
            if ((e = cudaMemsetAsync(flag, 0, 4, st)) != cudaSuccess) return status_from_cuda(e);
            prof::Scope ps(st, prof::kMisc2);
            finite_check_kernel<<<sms * 4, 256, 0, st>>>(xyz1, (long long)b * n * 3, xyz2, (long long)b * m * 3, flag);
            nonfinite_fixup_kernel<<<(unsigned)(((long long)b * (n + m) + 255) / 256), 256, 0, st>>>(flag, b, n, xyz1, m, xyz2, dist1, idx1,
                                                                                                  dist2, idx2, exact ? 1 : 0);
        }
        return status_from_cuda(cudaGetLastError());
    }

    if (n > 0) {
        e = (m_scan > 0) ? launch_oneway(b, xyz1, n, 0, n, xyz2, m, m_scan, dist1, idx1, zero1, exact ? 1 : 0, st)
                         : fill_zero(b, n, dist1, idx1, st);
        if (e != cudaSuccess) return status_from_cuda(e);
    }
    if (m > 0) {
        e = (n_scan > 0) ? launch_oneway(b, xyz2, m, 0, m, xyz1, n, n_scan, dist2, idx2, zero2, exact ? 1 : 0, st)
                         : fill_zero(b, m, dist2, idx2, st);
        if (e != cudaSuccess) return status_from_cuda(e);
    }
    return GEOM_OK;
}

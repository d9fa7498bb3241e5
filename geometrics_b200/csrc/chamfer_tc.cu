// chamfer_tc.cu -- tensor-core FILTERED exact nearest-neighbour search (tcgen05 + TMEM), sm_100a.
//
// Same results as the brute-force kernels of chamfer.cu (bit-identical distances and indices, lowest
// index on ties), at a fraction of the FP32 work.  BASELINE's north_star allows tensor cores "if the
// pairwise-distance tile is cast as a bf16 -2 P Q^T contraction"; a bf16 contraction alone cannot give
// bit-exact fp32 indices, so it is used as a conservative FILTER:
//
//   1. pack_kernel splits every coordinate into two bf16 terms (hi, lo) and every squared norm into
//      three, and lays them out so that ONE K=16 bf16 MMA row-times-row product is
//          dt(i,j) = |p_i|^2 + |q_j|^2 - 2 (ph.qh + pl.qh + ph.ql)            (fp32 accumulate)
//      an approximation of the squared distance with |dt - d| <= eps, eps = kEpsScale * R^2,
//      R = max |coordinate| of both point sets (derivation: DESIGN.md "Tensor-core filter").
//   2. search_kernel: a CTA owns 256 query rows (two M=128 accumulators in TMEM) and streams
//      128-candidate tiles: tcgen05.mma (one elected thread) -> tcgen05.ld -> every thread scans its
//      row's 128 approximate distances with 3-input minima (0.5 instruction per pair).  Only
//      candidates with  dt <= best_exact + eps  are re-evaluated with the reference's exact fp32
//      sequence (sqdist_ref); the true nearest neighbour always passes this test because
//      dt(j*) <= d(j*) + eps <= best_exact + eps.  Candidates are visited in ascending index order and
//      replaced only on a strictly smaller exact distance, so ties resolve to the lowest index
//      exactly as in the reference (chamfer_distance.cu:39,47).
//
// One launch per direction (rows = queries), like the reference's two launches, but each launch costs
// ~0.6 instructions per pair instead of ~10.
#include <cstdlib>

#include <cuda_bf16.h>

#include "common.cuh"

namespace geom {
namespace chamfer_tc {

constexpr int kRB = 4;                    // 128-row blocks per CTA (candidate-tile reuse factor)
constexpr int kRowsPerCta = 128 * kRB;
constexpr int kTileCols = 128;    // candidates per MMA
constexpr int kThreads = 288;  // 8 scanner warps + 1 producer warp
constexpr int kTmemCols = 2 * kTileCols;
constexpr float kEpsScale = 2.5e-4f;
constexpr float kBigNorm = 1.0e30f;  // padding candidates: approximate distance far above anything real

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor: K-major, no swizzle ("interleave"): 8-row x 16-byte core matrices;
// LBO = byte distance between the two 16-byte k-chunks, SBO = byte distance between 8-row groups.
__device__ __forceinline__ u64 make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    u64 d = 0;
    d |= (u64)((saddr & 0x3FFFFu) >> 4);
    d |= (u64)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (u64)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (u64)1 << 46;  // descriptor version 1 (sm_100)
    return d;
}

__device__ __forceinline__ unsigned short bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16_val(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) { return (unsigned)bf16_bits(lo) | ((unsigned)bf16_bits(hi) << 16); }

// Per point: two 32-byte operand rows.
//   A form (query):      -2ph(3) -2pl(3) -2ph(3)  n1 n2 n3   1  1  1   0
//   B form (candidate):    qh(3)   qh(3)   ql(3)   1  1  1   n1 n2 n3  0
// Rows past `cnt` (padding up to a multiple of kRowsPerCta): A form zero.  Rows past `cnt_scan` (padding and
// the reference's skipped tail): B form zero with n1 = kBigNorm, so they can never be a minimum.
__global__ void __launch_bounds__(256)
pack_kernel(int b, int cnt, int cnt_scan, int cnt_pad, const float *__restrict__ xyz, uint4 *__restrict__ formA,
            uint4 *__restrict__ formB, unsigned *__restrict__ rmax_bits)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    float amax = 0.f;
    if (t < (long long)b * cnt_pad) {
        const int bi = (int)(t / cnt_pad), r = (int)(t % cnt_pad);
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
        if (r < cnt) {
            const float *p = xyz + ((size_t)bi * cnt + r) * 3;
            const float x = p[0], y = p[1], z = p[2];
            amax = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
            const float xh = bf16_val(x), yh = bf16_val(y), zh = bf16_val(z);
            const float xl = x - xh, yl = y - yh, zl = z - zh;  // exact; rounded to bf16 when packed
            const float nn = fmaf(z, z, fmaf(x, x, y * y));
            const float n1 = bf16_val(nn), n2 = bf16_val(nn - n1), n3 = nn - n1 - bf16_val(nn - n1);
            // k = 0..7 in the first 16-byte chunk, 8..15 in the second
            a0 = make_uint4(pack_bf16x2(-2.f * xh, -2.f * yh), pack_bf16x2(-2.f * zh, -2.f * xl),
                            pack_bf16x2(-2.f * yl, -2.f * zl), pack_bf16x2(-2.f * xh, -2.f * yh));
            a1 = make_uint4(pack_bf16x2(-2.f * zh, n1), pack_bf16x2(n2, n3), pack_bf16x2(1.f, 1.f), pack_bf16x2(1.f, 0.f));
            b0 = make_uint4(pack_bf16x2(xh, yh), pack_bf16x2(zh, xh), pack_bf16x2(yh, zh), pack_bf16x2(xl, yl));
            b1 = make_uint4(pack_bf16x2(zl, 1.f), pack_bf16x2(1.f, 1.f), pack_bf16x2(n1, n2), pack_bf16x2(n3, 0.f));
        }
        if (r >= cnt_scan) {  // not a candidate (padding, or the reference's skipped tail): never the minimum
            b0 = make_uint4(0, 0, 0, 0);
            b1 = make_uint4(0, 0, pack_bf16x2(kBigNorm, 0.f), 0);
        }
        formA[t * 2] = a0;
        formA[t * 2 + 1] = a1;
        formB[t * 2] = b0;
        formB[t * 2 + 1] = b1;
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 16));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 8));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(rmax_bits, __float_as_uint(amax));  // amax >= 0: bits order like values
}

struct SearchArgs {
    const float *qxyz;    // exact query points   (b, nq, 3)
    const float *cxyz;    // exact candidates     (b, nc_full, 3)
    const uint4 *formA;   // packed queries       (b, nq_pad, 2)
    const uint4 *formB;   // packed candidates    (b, nc_pad, 2)
    float *dist;          // (b, nq) or null
    int *idx;             // (b, nq)
    int nq, nq_pad, nc, nc_full, nc_pad;  // nc = scanned candidates (tail-trimmed), nc_full = array stride
    int zero_if_pos;      // reference "empty final tile" behaviour (chamfer_distance.cu:29-30,47)
};
struct SearchArgs2 {
    SearchArgs dir[2];    // blockIdx.z selects the direction
    const unsigned *rmax_bits;
    float eps_scale;      // kEpsScale (tests shrink it to measure the safety margin)
};

#define GEOM_TMEM_LD32(v, taddr)                                                                                      \
    asm volatile(                                                                                                     \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                     \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                     \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                     \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),      \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),     \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                   \
        : "r"(taddr))

// tcgen05.wait::ld tied to the registers of the load it completes ("+r"), so that no use of those
// registers can be scheduled above the wait.
#define GEOM_TMEM_WAIT(v)                                                                                              \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                      \
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),     \
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),            \
                   "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),          \
                   "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),          \
                   "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory")

// 64 consecutive columns of this warp's 32 TMEM lanes, load and wait in ONE statement (no register ties needed)
__device__ __forceinline__ void tmem_ld64_wait(unsigned (&v)[64], unsigned taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)  // suspend-time hint (ns): park instead of spinning
                     : "memory");
    }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A CTA owns kRB x 128 query rows and streams 128-candidate tiles.  Work unit gu = (pass, tile t, row block rb):
// one M=128 x N=128 x K=16 MMA into accumulator gu&1 (double-buffered in TMEM).
//   warp 8  (producer): cp.async of the candidate tiles, tcgen05.mma issue, tcgen05.commit -> full[gu&1]
//   warps 0-7 (scanners): wait full -> tcgen05.ld (64 columns each) -> arrive empty[gu&1] -> scan registers.
// No CTA-wide barrier in the steady state: the tensor core runs up to two units ahead of the scanners.
// Scanner (warp w, lane l) owns row 32*(w&3)+l of each block and columns [64*(w>>2), +64) of each tile.
__global__ void __launch_bounds__(kThreads, 2) search_kernel(const SearchArgs2 both)
{
    const SearchArgs &a = both.dir[blockIdx.z];
    __shared__ __align__(128) uint4 sA[kRB][2][128];      // [row block][k-chunk][row]   canonical no-swizzle K-major
    __shared__ __align__(128) uint4 sB[2][2][kTileCols];  // [buffer][k-chunk][row]
    __shared__ __align__(16) float sX[4][kTileCols * 3];  // exact fp32 coordinates of the last four candidate tiles
    __shared__ __align__(8) unsigned long long full_bar[2], empty_bar[2], x_bar[4];
    __shared__ unsigned tmem_slot;
    __shared__ float s_part[kRB][2][128];                 // per (row block, column half, row): partial minima
    __shared__ int s_pidx[kRB][128];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bi = blockIdx.y;
    const int row0 = blockIdx.x * kRowsPerCta;
    if (row0 >= a.nq_pad) return;  // the grid covers the longer of the two directions
    const int ntiles = (a.nc + kTileCols - 1) / kTileCols;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((unsigned)kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full_bar[i])), "r"(1u));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty_bar[i])), "r"(8u));
        }
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&x_bar[i])), "r"(32u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    const uint4 *gB = a.formB + (size_t)bi * a.nc_pad * 2;
    {
        const uint4 *ga = a.formA + ((size_t)bi * a.nq_pad + row0) * 2;
        for (int i = tid; i < kRowsPerCta * 2; i += kThreads) {
            const int r = i >> 1;
            cp_async16(&sA[r >> 7][i & 1][r & 127], &ga[i]);
        }
        cp_async_commit();
        cp_async_wait_all();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;

    if (warp == 8) {
        // ================================ producer ================================
        // D = f32, A = B = bf16, both K-major, N = 128, M = 128
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kTileCols >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const float *cexact = a.cxyz + (size_t)bi * a.nc_full * 3;
        auto load_tile = [&](int tile) {
            // packed operands: 128 rows x 2 chunks = 256 16-byte copies, 8 per lane
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = lane + 32 * i;
                cp_async16(&sB[tile & 1][c & 1][c >> 1], &gB[(size_t)tile * kTileCols * 2 + c]);
            }
            // exact coordinates: 384 floats, 12 per lane (4-byte copies: no alignment assumption on the caller's array)
            const int last = a.nc_full * 3 - 1;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const int e = lane + 32 * i;
                cp_async4(&sX[tile & 3][e], &cexact[min(tile * kTileCols * 3 + e, last)]);
            }
            cp_async_commit();
        };
        load_tile(0);
        for (int ts = 0; ts < ntiles; ++ts) {
            // this tile has landed (issued one tile ago): visible to the tensor core and to the scanners
            cp_async_wait_all();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&x_bar[ts & 3]);
            __syncwarp();
            const int gu0 = ts * kRB;
            if (ts + 1 < ntiles) {
                // the other operand buffer was read by the MMAs of tile ts-1; the last of them is unit gu0-1.
                // (sX[(ts+1)&3] was last used by tile ts-3, long retired: scanners lag the MMAs by < 2 units.)
                if (gu0 > 0) mbar_wait(&full_bar[(gu0 - 1) & 1], (unsigned)(((gu0 - 1) >> 1) & 1));
                load_tile(ts + 1);
            }
#pragma unroll
            for (int rb = 0; rb < kRB; ++rb) {
                const int gu = gu0 + rb;
                // accumulator gu&1 was last read by the scanners of unit gu-2
                if (gu >= 2) mbar_wait(&empty_bar[gu & 1], (unsigned)(((gu >> 1) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const u64 da = make_desc(smem_u32(&sA[rb][0][0]), 128 * 16, 128);
                    const u64 db = make_desc(smem_u32(&sB[ts & 1][0][0]), kTileCols * 16, 128);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (unsigned)((gu & 1) * kTileCols)),
                        "l"(da), "l"(db), "r"(idesc), "r"(0u)
                        : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&full_bar[gu & 1])) : "memory");
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ scanners ================================
        const int half = warp >> 2;               // column half of the tile this thread scans
        const int rloc = (warp & 3) * 32 + lane;  // row inside a 128-row block == TMEM lane
        float qx[kRB], qy[kRB], qz[kRB], mt[kRB], best[kRB];  // mt: running min of dt; best: exact distance so far
        int bidx[kRB];
#pragma unroll
        for (int rb = 0; rb < kRB; ++rb) {
            const int qrow = min(row0 + rb * 128 + rloc, a.nq - 1);
            const float *q = a.qxyz + ((size_t)bi * a.nq + qrow) * 3;
            qx[rb] = q[0], qy[rb] = q[1], qz[rb] = q[2];
            // padding rows (zero operand: every dt is 0) start at -inf so that nothing ever passes their filter
            mt[rb] = (row0 + rb * 128 + rloc < a.nq) ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
            best[rb] = __int_as_float(0x7f800000);
            bidx[rb] = 0;
        }
        const float rmax = __uint_as_float(*both.rmax_bits);
        // |dt - d| <= eps; inputs too large for the bf16 split to stay finite disable the filter (eps = inf)
        const float eps2 = (rmax < 1.0e15f) ? 2.0f * both.eps_scale * rmax * rmax : __int_as_float(0x7f800000);
        // block minimum of 32 approximate distances with 3-input minima (0.5 instruction per pair)
        auto block_min = [&](const unsigned(&v)[32]) {
            float bm = min3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
#pragma unroll
            for (int k = 3; k + 1 < 32; k += 2) bm = min3(bm, __uint_as_float(v[k]), __uint_as_float(v[k + 1]));
            return fminf(bm, __uint_as_float(v[31]));
        };
        // One sweep.  mt = running minimum of dt over the candidates seen so far (this block included); a
        // candidate is re-evaluated exactly when dt <= mt + 2 eps.  The final survivors of a fixed-threshold
        // filter {dt <= mt_final + 2 eps} are a subset (mt only decreases), and the true nearest neighbour
        // j* is among them: dt(j*) <= d(j*) + eps <= d(jt) + eps <= dt(jt) + 2 eps = mt_final + 2 eps with
        // jt = argmin dt.  Exact coordinates of the tile come from shared memory (staged by the producer).
        for (int t = 0; t < ntiles; ++t) {
            mbar_wait(&x_bar[t & 3], (unsigned)((t >> 2) & 1));
            const float *xs = &sX[t & 3][half * 64 * 3];
            const int c0 = t * kTileCols + half * 64;
#pragma unroll
            for (int rb = 0; rb < kRB; ++rb) {
                const int gu = t * kRB + rb;
                mbar_wait(&full_bar[gu & 1], (unsigned)((gu >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tbase = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)((gu & 1) * kTileCols + half * 64);
                unsigned va[32], vb[32];
                GEOM_TMEM_LD32(va, tbase);
                GEOM_TMEM_LD32(vb, tbase + 32u);
                GEOM_TMEM_WAIT(va);
                GEOM_TMEM_WAIT(vb);
                // the values are in registers: hand the accumulator back to the tensor core
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[gu & 1]);
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const unsigned(&v)[32] = hb ? vb : va;
                    const float bm = block_min(v);
                    mt[rb] = fminf(mt[rb], bm);
                    const float thr = mt[rb] + eps2;
                    if (!(bm > thr)) {  // a candidate inside the window (NaN counts as inside)
                        unsigned mask = 0;
#pragma unroll
                        for (int k = 0; k < 32; ++k) mask |= (!(__uint_as_float(v[k]) > thr)) ? (1u << k) : 0u;
                        while (mask) {  // ascending index order
                            const int k = hb * 32 + __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int j = c0 + k;
                            if (j < a.nc) {
                                const float *cx = xs + k * 3;
                                const float d = sqdist_ref(qx[rb], qy[rb], qz[rb], cx[0], cx[1], cx[2]);
                                if (d < best[rb]) {  // ascending j + strict '<': lowest index among exact ties
                                    best[rb] = d;
                                    bidx[rb] = j;
                                }
                            }
                        }
                    }
                }
            }
        }
        float (&acc)[kRB] = best;
        // merge the two column halves of every row: smaller exact distance, lower index on ties
        if (half == 1) {
#pragma unroll
            for (int rb = 0; rb < kRB; ++rb) {
                s_part[rb][1][rloc] = acc[rb];
                s_pidx[rb][rloc] = bidx[rb];
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0) {
#pragma unroll
            for (int rb = 0; rb < kRB; ++rb) {
                const int qrow = row0 + rb * 128 + rloc;
                if (qrow >= a.nq) continue;
                float best = acc[rb];
                int bi_ = bidx[rb];
                const float o = s_part[rb][1][rloc];
                const int oi = s_pidx[rb][rloc];
                if (o < best || (o == best && oi < bi_)) {
                    best = o;
                    bi_ = oi;
                }
                if (a.zero_if_pos && best > 0.0f) {
                    best = 0.0f;
                    bi_ = 0;
                }
                const size_t off = (size_t)bi * a.nq + qrow;
                if (a.dist) a.dist[off] = best;
                a.idx[off] = bi_;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)kTmemCols));
}


// ------------------------------------------------------------------------------------------------------
// scan_kernel: the same filter with the survivor work DEFERRED, one accumulator per warp group.
//
// A CTA owns 4 x 128 query rows (A operands resident in shared memory) and streams 128-candidate tiles.
// Row block g has its own TMEM accumulator (128 columns) and its own group of four scanner warps (one per
// TMEM lane quarter), so a thread owns one query row for the whole sweep:
//   warp 16 (producer): cp.async candidate tiles into a 4-stage ring; per tile and group: wait empty[g] ->
//                       tcgen05.mma M128 N128 K16 -> tcgen05.commit -> full[g]
//   warps 0-15 (scanners, g = warp>>2): wait full[g] -> tcgen05.ld 2 x 64 columns -> arrive empty[g] ->
//                       minima of the four 32-candidate blocks with 3-input min trees (0.5 instr / pair).
// No exact arithmetic inside the sweep.  With mt = running minimum of dt for the row, a block whose
// minimum bm satisfies bm <= mt + 2 eps is appended to the thread's EVENT list (predicated store, no
// branch).  After the sweep (or when the list is nearly full) the thread re-evaluates with the exact
// fp32 sequence every candidate of the listed blocks that still satisfy bm <= mt + 2 eps, in ascending
// index order with strict '<'.  Correctness: the true nearest neighbour j* has dt(j*) <= d(j*) + eps <=
// d(jt) + eps <= dt(jt) + 2 eps = mt_final + 2 eps (jt = argmin dt), mt only decreases, so the block of j*
// is listed when it is scanned and survives every later re-check; candidates with the same exact distance
// as j* satisfy the same bound, so the lowest index among ties is found too.
// ------------------------------------------------------------------------------------------------------
constexpr int kGroups = 4;
constexpr int kScanThreads = kGroups * 128;
constexpr int kScanKernelThreads = kScanThreads + 32;
constexpr int kStages = 4;   // candidate-tile ring
constexpr int kEvCap = 24;   // events per thread held in shared memory
constexpr int kScanTmemCols = kGroups * kTileCols;
struct ScanSmem {
    uint4 sA[kGroups][2][128];           // [row block][k-chunk][row]  canonical no-swizzle K-major
    uint4 sB[kStages][2][kTileCols];     // [stage][k-chunk][row]
    uint2 ev[kEvCap][kScanThreads];      // (block index, block minimum of dt)
    unsigned long long full_bar[kGroups], empty_bar[kGroups];
    unsigned tmem_slot;
};

// minimum of 32 approximate distances: 16 three-input minima as a depth-4 tree
__device__ __forceinline__ float tree_min32(const unsigned (&v)[32])
{
    float l1[12];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        l1[i] = min3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
    l1[10] = __uint_as_float(v[30]);
    l1[11] = __uint_as_float(v[31]);
    const float a0 = min3(l1[0], l1[1], l1[2]), a1 = min3(l1[3], l1[4], l1[5]);
    const float a2 = min3(l1[6], l1[7], l1[8]), a3 = min3(l1[9], l1[10], l1[11]);
    return fminf(min3(a0, a1, a2), a3);
}
__device__ __forceinline__ float tree_min32o(const unsigned (&v)[64], int o)
{
    float l1[12];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        l1[i] = min3(__uint_as_float(v[o + 3 * i]), __uint_as_float(v[o + 3 * i + 1]), __uint_as_float(v[o + 3 * i + 2]));
    l1[10] = __uint_as_float(v[o + 30]);
    l1[11] = __uint_as_float(v[o + 31]);
    const float a0 = min3(l1[0], l1[1], l1[2]), a1 = min3(l1[3], l1[4], l1[5]);
    const float a2 = min3(l1[6], l1[7], l1[8]), a3 = min3(l1[9], l1[10], l1[11]);
    return fminf(min3(a0, a1, a2), a3);
}

__global__ void __launch_bounds__(kScanKernelThreads, 1) scan_kernel(const SearchArgs2 both)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScanSmem &S = *reinterpret_cast<ScanSmem *>(smem_raw);
    const SearchArgs &a = both.dir[blockIdx.z];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bi = blockIdx.y;
    const int row0 = blockIdx.x * kRowsPerCta;
    if (row0 >= a.nq_pad) return;  // the grid covers the longer of the two directions
    const int ntiles = (a.nc + kTileCols - 1) / kTileCols;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_slot)), "r"((unsigned)kScanTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < kGroups; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.full_bar[i])), "r"(1u));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.empty_bar[i])), "r"(4u));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    const uint4 *gB = a.formB + (size_t)bi * a.nc_pad * 2;
    {
        const uint4 *ga = a.formA + ((size_t)bi * a.nq_pad + row0) * 2;
        for (int i = tid; i < kRowsPerCta * 2; i += kScanKernelThreads) {
            const int r = i >> 1;
            cp_async16(&S.sA[r >> 7][i & 1][r & 127], &ga[i]);
        }
        cp_async_commit();
        cp_async_wait_all();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = S.tmem_slot;

    if (warp == kGroups * 4) {
        // ================================ producer ================================
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kTileCols >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        auto load_tile = [&](int tile) {
            if (tile < ntiles) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {  // 128 rows x 2 chunks = 256 16-byte copies, 8 per lane
                    const int c = lane + 32 * i;
                    cp_async16(&S.sB[tile % kStages][c & 1][c >> 1], &gB[(size_t)tile * kTileCols * 2 + c]);
                }
            }
            cp_async_commit();  // always: keeps the group count uniform
        };
        load_tile(0);
        load_tile(1);
        for (int t = 0; t < ntiles; ++t) {
            // stage (t+2)%4 was last read by the MMAs of tile t-2; those completed before the scanners could
            // release the accumulators that the MMAs of tile t-1 (already issued) waited for.
            load_tile(t + 2);
            asm volatile("cp.async.wait_group 2;" ::: "memory");  // tile t has landed
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                if (t >= 1) mbar_wait(&S.empty_bar[g], (unsigned)((t - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const u64 da = make_desc(smem_u32(&S.sA[g][0][0]), 128 * 16, 128);
                    const u64 db = make_desc(smem_u32(&S.sB[t % kStages][0][0]), kTileCols * 16, 128);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (unsigned)(g * kTileCols)),
                        "l"(da), "l"(db), "r"(idesc), "r"(0u)
                        : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.full_bar[g])) : "memory");
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ scanners ================================
        const int g = warp >> 2;         // row block == accumulator == warp group
        const int rloc = (warp & 3) * 32 + lane;  // row inside the block == TMEM lane
        const int qrow = row0 + g * 128 + rloc;
        const bool valid = qrow < a.nq;
        float qx, qy, qz;
        {
            const float *q = a.qxyz + ((size_t)bi * a.nq + min(qrow, a.nq - 1)) * 3;
            qx = q[0], qy = q[1], qz = q[2];
        }
        // padding rows (zero operand: every dt is 0) start at -inf so that nothing ever passes their filter
        float mt = valid ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
        float best = __int_as_float(0x7f800000);
        int bidx = 0;
        int cnt = 0;
        const float rmax = __uint_as_float(*both.rmax_bits);
        // |dt - d| <= eps; inputs too large for the bf16 split to stay finite disable the filter (eps = inf)
        const float eps2 = (rmax < 1.0e15f) ? 2.0f * both.eps_scale * rmax * rmax : __int_as_float(0x7f800000);
        const float *cb = a.cxyz + (size_t)bi * a.nc_full * 3;
        const bool vec_ok = (((uintptr_t)cb) & 15u) == 0;
        uint2 *evp = &S.ev[0][tid];

        // exact re-evaluation of the listed blocks that are still inside the window
        auto flush = [&]() {
            const float thr = mt + eps2;
            for (int e = 0; e < cnt; ++e) {
                const uint2 ev = evp[(size_t)e * kScanThreads];
                if (__uint_as_float(ev.y) > thr) continue;  // NaN counts as inside
                const int j0 = (int)ev.x * 32;
                auto cand = [&](int j, float cx, float cy, float cz) {
                    const float d = sqdist_ref(qx, qy, qz, cx, cy, cz);
                    if (j < a.nc && d < best) {  // ascending j + strict '<': lowest index among exact ties
                        best = d;
                        bidx = j;
                    }
                };
                if (vec_ok && j0 + 32 <= a.nc_full) {
                    const float4 *p = reinterpret_cast<const float4 *>(cb + (size_t)j0 * 3);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 u0 = __ldg(p + 3 * i), u1 = __ldg(p + 3 * i + 1), u2 = __ldg(p + 3 * i + 2);
                        cand(j0 + 4 * i, u0.x, u0.y, u0.z);
                        cand(j0 + 4 * i + 1, u0.w, u1.x, u1.y);
                        cand(j0 + 4 * i + 2, u1.z, u1.w, u2.x);
                        cand(j0 + 4 * i + 3, u2.y, u2.z, u2.w);
                    }
                } else {
                    const int jend = min(j0 + 32, a.nc);
                    for (int j = j0; j < jend; ++j) cand(j, __ldg(cb + (size_t)j * 3), __ldg(cb + (size_t)j * 3 + 1), __ldg(cb + (size_t)j * 3 + 2));
                }
            }
            cnt = 0;
        };

        const unsigned tbase = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(g * kTileCols);
        for (int t = 0; t < ntiles; ++t) {
            mbar_wait(&S.full_bar[g], (unsigned)(t & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned v[64];
            tmem_ld64_wait(v, tbase);
            const float bm0 = tree_min32o(v, 0), bm1 = tree_min32o(v, 32);
            tmem_ld64_wait(v, tbase + 64u);
            // the values are in registers: hand the accumulator back to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty_bar[g]);
            const float bm2 = tree_min32o(v, 0), bm3 = tree_min32o(v, 32);
            mt = min3(mt, min3(bm0, bm1, bm2), bm3);
            const float thr = mt + eps2;
            const float bms[4] = {bm0, bm1, bm2, bm3};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!(bms[i] > thr)) {  // predicated append
                    evp[(size_t)cnt * kScanThreads] = make_uint2((unsigned)(t * 4 + i), __float_as_uint(bms[i]));
                    ++cnt;
                }
            }
            if (cnt > kEvCap - 4) flush();
        }
        flush();
        if (valid) {
            if (a.zero_if_pos && best > 0.0f) {
                best = 0.0f;
                bidx = 0;
            }
            const size_t off = (size_t)bi * a.nq + qrow;
            if (a.dist) a.dist[off] = best;
            a.idx[off] = bidx;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)kScanTmemCols));
}

static size_t pad_rows(int cnt) { return ((size_t)cnt + kRowsPerCta - 1) / kRowsPerCta * kRowsPerCta; }

size_t workspace_bytes(int b, int n, int m)
{
    // four packed arrays (A and B form of both sets, 32 B per padded row each) + the R scalar
    return ((size_t)b * (pad_rows(n) + pad_rows(m)) * 2) * 32 + 256;
}

// Both directions.  n_scan / m_scan: candidates actually scanned (reference tail skip); zero1 / zero2: the
// reference's empty-final-tile behaviour.  Requires n_scan > 0 and m_scan > 0.
cudaError_t launch(int b, int n, const float *xyz1, int m, const float *xyz2, int n_scan, int m_scan, int zero1,
                   int zero2, float *dist1, int *idx1, float *dist2, int *idx2, void *workspace, cudaStream_t st)
{
    const size_t np = pad_rows(n), mp = pad_rows(m);
    char *w = (char *)workspace;
    unsigned *rmax = (unsigned *)w;
    uint4 *a1 = (uint4 *)(w + 256);
    uint4 *b1 = a1 + (size_t)b * np * 2;
    uint4 *a2 = b1 + (size_t)b * np * 2;
    uint4 *b2 = a2 + (size_t)b * mp * 2;
    cudaError_t e = cudaMemsetAsync(rmax, 0, 256, st);
    if (e != cudaSuccess) return e;
    {
        prof::Scope ps(st, prof::kMisc2);
        pack_kernel<<<(unsigned)(((size_t)b * np + 255) / 256), 256, 0, st>>>(b, n, n_scan, (int)np, xyz1, a1, b1, rmax);
        pack_kernel<<<(unsigned)(((size_t)b * mp + 255) / 256), 256, 0, st>>>(b, m, m_scan, (int)mp, xyz2, a2, b2, rmax);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    SearchArgs2 s;
    s.rmax_bits = rmax;
    s.eps_scale = kEpsScale;
    if (const char *env = getenv("GEOM_TC_EPS_SCALE")) {  // tests only: multiply the filter margin
        const float f = (float)atof(env);
        if (f > 0.f) s.eps_scale = kEpsScale * f;
    }
    // direction 0: every xyz1 point over xyz2[0, m_scan);  direction 1: every xyz2 point over xyz1[0, n_scan)
    SearchArgs &d0 = s.dir[0], &d1 = s.dir[1];
    d0.qxyz = xyz1, d0.cxyz = xyz2, d0.formA = a1, d0.formB = b2, d0.dist = dist1, d0.idx = idx1;
    d0.nq = n, d0.nq_pad = (int)np, d0.nc = m_scan, d0.nc_full = m, d0.nc_pad = (int)mp, d0.zero_if_pos = zero1;
    d1.qxyz = xyz2, d1.cxyz = xyz1, d1.formA = a2, d1.formB = b1, d1.dist = dist2, d1.idx = idx2;
    d1.nq = m, d1.nq_pad = (int)mp, d1.nc = n_scan, d1.nc_full = n, d1.nc_pad = (int)np, d1.zero_if_pos = zero2;
    {
        prof::Scope ps(st, prof::kChamferTwoWay);
        const size_t rows = np > mp ? np : mp;
        static const bool legacy = getenv("GEOM_TC_LEGACY") != nullptr;  // A/B switch: the inline-survivor kernel
        if (legacy) {
            search_kernel<<<dim3((unsigned)(rows / kRowsPerCta), b, 2), kThreads, 0, st>>>(s);
        } else {
            if ((e = cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem))) != cudaSuccess) return e;
            scan_kernel<<<dim3((unsigned)(rows / kRowsPerCta), b, 2), kScanKernelThreads, sizeof(ScanSmem), st>>>(s);
        }
    }
    return cudaGetLastError();
}

}  // namespace chamfer_tc
}  // namespace geom

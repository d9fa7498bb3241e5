// chamfer_tc.cu -- tensor-core FILTERED exact nearest-neighbour search (tcgen05 + TMEM + TMA), sm_100a.
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
//      The packed operands are stored in the tensor core's shared-memory order (128-row blocks,
//      k-chunk major), so a block is ONE contiguous 4 KB TMA bulk copy.
//   2. scan_kernel: a CTA owns 4 x 128 query rows and streams 128-candidate tiles through a TMA ring.
//      Row block g has two TMEM accumulators (64 candidates each), a producer warp that issues
//      tcgen05.mma (M128 N64 K16) into whichever of the two is free (one elected lane, warp-uniform operands), and
//      four scanner warps (one per TMEM lane quarter) that pull 64 columns with tcgen05.ld, hand the accumulator
//      back and reduce the values with 3-input minima (0.5 instruction per pair).  A thread owns one query row
//      for the whole sweep.  No exact arithmetic inside the sweep: with mt = running minimum of dt for the row,
//      a tile that holds a 32-candidate block with minimum bm <= mt + 2 eps is appended to the thread's EVENT
//      list in shared memory (predicated store, no branch); a full list is compacted (tiles that left the
//      window are dropped).  After the sweep the warp re-evaluates, with the reference's exact fp32 sequence
//      (sqdist_ref), every candidate of the listed blocks that still satisfy bm <= mt + 2 eps: lane c takes
//      candidate c of a block, the minimum and its lowest lane come from REDUX + ballot, the row's owner lane
//      keeps the result with strict '<' (blocks of a row in ascending order).
//      Correctness: the true nearest neighbour j* has dt(j*) <= d(j*) + eps <= d(jt) + eps <=
//      dt(jt) + 2 eps = mt_final + 2 eps (jt = argmin dt); mt only decreases, so the block of j* is
//      listed when it is scanned and survives every later re-check.  Candidates with the same exact
//      distance as j* satisfy the same bound, so the lowest index among exact ties is found too
//      (chamfer_distance.cu:39,47).
//
// One launch covers both directions (1-D grid, see SearchArgs2); ~1 instruction per pair instead of ~10.
#include "tc_common.cuh"

namespace geom {
namespace chamfer_tc {

using namespace geom::tc;

constexpr int kGroups = 4;                      // 128-row blocks per CTA == scanner warp groups
constexpr int kRowsPerCta = 128 * kGroups;
constexpr int kTileCols = 128;                  // candidates per ring stage (one packed block)
constexpr int kHalfCols = 64;                   // candidates per MMA / accumulator
constexpr int kScanThreads = kGroups * 128;     // scanner warps come first
constexpr int kThreads = kScanThreads + kGroups * 32;  // + one producer warp per row block
constexpr int kStages = 8;                      // candidate-tile ring (TMA)
constexpr int kPrefetch = kStages - 2;          // tiles in flight ahead of the MMAs
constexpr int kEvCap = 16;                      // events (listed tiles) per thread held in shared memory
constexpr int kTmemCols = kGroups * 2 * kHalfCols;
constexpr unsigned kBlockBytes = 2 * 128 * 16;  // one packed 128-row block
constexpr float kEpsScale = 2.5e-4f;
constexpr float kBigNorm = 1.0e30f;  // padding candidates: approximate distance far above anything real

// Per point: two 16-byte k-chunks per form.
//   A form (query):      -2ph(3) -2pl(3) -2ph(3)  n1 n2 n3   1  1  1   0
//   B form (candidate):    qh(3)   qh(3)   ql(3)   1  1  1   n1 n2 n3  0
// Rows past `cnt` (padding up to a multiple of kRowsPerCta): A form zero.  Rows past `cnt_scan` (padding and
// the reference's skipped tail): B form zero with n1 = kBigNorm, so they can never be a minimum.
// Storage order (uint4 units): [batch item][128-row block][k-chunk][row in block].
struct PackSet {
    const float *xyz;
    uint4 *formA, *formB;
    int cnt, cnt_scan, cnt_pad;
    unsigned blocks;  // thread blocks of this set
};
// One launch packs both point sets (blockIdx.x < sets[0].blocks: the first).
__global__ void __launch_bounds__(256) pack_kernel(int b, const PackSet s0, const PackSet s1, unsigned *__restrict__ rmax_bits)
{
    const bool second = blockIdx.x >= s0.blocks;
    const PackSet &ps = second ? s1 : s0;
    const int cnt = ps.cnt, cnt_scan = ps.cnt_scan, cnt_pad = ps.cnt_pad;
    const float *__restrict__ xyz = ps.xyz;
    uint4 *__restrict__ formA = ps.formA;
    uint4 *__restrict__ formB = ps.formB;
    const long long t = (long long)(blockIdx.x - (second ? s0.blocks : 0u)) * 256 + threadIdx.x;
    float amax = 0.f;
    if (t < (long long)b * cnt_pad) {
        const int bi = (int)(t / cnt_pad), r = (int)(t % cnt_pad);
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
        if (r < cnt) {
            const float *p = xyz + ((size_t)bi * cnt + r) * 3;
            const float x = p[0], y = p[1], z = p[2];
            amax = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
            if (!(fabsf(x) + fabsf(y) + fabsf(z) < __int_as_float(0x7f800000))) amax = __int_as_float(0x7f800000);  // NaN too (fmaxf drops NaN)
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
        const size_t o = ((size_t)bi * cnt_pad + (size_t)(r & ~127)) * 2 + (size_t)(r & 127);
        formA[o] = a0;
        formA[o + 128] = a1;
        formB[o] = b0;
        formB[o + 128] = b1;
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 16));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 8));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(rmax_bits, __float_as_uint(amax));  // amax >= 0: bits order like values
}

#ifdef CH_TIMING
__device__ unsigned long long g_ch_timing[256 * 20 * 8];
#define CT_DECL unsigned long long ct_[8] = {}; unsigned long long ct_t_ = clock64()
#define CT_LAP(slot) do { const unsigned long long n_ = clock64(); ct_[slot] += n_ - ct_t_; ct_t_ = n_; } while (0)
#define CT_DUMP do { if (lane == 0 && blockIdx.x < 256) for (int i_ = 0; i_ < 8; ++i_) g_ch_timing[((size_t)blockIdx.x * 20 + warp) * 8 + i_] = ct_[i_]; } while (0)
// absolute time line of the CTA (thread 0): globaltimer stamps + SM id
__device__ unsigned long long g_ch_stamps[2048 * 8];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CT_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 2048) g_ch_stamps[(size_t)blockIdx.x * 8 + (i)] = gtime(); } while (0)
#define CT_SMID do { if (threadIdx.x == 0 && blockIdx.x < 2048) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); g_ch_stamps[(size_t)blockIdx.x * 8 + 7] = sm_; } } while (0)
#else
#define CT_STAMP(i) do { } while (0)
#define CT_SMID do { } while (0)
#define CT_DECL do { } while (0)
#define CT_LAP(slot) do { } while (0)
#define CT_DUMP do { } while (0)
#endif

struct SearchArgs {
    const float *qxyz;    // exact query points   (b, nq, 3)
    const float *cxyz;    // exact candidates     (b, nc_full, 3)
    const uint4 *formA;   // packed queries       (b, nq_pad / 128 blocks)
    const uint4 *formB;   // packed candidates    (b, nc_pad / 128 blocks)
    float *dist;          // (b, nq) or null
    int *idx;             // (b, nq)
    int nq, nq_pad, nc, nc_full, nc_pad;  // nc = scanned candidates (tail-trimmed), nc_full = array stride
    int zero_if_pos;      // reference "empty final tile" behaviour (chamfer_distance.cu:29-30,47)
    u64 *keys;            // (b, nq) merge keys (distance bits << 32 | index), 0xff..: only when the candidates are split
    unsigned *done;       // (b, nq_pad / 512) arrival counters of the row blocks, 0: only when the candidates are split
};
struct SearchArgs2 {
    SearchArgs dir[2];    // blockIdx.z selects the direction
    const unsigned *rmax_bits;
    float eps_scale;      // kEpsScale (tests shrink it to measure the safety margin)
    int exact;            // GEOM_FLAG_EXACT_TAIL (only the non-finite slow path needs to know)
    // Grid layout (1-D).  Row blocks (512 query rows of one batch item and direction) are numbered direction 0 first;
    // the first n_whole of them get ONE CTA each that sweeps every candidate tile; each of the others gets `chunks` CTAs
    // that sweep 1/chunks of the tiles and merge through 64-bit keys -- the remainder that would otherwise run as a nearly
    // empty last wave (small batches) is spread over the whole GPU instead.
    int b, blocks0, blocks1;  // row blocks per item of direction 0 / 1
    int n_whole, chunks;
};

struct ScanSmem {
    uint4 sA[kGroups][2][128];           // [row block][k-chunk][row]  canonical no-swizzle K-major
    uint4 sB[kStages][2][kTileCols];     // [stage][k-chunk][row]
    uint4 evm[kEvCap][kScanThreads];     // events: the four 32-candidate block minima of a listed tile (float bits; the
                                         // flush reuses the slots for block ids, so every access is through unsigned) ...
    unsigned evt[kEvCap][kScanThreads];  // ... and the tile index
    unsigned long long full_bar[kGroups][2], empty_bar[kGroups][2];  // accumulator hand-over, per group and half
    unsigned long long b_full[kStages], b_free[kStages];             // ring stage landed / no longer read
    unsigned long long a_full;
    unsigned tmem_slot;
    unsigned last_flag;   // candidate split: this CTA is the last of its row block to arrive
};

// blockIdx.x -> (direction, batch item, row block, candidate chunk, chunks of this row block); see SearchArgs2
__device__ __forceinline__ void decode_cta(const SearchArgs2 &both, int rb_lin, int &dirsel, int &bi, int &rblock, int &chunk, int &nchunks)
{
    chunk = 0, nchunks = 1;
    if (rb_lin >= both.n_whole) {
        const int j = rb_lin - both.n_whole;
        rb_lin = both.n_whole + j / both.chunks;
        chunk = j % both.chunks;
        nchunks = both.chunks;
    }
    dirsel = rb_lin >= both.b * both.blocks0 ? 1 : 0;
    const int per_item = dirsel ? both.blocks1 : both.blocks0;
    const int rb_dir = rb_lin - (dirsel ? both.b * both.blocks0 : 0);
    bi = rb_dir / per_item, rblock = rb_dir % per_item;
}

// Everything a scanner thread needs about its row OUTSIDE the sweep (exact re-evaluation, result write-back), recomputed
// from opaque copies of the block / thread index whenever it is needed: kept in registers through the sweep it would push the
// loop over the 96-register budget of a 20-warp CTA and ptxas would rematerialise loop invariants every tile instead.
struct RowCtx {
    int dirsel, bi, rblock, chunk, nchunks;
    int qrow;
    bool valid, vec_ok;
    float qx, qy, qz;
    const float *cb;       // exact candidates of this CTA's first tile
    int nc_loc, nc_full_loc, tile0;
};
__device__ __forceinline__ RowCtx row_ctx(const SearchArgs2 &both)
{
    int bx = (int)blockIdx.x, tx = (int)threadIdx.x;
    asm volatile("" : "+r"(bx), "+r"(tx));  // opaque: no common subexpressions with the values computed before the sweep
    RowCtx c;
    decode_cta(both, bx, c.dirsel, c.bi, c.rblock, c.chunk, c.nchunks);
    const SearchArgs &a = both.dir[c.dirsel];
    c.qrow = c.rblock * kRowsPerCta + tx;  // scanner thread tx owns row tx of the CTA's 512
    c.valid = c.qrow < a.nq;
    const float *q = a.qxyz + ((size_t)c.bi * a.nq + min(c.qrow, a.nq - 1)) * 3;
    c.qx = q[0], c.qy = q[1], c.qz = q[2];
    c.tile0 = (int)((long long)((a.nc + kTileCols - 1) / kTileCols) * c.chunk / c.nchunks);
    c.cb = a.cxyz + ((size_t)c.bi * a.nc_full + (size_t)c.tile0 * kTileCols) * 3;
    c.nc_loc = a.nc - c.tile0 * kTileCols, c.nc_full_loc = a.nc_full - c.tile0 * kTileCols;
    c.vec_ok = (((uintptr_t)c.cb) & 15u) == 0;
    return c;
}

__global__ void __launch_bounds__(kThreads, 1) scan_kernel(const SearchArgs2 both)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScanSmem &S = *reinterpret_cast<ScanSmem *>(smem_raw);
    CT_STAMP(0);
    CT_SMID;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int dirsel, bi, rblock, chunk, nchunks;
    decode_cta(both, (int)blockIdx.x, dirsel, bi, rblock, chunk, nchunks);
    const SearchArgs &a = both.dir[dirsel];
    const int row0 = rblock * kRowsPerCta;
    // this CTA sweeps candidate tiles [tile0, tile0 + ntiles) of the direction's tiles
    const int all_tiles = (a.nc + kTileCols - 1) / kTileCols;
    const int tile0 = (int)((long long)all_tiles * chunk / nchunks);
    const int ntiles = (int)((long long)all_tiles * (chunk + 1) / nchunks) - tile0;
    const uint4 *gB = a.formB + ((size_t)bi * a.nc_pad + (size_t)tile0 * kTileCols) * 2;

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
            mbar_init(&S.b_free[s], kGroups);
        }
        mbar_init(&S.a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // barrier words visible to the TMA unit
        // the CTA's packed query blocks are contiguous: one copy; then the head of the candidate ring
        tma_load(&S.sA[0][0][0], a.formA + ((size_t)bi * a.nq_pad + row0) * 2, kGroups * kBlockBytes, &S.a_full);
        for (int t = 0; t < kPrefetch && t < ntiles; ++t)
            tma_load(&S.sB[t][0][0], gB + (size_t)t * 256, kBlockBytes, &S.b_full[t]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = S.tmem_slot;
    CT_STAMP(1);

    if (warp >= kGroups * 4) {
        // ================================ producers (one warp per row block) ================================
        // warp-uniform by construction (and visibly so to ptxas): the MMA operands then live in uniform registers and the
        // issue is one predicated instruction instead of a per-lane R2UR loop
        const int g = __shfl_sync(0xffffffffu, warp, 0) - kGroups * 4;
        const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        // D = f32, A = B = bf16, both K-major, N = 64, M = 128
        const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kHalfCols >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const u64 da = make_desc(smem_u32(&S.sA[g][0][0]), 128 * 16, 128);
        mbar_wait(&S.a_full, 0);
        CT_DECL;
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % kStages;
            if ((t & (kGroups - 1)) == g) {
                // ring refill, taken in turns by the producers (one doing all of it would be the slowest chain of the CTA):
                // tile t+kPrefetch goes to the stage of tile t-2, once every group is done with it
                const int tl = t + kPrefetch;
                if (tl < ntiles) {
                    if (t >= 2) mbar_wait_fast(&S.b_free[tl % kStages], (unsigned)(((t - 2) / kStages) & 1));
                    tma_load_elect(&S.sB[tl % kStages][0][0], gB + (size_t)tl * 256, kBlockBytes, &S.b_full[tl % kStages]);
                }
            }
            CT_LAP(0);
            mbar_wait(&S.b_full[s], (unsigned)((t / kStages) & 1));
            CT_LAP(1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // accumulator (g, h) was last read by the scanners of tile t-1
                if (t >= 1) mbar_wait(&S.empty_bar[g][h], (unsigned)((t - 1) & 1));
                CT_LAP(2 + 2 * h);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
                    const u64 db = make_desc(smem_u32(&S.sB[s][0][h * kHalfCols]), kTileCols * 16, 128);
                    asm volatile(
                        "{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}" ::"r"(tmem_u + (unsigned)((g * 2 + h) * kHalfCols)),
                        "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(smem_u32(&S.full_bar[g][h]))
                        : "memory");
                    // both MMAs of this row block on the tile are issued: their completion also releases the ring stage
                    if (h == 1)
                        asm volatile(
                            "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                            "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&S.b_free[s]))
                            : "memory");
                }
                CT_LAP(3 + 2 * h);
            }
        }
        CT_DUMP;
    } else {
        // ================================ scanners ================================
        const int g = warp >> 2;                  // row block == warp group; TMEM lane == tid & 127
        // padding rows (zero operand: every dt is 0) start at -inf so that nothing ever passes their filter
        float mt = (row0 + tid < a.nq) ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
        float best = __int_as_float(0x7f800000);
        int bidx = 0;
        int cnt = 0;
        const float rmax = __uint_as_float(*both.rmax_bits);
        // |dt - d| <= eps; inputs too large for the bf16 split to stay finite disable the filter (eps = inf)
        // (padding candidates carry a norm of kBigNorm = 1e30: they can only stay above every real approximate distance
        // while 12 R^2 << 1e30, hence the 1e13 bound)
        const float eps2 = (rmax < 1.0e13f) ? 2.0f * both.eps_scale * rmax * rmax : __int_as_float(0x7f800000);
        // candidate indices are local to this CTA's tile range inside the sweep (tile0 is added back at the end)
        uint4 *evm = &S.evm[0][tid];
        unsigned *evt = &S.evt[0][tid];

        CT_DECL;
        // ---- exact re-evaluation of the listed tiles' 32-candidate blocks that are still inside the window ----
        unsigned *blk = &S.evm[0][0].x;  // surviving block ids: evm[e][tid] reused as 4 x u32
        auto blk_at = [&](int e) -> unsigned & { return blk[((size_t)(e >> 2) * kScanThreads + tid) * 4 + (e & 3)]; };
        // Step 1: the thread's blocks with bm <= mt + 2 eps, compacted to the front of its list (in place: at most four block
        // ids per event, written behind the read position).
        auto list_blocks = [&]() -> int {
            const float thr = mt + eps2;
            int keep = 0;
            for (int e = 0; e < cnt; ++e) {
                const uint4 mb = evm[(size_t)e * kScanThreads];
                const unsigned t4 = evt[(size_t)e * kScanThreads] * 4u;
                // block ids of event e go to evm[keep >> 2]: keep <= 4 e before the event, and event e has already been
                // read, so slot (keep >> 2) <= e is free to overwrite
                const float bms[4] = {__uint_as_float(mb.x), __uint_as_float(mb.y), __uint_as_float(mb.z), __uint_as_float(mb.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!(bms[i] > thr)) {  // NaN counts as inside
                        blk_at(keep) = t4 + i;
                        ++keep;
                    }
                }
            }
            return keep;
        };
        // Step 2 for one block: the reference's arithmetic on its 32 candidates, ascending index + strict '<' (lowest index
        // among exact ties), for the query (x, y, z).
        auto eval_block = [&](const RowCtx &c, int j0, float x, float y, float z, float &bd, int &bj) {
            const float *cb = c.cb;
            const int nc_loc = c.nc_loc;
            auto cand = [&](int j, float cx, float cy, float cz) {
                const float d = sqdist_ref(x, y, z, cx, cy, cz);
                if (j < nc_loc && d < bd) {
                    bd = d;
                    bj = j;
                }
            };
            if (c.vec_ok && j0 + 32 <= c.nc_full_loc) {
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
                const int jend = min(j0 + 32, nc_loc);
                for (int j = j0; j < jend; ++j) cand(j, __ldg(cb + (size_t)j * 3), __ldg(cb + (size_t)j * 3 + 1), __ldg(cb + (size_t)j * 3 + 2));
            }
        };
        // Every thread evaluates its own blocks: max-over-lanes(blocks) rounds per warp.  Only the fallback.
        auto flush_own = [&]() {
            const int keep = list_blocks();
            const RowCtx c = row_ctx(both);
            for (int e = 0; e < keep; ++e) eval_block(c, (int)blk_at(e) * 32, c.qx, c.qy, c.qz, best, bidx);
            cnt = 0;
        };
        // Inside the sweep a full list is only COMPACTED: events whose tile left the window since they were listed are
        // dropped (mt only decreases), no candidate is touched.  Data with very many near-ties can keep the list full; then
        // the thread evaluates what it has.
        auto make_room = [&]() {
            const float thr = mt + eps2;
            int keep = 0;
            for (int e = 0; e < cnt; ++e) {
                const uint4 mb = evm[(size_t)e * kScanThreads];
                const unsigned te = evt[(size_t)e * kScanThreads];
                const float tmin = fminf(fminf(__uint_as_float(mb.x), __uint_as_float(mb.y)), fminf(__uint_as_float(mb.z), __uint_as_float(mb.w)));
                if (!(tmin > thr)) {
                    evm[(size_t)keep * kScanThreads] = mb;
                    evt[(size_t)keep * kScanThreads] = te;
                    ++keep;
                }
            }
            cnt = keep;
            if (cnt > kEvCap / 2) flush_own();
        };
        // After the sweep the warp works through the surviving (row, block) pairs of its 32 rows TOGETHER: lane c takes
        // candidate c of the block, so a block is three coalesced 128-byte lines (per-thread evaluation makes every lane
        // of every load touch its own line, and the L1 tag stage, one line per cycle, becomes the bound); the minimum
        // and its lowest lane come from one REDUX + one ballot, and the row's owner lane keeps the result.  A row's pairs
        // are consecutive and ascending, so strict '<' keeps the lowest index among exact ties.
        auto flush_shared = [&]() {
            const int keep = list_blocks();
            CT_LAP(6);
            int incl = keep;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const RowCtx c = row_ctx(both);
            const float qx = c.qx, qy = c.qy, qz = c.qz;
            const float *cb = c.cb;
            const int nc_loc = c.nc_loc;
            if (total > kEvCap * 32) {  // does not fit the pair list (warp-uniform)
                for (int e = 0; e < keep; ++e) eval_block(c, (int)blk_at(e) * 32, qx, qy, qz, best, bidx);
            } else {
                // every lane is past its events: the warp's evt slots become the pair list (block id << 5 | owner lane)
                unsigned *list = &S.evt[0][warp * 32];
                __syncwarp();
                for (int e = 0; e < keep; ++e) {
                    const int p = incl - keep + e;
                    list[(size_t)(p >> 5) * kScanThreads + (p & 31)] = (blk_at(e) << 5) | (unsigned)lane;
                }
                __syncwarp();
                const float kInfF = __int_as_float(0x7f800000);
                constexpr int kBatch = 8;  // blocks whose loads are in flight together (one L2 round trip per batch)
                for (int base = 0; base < total; base += kBatch) {
                    unsigned pr[kBatch];
                    float cx[kBatch], cy[kBatch], cz[kBatch];
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int p = min(base + u, total - 1);  // the tail repeats the last pair (harmless: strict '<')
                        pr[u] = list[(size_t)(p >> 5) * kScanThreads + (p & 31)];
                        const int j = min((int)(pr[u] >> 5) * 32 + lane, nc_loc - 1);
                        const float *cp = cb + (size_t)j * 3;
                        cx[u] = __ldg(cp), cy[u] = __ldg(cp + 1), cz[u] = __ldg(cp + 2);
                    }
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int r = (int)(pr[u] & 31u), j0 = (int)(pr[u] >> 5) * 32;
                        const float x = __shfl_sync(0xffffffffu, qx, r), y = __shfl_sync(0xffffffffu, qy, r), z = __shfl_sync(0xffffffffu, qz, r);
                        const float d = (j0 + lane < nc_loc) ? sqdist_ref(x, y, z, cx[u], cy[u], cz[u]) : kInfF;
                        // d >= +0 (or +inf): the bit patterns order like the values
                        const unsigned dm = __reduce_min_sync(0xffffffffu, __float_as_uint(d));
                        const unsigned who = __ballot_sync(0xffffffffu, __float_as_uint(d) == dm);
                        // branch-free owner update: the eight blocks of a batch stay one straight-line instruction stream
                        const bool upd = (lane == r) & (__uint_as_float(dm) < best);
                        best = upd ? __uint_as_float(dm) : best;
                        bidx = upd ? j0 + (__ffs(who) - 1) : bidx;
                    }
                }
            }
            cnt = 0;
        };

        // warp-uniform, visibly to ptxas (stays in a uniform register: no R2UR in front of every TMEM load)
        const unsigned tbase = __shfl_sync(0xffffffffu, tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(g * 2 * kHalfCols), 0);
        for (int t = 0; t < ntiles; ++t) {
            float bm0, bm1, bm2, bm3;
            {
                CT_LAP(5);
                // ONE wait per tile: the MMAs complete in order, so half 1 done implies half 0 done
                mbar_wait(&S.full_bar[g][1], (unsigned)(t & 1));
                if (t == 0) CT_STAMP(2);
                CT_LAP(0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned v[64];
                tmem_ld64_wait(v, tbase);
                CT_LAP(1);
                // the values are in registers: hand the accumulator back to the tensor core
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.empty_bar[g][0]);
                bm0 = tree_min32(v, 0);
                bm1 = tree_min32(v, 32);
            }
            {
                CT_LAP(2);
                CT_LAP(3);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned v[64];
                tmem_ld64_wait(v, tbase + (unsigned)kHalfCols);
                CT_LAP(4);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.empty_bar[g][1]);
                bm2 = tree_min32(v, 0);
                bm3 = tree_min32(v, 32);
            }
            const float tm = fminf(min3(bm0, bm1, bm2), bm3);
            mt = fminf(mt, tm);
            if (!(tm > mt + eps2)) {  // predicated append: the tile holds a candidate inside the window
                evm[(size_t)cnt * kScanThreads] = make_uint4(__float_as_uint(bm0), __float_as_uint(bm1), __float_as_uint(bm2), __float_as_uint(bm3));
                evt[(size_t)cnt * kScanThreads] = (unsigned)t;
                ++cnt;
            }
            if (cnt >= kEvCap) make_room();
        }
        CT_LAP(5);
        CT_STAMP(3);
        flush_shared();
        CT_LAP(7);
        CT_STAMP(4);
        CT_DUMP;
        const RowCtx c = row_ctx(both);  // nothing of the grid decode stayed live through the sweep
        const SearchArgs &a = both.dir[c.dirsel];
        const int bi = c.bi, rblock = c.rblock, chunk = c.chunk, nchunks = c.nchunks;
        const bool valid = c.valid;
        const size_t off = (size_t)bi * a.nq + c.qrow;
        bool write = valid;
        bidx += c.tile0 * kTileCols;
        if (!(rmax < __int_as_float(0x7f800000))) {
            // Some input coordinate is NaN / Inf: the reference's answer depends on WHERE it sits in the 512-tiles
            // (chamfer_distance.cu:39,47), so every row is redone by the transcription of the reference loop.
            write = false;
            if (valid && chunk == 0) {
                chamfer_ref_row(c.qx, c.qy, c.qz, a.cxyz + (size_t)bi * a.nc_full * 3, a.nc_full, both.exact != 0, best, bidx);
                if (a.dist) a.dist[off] = best;
                a.idx[off] = bidx;
            }
        } else if (nchunks > 1) {
            // The row block's candidates are split over several CTAs: results meet in 64-bit keys (smaller distance, then
            // lower index = the reference's tie rule); the CTA that arrives last decodes them.
            if (valid) atomicMin(&a.keys[off], make_key(best, (unsigned)bidx));
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(kScanThreads) : "memory");
            if (tid == 0) S.last_flag = atomicAdd(&a.done[(size_t)bi * (a.nq_pad / kRowsPerCta) + rblock], 1u) == (unsigned)nchunks - 1u;
            asm volatile("bar.sync 1, %0;" ::"n"(kScanThreads) : "memory");
            write = valid && S.last_flag;
            if (write) {
                __threadfence();
                const u64 k = *(volatile u64 *)&a.keys[off];
                best = __uint_as_float((unsigned)(k >> 32));
                bidx = (int)(unsigned)k;
            }
        }
        if (write) {
            if (a.zero_if_pos && best > 0.0f) {
                best = 0.0f;
                bidx = 0;
            }
            if (a.dist) a.dist[off] = best;
            a.idx[off] = bidx;
        }
    }
    CT_STAMP(5);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)kTmemCols));
    CT_STAMP(6);
}

// Test hook (geom_debug_tc_dump): the approximate squared distances exactly as the tensor core delivers them -- one CTA,
// one 128-row query block against every 64-candidate half tile, accumulators read back from TMEM and stored.  Lets a test
// assert the filter's error bound |dt - d| <= eps on the hardware itself instead of on a numpy model of it.
__global__ void __launch_bounds__(128) dump_kernel(const uint4 *__restrict__ formA, const uint4 *__restrict__ formB, int nc_pad,
                                                   float *__restrict__ out)
{
    __shared__ __align__(128) uint4 sA[2][128];
    __shared__ __align__(128) uint4 sB[2][kHalfCols];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    sA[0][tid] = formA[tid];
    sA[1][tid] = formA[128 + tid];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kHalfCols >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    unsigned phase = 0;
    for (int c0 = 0; c0 < nc_pad; c0 += kHalfCols) {
        // half tile c0..c0+63 of the packed candidates: block (c0 / 128), rows (c0 % 128)..+63 of both k-chunks
        const uint4 *blk = formB + (size_t)(c0 / 128) * 256 + (c0 % 128);
        if (tid < kHalfCols) {
            sB[0][tid] = blk[tid];
            sB[1][tid] = blk[128 + tid];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const u64 da = make_desc(smem_u32(&sA[0][0]), 128 * 16, 128);
            const u64 db = make_desc(smem_u32(&sB[0][0]), kHalfCols * 16, 128);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(0u)
                         : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        unsigned v[64];
        tmem_ld64_wait(v, tmem + ((unsigned)(warp * 32) << 16));
        for (int k = 0; k < 64; ++k) out[(size_t)(warp * 32 + lane) * nc_pad + c0 + k] = __uint_as_float(v[k]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}

}  // namespace chamfer_tc
}  // namespace geom
#ifdef CH_TIMING
extern "C" __attribute__((visibility("default"))) int geom_debug_read_ch_timing(unsigned long long *host, size_t count)
{
    return (int)cudaMemcpyFromSymbol(host, geom::chamfer_tc::g_ch_timing, count * sizeof(unsigned long long));
}
extern "C" __attribute__((visibility("default"))) int geom_debug_read_ch_stamps(unsigned long long *host, size_t count)
{
    return (int)cudaMemcpyFromSymbol(host, geom::chamfer_tc::g_ch_stamps, count * sizeof(unsigned long long));
}
#endif
namespace geom {
namespace chamfer_tc {

static size_t pad_rows(int cnt) { return ((size_t)cnt + kRowsPerCta - 1) / kRowsPerCta * kRowsPerCta; }

size_t workspace_bytes(int b, int n, int m)
{
    // four packed arrays (A and B form of both sets, 32 B per padded row each) + the R scalar
    // + merge keys (8 B per padded row) and row-block arrival counters for the candidate split of small batches
    const size_t rows = (size_t)b * (pad_rows(n) + pad_rows(m));
    return rows * 2 * 32 + 256 + rows * 8 + (rows / kRowsPerCta) * 4 + 256;
}

// Grid plan: row blocks that fill whole waves of the one-CTA-per-SM grid sweep all their candidate tiles; the remaining
// row blocks are split `chunks` ways so that the last wave covers the GPU too.  Costs in tile-steps, fitted on B200
// (profiles/r02_tc_chunks.txt, table 3): a CTA costs its tiles + ~26 tile-steps of prologue, pipeline fill, final
// re-evaluation and key merge (the absolute overhead, ~11 us, has not changed since the first fit; the tile-step got faster).
static void plan_grid(long long row_blocks, int tiles, int sms, int &n_whole, int &chunks)
{
    const double kCtaOverhead = 26.0;
    const long long rem = row_blocks % sms;
    n_whole = (int)row_blocks;
    chunks = 1;
    if (rem == 0 || row_blocks > 0x3fffffffLL / 8) return;
    const double whole_cost = (double)tiles + kCtaOverhead;  // the remainder as one more wave of whole sweeps
    double best = whole_cost;
    for (int c = 2; c <= 8; ++c) {
        if (tiles / c < 8) break;
        const double waves = (double)((rem * c + sms - 1) / sms);
        const double cost = waves * ((double)tiles / c + kCtaOverhead);
        if (cost < best * 0.97) {
            best = cost;
            chunks = c;
        }
    }
    if (chunks > 1) n_whole = (int)(row_blocks - rem);
}

int ctas_per_item(int n, int m)
{
    const size_t np = pad_rows(n), mp = pad_rows(m);
    return (int)((np + mp) / kRowsPerCta);
}

// Both directions.  n_scan / m_scan: candidates actually scanned (reference tail skip); zero1 / zero2: the
// reference's empty-final-tile behaviour.  Requires n_scan > 0 and m_scan > 0.
cudaError_t launch(int b, int n, const float *xyz1, int m, const float *xyz2, int n_scan, int m_scan, int zero1,
                   int zero2, int exact, float *dist1, int *idx1, float *dist2, int *idx2, void *workspace, cudaStream_t st)
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
        prof::Scope ps(st, prof::kChamferTcPack);
        PackSet p0, p1;
        p0.xyz = xyz1, p0.formA = a1, p0.formB = b1, p0.cnt = n, p0.cnt_scan = n_scan, p0.cnt_pad = (int)np;
        p0.blocks = (unsigned)(((size_t)b * np + 255) / 256);
        p1.xyz = xyz2, p1.formA = a2, p1.formB = b2, p1.cnt = m, p1.cnt_scan = m_scan, p1.cnt_pad = (int)mp;
        p1.blocks = (unsigned)(((size_t)b * mp + 255) / 256);
        pack_kernel<<<p0.blocks + p1.blocks, 256, 0, st>>>(b, p0, p1, rmax);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    SearchArgs2 s;
    s.rmax_bits = rmax;
    s.eps_scale = kEpsScale;
    s.exact = exact;
    if (const float f = debug::knob(GEOM_DEBUG_TC_EPS_FACTOR); f > 0.f) s.eps_scale = kEpsScale * f;  // tests: margin study
    // direction 0: every xyz1 point over xyz2[0, m_scan);  direction 1: every xyz2 point over xyz1[0, n_scan)
    SearchArgs &d0 = s.dir[0], &d1 = s.dir[1];
    d0.qxyz = xyz1, d0.cxyz = xyz2, d0.formA = a1, d0.formB = b2, d0.dist = dist1, d0.idx = idx1;
    d0.nq = n, d0.nq_pad = (int)np, d0.nc = m_scan, d0.nc_full = m, d0.nc_pad = (int)mp, d0.zero_if_pos = zero1;
    d1.qxyz = xyz2, d1.cxyz = xyz1, d1.formA = a2, d1.formB = b1, d1.dist = dist2, d1.idx = idx2;
    d1.nq = m, d1.nq_pad = (int)mp, d1.nc = n_scan, d1.nc_full = n, d1.nc_pad = (int)np, d1.zero_if_pos = zero2;
    {
        int dev = 0, sms = 148;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        const int tiles = (int)(((n_scan > m_scan ? n_scan : m_scan) + kTileCols - 1) / kTileCols);
        s.b = b;
        s.blocks0 = (int)(np / kRowsPerCta);
        s.blocks1 = (int)(mp / kRowsPerCta);
        const long long row_blocks = (long long)b * (s.blocks0 + s.blocks1);
        if (row_blocks > 0x3fffffffLL / 8) return cudaErrorInvalidValue;
        plan_grid(row_blocks, tiles, sms, s.n_whole, s.chunks);
        if (const float f = debug::knob(GEOM_DEBUG_TC_CHUNKS); f >= 1.f && f <= 8.f) {  // tests: split EVERY row block
            s.chunks = (int)f;
            s.n_whole = s.chunks > 1 ? 0 : (int)row_blocks;
        }
        d0.keys = d1.keys = nullptr;
        d0.done = d1.done = nullptr;
        if (s.chunks > 1) {
            char *kbase = (char *)(b2 + (size_t)b * mp * 2);  // behind the packed operands (16-byte aligned)
            const size_t key_bytes = (size_t)b * (np + mp) * 8, cnt_bytes = (size_t)b * ((np + mp) / kRowsPerCta) * 4;
            d0.keys = (u64 *)kbase;
            d1.keys = d0.keys + (size_t)b * n;
            d0.done = (unsigned *)(kbase + key_bytes);
            d1.done = d0.done + (size_t)b * (np / kRowsPerCta);
            if ((e = cudaMemsetAsync(kbase, 0xff, key_bytes, st)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(kbase + key_bytes, 0, cnt_bytes, st)) != cudaSuccess) return e;
        }
    }
    {
        prof::Scope ps(st, prof::kChamferTcScan);
        static std::atomic<unsigned long long> smem_done{0};
        if ((e = ensure_dynamic_smem(scan_kernel, (int)sizeof(ScanSmem), smem_done)) != cudaSuccess) return e;
        const long long row_blocks = (long long)b * (s.blocks0 + s.blocks1);
        const unsigned ctas = (unsigned)(s.n_whole + (row_blocks - s.n_whole) * s.chunks);
        scan_kernel<<<ctas, kThreads, sizeof(ScanSmem), st>>>(s);
    }
    return cudaGetLastError();
}

// dt (128 x pad128(m)) = approximate squared distances of the first min(n, 128) points of xyz1 (one batch item) against the
// m points of xyz2, as the MMA computes them; *rmax = R.  workspace >= workspace_bytes(1, 128, m).
cudaError_t debug_dump(int n, const float *xyz1, int m, const float *xyz2, float *dt, float *rmax_out, void *workspace,
                       cudaStream_t st)
{
    const size_t np = pad_rows(n < 128 ? n : 128), mp = pad_rows(m);
    char *w = (char *)workspace;
    unsigned *rmax = (unsigned *)w;
    uint4 *a1 = (uint4 *)(w + 256);
    uint4 *b1 = a1 + np * 2;
    uint4 *a2 = b1 + np * 2;
    uint4 *b2 = a2 + mp * 2;
    cudaError_t e = cudaMemsetAsync(rmax, 0, 256, st);
    if (e != cudaSuccess) return e;
    PackSet p0, p1;
    p0.xyz = xyz1, p0.formA = a1, p0.formB = b1, p0.cnt = n < 128 ? n : 128, p0.cnt_scan = p0.cnt, p0.cnt_pad = (int)np;
    p0.blocks = (unsigned)((np + 255) / 256);
    p1.xyz = xyz2, p1.formA = a2, p1.formB = b2, p1.cnt = m, p1.cnt_scan = m, p1.cnt_pad = (int)mp;
    p1.blocks = (unsigned)((mp + 255) / 256);
    pack_kernel<<<p0.blocks + p1.blocks, 256, 0, st>>>(1, p0, p1, rmax);
    const int nc_pad = (int)((m + 127) / 128 * 128);
    dump_kernel<<<1, 128, 0, st>>>(a1, b2, nc_pad, dt);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return cudaMemcpyAsync(rmax_out, rmax, 4, cudaMemcpyDeviceToDevice, st);
}

}  // namespace chamfer_tc
}  // namespace geom

extern "C" int geom_debug_tc_dump(int n, const float *xyz1, int m, const float *xyz2, float *dt, float *rmax, void *workspace,
                                  size_t workspace_bytes, geom_stream_t stream)
{
    if (n <= 0 || m <= 0 || !xyz1 || !xyz2 || !dt || !rmax) return GEOM_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < geom::chamfer_tc::workspace_bytes(1, 128, m) || ((uintptr_t)workspace & 15u)) return GEOM_ERR_WORKSPACE;
    return geom::status_from_cuda(geom::chamfer_tc::debug_dump(n, xyz1, m, xyz2, dt, rmax, workspace, (cudaStream_t)stream));
}

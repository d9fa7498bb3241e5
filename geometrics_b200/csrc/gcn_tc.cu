// gcn_tc.cu -- tensor-core feature GEMM for the 0N-GCN layer: fp32 accuracy from three TF32 tcgen05
// MMAs per product term (3xTF32 split), accumulators in TMEM.
//
//   S = X (M,K) . W (K,N) with  x = xh + xl,  w = wh + wl  (xh = tf32(x), xl = tf32(x - xh)):
//   S ~= xh.wh + xh.wl + xl.wh   accumulated in fp32 by the tensor core; the dropped xl.wl term and the
//   split residuals are ~2^-21 relative per product, i.e. the same order as fp32 accumulation error,
//   which keeps the layer inside the 1e-5 parity budget that rules out a plain TF32 GEMM.
//
// One CTA computes a 128 x Npad output tile (Npad = N rounded up to 16, <= 256: one M=128 accumulator of
// Npad TMEM columns), two CTAs per SM.  K is consumed in stages of 16 through TWO shared-memory buffers, so
// the tensor core works on stage s while stage s+1 is being staged:
//   * all threads load the X tile of the NEXT stage into registers (64-byte row segments) before they wait
//     for anything, split it into hi/lo TF32 and store both in the canonical no-swizzle K-major layout
//     ([16-byte k-chunk][row][16 B]) as soon as the MMAs that last read that buffer have completed;
//   * the W tile comes pre-split and pre-arranged in that layout (wsplit_kernel, once per call: W is
//     small); warp 0 fetches its eight 4 KB planes with TMA bulk copies (mbarrier complete_tx);
//   * warp 0 (one elected lane, warp-uniform operands) issues 2 k-substeps x 3 tcgen05.mma (kind::tf32, K = 8) and
//     commits to the buffer's mbarrier.
// Epilogue: tcgen05.ld -> shared-memory transpose -> coalesced row writes with the 0N-GCN routing
// (columns < ksplit to the smoothing scratch T, the rest straight to Y with bias / ReLU).
#include "common.cuh"

namespace geom {
namespace gcn_tc {

constexpr int BK = 16;              // fp32 k-elements per stage
constexpr int kChunks = BK / 4;     // 16-byte chunks per row per stage
constexpr int kThreads = 256;
constexpr int kMaxN = 256;
constexpr int kAPlane = 128 * 16 + 16;  // bytes between k-chunk planes of the A tile (+16: bank spread)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u64 make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    u64 d = 0;
    d |= (u64)((saddr & 0x3FFFFu) >> 4);
    d |= (u64)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (u64)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (u64)1 << 46;  // descriptor version 1 (sm_100); no swizzle
    return d;
}
__device__ __forceinline__ unsigned to_tf32(float v)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void split4(float4 v, uint4 &hi, uint4 &lo)
{
    hi.x = to_tf32(v.x), hi.y = to_tf32(v.y), hi.z = to_tf32(v.z), hi.w = to_tf32(v.w);
    lo.x = to_tf32(v.x - __uint_as_float(hi.x)), lo.y = to_tf32(v.y - __uint_as_float(hi.y));
    lo.z = to_tf32(v.z - __uint_as_float(hi.z)), lo.w = to_tf32(v.w - __uint_as_float(hi.w));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

// Called by a whole (converged) warp with warp-uniform operands: ONE elected lane issues.  Under `if (tid == 0)` ptxas wraps
// every tcgen05 / bulk-copy instruction in a per-lane R2UR.BROADCAST loop (~240 cycles per MMA + commit).
__device__ __forceinline__ void mma_tf32(unsigned tmem_d, u64 da, u64 db, unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long *bar)
{
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

// W (K,N) row-major -> hi / lo TF32 in the operand layout [stage][chunk][n (Npad)][4 elements]; zero padded.
__global__ void __launch_bounds__(256)
wsplit_kernel(int K, int N, int nstages, int Npad, const float *__restrict__ W, uint4 *__restrict__ Bhi, uint4 *__restrict__ Blo)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)nstages * kChunks * Npad;
    if (t >= total) return;
    const int n = (int)(t % Npad);
    const int kc = (int)(t / Npad);  // global chunk index: k = 4*kc .. 4*kc+3
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
        const int k = kc * 4;
        if (k + 0 < K) v.x = W[(size_t)(k + 0) * N + n];
        if (k + 1 < K) v.y = W[(size_t)(k + 1) * N + n];
        if (k + 2 < K) v.z = W[(size_t)(k + 2) * N + n];
        if (k + 3 < K) v.w = W[(size_t)(k + 3) * N + n];
    }
    uint4 hi, lo;
    split4(v, hi, lo);
    Bhi[t] = hi;
    Blo[t] = lo;
}

struct Args {
    const float *X;
    const uint4 *Bhi, *Blo;
    const float *bias;
    float *Y, *T;
    int M, K, N, Npad, ksplit, ldt, relu, nstages;
};

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar` (expect_tx posted by the caller)
// (whole warp, one elected lane copies)
__device__ __forceinline__ void tma_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(kThreads, 2) gemm_kernel(const Args g)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int bplane = g.Npad * 16 + 16;  // bytes between k-chunk planes of the B tile
    const int abytes = 2 * kChunks * kAPlane, bbytes = 2 * kChunks * bplane;
    // buffer i: [A hi planes | A lo planes | B hi planes | B lo planes]
    auto buf_of = [&](int s) { return smem + (size_t)(s & 1) * (size_t)(abytes + bbytes); };
    __shared__ __align__(8) unsigned long long mma_done[2], b_full[2];
    __shared__ unsigned tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128;
    unsigned ncols = 32;  // TMEM allocation: power of two >= Npad
    while (ncols < (unsigned)g.Npad) ncols <<= 1;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    auto load_b = [&](int s) {  // warp 0, one elected lane: eight plane copies of stage s into buffer s & 1
        unsigned char *sb = buf_of(s) + abytes;
        const unsigned plane = (unsigned)g.Npad * 16u;
        asm volatile(
            "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
            "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(&b_full[s & 1])), "r"(2u * kChunks * plane)
            : "memory");
        const uint4 *ghi = g.Bhi + (size_t)s * kChunks * g.Npad;
        const uint4 *glo = g.Blo + (size_t)s * kChunks * g.Npad;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            tma_load(sb + c * bplane, ghi + (size_t)c * g.Npad, plane, &b_full[s & 1]);
            tma_load(sb + (kChunks + c) * bplane, glo + (size_t)c * g.Npad, plane, &b_full[s & 1]);
        }
    };
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mma_done[i])), "r"(1u));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&b_full[i])), "r"(1u));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();  // the barriers are initialised (by lane 0, which is also the lane elect.sync picks)
        load_b(0);
    }
    // D = f32 (1<<4), A = B = tf32 (2<<7, 2<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(g.Npad >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const bool xvec = (g.K % 4 == 0) && (((uintptr_t)g.X & 15u) == 0);
    constexpr int kLoads = 128 * kChunks / kThreads;  // float4 per thread per stage

    // X tile: 128 rows x 4 chunks; thread -> (row = q / 4, chunk = q % 4): 64-byte row segments
    auto load_x = [&](int s, float4 (&xa)[kLoads]) {
#pragma unroll
        for (int i = 0; i < kLoads; ++i) {
            const int q = tid + kThreads * i;
            const int row = q >> 2, c = q & 3;
            const int r = m0 + row, k = s * BK + c * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < g.M) {
                const float *src = g.X + (size_t)r * g.K + k;
                if (xvec && k + 3 < g.K) {
                    v = *reinterpret_cast<const float4 *>(src);
                } else {
                    if (k + 0 < g.K) v.x = src[0];
                    if (k + 1 < g.K) v.y = src[1];
                    if (k + 2 < g.K) v.z = src[2];
                    if (k + 3 < g.K) v.w = src[3];
                }
            }
            xa[i] = v;
        }
    };
    float4 xa[kLoads];
    load_x(0, xa);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;

    for (int s = 0; s < g.nstages; ++s) {
        unsigned char *sAhi = buf_of(s), *sAlo = sAhi + kChunks * kAPlane;
        unsigned char *sBhi = sAhi + abytes, *sBlo = sBhi + kChunks * bplane;
        // buffer s & 1 was last read by the MMAs of stage s-2
        if (s >= 2) mbar_wait(&mma_done[s & 1], (unsigned)(((s >> 1) - 1) & 1));
#pragma unroll
        for (int i = 0; i < kLoads; ++i) {
            const int q = tid + kThreads * i;
            const int row = q >> 2, c = q & 3;
            uint4 hi, lo;
            split4(xa[i], hi, lo);
            *reinterpret_cast<uint4 *>(sAhi + c * kAPlane + row * 16) = hi;
            *reinterpret_cast<uint4 *>(sAlo + c * kAPlane + row * 16) = lo;
        }
        if (s + 1 < g.nstages) load_x(s + 1, xa);  // in flight while this stage's MMAs are issued and run
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 0) {
            const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);  // warp-uniform, visibly to ptxas
            mbar_wait(&b_full[s & 1], (unsigned)((s >> 1) & 1));
#pragma unroll
            for (int sub = 0; sub < BK / 8; ++sub) {  // one K = 8 MMA covers chunks 2*sub, 2*sub+1
                const u64 ah = make_desc(smem_u32(sAhi + 2 * sub * kAPlane), kAPlane, 128);
                const u64 al = make_desc(smem_u32(sAlo + 2 * sub * kAPlane), kAPlane, 128);
                const u64 bh = make_desc(smem_u32(sBhi + 2 * sub * bplane), bplane, 128);
                const u64 bl = make_desc(smem_u32(sBlo + 2 * sub * bplane), bplane, 128);
                const unsigned first = (s == 0 && sub == 0) ? 0u : 1u;
                mma_tf32(tmem_u, al, bh, idesc, first);  // small terms first
                mma_tf32(tmem_u, ah, bl, idesc, 1u);
                mma_tf32(tmem_u, ah, bh, idesc, 1u);
            }
            mma_commit(&mma_done[s & 1]);
            if (s + 1 < g.nstages) {
                // the other buffer's W planes are read by the MMAs of stage s-1: refill it once they are done
                if (s >= 1) mbar_wait(&mma_done[(s + 1) & 1], (unsigned)(((s - 1) >> 1) & 1));
                load_b(s + 1);
            }
        }
    }
    // all MMAs complete (the last commit covers every earlier one)
    {
        const int last = g.nstages - 1;
        mbar_wait(&mma_done[last & 1], (unsigned)((last >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w reads TMEM lanes 32*(w&3).. (rows) and the column chunks c0 = 32*(2*j + (w>>2))
    float *tile = reinterpret_cast<float *>(smem) + warp * (32 * 33);  // staging area is free now
    __syncthreads();
    const int lane_base = (warp & 3) * 32;
    for (int c0 = (warp >> 2) * 32; c0 < g.Npad; c0 += 64) {
        unsigned v[32];
        const unsigned taddr = tmem + ((unsigned)lane_base << 16) + (unsigned)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                       "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                       "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                       "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory");
        // transpose through shared memory: lane l holds row (lane_base + l), 32 columns -> rows written coalesced
#pragma unroll
        for (int k = 0; k < 32; ++k) tile[lane * 33 + k] = __uint_as_float(v[k]);
        __syncwarp();
        const int c = c0 + lane;
        const float bv = (g.bias && c < g.N) ? g.bias[c] : 0.f;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
            const int r = m0 + lane_base + rr;
            if (r < g.M && c < g.N) {
                const float val = tile[rr * 33 + lane];
                if (c < g.ksplit) {
                    g.T[(size_t)r * g.ldt + c] = val;
                } else {
                    float o = val + bv;
                    if (g.relu) o = fmaxf(o, 0.f);
                    g.Y[(size_t)r * g.N + c] = o;
                }
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
}

static int stages_for(int K) { return (K + BK - 1) / BK; }
static int npad_for(int N) { return (N + 15) / 16 * 16; }

bool supported(int N) { return N >= 16 && npad_for(N) <= kMaxN; }

size_t workspace_bytes(int K, int N)
{
    return 2 * (size_t)stages_for(K) * kChunks * npad_for(N) * 16 + 256;
}

// S = X W with the 0N-GCN routing.  `wsplit` >= workspace_bytes(K, N), 16-byte aligned.
cudaError_t launch(int M, int K, int N, const float *X, const float *W, const float *bias, float *Y, float *T,
                   int ksplit, int ldt, int relu, void *wsplit, cudaStream_t st)
{
    Args g;
    g.nstages = stages_for(K);
    g.Npad = npad_for(N);
    uint4 *bhi = (uint4 *)wsplit;
    uint4 *blo = bhi + (size_t)g.nstages * kChunks * g.Npad;
    const long long total = (long long)g.nstages * kChunks * g.Npad;
    wsplit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(K, N, g.nstages, g.Npad, W, bhi, blo);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    g.X = X, g.Bhi = bhi, g.Blo = blo, g.bias = bias, g.Y = Y, g.T = T;
    g.M = M, g.K = K, g.N = N, g.ksplit = ksplit, g.ldt = ldt, g.relu = relu;
    const size_t smem = 2 * (2 * (size_t)kChunks * kAPlane + 2 * (size_t)kChunks * (g.Npad * 16 + 16));  // two buffers
    const size_t epi = 8 * 32 * 33 * sizeof(float);
    const size_t dyn = smem > epi ? smem : epi;
    static std::atomic<unsigned long long> smem_done{0};
    if ((e = ensure_dynamic_smem(gemm_kernel, 110 * 1024, smem_done)) != cudaSuccess) return e;
    gemm_kernel<<<(unsigned)((M + 127) / 128), kThreads, dyn, st>>>(g);
    return cudaGetLastError();
}

}  // namespace gcn_tc
}  // namespace geom

// tmem_probe.cu -- microbenchmark of TMEM read bandwidth (tcgen05.ld) as seen by the Chamfer scan kernel:
// W warps of one CTA per SM each read their lane quarter's columns in a loop; reports bytes/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_probe tools/tmem_probe.cu && ./tools/tmem_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ unsigned ld_cols(unsigned taddr);

template <>
__device__ __forceinline__ unsigned ld_cols<32>(unsigned taddr)
{
    unsigned v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) x ^= v[i];
    return x;
}

template <>
__device__ __forceinline__ unsigned ld_cols<16>(unsigned taddr)
{
    unsigned v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < 16; i += 8) x ^= v[i];
    return x;
}

// two x32 loads in flight before one wait
__device__ __forceinline__ unsigned ld_cols_2x32(unsigned taddr)
{
    unsigned v[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
          "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
          "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
          "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
          "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < 64; i += 8) x ^= v[i];
    return x;
}

// mode 0: x16, 1: x32, 2: x64
__global__ void __launch_bounds__(512) probe(int mode, int iters, long long *cycles, unsigned *sink)
{
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    const unsigned base = tmem + ((unsigned)((warp & 3) * 32) << 16);
    unsigned x = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const unsigned col = (unsigned)(((warp >> 2) * 64 + it * 64) & 447);
        if (mode == 0) {
            x ^= ld_cols<16>(base + col) ^ ld_cols<16>(base + col + 16) ^ ld_cols<16>(base + col + 32) ^ ld_cols<16>(base + col + 48);
        } else if (mode == 1) {
            x ^= ld_cols<32>(base + col) ^ ld_cols<32>(base + col + 32);
        } else {
            x ^= ld_cols_2x32(base + col);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (x == 0x12345678u) sink[0] = x;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main()
{
    long long *cyc;
    unsigned *sink;
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&sink, 4);
    const int iters = 2000;
    for (int mode = 0; mode < 3; ++mode) {
        for (int warps = 4; warps <= 16; warps *= 2) {
            probe<<<148, warps * 32, 0>>>(mode, iters, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e));
                return 1;
            }
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += (double)h[i];
            avg /= 148;
            const double bytes = (double)warps * iters * 64 * 32 * 4;
            printf("{\"ld\": \"%s\", \"warps\": %d, \"cycles_per_64col_load\": %.1f, \"bytes_per_clk_per_sm\": %.1f}\n",
                   mode == 0 ? "4 x (x16+wait)" : mode == 1 ? "2 x (x32+wait)" : "x64+wait", warps, avg / iters, bytes / avg);
        }
    }
    return 0;
}

// Microbenchmark: issue rates of packed fp32x2 vs scalar fp32 ops and FMNMX3 mixes on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_fp32x2.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){u64 r;asm volatile("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(r):"l"(a),"l"(b),"l"(c));return r;}
__device__ __forceinline__ u64 add2(u64 a, u64 b){u64 r;asm volatile("add.rn.f32x2 %0,%1,%2;":"=l"(r):"l"(a),"l"(b));return r;}
__device__ __forceinline__ u64 mul2(u64 a, u64 b){u64 r;asm volatile("mul.rn.f32x2 %0,%1,%2;":"=l"(r):"l"(a),"l"(b));return r;}
__device__ __forceinline__ float min3(float a,float b,float c){float r;asm volatile("min.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c));return r;}
__device__ __forceinline__ float fmas(float a,float b,float c){float r;asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c));return r;}

#define ITERS 2048
template<int MODE> __global__ void k(u64* out, u64 seed)
{
    u64 a[8]; float f[8];
    for(int i=0;i<8;i++){a[i]=seed+i+threadIdx.x; f[i]=(float)(threadIdx.x+i);}
    u64 b=seed*3+1; float fb=1.0001f;
    for(int it=0;it<ITERS;it++){
        if(MODE==0){ // 8 independent FFMA2
#pragma unroll
            for(int i=0;i<8;i++) a[i]=fma2(a[i],b,a[i]);
        } else if(MODE==1){ // 8 FADD2
#pragma unroll
            for(int i=0;i<8;i++) a[i]=add2(a[i],b);
        } else if(MODE==2){ // 8 scalar FFMA x2 (16 scalar)
#pragma unroll
            for(int i=0;i<8;i++){ f[i]=fmas(f[i],fb,f[i]); }
#pragma unroll
            for(int i=0;i<8;i++){ f[i]=fmas(f[i],fb,f[i]); }
        } else if(MODE==3){ // chamfer mix: 3 add2,1 mul2,2 fma2 (x2 chains) + 2 min3
#pragma unroll
            for(int i=0;i<8;i+=4){
                u64 dx=add2(a[i],b), dy=add2(a[i+1],b), dz=add2(a[i+2],b);
                u64 d=fma2(dz,dz,fma2(dx,dx,mul2(dy,dy)));
                float lo=__uint_as_float((unsigned)d), hi=__uint_as_float((unsigned)(d>>32));
                f[i]=min3(f[i],lo,hi); 
                a[i+3]^=d;
            }
        } else if(MODE==4){ // 8 FFMA2 + 8 FMNMX3 independent
#pragma unroll
            for(int i=0;i<8;i++){ a[i]=fma2(a[i],b,a[i]); f[i]=min3(f[i],fb,f[(i+1)&7]); }
        } else if(MODE==5){ // 8 FMNMX3 only
#pragma unroll
            for(int i=0;i<8;i++){ f[i]=min3(f[i],fb,f[(i+1)&7]); }
        } else if(MODE==6){ // 8 FFMA2 + 4 FMNMX3
#pragma unroll
            for(int i=0;i<8;i++){ a[i]=fma2(a[i],b,a[i]); if(i&1) f[i]=min3(f[i],fb,f[(i+1)&7]); }
        } else if(MODE==7){ // 4 FFMA2 + 4 FADD2 (2-operand vs 3-operand mix)
#pragma unroll
            for(int i=0;i<8;i++){ if(i&1) a[i]=fma2(a[i],b,a[i]); else a[i]=add2(a[i],b); }
        }
    }
    u64 s=0; for(int i=0;i<8;i++) s+=a[i]+(u64)__float_as_uint(f[i]);
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name,int instr_per_iter,int warps_per_sm)
{
    int threads=128, blocks=148*(warps_per_sm/4);
    u64* out; cudaMalloc(&out,sizeof(u64)*threads*blocks);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks,threads>>>(out,1); cudaDeviceSynchronize();
    cudaEventRecord(e0); for(int r=0;r<5;r++) k<MODE><<<blocks,threads>>>(out,r); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=5;
    int clk; cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,0);
    double cyc=ms*1e-3*clk*1e3; // cycles at max clock
    double winstr=(double)ITERS*instr_per_iter*warps_per_sm; // warp-instr per SM
    printf("%-34s warps/SM=%2d  %.3f ms  warp-instr/clk/SM=%.3f (per SMSP %.3f)\n",name,warps_per_sm,ms,winstr/cyc,winstr/cyc/4);
    cudaFree(out);
}
int main(){
    for(int w: {4,8,16,32}){
        run<0>("8xFFMA2",8,w); run<1>("8xFADD2",8,w); run<2>("16xFFMA scalar",16,w);
        run<3>("chamfer mix 12 FP2 + 2 min3",14,w); run<4>("8 FFMA2 + 8 FMNMX3",16,w); run<5>("8 FMNMX3",8,w);
        run<6>("8 FFMA2 + 4 FMNMX3",12,w); run<7>("4 FFMA2 + 4 FADD2",8,w);
    }
    return 0;
}

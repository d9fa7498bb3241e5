// Microbenchmark: issue rates of the minimum-reduction candidates for the tensor-core scan on sm_100a:
// FMNMX3 (min.f32 3-input), FMNMX (2-input), VIMNMX3 / VIMNMX (integer, DPX), HMNMX2 (packed bf16/f16), and mixes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_min tools/ubench_min.cu && tools/ubench_min
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
__device__ __forceinline__ float fmin3(float a,float b,float c){float r;asm volatile("min.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c));return r;}
__device__ __forceinline__ float fmin2(float a,float b){float r;asm volatile("min.f32 %0,%1,%2;":"=f"(r):"f"(a),"f"(b));return r;}
__device__ __forceinline__ int imin3(int a,int b,int c){return __vimin3_s32(a,b,c);}
__device__ __forceinline__ int imin2(int a,int b){int r;asm volatile("min.s32 %0,%1,%2;":"=r"(r):"r"(a),"r"(b));return r;}
__device__ __forceinline__ unsigned umin3(unsigned a,unsigned b,unsigned c){return __vimin3_u32(a,b,c);}
__device__ __forceinline__ unsigned hmin2(unsigned a,unsigned b){unsigned r;asm volatile("min.f16x2 %0,%1,%2;":"=r"(r):"r"(a),"r"(b));return r;}
__device__ __forceinline__ unsigned bmin2(unsigned a,unsigned b){unsigned r;asm volatile("min.bf16x2 %0,%1,%2;":"=r"(r):"r"(a),"r"(b));return r;}
__device__ __forceinline__ unsigned hmin3(unsigned a,unsigned b,unsigned c){return hmin2(hmin2(a,b),c);}
__device__ __forceinline__ unsigned cvtbf2(float a,float b){unsigned r;asm volatile("cvt.rn.bf16x2.f32 %0,%1,%2;":"=r"(r):"f"(a),"f"(b));return r;}
__device__ __forceinline__ float ffma(float a,float b,float c){float r;asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c));return r;}
__device__ __forceinline__ unsigned setp_sel(float a,float t,unsigned m,unsigned bit){unsigned r;asm volatile("{.reg .pred p; setp.le.f32 p,%1,%2; @p or.b32 %0,%3,%4; @!p mov.b32 %0,%3;}":"=r"(r):"f"(a),"f"(t),"r"(m),"r"(bit));return r;}
#define ITERS 2048
template<int MODE> __global__ void k(unsigned* out, unsigned seed)
{
    float f[8]; int v[8]; unsigned u[8];
    for(int i=0;i<8;i++){f[i]=(float)(threadIdx.x+i+seed); v[i]=threadIdx.x*7+i+seed; u[i]=threadIdx.x*13+i+seed;}
    float fb=1.0001f+seed; int ib=seed+12345; unsigned ub=seed+999;
    for(int it=0;it<ITERS;it++){
#pragma unroll
        for(int i=0;i<8;i++){
            if(MODE==0) f[i]=fmin3(f[i],fb,f[(i+1)&7]);
            else if(MODE==1) f[i]=fmin2(f[i],f[(i+1)&7]);
            else if(MODE==2) v[i]=imin3(v[i],ib,v[(i+1)&7]);
            else if(MODE==3) v[i]=imin2(v[i],v[(i+1)&7]);
            else if(MODE==4) u[i]=umin3(u[i],ub,u[(i+1)&7]);
            else if(MODE==5) u[i]=hmin2(u[i],u[(i+1)&7]);
            else if(MODE==6) u[i]=bmin2(u[i],u[(i+1)&7]);
            else if(MODE==7) u[i]=hmin3(u[i],ub,u[(i+1)&7]);
            else if(MODE==8) u[i]=cvtbf2(f[i],f[(i+1)&7])^u[i];
            else if(MODE==9){ f[i]=fmin3(f[i],fb,f[(i+1)&7]); v[i]=imin3(v[i],ib,v[(i+1)&7]); }      // do the two overlap?
            else if(MODE==10){ f[i]=fmin3(f[i],fb,f[(i+1)&7]); f[(i+3)&7]=ffma(f[(i+3)&7],fb,fb); }   // ALU + FMA pipe
            else if(MODE==11){ u[i]=setp_sel(f[i],fb,u[i],1u<<i); }
        }
    }
    unsigned s=0; for(int i=0;i<8;i++) s+=__float_as_uint(f[i])+v[i]+u[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name,int instr_per_iter,int warps_per_sm)
{
    int threads=128, blocks=148*(warps_per_sm/4);
    unsigned* out; cudaMalloc(&out,sizeof(unsigned)*threads*blocks);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks,threads>>>(out,1); cudaDeviceSynchronize();
    cudaEventRecord(e0); for(int r=0;r<5;r++) k<MODE><<<blocks,threads>>>(out,r); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=5;
    int clk; cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,0);
    double cyc=ms*1e-3*clk*1e3;
    double winstr=(double)ITERS*instr_per_iter*warps_per_sm;
    printf("%-40s warps/SM=%2d  %.3f ms  warp-instr/clk/SMSP=%.3f\n",name,warps_per_sm,ms,winstr/cyc/4);
    cudaFree(out);
}
int main(){
    for(int w: {16,32}){
        run<0>("FMNMX3 (min.f32 x3)",8,w); run<1>("FMNMX (min.f32 x2)",8,w); run<2>("VIMNMX3 (min.s32 x3)",8,w); run<3>("VIMNMX (min.s32 x2)",8,w);
        run<4>("min.u32 x3",8,w); run<5>("HMNMX2 f16x2",8,w); run<6>("HMNMX2 bf16x2",8,w); run<7>("min.f16x2 x3",8,w); run<8>("cvt.rn.bf16x2.f32 (+xor)",16,w);
        run<9>("FMNMX3 + VIMNMX3 interleaved",16,w); run<10>("FMNMX3 + FFMA interleaved",16,w); run<11>("setp+predicated or (2-3 instr)",16,w);
    }
    return 0;
}

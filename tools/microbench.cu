// microbench.cu -- issue-rate probe for the FP32 instructions the exact kernels are built from
// (B200, sm_100a).  Prints warp-instructions per clock per SM for each op at 32 warps/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk_add(u64 a, u64 b){ u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 pk_mulz(u64 a, u64 b){ u64 r; const u64 z=0; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(z)); return r; }
__device__ __forceinline__ u64 pk_fma(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template<int OP> __global__ void k(float* out, long long* cyc, float seed, int iters){
  float a[8], b[8]; u64 p[8], q[8];
  #pragma unroll
  for(int i=0;i<8;i++){ a[i]=seed+i+threadIdx.x; b[i]=a[i]*0.25f; q[i]=((u64)__float_as_uint(a[i]*0.3f)<<32)|__float_as_uint(a[i]*0.7f); p[i]=((u64)__float_as_uint(a[i])<<32)|__float_as_uint(a[i]*0.5f); }
  float c = seed*0.999f; u64 pc = ((u64)__float_as_uint(c)<<32)|__float_as_uint(c);
  long long t0 = clock64();
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int u=0;u<8;u++){
      #pragma unroll
      for(int i=0;i<8;i++){
        if(OP==0) a[i]=__fadd_rn(a[i],c);
        if(OP==1) a[i]=__fmul_rn(a[i],c);
        if(OP==2) a[i]=__fmul_rn(a[i],0.99993f);
        if(OP==3) a[i]=__fmaf_rn(a[i],c,c);
        if(OP==4) p[i]=pk_add(p[i],pc);
        if(OP==5) p[i]=pk_mulz(p[i],pc);
        if(OP==6) p[i]=pk_fma(p[i],pc,pc);
        if(OP==7){ a[i]=__fmul_rn(a[i],c); a[i]=__fadd_rn(a[i],c);}          // 2 instr
        if(OP==8){ u64 t=pk_mulz(p[i],pc); p[i]=pk_add(p[(i+1)&7],t);}       // 2 instr
        if(OP==9){ a[i]=__fsub_rn(__fmul_rn(a[i],c), __fmul_rn(a[(i+1)&7],c)); } // 3 instr
        if(OP==10){ p[i]=pk_add(p[i],pc); a[i]=__fadd_rn(a[i],c); }          // 1 packed + 1 scalar
        if(OP==11){ p[i]=pk_add(p[i],pc); a[i]=__fadd_rn(a[i],c); b[i]=__fadd_rn(b[i],c); }  // 1 packed + 2 scalar
        if(OP==12){ p[i]=pk_mulz(p[i],pc); a[i]=__fmul_rn(a[i],c); b[i]=__fadd_rn(b[i],c); }
        if(OP==13){ p[i]=pk_add(p[i],pc); q[i]=pk_mulz(q[i],pc); a[i]=__fadd_rn(a[i],c); b[i]=__fmul_rn(b[i],c);}  // 2 packed + 2 scalar
      }
    }
  }
  long long t1 = clock64();
  float s=0; 
  #pragma unroll
  for(int i=0;i<8;i++){ s+=b[i]+__uint_as_float((unsigned)q[i])+__uint_as_float((unsigned)(q[i]>>32))+a[i]+__uint_as_float((unsigned)p[i])+__uint_as_float((unsigned)(p[i]>>32)); }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
template<int OP> void run(const char* name, int per){
  int sms=148, thr=1024, iters=2000; float* out; long long* cyc; cudaMalloc(&out, sms*thr*4); cudaMalloc(&cyc, sms*8);
  k<OP><<<sms,thr>>>(out,cyc,1.0001f,10); cudaDeviceSynchronize();
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<OP><<<sms,thr>>>(out,cyc,1.0001f,iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  long long h[148]; cudaMemcpy(h,cyc,sms*8,cudaMemcpyDeviceToHost); double avg=0; for(int i=0;i<sms;i++) avg+=h[i]; avg/=sms;
  double winstr = (double)(thr/32)*iters*64.0*per;   // warp-instr per SM
  printf("%-28s %7.3f warp-instr/clk/SM   (%.0f cyc, %.3f ms, eff clk %.0f MHz)\n", name, winstr/avg, avg, ms, avg/ms/1e3);
  cudaFree(out); cudaFree(cyc);
}
int main(){
  run<0>("FADD reg",1); run<1>("FMUL reg",1); run<2>("FMUL imm",1); run<3>("FFMA reg",1);
  run<4>("FADD2",1); run<5>("FFMA2 (mul, +0)",1); run<6>("FFMA2 full",1);
  run<10>("FADD2 + FADD",2); run<11>("FADD2 + 2 FADD",3); run<12>("FFMA2z + FMUL + FADD",3); run<13>("FADD2+FFMA2z+FADD+FMUL",4);
  run<7>("FMUL+FADD pair",2); run<8>("FFMA2z+FADD2 pair",2); run<9>("2FMUL+FSUB",3);
  return 0;
}

// microbenchmark: scalar FFMA vs packed fma.rn.f32x2 / add.f32x2 issue throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint64_t pk(float a, float b){ uint64_t r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(uint64_t v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c){ uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b){ uint64_t d; asm("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
template<int MODE> __global__ void k(float* out, int iters, float s){
  float a[16]; uint64_t p[8];
  for(int i=0;i<16;i++) a[i]=threadIdx.x*0.001f+i;
  for(int i=0;i<8;i++) p[i]=pk(a[2*i],a[2*i+1]);
  uint64_t ps=pk(s,s), pc=pk(0.5f,0.25f);
  long long t0=clock64();
  for(int it=0;it<iters;it++){
    if(MODE==0){
#pragma unroll
      for(int i=0;i<16;i++) a[i]=fmaf(a[i],s,0.5f);
    } else if(MODE==1){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=fma2(p[i],ps,pc);
    } else if(MODE==2){
#pragma unroll
      for(int i=0;i<16;i++) a[i]=a[i]+s;
    } else if(MODE==3){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=add2(p[i],ps);
    } else if(MODE==4){ // mixed: 8 FFMA2 + 8 independent ALU ops (IADD/LOP)
#pragma unroll
      for(int i=0;i<8;i++) p[i]=fma2(p[i],ps,pc);
      int* ai=(int*)a;
#pragma unroll
      for(int i=0;i<16;i++) ai[i]=(ai[i]^it)+i;
    } else if(MODE==5){ // 16 FFMA + 16 ALU
#pragma unroll
      for(int i=0;i<8;i++){ float x,y; upk(p[i],x,y); x=fmaf(x,s,0.5f); y=fmaf(y,s,0.25f); p[i]=pk(x,y);}
      int* ai=(int*)a;
#pragma unroll
      for(int i=0;i<16;i++) ai[i]=(ai[i]^it)+i;
    }
  }
  long long t1=clock64();
  float acc=0; for(int i=0;i<16;i++) acc+=a[i]; for(int i=0;i<8;i++){float x,y; upk(p[i],x,y); acc+=x+y;}
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
  if(threadIdx.x==0&&blockIdx.x==0) ((long long*)out)[100000]=t1-t0;
}
int main(){
  float* d; cudaMalloc(&d, 8<<20);
  const char* names[]={"FFMA x16","FFMA2 x8","FADD x16","FADD2 x8","FFMA2x8+ALUx16","FFMAx16+ALUx16"};
  for(int mode=0;mode<6;mode++){
    for(int warps: {4,8,16}){
      int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      auto launch=[&](){ switch(mode){case 0:k<0><<<148,warps*32>>>(d,iters,1.0001f);break;case 1:k<1><<<148,warps*32>>>(d,iters,1.0001f);break;case 2:k<2><<<148,warps*32>>>(d,iters,1.0001f);break;case 3:k<3><<<148,warps*32>>>(d,iters,1.0001f);break;case 4:k<4><<<148,warps*32>>>(d,iters,1.0001f);break;case 5:k<5><<<148,warps*32>>>(d,iters,1.0001f);break;} };
      launch(); cudaDeviceSynchronize();
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms,e0,e1);
      long long cyc; cudaMemcpy(&cyc,((long long*)d)+100000,8,cudaMemcpyDeviceToHost);
      double flop_inst = (mode==4||mode==5)?16.0:16.0; // scalar-equivalent fp ops per iter per thread
      double cyc_per_iter=(double)cyc/iters;
      printf("%-18s warps/SM=%2d  cycles/iter=%7.2f  fp-lane-ops/clk/SM=%7.1f  ms=%.3f  err=%s\n",names[mode],warps,cyc_per_iter, flop_inst*warps*32/cyc_per_iter, ms, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}

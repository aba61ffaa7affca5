// Micro-benchmark: FP64 DMMA.8x8x4 vs DFMA issue throughput on sm_100a.
// Used once to pick the FP64 GEMM inner loop and to record the FP64 ceilings
// (profiles/r01_fp64_peaks.md). Not part of the product path.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template<int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters){
  double c[NACC][2];
  double a = threadIdx.x*1e-9, b = 1.0 + threadIdx.x*1e-9;
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters){
  double c[NACC];
  double a = 1.0+threadIdx.x*1e-9, b = threadIdx.x*1e-9;
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  int dev=0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int sms=p.multiProcessorCount; printf("device %s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*1024));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=20000;
  for(int warps : {4,8,16,32}){
    for(int rep=0;rep<2;rep++){
      dmma_kernel<8><<<sms, warps*32>>>(out, iters);
      CK(cudaDeviceSynchronize());
    }
    cudaEventRecord(e0);
    dmma_kernel<8><<<sms, warps*32>>>(out, iters);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*256*8*(double)iters*warps*sms;
    printf("DMMA.8x8x4 warps/SM=%d  %.3f ms  %.2f TFLOP/s\n", warps, ms, flops/ms*1e-9);
  }
  for(int warps : {8,16,32}){
    for(int rep=0;rep<2;rep++){ dfma_kernel<16><<<sms, warps*32>>>(out, iters); CK(cudaDeviceSynchronize()); }
    cudaEventRecord(e0);
    dfma_kernel<16><<<sms, warps*32>>>(out, iters);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*32*16*(double)iters*warps*sms;
    printf("DFMA warps/SM=%d  %.3f ms  %.2f TFLOP/s\n", warps, ms, flops/ms*1e-9);
  }
  // long run to see sustained clocks (approx 2 s)
  cudaEventRecord(e0);
  for(int r=0;r<20;r++) dmma_kernel<8><<<sms, 512>>>(out, iters*4);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  printf("DMMA sustained 20 launches: %.1f ms  %.2f TFLOP/s\n", ms, 2.0*256*8*(double)iters*4*16*sms*20/ms*1e-9);
  return 0;
}

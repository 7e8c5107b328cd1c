#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b){ unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template<int MODE>
__global__ void k(float* out, int iters, float seed) {
  // 16 independent chains
  float a[16], b[16];
  for (int i=0;i<16;i++){ a[i]=seed+i+threadIdx.x; b[i]=seed*0.5f+i; }
  if (MODE==0) {
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int i=0;i<16;i++){ a[i] = a[i] + b[i]; b[i] = b[i] + a[i]; }   // 32 FADD
    }
  } else if (MODE==1) {
    unsigned long long x[8], y[8];
    for (int i=0;i<8;i++){ x[i]=pk(a[2*i],a[2*i+1]); y[i]=pk(b[2*i],b[2*i+1]); }
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int i=0;i<8;i++){ x[i] = add2(x[i], y[i]); y[i] = add2(y[i], x[i]); }   // 16 FADD2 = 32 flops-lanes
    }
    for (int i=0;i<8;i++){ upk(x[i],a[2*i],a[2*i+1]); upk(y[i],b[2*i],b[2*i+1]); }
  } else if (MODE==2) {
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int i=0;i<16;i++){ a[i] = fmaf(a[i], 1.0001f, b[i]); b[i] = fmaf(b[i], 0.9999f, a[i]); }
    }
  } else {
    unsigned long long x[8], y[8]; unsigned long long c1 = pk(1.0001f,1.0001f), c2 = pk(0.9999f,0.9999f);
    for (int i=0;i<8;i++){ x[i]=pk(a[2*i],a[2*i+1]); y[i]=pk(b[2*i],b[2*i+1]); }
    for (int it=0; it<iters; ++it) {
      #pragma unroll
      for (int i=0;i<8;i++){ x[i] = fma2(x[i], c1, y[i]); y[i] = fma2(y[i], c2, x[i]); }
    }
    for (int i=0;i<8;i++){ upk(x[i],a[2*i],a[2*i+1]); upk(y[i],b[2*i],b[2*i+1]); }
  }
  float s=0; for (int i=0;i<16;i++) s+=a[i]+b[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, int warps_per_sm_x4) {
  float* out; cudaMalloc(&out, 148*1024*4);
  int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int threads = warps_per_sm_x4*32;
  k<MODE><<<148,threads>>>(out,100,1.f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148,threads>>>(out,iters,1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double lane_ops = 148.0*threads*iters*32.0; // scalar-equivalent ops
  printf("%s threads/SM=%d: %.3f ms, %.2f Tops/s scalar-equivalent (x2 flops for fma)\n", name, threads, ms, lane_ops/ms/1e9);
  cudaFree(out);
}
int main(){ for (int w : {4,8,16,32}) { run<0>("FADD  ",w); run<1>("FADD2 ",w); run<2>("FFMA  ",w); run<3>("FFMA2 ",w);} return 0; }

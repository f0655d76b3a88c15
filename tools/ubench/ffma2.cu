// Microbenchmark: FP32 FMA issue rate on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2), at several
// degrees of instruction-level parallelism.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int ILP, int MODE>  // MODE 0: FFMA, 1: FFMA2 vector*vector, 2: FFMA2 vector*broadcast scalar
__global__ void __launch_bounds__(256) k(float* out, int iters, float w) {
  float acc[2 * ILP];
  u64 acc2[ILP];
  for (int i = 0; i < ILP; ++i) { acc[2 * i] = threadIdx.x + i; acc[2 * i + 1] = threadIdx.x - i; acc2[i] = pk(acc[2 * i], acc[2 * i + 1]); }
  const u64 wv = pk(w, w * 1.0001f);
  const u64 wb = pk(w, w);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) asm volatile("fma.rn.f32 %0,%1,%2,%0;" : "+f"(acc[i]) : "f"(acc[i]), "f"(w));
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc2[i] = fma2(acc2[i], wv, acc2[i]);
      } else {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc2[i] = fma2(acc2[i], wb, acc2[i]);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 2 * ILP; ++i) s += acc[i];
  for (int i = 0; i < ILP; ++i) s += __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32));
  if (s == 12345.678f) out[0] = s;
}

template <int ILP, int MODE>
void run(const char* name, int warps_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, 4);
  const int iters = 4096, threads = 256;
  const int blocks = sms * warps_per_sm * 32 / threads;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<ILP, MODE><<<blocks, threads>>>(out, 16, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<ILP, MODE><<<blocks, threads>>>(out, iters, 1.0f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double fmas = (double)blocks * threads * iters * 8 * 2 * ILP;  // scalar-equivalent FMAs
  printf("%-28s ILP=%d warps/SM=%2d  %.3f ms  %.1f TFLOP/s\n", name, ILP, warps_per_sm, ms, 2 * fmas / ms * 1e-9);
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    run<1, 0>("FFMA  (2 chains)", w); run<1, 1>("FFMA2 vec*vec (1 chain)", w); run<1, 2>("FFMA2 vec*bcast (1 chain)", w);
    run<2, 0>("FFMA  (4 chains)", w); run<2, 1>("FFMA2 vec*vec (2 chains)", w); run<2, 2>("FFMA2 vec*bcast (2 chains)", w);
    run<4, 0>("FFMA  (8 chains)", w); run<4, 1>("FFMA2 vec*vec (4 chains)", w); run<4, 2>("FFMA2 vec*bcast (4 chains)", w);
    run<8, 0>("FFMA  (16 chains)", w); run<8, 1>("FFMA2 vec*vec (8 chains)", w); run<8, 2>("FFMA2 vec*bcast (8 chains)", w);
  }
  return 0;
}

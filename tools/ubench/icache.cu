// Microbenchmark: issue rate of LONG straight-line code (no reuse inside a warp) on sm_100a, as a function of
// resident warps per SM, CTA size, and whether the warps of a CTA are kept in lockstep with barriers.
// Answers: how much does instruction fetch cost a 100-200 KB straight-line kernel, and does lockstep help?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NINST, int SYNC_EVERY>  // NINST straight-line FFMAs (16 B each); barrier every SYNC_EVERY (0 = never)
__global__ void k(float* out, float w, int reps) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int n = 0; n < NINST / 8; ++n) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0,%1,%2,%0;" : "+f"(a[i]) : "f"(a[i]), "f"(w));
      if (SYNC_EVERY > 0 && (n * 8) % SYNC_EVERY == SYNC_EVERY - 8) __syncthreads();
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678f) out[0] = s;
}

template <int NINST, int SYNC_EVERY>
void run(int nt, int ctas_per_sm, int reps) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = sms * ctas_per_sm;
  k<NINST, SYNC_EVERY><<<blocks, nt>>>(out, 1.0f, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<NINST, SYNC_EVERY><<<blocks, nt>>>(out, 1.0f, reps);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warp_inst = (double)blocks * (nt / 32) * NINST * reps;
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("code %4d KB  sync %5d  nt %4d  ctas/SM %2d  warps/SM %2d  reps %d  %.3f ms  IPC/SM %.2f (peak 4)\n", NINST * 16 / 1024,
         SYNC_EVERY, nt, ctas_per_sm, nt / 32 * ctas_per_sm, reps, ms, warp_inst / sms / cycles);
  cudaFree(out);
}

int main() {
  // small code (fits every I-cache level) as the reference point
  run<1024, 0>(64, 8, 64);
  // 192 KB of straight-line code, executed once per warp (reps = 1) or looped (reps = 4)
  for (int reps : {1, 4}) {
    run<12288, 0>(32, 4, reps); run<12288, 0>(32, 8, reps); run<12288, 0>(32, 16, reps); run<12288, 0>(32, 32, reps);
    run<12288, 0>(64, 8, reps); run<12288, 0>(128, 4, reps); run<12288, 0>(256, 2, reps); run<12288, 0>(512, 1, reps);
    run<12288, 0>(1024, 1, reps);
    run<12288, 512>(64, 8, reps); run<12288, 512>(128, 4, reps); run<12288, 512>(256, 2, reps); run<12288, 512>(512, 1, reps);
    run<12288, 512>(1024, 1, reps);
    run<12288, 2048>(512, 1, reps);
  }
  // 48 KB and 24 KB bodies looped
  run<3072, 0>(64, 8, 16); run<1536, 0>(64, 8, 32); run<2048, 0>(64, 8, 32); run<6144, 0>(64, 8, 8);
  return 0;
}

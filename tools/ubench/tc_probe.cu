// tcgen05 probes for the wide path (sm_100a): what the design of vm_tc2.cuh rests on.
//   1. numerics of kind::tf32 with A from shared memory (SS) and A from tensor memory (TS), M=128 N=256 K=32
//   2. dense TF32 tcgen05.mma rate per SM and per chip (the denominator of roofline.frac: "TF32 peak / 3" for 3xTF32)
//   3. tcgen05.ld (TMEM -> registers) bandwidth per SM
//   4. rounding of the accumulator add (what bounds how many K steps may be chained inside the tensor core)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I surjectors_b200/csrc -I include tools/ubench/tc_probe.cu -o tools/ubench/tc_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vm_tc.cuh"

using namespace vmtc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

constexpr int PN = 256;                 // N of the probes
constexpr int PA = 128 * 32, PB = PN * 32;  // floats of one K = 32 operand block
constexpr size_t PROBE_SMEM = (size_t)(PA + PB) * 4 + 1024 + 64;

struct PSm { float* a; float* b; uint64_t* bar; uint32_t* slot; };
__device__ __forceinline__ PSm pcarve(unsigned char* raw) {
  uintptr_t p = ((uintptr_t)raw + 1023) & ~(uintptr_t)1023;
  PSm s;
  s.a = (float*)p; s.b = s.a + PA; s.bar = (uint64_t*)(s.b + PB); s.slot = (uint32_t*)(s.bar + 2);
  return s;
}
__device__ __forceinline__ uint32_t palloc(const PSm& sm) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(sm.bar, 1); mbar_init(sm.bar + 1, 1); }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sm.slot;
}
__device__ __forceinline__ void pfree(uint32_t t) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "n"(512) : "memory");
}
__device__ __forceinline__ uint32_t idesc_of(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// ---- 1. numerics: D[128][256] = A[128][32] B[256][32]^T, A from smem (ts = 0) or TMEM (ts = 1)
__global__ void __launch_bounds__(128) k_numerics(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int ts) {
  extern __shared__ unsigned char raw[];
  const PSm sm = pcarve(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 32; i += 128) sm.a[tile_off(i / 32, i % 32)] = A[i];
  for (int i = tid; i < PN * 32; i += 128) sm.b[tile_off(i / 32, i % 32)] = B[i];
  const uint32_t t0 = palloc(sm);
  const uint32_t a_col = 256 + 64;  // TS: A block at columns [320, 352), D at [0, 256)
  if (ts) {
    float v[32];
    for (int k = 0; k < 32; ++k) v[k] = A[(warp * 32 + lane) * 32 + k];
    tmem_st32(t0 + ((uint32_t)(warp * 32) << 16) + a_col, v);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint64_t da = make_desc(smem_u32(sm.a)), db = make_desc(smem_u32(sm.b));
    for (int ks = 0; ks < 4; ++ks) {
      if (ts) mma_tf32_ts(t0, t0 + a_col + ks * 8, db + (uint64_t)(ks * 16), idesc_of(PN), ks > 0);
      else mma_tf32(t0, da + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), idesc_of(PN), ks > 0);
    }
    umma_commit(sm.bar);
  }
  if (tid < 32) mbar_wait(sm.bar, 0);
  __syncthreads();
  tc_fence_after();
  for (int c0 = 0; c0 < PN; c0 += 32) {
    float v[32];
    tmem_ld32(t0 + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int q = 0; q < 32; ++q) D[(warp * 32 + lane) * PN + c0 + q] = v[q];
  }
  pfree(t0);
}

// ---- 2. rate: `iters` K = 32 blocks (4 MMAs each) of M=128 x N=n into alternating accumulators
__global__ void __launch_bounds__(128) k_rate(int iters, int n, int ts, float* __restrict__ sink) {
  extern __shared__ unsigned char raw[];
  const PSm sm = pcarve(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < PA; i += 128) sm.a[i] = 1.0f + (float)(i & 7) * 0.125f;
  for (int i = tid; i < PB; i += 128) sm.b[i] = 0.5f - (float)(i & 3) * 0.25f;
  const uint32_t t0 = palloc(sm);
  const uint32_t a_col = 480;
  if (ts) {
    float v[32];
    for (int k = 0; k < 32; ++k) v[k] = 1.0f + 0.125f * (float)(k & 7);
    tmem_st32(t0 + ((uint32_t)(warp * 32) << 16) + a_col, v);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint64_t da = make_desc(smem_u32(sm.a)), db = make_desc(smem_u32(sm.b));
    const uint32_t id = idesc_of(n);
    const int nbuf = ts ? 1 : (512 / n);
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = t0 + (uint32_t)((it % nbuf) * n);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (ts) mma_tf32_ts(d, t0 + a_col + ks * 8, db + (uint64_t)(ks * 16), id, 1u);
        else mma_tf32(d, da + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), id, 1u);
      }
    }
    umma_commit(sm.bar);
  }
  if (tid < 32) mbar_wait(sm.bar, 0);
  __syncthreads();
  tc_fence_after();
  float v[32];
  tmem_ld32(t0 + ((uint32_t)(warp * 32) << 16), v);
  if (v[0] == 123.456f) sink[0] = v[1];
  (void)lane;
  pfree(t0);
}

// ---- 3. tcgen05.ld bandwidth: nw warps read columns [0, 256) `iters` times
__global__ void __launch_bounds__(256) k_ldtm(int iters, long long* __restrict__ cycles, float* __restrict__ sink) {
  extern __shared__ unsigned char raw[];
  const PSm sm = pcarve(raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t t0 = palloc(sm);
  float acc = 0.0f;
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 256; c += 32) {
      float v[32];
      tmem_ld32(t0 + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c + (warp >> 2) * 256), v);
#pragma unroll
      for (int q = 0; q < 32; ++q) acc += v[q];
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cycles[0] = c1 - c0;
  if (acc == 123.456f) sink[0] = acc;
  pfree(t0);
}

// ---- 4. accumulator rounding.  D[0][0] only: MMA 0 sets it to `big` (a = big, b = 1), then n_add MMAs add c = ca * cb.
// mode 1: ONE extra MMA whose K = 8 products are 7 x c (intra-instruction sum).
__global__ void __launch_bounds__(128) k_round(float big, float ca, float cb, int n_add, int mode, float* __restrict__ out) {
  extern __shared__ unsigned char raw[];
  const PSm sm = pcarve(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < PA; i += 128) sm.a[i] = 0.0f;
  for (int i = tid; i < PB; i += 128) sm.b[i] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    sm.a[tile_off(0, 0)] = big; sm.b[tile_off(0, 0)] = 1.0f;          // k slice 0
    if (mode == 0) { sm.a[tile_off(0, 8)] = ca; sm.b[tile_off(0, 8)] = cb; }  // k slice 1: one product
    else for (int k = 8; k < 15; ++k) { sm.a[tile_off(0, k)] = ca; sm.b[tile_off(0, k)] = cb; }  // seven products
  }
  const uint32_t t0 = palloc(sm);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint64_t da = make_desc(smem_u32(sm.a)), db = make_desc(smem_u32(sm.b));
    mma_tf32(t0, da, db, idesc_of(PN), 0u);
    for (int i = 0; i < n_add; ++i) mma_tf32(t0, da + 16, db + 16, idesc_of(PN), 1u);
    umma_commit(sm.bar);
  }
  if (tid < 32) mbar_wait(sm.bar, 0);
  __syncthreads();
  tc_fence_after();
  float v[32];
  tmem_ld32(t0 + ((uint32_t)(warp * 32) << 16), v);
  if (warp == 0 && lane == 0) out[0] = v[0];
  pfree(t0);
}

static float tf32r(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  CK(cudaFuncSetAttribute(k_numerics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROBE_SMEM));
  CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROBE_SMEM));
  CK(cudaFuncSetAttribute(k_ldtm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROBE_SMEM));
  CK(cudaFuncSetAttribute(k_round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROBE_SMEM));
  // 1. numerics
  {
    std::vector<float> A(128 * 32), B(PN * 32), D(128 * PN);
    srand(1);
    for (auto& v : A) v = tf32r((float)rand() / RAND_MAX - 0.5f);
    for (auto& v : B) v = tf32r((float)rand() / RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int ts = 0; ts < 2; ++ts) {
      CK(cudaMemset(dD, 0, D.size() * 4));
      k_numerics<<<1, 128, PROBE_SMEM>>>(dA, dB, dD, ts);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("numerics ts=%d: CUDA error %s\n", ts, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double worst = 0.0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < PN; ++n) {
          double r = 0.0;
          for (int k = 0; k < 32; ++k) r += (double)A[m * 32 + k] * B[n * 32 + k];
          worst = fmax(worst, fabs(r - D[m * PN + n]));
        }
      printf("numerics %s: max |D - ref| = %.3e (D[0][0] = %.6f, D[127][255] = %.6f)\n", ts ? "TS (A in TMEM)" : "SS", worst, D[0], D[127 * PN + 255]);
    }
  }
  float* sink;
  CK(cudaMalloc(&sink, 64));
  // 2. rate
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    struct Cfg { int n, ts, ctas_per_sm; };
    const Cfg cfgs[] = {{128, 0, 1}, {256, 0, 1}, {256, 1, 1}, {128, 0, 2}, {64, 0, 1}, {208, 0, 1}};
    for (const Cfg& c : cfgs) {
      const int grid = prop.multiProcessorCount * c.ctas_per_sm;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        k_rate<<<grid, 128, PROBE_SMEM>>>(iters, c.n, c.ts, sink);
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = (double)grid * iters * 4.0 * 128.0 * c.n * 8.0 * 2.0;
        if (rep == 1)
          printf("rate N=%d %s %d CTA/SM: %.3f ms, %.1f TFLOP/s dense TF32 (chip), %.0f FLOP/clk/SM at %d MHz nominal\n", c.n,
                 c.ts ? "TS" : "SS", c.ctas_per_sm, ms, fl / (ms * 1e-3) / 1e12,
                 fl / (ms * 1e-3) / prop.multiProcessorCount / (prop.clockRate * 1e3), prop.clockRate / 1000);
      }
    }
  }
  // 3. tcgen05.ld bandwidth
  {
    long long* cyc;
    CK(cudaMalloc(&cyc, 8));
    for (int nthr : {128, 256}) {
      const int iters = 2000;
      k_ldtm<<<1, nthr, PROBE_SMEM>>>(iters, cyc, sink);
      CK(cudaDeviceSynchronize());
      k_ldtm<<<prop.multiProcessorCount, nthr, PROBE_SMEM>>>(iters, cyc, sink);
      CK(cudaDeviceSynchronize());
      long long h;
      CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
      const double bytes = (double)(nthr / 32) * 32 * 256 * 4 * iters;
      printf("tcgen05.ld 32x32b.x32 + 32 FADD, %d warps: %.1f B/clk/SM (%lld cycles)\n", nthr / 32, bytes / (double)h, h);
    }
  }
  // 4. accumulator rounding
  {
    float* out;
    CK(cudaMalloc(&out, 4));
    const float ulp = ldexpf(1.0f, -11);  // ulp of 4096
    struct R { float big, ca, cb; int n, mode; const char* what; };
    const R rs[] = {
        {4096.0f, 3.0f, ldexpf(1.0f, -13), 16, 0, "4096 + 16 x 0.75 ulp (RN: +16 ulp, RZ: +0)"},
        {4096.0f, 7.0f, ldexpf(1.0f, -13), 16, 0, "4096 + 16 x 1.75 ulp (RN: +32 ulp, RZ: +16)"},
        {4096.0f, 1.0f, ldexpf(1.0f, -13), 16, 0, "4096 + 16 x 0.25 ulp (RN: +0, RZ: +0, wide acc: +4)"},
        {-4096.0f, 3.0f, ldexpf(1.0f, -13), 16, 0, "-4096 + 16 x 0.75 ulp (RN: +16 ulp, RZ toward zero: +16, floor: +0)"},
        {4096.0f, -3.0f, ldexpf(1.0f, -13), 16, 0, "4096 - 16 x 0.75 ulp (RN: -16 ulp, RZ: -16, trunc of addend: -0..)"},
        {4096.0f, 3.0f, ldexpf(1.0f, -13), 1, 1, "4096 + one MMA with 7 products of 0.75 ulp (exact 5.25 ulp)"},
        {4096.0f, 1.0f, ldexpf(1.0f, -13), 1, 1, "4096 + one MMA with 7 products of 0.25 ulp (exact 1.75 ulp)"},
    };
    for (const R& r : rs) {
      k_round<<<1, 128, PROBE_SMEM>>>(r.big, r.ca, r.cb, r.n, r.mode, out);
      CK(cudaDeviceSynchronize());
      float h;
      CK(cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost));
      printf("round: %-75s -> %.7f = big %+.2f ulp\n", r.what, h, (h - r.big) / ulp);
    }
  }
  printf("done\n");
  return 0;
}

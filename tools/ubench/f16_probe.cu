// tcgen05 kind::f16 probes for the fused conditioner + spline kernel (vm_fused.cuh), sm_100a:
//   1. numerics of kind::f16 (fp16 operands, fp32 accumulate) with A from shared memory (SS) and from tensor memory (TS)
//      in the K-major no-swizzle layout, N in {96, 112, 208, 256}: validates descriptor, instruction descriptor and the
//      packing of two fp16 per 32-bit TMEM column
//   2. FP32-grade products from fp16 pairs ("fp16x3": hi*hi + hi*lo + lo*hi with power-of-two row / tensor scales):
//      error and BIAS against float64 for K = 256, single accumulator (48 truncating adds) vs cross terms in their own
//      accumulator (16 truncating adds on the main one)
//   3. dense fp16 tcgen05.mma rate, SS and TS, per N
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I surjectors_b200/csrc -I include tools/ubench/f16_probe.cu -o tools/ubench/f16_probe
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vm_tc.cuh"

using namespace vmtc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// fp16 K-major no-swizzle operand chunk of 64 k: element (r, k) at byte (r/8)*1024 + (k/8)*128 + (r%8)*16 + (k%8)*2
__device__ __host__ __forceinline__ int h_off(int r, int k) { return (r >> 3) * 512 + (k >> 3) * 64 + (r & 7) * 8 + (k & 7); }  // in halfs
__device__ __forceinline__ uint32_t idesc_f16(int n) {  // D = f32 (1 << 4), A = B = f16 (format 0), K-major both
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int KC = 64;  // k per chunk
constexpr int MAXN = 256;
constexpr size_t SMEM = (size_t)(2 * 128 * KC + 2 * MAXN * KC) * 2 + 1024 + 64;

struct Sm { __half* a_hi; __half* a_lo; __half* b_hi; __half* b_lo; uint64_t* bar; uint32_t* slot; };
__device__ __forceinline__ Sm carve_h(unsigned char* raw) {
  uintptr_t p = ((uintptr_t)raw + 1023) & ~(uintptr_t)1023;
  Sm s;
  s.a_hi = (__half*)p; s.a_lo = s.a_hi + 128 * KC; s.b_hi = s.a_lo + 128 * KC; s.b_lo = s.b_hi + MAXN * KC;
  s.bar = (uint64_t*)(s.b_lo + MAXN * KC); s.slot = (uint32_t*)(s.bar + 2);
  return s;
}
__device__ __forceinline__ uint32_t alloc512(const Sm& sm) {
  if (threadIdx.x == 0) { mbar_init(sm.bar, 1); mbar_init(sm.bar + 1, 1); }
  __syncwarp();
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sm.slot;
}
__device__ __forceinline__ void free512(uint32_t t) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "n"(512) : "memory");
}

// mode bit 0: TS (A from TMEM);  bit 1: cross terms (hi*lo, lo*hi) into their own accumulator;  bit 2: hi*hi only (plain fp16)
// A[128][K] and B[N][K] arrive as float already scaled; split into fp16 hi / lo here.  D[128][N] = sum of the requested terms.
__global__ void __launch_bounds__(128) k_gemm(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                                              int mode) {
  extern __shared__ unsigned char raw[];
  const Sm sm = carve_h(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool ts = mode & 1, split = mode & 2, plain = mode & 4;
  const uint32_t t0 = alloc512(sm);
  // TMEM map: D main [0, 256), then TS: A hi [256, 256 + K/2), A lo [256 + K/2, 256 + K)  (K <= 256; with split and TS N <= 128)
  //           SS: D cross [256, 512)
  const uint32_t d_main = t0, d_cross = ts ? t0 + 128 : t0 + 256;
  const uint32_t a_hi_col = 256, a_lo_col = 256 + K / 2;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  if (ts) {  // my row of A -> TMEM, two fp16 per column (k even in the low half)
    const int m = warp * 32 + lane;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t vh[8], vl[8];
      for (int q = 0; q < 8; ++q) {
        const float x0 = A[m * K + k0 + 2 * q], x1 = A[m * K + k0 + 2 * q + 1];
        const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
        const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
        vh[q] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        vl[q] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      tmem_st8(t0 + lane_addr + a_hi_col + k0 / 2, vh);
      tmem_st8(t0 + lane_addr + a_lo_col + k0 / 2, vl);
    }
    tmem_st_wait();
  }
  const uint32_t id = idesc_f16(N);
  uint32_t parity = 0;
  for (int kc = 0; kc < K; kc += KC) {
    for (int i = tid; i < 128 * KC; i += 128) {
      const int r = i / KC, k = i % KC;
      const float x = A[r * K + kc + k];
      const __half h = __float2half_rn(x);
      sm.a_hi[h_off(r, k)] = h;
      sm.a_lo[h_off(r, k)] = __float2half_rn(x - __half2float(h));
    }
    for (int i = tid; i < N * KC; i += 128) {
      const int r = i / KC, k = i % KC;
      const float x = B[r * K + kc + k];
      const __half h = __float2half_rn(x);
      sm.b_hi[h_off(r, k)] = h;
      sm.b_lo[h_off(r, k)] = __float2half_rn(x - __half2float(h));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t da_hi = make_desc(smem_u32(sm.a_hi)), da_lo = make_desc(smem_u32(sm.a_lo));
      const uint64_t db_hi = make_desc(smem_u32(sm.b_hi)), db_lo = make_desc(smem_u32(sm.b_lo));
      for (int ks = 0; ks < KC / 16; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 16);  // 2 core matrices = 256 B >> 4
        const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
        const uint32_t ka = (uint32_t)((kc + ks * 16) / 2);
        const uint32_t dx = split ? d_cross : d_main;
        if (!plain) {
          if (ts) {
            mma_f16_ts(dx, t0 + a_lo_col + ka, db_hi + adv, id, first);
            mma_f16_ts(dx, t0 + a_hi_col + ka, db_lo + adv, id, 1u);
          } else {
            mma_f16_ss(dx, da_lo + adv, db_hi + adv, id, first);
            mma_f16_ss(dx, da_hi + adv, db_lo + adv, id, 1u);
          }
        }
        const uint32_t acc_main = (split || plain) ? first : 1u;
        if (ts) mma_f16_ts(d_main, t0 + a_hi_col + ka, db_hi + adv, id, acc_main);
        else mma_f16_ss(d_main, da_hi + adv, db_hi + adv, id, acc_main);
      }
      umma_commit(sm.bar);
    }
    if (tid < 32) mbar_wait(sm.bar, parity);
    parity ^= 1u;
    __syncthreads();
    tc_fence_after();
  }
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16], r2[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(d_main + lane_addr + c0));
    if (split)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                   : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]), "=r"(r2[8]),
                     "=r"(r2[9]), "=r"(r2[10]), "=r"(r2[11]), "=r"(r2[12]), "=r"(r2[13]), "=r"(r2[14]), "=r"(r2[15])
                   : "r"(d_cross + lane_addr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int q = 0; q < 16; ++q) {
      float v = __uint_as_float(r[q]);
      if (split) v += __uint_as_float(r2[q]);
      D[(warp * 32 + lane) * N + c0 + q] = v;
    }
  }
  free512(t0);
}

// rate: iters x (KC / 16) MMAs of M=128 x N x K=16 into alternating accumulators
__global__ void __launch_bounds__(128) k_rate(int iters, int n, int ts, float* __restrict__ sink) {
  extern __shared__ unsigned char raw[];
  const Sm sm = carve_h(raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * KC; i += 128) sm.a_hi[i] = __float2half_rn(1.0f + (float)(i & 7) * 0.125f);
  for (int i = tid; i < MAXN * KC; i += 128) sm.b_hi[i] = __float2half_rn(0.5f - (float)(i & 3) * 0.25f);
  const uint32_t t0 = alloc512(sm);
  const uint32_t a_col = 480;
  if (ts) {
    uint32_t v[8];
    for (int q = 0; q < 8; ++q) v[q] = 0x3c003c00u;
    for (int c = 0; c < 32; c += 8) tmem_st8(t0 + ((uint32_t)(warp * 32) << 16) + a_col + c, v);
    tmem_st_wait();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint64_t da = make_desc(smem_u32(sm.a_hi)), db = make_desc(smem_u32(sm.b_hi));
    const uint32_t id = idesc_f16(n);
    const int nbuf = ts ? (448 / n) : (512 / n);
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = t0 + (uint32_t)((it % nbuf) * n);
#pragma unroll
      for (int ks = 0; ks < KC / 16; ++ks) {
        if (ts) mma_f16_ts(d, t0 + a_col + ks * 8, db + (uint64_t)(ks * 16), id, 1u);
        else mma_f16_ss(d, da + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), id, 1u);
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
  free512(t0);
}

static double urand() { return (double)rand() / RAND_MAX; }
static double nrand() { return sqrt(-2.0 * log(urand() + 1e-12)) * cos(6.283185307179586 * urand()); }

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  float *dA, *dB, *dD, *sink;
  CK(cudaMalloc(&dA, 128 * 256 * 4)); CK(cudaMalloc(&dB, 256 * 256 * 4)); CK(cudaMalloc(&dD, 128 * 256 * 4)); CK(cudaMalloc(&sink, 64));
  // 1. plain fp16 numerics (operands exactly representable in fp16): layout / descriptor / TMEM packing check
  srand(1);
  for (int N : {96, 112, 208, 256}) {
    const int K = 64;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    for (auto& v : A) v = __half2float(__float2half_rn((float)(urand() - 0.5)));
    for (auto& v : B) v = __half2float(__float2half_rn((float)(urand() - 0.5)));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int ts = 0; ts < 2; ++ts) {
      CK(cudaMemset(dD, 0, 128 * 256 * 4));
      k_gemm<<<1, 128, SMEM>>>(dA, dB, dD, N, K, 4 | ts);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("numerics N=%d ts=%d: CUDA error %s\n", N, ts, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double worst = 0.0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double r = 0.0;
          for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
          worst = fmax(worst, fabs(r - D[m * N + n]));
        }
      printf("plain fp16 N=%d %s: max |D - ref| = %.3e (D[0][0] = %.6f, D[127][N-1] = %.6f)\n", N, ts ? "TS" : "SS", worst, D[0], D[127 * N + N - 1]);
    }
  }
  // 2. fp16x3 with power-of-two scales, K = 256: error and bias against float64
  {
    const int K = 256, N = 112;
    std::vector<float> A(128 * K), B(N * K), As(128 * K), Bs(N * K), D(128 * N);
    std::vector<int> ea(128);
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < K; ++k) {
        const double pre = nrand() * 1.5;
        A[m * K + k] = (float)(pre / (1.0 + exp(-pre)));  // swish of a normal pre-activation
      }
    for (auto& v : B) v = (float)(nrand() * 0.06);
    float bmax = 0.0f;
    for (auto& v : B) bmax = fmaxf(bmax, fabsf(v));
    int eb; frexpf(bmax, &eb);  // bmax = f * 2^eb, f in [0.5, 1)
    for (size_t i = 0; i < B.size(); ++i) Bs[i] = ldexpf(B[i], 15 - eb);
    for (int m = 0; m < 128; ++m) {
      float amax = 0.0f;
      for (int k = 0; k < K; ++k) amax = fmaxf(amax, fabsf(A[m * K + k]));
      frexpf(amax, &ea[m]);
      for (int k = 0; k < K; ++k) As[m * K + k] = ldexpf(A[m * K + k], 15 - ea[m]);
    }
    CK(cudaMemcpy(dA, As.data(), As.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bs.data(), Bs.size() * 4, cudaMemcpyHostToDevice));
    const char* names[] = {"SS one accumulator (48 adds)", "TS one accumulator", "SS cross terms apart (16 adds)", "TS cross terms apart", "SS hi*hi only (plain fp16)"};
    const int modes[] = {0, 1, 2, 3, 4};
    for (int v = 0; v < 5; ++v) {
      CK(cudaMemset(dD, 0, 128 * 256 * 4));
      k_gemm<<<1, 128, SMEM>>>(dA, dB, dD, N, K, modes[v]);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("fp16x3 %s: CUDA error %s\n", names[v], cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double bias = 0.0, rms = 0.0, worst = 0.0, bias32 = 0.0, rms32 = 0.0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double r = 0.0, nrm = 0.0;
          float f = 0.0f;
          for (int k = 0; k < K; ++k) {
            r += (double)A[m * K + k] * B[n * K + k];
            nrm += fabs((double)A[m * K + k] * B[n * K + k]);
            f = fmaf(A[m * K + k], B[n * K + k], f);
          }
          const double got = ldexp((double)D[m * N + n], (ea[m] - 15) + (eb - 15));
          const double sgn = r >= 0 ? 1.0 : -1.0;
          const double e1 = (got - r) / fabs(r), e32 = ((double)f - r) / fabs(r);
          bias += sgn * (got - r) / fabs(r); rms += e1 * e1; worst = fmax(worst, fabs(got - r) / nrm);
          bias32 += sgn * ((double)f - r) / fabs(r); rms32 += e32 * e32;
        }
      const double cnt = 128.0 * N;
      printf("fp16x3 K=256 %-32s: mean signed rel err (toward |D| growth) %+.3e, rms rel %.3e, max |err| / sum|terms| %.3e   [fp32 fma chain: %+.3e, %.3e]\n",
             names[v], bias / cnt, sqrt(rms / cnt), worst, bias32 / cnt, sqrt(rms32 / cnt));
    }
  }
  // 3. rate
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    struct Cfg { int n, ts; };
    const Cfg cfgs[] = {{256, 0}, {208, 0}, {128, 0}, {112, 0}, {256, 1}, {208, 1}, {128, 1}, {112, 1}, {96, 1}};
    for (const Cfg& c : cfgs) {
      const int grid = prop.multiProcessorCount;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        k_rate<<<grid, 128, SMEM>>>(iters, c.n, c.ts, sink);
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = (double)grid * iters * (KC / 16) * 128.0 * c.n * 16.0 * 2.0;
        if (rep == 1)
          printf("rate fp16 N=%d %s: %.3f ms, %.1f TFLOP/s dense (chip), %.0f FLOP/clk/SM at %d MHz nominal\n", c.n, c.ts ? "TS" : "SS", ms,
                 fl / (ms * 1e-3) / 1e12, fl / (ms * 1e-3) / prop.multiProcessorCount / (prop.clockRate * 1e3), prop.clockRate / 1000);
      }
    }
  }
  printf("done\n");
  return 0;
}

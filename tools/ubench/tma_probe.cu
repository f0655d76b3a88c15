// TMA tensor loads feeding tcgen05 kind::tf32 straight from float32 rows (sm_100a): validates, with exact small-integer
// data, the two shared-memory layouts vm_gemm3.cuh's TMA loaders rely on
//   1. K-major operands (rows [r][k], k contiguous): box {16 k, rows} with CU_TENSOR_MAP_SWIZZLE_64B -> UMMA descriptor
//      layout SWIZZLE_64B, SBO = 512 B (8 rows x 64 B), k step of 8 = +32 B on the start address
//   2. an MN-major A operand (rows [k][m], m contiguous -- the activations of LINEAR as they live in the workspace):
//      four boxes {32 m, 16 k} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32-byte chunks swizzled within 128 bytes: chunk
//      (m % 32) / 8 of k row r lands at chunk position ((m % 32) / 8) ^ (r % 4)) -> descriptor layout code 1
//      (SWIZZLE_128B_BASE32B), LBO = 2048 B (next 32 m), SBO = 512 B (next 4 k), instruction-descriptor bit 15 (A is
//      MN-major), k step of 8 = +1024 B.  The plain 128-byte swizzle (layout code 2) does NOT work for a 32-bit MN-major
//      operand: those variants are run too and must mismatch (they document what was tried).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I surjectors_b200/csrc -I include tools/ubench/tma_probe.cu -o tools/ubench/tma_probe
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vm_tc.cuh"

using namespace vmtc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, const float* base, uint64_t inner, uint64_t outer, uint64_t pitch_floats, uint32_t box_inner,
                            uint32_t box_outer, CUtensorMapSwizzle sw) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer}, strides[1] = {pitch_floats * 4};
  cuuint32_t box[2] = {box_inner, box_outer}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}

// MODE 0: A K-major (rows [m][k]); MODE 1: A MN-major (rows [k][m]).  B always K-major, 256 rows.  K = 16 at k offset k0.
template <int MODE>
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, int k0, int m0, int n_rows_b, float* out,
      uint32_t v_lbo, uint32_t v_sbo, uint32_t v_layout, uint32_t v_idesc, uint32_t v_kadv, float* dump) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sa = base;            // 8 KB
  unsigned char* sb = base + 8192;     // 16 KB
  uint64_t* bar = (uint64_t*)(base + 8192 + 16384);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    mbar_expect_tx(bar, 8192 + 16384);
    if (MODE == 0) {
      tma_2d(sa, &ma, k0, m0, bar);
    } else {
      for (int c = 0; c < 4; ++c) tma_2d(sa + 2048 * c, &ma, m0 + 32 * c, k0, bar);
    }
    tma_2d(sb, &mb, k0, 0, bar);
    mbar_wait(bar, 0);
    tc_fence_after();
    if (dump) for (int i = 0; i < 2048; ++i) dump[i] = ((const float*)sa)[i];
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (MODE == 1 ? v_idesc : 0u) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t da = MODE == 0 ? desc_sw(smem_u32(sa) + 32 * ks, 16, 512, 4) : desc_sw(smem_u32(sa) + v_kadv * ks, v_lbo, v_sbo, v_layout);
      const uint64_t db = desc_sw(smem_u32(sb) + 32 * ks, 16, 512, 4);
      mma_tf32(tmem, da, db, idesc, ks);
    }
    umma_commit(bar + 1);
  }
  __syncwarp();
  mbar_wait(bar + 1, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 256; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) out[(size_t)tid * 256 + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int M = 256, N = 200, K = 64;  // A: 256 rows (tile at m0 = 128), B: 200 valid rows (rows 200..255 of the tile out of bounds -> zeros)
  std::vector<float> A((size_t)M * K), At((size_t)K * M), B((size_t)N * K);
  srand(7);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) { A[(size_t)m * K + k] = (float)(rand() % 17 - 8); At[(size_t)k * M + m] = A[(size_t)m * K + k]; }
  for (auto& v : B) v = (float)(rand() % 17 - 8);
  float *dA, *dAt, *dB, *dO;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dAt, At.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, 128 * 256 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAt, At.data(), At.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const int k0 = 32, m0 = 128;
  const size_t smem = 1024 + 8192 + 16384 + 64;
  CK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap mb = make_map(enc, dB, K, N, K, 16, 256, CU_TENSOR_MAP_SWIZZLE_64B);
  int bad_total = 0;
  float* dD;
  CK(cudaMalloc(&dD, 8192));
  struct Var { int mode; uint32_t lbo, sbo, layout, idesc, kadv; const char* name; int atom32; };
  const Var vars[] = {{0, 0, 0, 0, 0, 0, "K-major, 64B swizzle (layout 4, SBO 512)", 0},
                      {1, 2048, 512, 1, 1u << 15, 1024, "MN-major, 128B swizzle on 32B atoms (layout 1, LBO 2048, SBO 512)", 1},
                      {1, 2048, 1024, 1, 1u << 15, 1024, "[expected to mismatch] same, SBO 1024", 1},
                      {1, 2048, 1024, 2, 1u << 15, 1024, "[expected to mismatch] plain 128B swizzle image, layout 2", 0},
                      {1, 2048, 1024, 1, 1u << 15, 1024, "[expected to mismatch] plain 128B swizzle image, layout 1", 0}};
  for (const Var& v : vars) {
    const int mode = v.mode;
    CUtensorMap ma = mode == 0 ? make_map(enc, dA, K, M, K, 16, 128, CU_TENSOR_MAP_SWIZZLE_64B) : make_map(enc, dAt, M, K, M, 32, 16, v.atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
    CK(cudaMemset(dO, 0xff, 128 * 256 * 4));
    if (mode == 0) probe<0><<<1, 128, smem>>>(ma, mb, k0, m0, N, dO, v.lbo, v.sbo, v.layout, v.idesc, v.kadv, dD);
    else probe<1><<<1, 128, smem>>>(ma, mb, k0, m0, N, dO, v.lbo, v.sbo, v.layout, v.idesc, v.kadv, dD);
    CK(cudaDeviceSynchronize());
    std::vector<float> O(128 * 256), Dm(2048);
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Dm.data(), dD, 8192, cudaMemcpyDeviceToHost));
    if (mode == 1) {  // where TMA put the tile: element (m, k) of 32-m chunk c at c*2048 + k*128 + ((m%32)/4 ^ k%8)*16 + (m%4)*4 bytes
      int sb = 0;
      for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 16; ++k) {
          const int off = v.atom32 ? (m / 32) * 2048 + k * 128 + ((((m % 32) / 8) ^ (k % 4)) * 32) + (m % 8) * 4
                                   : (m / 32) * 2048 + k * 128 + ((((m % 32) / 4) ^ (k % 8)) * 16) + (m % 4) * 4;
          if (Dm[off / 4] != A[(size_t)(m0 + m) * K + k0 + k]) ++sb;
        }
      printf("  [next line's variant] shared-memory image of A: %d of 2048 elements not where the assumed swizzle puts them\n", sb);
    }
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 256; ++n) {
        float ref = 0.0f;
        if (n < N)
          for (int k = 0; k < 16; ++k) ref += A[(size_t)(m0 + m) * K + k0 + k] * B[(size_t)n * K + k0 + k];
        if (O[(size_t)m * 256 + n] != ref && bad++ < 3) printf("  (%d, %d): got %g want %g\n", m, n, O[(size_t)m * 256 + n], ref);
      }
    printf("%s: %d mismatches of %d\n", v.name, bad, 128 * 256);
    if (&v == &vars[0] || &v == &vars[1]) bad_total += bad;
  }
  printf(bad_total == 0 ? "TMA PROBE OK\n" : "TMA PROBE FAILED\n");
  return bad_total != 0;
}

// tcgen05 (5th-gen tensor core) matrix products of the per-op executor, FP32-grade via 3xTF32.
//
//   D[m][n] = sum_k A[m][k] * B[n][k]        M tile = 128 (TMEM lanes), N tile <= 256 (TMEM columns)
//
// issued as tcgen05.mma.cta_group::1.kind::tf32 by ONE thread, operands in shared memory in the
// canonical K-major no-swizzle ("interleave") UMMA layout, accumulator in TMEM, read back with
// tcgen05.ld for the epilogue.  Every fp32 operand x is split into hi = tf32(x) and
// lo = x - hi (truncated to tf32 by the tensor core itself), and each K step issues lo*hi + hi*lo + hi*hi, which keeps the
// product error near 2^-21 relative -- the 1e-5 / 1e-4 parity bounds rule out plain TF32 (2^-11).
// The tensor core's accumulator add truncates, which shows up as a systematic bias once hundreds
// of K steps are chained (measured: 1.7e-5 on the 16-layer MADE config); so only ONE K block (32) is
// accumulated inside the tensor core, into one of two TMEM buffers, and the blocks are summed in
// FP32 registers with round-to-nearest adds while the next block's MMAs run.
//
// Two instantiations (vm_global.cu):
//   LINEAR     out[j][s] = act(b[j] + sum_i W[i][j] in[i][s])   M = samples, N = outputs, K = inputs
//              A = activations rows [i][s] (staged with a transpose), B = the [j][i] copy of W
//   LINEAR_DW  grad[i][j] += sum_s in[i][s] g[j][s]             M = inputs (+ ones row), N = outputs,
//              K = samples: both operands are K-major as stored (the [row][sample] layout)
//
// Shared-memory operand tile, K block of 32 floats: element (r, k) lives at byte
//   (r / 8) * 1024 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4
// i.e. 8x(16 B) core matrices, K-adjacent core matrices 128 B apart (LBO), 8-row groups 1024 B apart
// (SBO); one MMA (K = 8 tf32) consumes two K-adjacent core matrices, so stepping K by 8 adds 256 B
// to the descriptor start address.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vm.cuh"
#include "vm_plan.h"

namespace vmtc {

constexpr int TM = 128;      // M tile (TMEM lanes)
constexpr int TN = 128;      // max N tile; two TMEM accumulator buffers of TN columns
constexpr int BK = 32;       // K block staged per round
constexpr int A_TILE = TM * BK;  // floats
constexpr int B_TILE = TN * BK;
constexpr size_t SMEM_BYTES = (size_t)(2 * A_TILE + 2 * B_TILE) * 4 + 1024 + 64;
constexpr int TMEM_COLS = 2 * TN;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// hi = x rounded to TF32 (10 mantissa bits, round-half-away: add half an ulp, clear the 13 low bits); the PTX
// cvt.rna.tf32.f32 compiles to the same two integer ops plus an inf / NaN guard (FSETP + SEL) that adds nothing here
// (the mask leaves inf as inf, NaN as NaN).  lo = x - hi is exact and is handed to the tensor core as it is: kind::tf32
// ignores the 13 low mantissa bits of its inputs, i.e. truncates lo, which leaves the 3xTF32 product error at 2^-21.
// Split cost: 3 instructions per element instead of 9.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ int tile_off(int r, int k) {  // float offset of element (r, k) inside an operand tile
  return (r >> 3) * 256 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // start address (>>4) | LBO = 128 B (>>4 = 8) at bit 16 | SBO = 1024 B (>>4 = 64) at bit 32 | version 1 at bit 46
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (64ull << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct Smem {
  float* a_hi; float* a_lo; float* b_hi; float* b_lo;
  uint64_t* bar;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(unsigned char* raw) {
  uintptr_t p = ((uintptr_t)raw + 1023) & ~(uintptr_t)1023;
  Smem s;
  s.a_hi = (float*)p;
  s.a_lo = s.a_hi + A_TILE;
  s.b_hi = s.a_lo + A_TILE;
  s.b_lo = s.b_hi + B_TILE;
  s.bar = (uint64_t*)(s.b_lo + B_TILE);
  s.tmem_slot = (uint32_t*)(s.bar + 2);
  return s;
}

constexpr int NTHR = 256;    // 8 warps: all stage operands; warps w and w + 4 share TMEM lanes 32 (w % 4) and split the columns
constexpr int NCOL = TN / 2;  // accumulator columns per thread

// split-store 4 consecutive-k values of row r into the hi / lo tiles (one 16-byte core-matrix row each)
__device__ __forceinline__ void put4(float* hi, float* lo, int r, int k, float4 v) {
  float4 h, l;
  h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
  l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
  const int o = tile_off(r, k);
  *reinterpret_cast<float4*>(hi + o) = h;
  *reinterpret_cast<float4*>(lo + o) = l;
}

// Operand loaders: `load` issues the global loads of one K block into registers (so their latency
// hides behind the MMAs of the previous block), `store` splits and writes them to shared memory.
// K-major source: src[r * pitch + k], rows r < n_valid_rows of a tile of tile_rows (<= 128) rows.  One
// row may come from another pointer (the ones row that carries the bias gradient).  A warp's 32
// loads cover 8 rows x 64 contiguous bytes; its 8-lane store phases are 128 contiguous bytes.
struct LoadKMajor {
  const float* src; size_t pitch; int tile_rows, n_valid_rows, special_row; const float* special_src;
  float4 v[4];
  float rs[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // running sum over K of this thread's part of its four rows (row = row_of(tid, it))
  __device__ __forceinline__ static int row_of(int tid, int it) { const int idx = tid + it * NTHR; return (idx & 7) + 8 * (idx >> 6); }
  __device__ __forceinline__ void load(int k0, int k_valid, int tid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * NTHR;
      const int r = (idx & 7) + 8 * (idx >> 6), c4 = (idx >> 3) & 7;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < n_valid_rows) {
        const float* p = (r == special_row ? special_src : src + (size_t)r * pitch) + k0 + 4 * c4;
        if (4 * c4 + 3 < k_valid) t = *reinterpret_cast<const float4*>(p);
        else {
          if (4 * c4 + 0 < k_valid) t.x = p[0];
          if (4 * c4 + 1 < k_valid) t.y = p[1];
          if (4 * c4 + 2 < k_valid) t.z = p[2];
        }
      }
      v[it] = t;
      rs[it] += (t.x + t.y) + (t.z + t.w);
    }
  }
  __device__ __forceinline__ void store(float* hi, float* lo, int tid) const {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * NTHR;
      const int r = (idx & 7) + 8 * (idx >> 6), c4 = (idx >> 3) & 7;
      if (r < tile_rows) put4(hi, lo, r, 4 * c4, v[it]);
    }
  }
};

// MN-major source: src[k * pitch + r] (rows contiguous), 128 rows x 32 k: coalesced along rows for every k.
struct LoadTransposed {
  const float* src; size_t pitch;
  float4 v[4];
  __device__ __forceinline__ void load(int k0, int k_valid, int tid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * NTHR;
      const int r = idx & (TM - 1), c4 = idx >> 7;
      const float* p = src + (size_t)(k0 + 4 * c4) * pitch + r;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * c4 + 0 < k_valid) t.x = p[0];
      if (4 * c4 + 1 < k_valid) t.y = p[pitch];
      if (4 * c4 + 2 < k_valid) t.z = p[2 * pitch];
      if (4 * c4 + 3 < k_valid) t.w = p[3 * pitch];
      v[it] = t;
    }
  }
  __device__ __forceinline__ void store(float* hi, float* lo, int tid) const {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * NTHR;
      put4(hi, lo, idx & (TM - 1), 4 * (idx >> 7), v[it]);
    }
  }
};

// acc[c] += TMEM buffer: this warp's 32 lanes, columns [col0, col0 + NCOL) clipped to n_tile
__device__ __forceinline__ void drain(uint32_t tmem_buf, int lane_base, int col0, int n_tile, float (&acc)[NCOL]) {
#pragma unroll
  for (int c0 = 0; c0 < NCOL; c0 += 32) {
    if (col0 + c0 < n_tile) {
      float v[32];
      tmem_ld32(tmem_buf + ((uint32_t)lane_base << 16) + (uint32_t)(col0 + c0), v);
#pragma unroll
      for (int q = 0; q < 32; ++q) acc[c0 + q] += v[q];
    }
  }
}

// K loop shared by both kernels.  Per K block: (wait for the previous block's MMAs) -> split-store
// the prefetched operands -> 12 MMAs into TMEM buffer kb & 1 (fresh accumulator) -> commit -> prefetch
// the next block's operands into registers and drain the previous block's buffer into `acc`, both
// while the MMAs run.  On return `acc` holds columns [col0, col0 + NCOL) of this thread's row of D.
template <class LA, class LB>
__device__ __forceinline__ void k_loop(const Smem& sm, uint32_t tmem_base, int n_tile, int K, int lane_base, int col0,
                                       LA& la, LB& lb, float (&acc)[NCOL]) {
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  const uint64_t da_hi = make_desc(smem_u32(sm.a_hi)), da_lo = make_desc(smem_u32(sm.a_lo));
  const uint64_t db_hi = make_desc(smem_u32(sm.b_hi)), db_lo = make_desc(smem_u32(sm.b_lo));
  const int tid = threadIdx.x;
#pragma unroll
  for (int c = 0; c < NCOL; ++c) acc[c] = 0.0f;
  uint32_t parity[2] = {0u, 0u};
  const int nkb = (K + BK - 1) / BK;
  if (nkb > 0) {
    la.load(0, min(BK, K), tid);
    lb.load(0, min(BK, K), tid);
  }
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb > 0) {  // MMAs of block kb - 1 done: shared memory is free, TMEM buffer buf ^ 1 is complete
      // one warp polls the mbarrier, the others sleep in the hardware barrier (8 polling warps were
      // measured to spend 2/3 of the kernel's issue slots on the try_wait loop)
      if (tid < 32) mbar_wait(sm.bar + (buf ^ 1), parity[buf ^ 1]);
      parity[buf ^ 1] ^= 1u;
      __syncthreads();
      tc_fence_after();
    }
    la.store(sm.a_hi, sm.a_lo, tid);
    lb.store(sm.b_hi, sm.b_lo, tid);
    fence_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
    tc_fence_before();    // orders the tcgen05.ld of the previous drain before the MMAs that reuse that buffer
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(buf * TN);
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 16);  // 256 B >> 4
        mma_tf32(d, da_lo + adv, db_hi + adv, idesc, ks > 0 ? 1u : 0u);
        mma_tf32(d, da_hi + adv, db_lo + adv, idesc, 1u);
        mma_tf32(d, da_hi + adv, db_hi + adv, idesc, 1u);
      }
      umma_commit(sm.bar + buf);  // arrives when the MMAs above have finished (and released shared memory)
    }
    if (kb + 1 < nkb) {
      const int k1 = (kb + 1) * BK;
      la.load(k1, min(BK, K - k1), tid);
      lb.load(k1, min(BK, K - k1), tid);
    }
    if (kb > 0) drain(tmem_base + (uint32_t)((buf ^ 1) * TN), lane_base, col0, n_tile, acc);
  }
  if (nkb > 0) {
    const int buf = (nkb - 1) & 1;
    if (tid < 32) mbar_wait(sm.bar + buf, parity[buf]);
    __syncthreads();
    tc_fence_after();
    drain(tmem_base + (uint32_t)(buf * TN), lane_base, col0, n_tile, acc);
  }
}

__device__ __forceinline__ uint32_t tmem_alloc(const Smem& sm, int warp) {
  if (threadIdx.x == 0) {
    mbar_init(sm.bar, 1);
    mbar_init(sm.bar + 1, 1);
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sm.tmem_slot;
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(TMEM_COLS) : "memory");
}

// ---- LINEAR: grid (cols / 128, ceil(n_out_pad / 128)); 256 threads.
// op fields as OP_LINEAR plus a10 / a11 = offset / leading dimension of the [out][in] copy of the weights,
// a12 = number of real (unpadded) outputs.
__global__ void __launch_bounds__(NTHR, 2)
vg_linear_tc(vm::Op op, float* __restrict__ rows, int C, const float* __restrict__ par) {
  extern __shared__ unsigned char smem_raw[];
  const Smem sm = carve(smem_raw);
  const int in_row = op.a[0], n_in = op.a[1], out_row = op.a[4], n_out = op.a[5], b_off = op.a[6], act = op.a[7],
            accumulate = op.a[8], pre_row = op.a[9], tw_off = op.a[10], tw_ld = op.a[11], n_real = op.a[12];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * TM, j0 = blockIdx.y * TN;
  const int n_valid = min(TN, n_out - j0);                    // padded outputs this tile writes
  const int n_rows_b = max(0, min(TN, n_real - j0));          // rows of the [out][in] weight copy that exist
  const int n_tile = (n_valid + 15) & ~15;
  const uint32_t tmem_d = tmem_alloc(sm, warp);
  LoadTransposed la{rows + (size_t)in_row * C + s0, (size_t)C};                                   // A[s][i] from in[i][s]
  LoadKMajor lb{par + tw_off + (size_t)j0 * tw_ld, (size_t)tw_ld, n_tile, n_rows_b, -1, nullptr};  // B[j][i]
  const int lane_base = 32 * (warp & 3), col0 = NCOL * (warp >> 2);
  float acc[NCOL];
  k_loop(sm, tmem_d, n_tile, n_in, lane_base, col0, la, lb, acc);
  // epilogue: this thread holds sample s0 + lane_base + lane (its TMEM lane); acc[c] = output j0 + col0 + c
  const size_t col = (size_t)s0 + lane_base + lane;
  const int acc_row = pre_row >= 0 ? pre_row : out_row;
  vm::with_act(act, [&](auto tag) {
    constexpr int ACT = decltype(tag)::value;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      const int j = j0 + col0 + c;
      if (col0 + c < n_valid) {
        float x = acc[c] + (b_off >= 0 ? __ldg(par + b_off + j) : 0.0f);
        if (accumulate) x += rows[(size_t)(acc_row + j) * C + col];
        if (pre_row >= 0) rows[(size_t)(pre_row + j) * C + col] = x;
        rows[(size_t)(out_row + j) * C + col] = vm::act_fwd(ACT, x);
      }
    }
  });
  tmem_free(tmem_d, warp);
}

// ---- LINEAR_DW: grid (ceil(n_in / 128), ceil(n_out / 128), slabs); 256 threads; K = samples of the slab.
// grad[w] += in^T g over the slab.  The bias gradient (column sums of g) is NOT a 129th / 257th / 513th operand row --
// that would add a whole extra tile row re-reading all of g -- but falls out of staging the g operand: every thread
// keeps the running sums of the g values it loads, and the CTAs of tile row 0 add them to grad[b].
__global__ void __launch_bounds__(NTHR, 2)
vg_dw_tc(vm::Op op, const float* __restrict__ rows, int C, int cols, int slab, float* __restrict__ grad,
         unsigned long long skip_tiles) {
  extern __shared__ unsigned char smem_raw[];
  const Smem sm = carve(smem_raw);
  const int in_row = op.a[0], n_in = op.a[1], g_row = op.a[2], n_out = op.a[3], w_off = op.a[4], ld = op.a[5],
            b_off = op.a[6];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.x * TM, j0 = blockIdx.y * TN;
  const int s_begin = blockIdx.z * slab, s_end = min(cols, s_begin + slab);
  const int m_valid = min(TM, n_in - i0), n_valid = min(TN, n_out - j0);
  const int n_tile = (n_valid + 15) & ~15;
  const bool bias_cta = blockIdx.x == 0 && b_off >= 0;
  if ((skip_tiles >> (blockIdx.x * gridDim.y + blockIdx.y)) & 1ull) {  // every weight of this tile is masked: exact-zero gradient
    if (!bias_cta) return;
    // still owes the bias gradient of its columns: plain column sums, no tensor core
    for (int j = warp; j < n_valid; j += NTHR / 32) {
      const float* g = rows + (size_t)(g_row + j0 + j) * C;
      float a = 0.0f;
      for (int s = s_begin + lane; s < s_end; s += 32) a += g[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) atomicAdd(grad + b_off + j0 + j, a);
    }
    return;
  }
  const uint32_t tmem_d = tmem_alloc(sm, warp);
  LoadKMajor la{rows + (size_t)(in_row + i0) * C + s_begin, (size_t)C, TM, m_valid, -1, nullptr};
  LoadKMajor lb{rows + (size_t)(g_row + j0) * C + s_begin, (size_t)C, n_tile, n_valid, -1, nullptr};
  const int lane_base = 32 * (warp & 3), col0 = NCOL * (warp >> 2);
  float acc[NCOL];
  k_loop(sm, tmem_d, n_tile, s_end - s_begin, lane_base, col0, la, lb, acc);
  const int i = i0 + lane_base + lane;  // TMEM lane = input index
  if (i < n_in) {
    float* dst = grad + w_off + (size_t)i * ld + j0 + col0;
#pragma unroll
    for (int c = 0; c < NCOL; ++c)
      if (col0 + c < n_valid) atomicAdd(dst + c, acc[c]);
  }
  if (bias_cta) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = LoadKMajor::row_of(tid, it);
      if (r < n_valid) atomicAdd(grad + b_off + j0 + r, lb.rs[it]);
    }
  }
  tmem_free(tmem_d, warp);
}

// =====================================================================================
// Pipelined LINEAR: operands are split (hi / lo TF32) and tiled ONCE into the canonical UMMA layout in
// global memory, so the GEMM's K loop is pure data movement + MMA:
//   producer thread   cp.async.bulk (TMA unit, no tensor map needed: a tile is 16 KB contiguous) into a
//                     3-stage shared-memory ring, completion on the stage's `full` mbarrier
//   MMA thread        waits `full`, issues the 12 tcgen05.mma of the K block into TMEM buffer kb & 1,
//                     tcgen05.commit -> the stage's `empty` mbarrier and the buffer's `tmem_full` mbarrier
//   4 drain warps     tcgen05.ld the finished buffer, add it into FP32 registers (round-to-nearest; see the
//                     note on the truncating tensor-core accumulator above), arrive on `tmem_empty`
// Tiled operand layout: block (row_tile, k_block) = [hi 4096 floats | lo 4096 floats], element (r, k) of a
// plane at tile_off(r, k); blocks ordered [row_tile][k_block].
constexpr int PIPE_STAGES = 3;
constexpr int PBK = 16;                              // K block of the pipelined kernel: 32 KB per stage -> 2 CTAs per SM,
                                                     // so one CTA's prologue / epilogue hides behind the other's MMAs
constexpr int TILE_FLOATS = TM * PBK;                // one plane of one block
constexpr int BLOCK_FLOATS = 2 * TILE_FLOATS;        // hi + lo
constexpr size_t PIPE_SMEM = (size_t)PIPE_STAGES * 2 * BLOCK_FLOATS * 4 + 1024 + 256;
constexpr int PIPE_THREADS = 192;                    // warp 0 producer, warp 1 MMA, warps 2..5 drain (lanes 32 * (w % 4))

__device__ __forceinline__ int ptile_off(int r, int k) {  // canonical K-major layout of a 128 x PBK plane
  return (r >> 3) * (PBK * 8) + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
}
__device__ __forceinline__ uint64_t pmake_desc(uint32_t saddr) {  // LBO 128 B, SBO = PBK / 4 core matrices = PBK * 32 B
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | ((uint64_t)(PBK * 2) << 32) | (1ull << 46);
}

// rows x K source -> tiled planes.  TRANSPOSED: src[k * pitch + r] (rows contiguous), else src[r * pitch + k].
// grid (row tiles, ceil(K / 32)): a CTA converts 128 rows x 32 k = two K blocks.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256)
vg_tile(const float* __restrict__ src, size_t pitch, int n_rows, int K, int nkb, float* __restrict__ dst) {
  const int rt = blockIdx.x;
  const int tid = threadIdx.x;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = tid + it * 256;
    int r, c4;
    if (TRANSPOSED) { r = idx & (TM - 1); c4 = idx >> 7; }
    else { r = (idx & 7) + 8 * (idx >> 6); c4 = (idx >> 3) & 7; }
    const int row = rt * TM + r, k0 = blockIdx.y * 32 + 4 * c4;
    const int kb = k0 / PBK;
    if (kb >= nkb) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n_rows) {
      if (TRANSPOSED) {
        const float* p = src + (size_t)k0 * pitch + row;
        if (k0 + 0 < K) v.x = p[0];
        if (k0 + 1 < K) v.y = p[pitch];
        if (k0 + 2 < K) v.z = p[2 * pitch];
        if (k0 + 3 < K) v.w = p[3 * pitch];
      } else {
        const float* p = src + (size_t)row * pitch + k0;
        if (k0 + 3 < K) v = *reinterpret_cast<const float4*>(p);
        else {
          if (k0 + 0 < K) v.x = p[0];
          if (k0 + 1 < K) v.y = p[1];
          if (k0 + 2 < K) v.z = p[2];
        }
      }
    }
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    float* blk = dst + ((size_t)rt * nkb + kb) * BLOCK_FLOATS;
    const int o = ptile_off(r, k0 - kb * PBK);
    *reinterpret_cast<float4*>(blk + o) = h;
    *reinterpret_cast<float4*>(blk + TILE_FLOATS + o) = l;
  }
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// out[j][s] = act(b[j] (+ prev) + sum_i W[i][j] in[i][s]) from tiled operands.
// a_tiles: activations, row tiles = 128-sample tiles;  b_tiles: the [out][in] weight copy, row tiles = 128 outputs.
// grid (sample tiles, output tiles); op fields as OP_LINEAR (a5 = padded outputs, a12 = real outputs).
__global__ void __launch_bounds__(PIPE_THREADS, 2)
vg_linear_pipe(vm::Op op, float* __restrict__ rows, int C, const float* __restrict__ par, const float* __restrict__ a_tiles,
               const float* __restrict__ b_tiles, int nkb, VmKRange kr) {
  extern __shared__ unsigned char smem_raw[];
  uintptr_t base = ((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023;
  float* stage_mem = (float*)base;  // [PIPE_STAGES][A block | B block]
  uint64_t* bars = (uint64_t*)(stage_mem + (size_t)PIPE_STAGES * 2 * BLOCK_FLOATS);
  uint64_t* full = bars;                       // [PIPE_STAGES]
  uint64_t* empty = bars + PIPE_STAGES;        // [PIPE_STAGES]
  uint64_t* tmem_full = bars + 2 * PIPE_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  const int out_row = op.a[4], n_out = op.a[5], b_off = op.a[6], act = op.a[7], accumulate = op.a[8], pre_row = op.a[9];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * TM, j0 = blockIdx.y * TN;
  const int n_valid = min(TN, n_out - j0);
  const int n_tile = (n_valid + 15) & ~15;
  // K blocks that can contribute to this output tile (structural zeros of the weights skipped at tile granularity)
  const int kb0 = kr.enabled ? (int)kr.lo[blockIdx.y] : 0;
  const int nk = (kr.enabled ? (int)kr.hi[blockIdx.y] : nkb) - kb0;
  if (tid == 0) {
    for (int s = 0; s < PIPE_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full + b, 1); mbar_init(tmem_empty + b, 4); }
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- producer
      const float* a_src = a_tiles + (size_t)blockIdx.x * nkb * BLOCK_FLOATS;
      const float* b_src = b_tiles + (size_t)blockIdx.y * nkb * BLOCK_FLOATS;
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % PIPE_STAGES;
        if (kb >= PIPE_STAGES) mbar_wait(empty + s, ((kb / PIPE_STAGES) - 1) & 1);
        float* sa = stage_mem + (size_t)s * 2 * BLOCK_FLOATS;
        mbar_expect_tx(full + s, 2 * BLOCK_FLOATS * 4);
        bulk_g2s(sa, a_src + (size_t)(kb0 + kb) * BLOCK_FLOATS, BLOCK_FLOATS * 4, full + s);
        bulk_g2s(sa + BLOCK_FLOATS, b_src + (size_t)(kb0 + kb) * BLOCK_FLOATS, BLOCK_FLOATS * 4, full + s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer
      const uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % PIPE_STAGES, buf = kb & 1;
        mbar_wait(full + s, (kb / PIPE_STAGES) & 1);
        if (kb >= 2) mbar_wait(tmem_empty + buf, ((kb >> 1) - 1) & 1);
        tc_fence_after();
        const float* sa = stage_mem + (size_t)s * 2 * BLOCK_FLOATS;
        const uint64_t da_hi = pmake_desc(smem_u32(sa)), da_lo = pmake_desc(smem_u32(sa + TILE_FLOATS));
        const uint64_t db_hi = pmake_desc(smem_u32(sa + BLOCK_FLOATS)), db_lo = pmake_desc(smem_u32(sa + BLOCK_FLOATS + TILE_FLOATS));
        const uint32_t d = tmem_base + (uint32_t)(buf * TN);
#pragma unroll
        for (int ks = 0; ks < PBK / 8; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 16);
          mma_tf32(d, da_lo + adv, db_hi + adv, idesc, ks > 0 ? 1u : 0u);
          mma_tf32(d, da_hi + adv, db_lo + adv, idesc, 1u);
          mma_tf32(d, da_hi + adv, db_hi + adv, idesc, 1u);
        }
        umma_commit(empty + s);          // shared-memory stage free once these MMAs have read it
        umma_commit(tmem_full + buf);    // accumulator of this K block complete
      }
    }
  } else {  // ---- drain warps 2..5
    const int lane_base = 32 * (warp & 3);
    float acc[TN];
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[c] = 0.0f;
    for (int kb = 0; kb < nk; ++kb) {
      const int buf = kb & 1;
      mbar_wait(tmem_full + buf, (kb >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        if (c0 < n_tile) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(buf * TN + c0), v);
#pragma unroll
          for (int q = 0; q < 32; ++q) acc[c0 + q] += v[q];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + buf);
    }
    const size_t col = (size_t)s0 + lane_base + lane;
    const int acc_row = pre_row >= 0 ? pre_row : out_row;
    vm::with_act(act, [&](auto tag) {
      constexpr int ACT = decltype(tag)::value;
      // the rows an accumulating product adds to are requested 32 at a time, BEFORE that batch's stores: a load placed after
      // a store to the same array cannot be hoisted by the compiler, and one dependent global load per output (128 per
      // thread, ~700 cycles each) cost an accumulating LINEAR four times its mainloop (cfg5: the masked first layer adds
      // onto the context Linear)
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        float prev[32];
        if (accumulate) {
#pragma unroll
          for (int q = 0; q < 32; ++q) prev[q] = c0 + q < n_valid ? rows[(size_t)(acc_row + j0 + c0 + q) * C + col] : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int c = c0 + q, j = j0 + c;
          if (c < n_valid) {
            float x = acc[c] + (b_off >= 0 ? __ldg(par + b_off + j) : 0.0f);
            if (accumulate) x += prev[q];
            if (pre_row >= 0) rows[(size_t)(pre_row + j) * C + col] = x;
            rows[(size_t)(out_row + j) * C + col] = vm::act_fwd(ACT, x);
          }
        }
      }
    });
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

}  // namespace vmtc

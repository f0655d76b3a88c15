// FP32-grade tcgen05 matrix products of the per-op executor, second generation (sm_100a): the products the fused
// conditioner + spline kernel (vm_fused.cuh) leaves over -- the weight-gradient contractions and input-gradient products
// of the backward pass, decoders, MADE layers:
//
//   LINEAR     out[j][s] = act(b[j] (+ prev) + sum_i Wt[j][i] in[i][s])    M = 128 samples, N <= 256 outputs, K = inputs
//              (hk.Linear / MaskedLinear, conditioners/mlp.py:24-29, nn/masked_linear.py:35-58; with the transposed copy
//              of W it is the input-gradient product g W^T of SURVEY appendix C)
//   LINEAR_DW  grad[i][j] += sum_s in[i][s] g[j][s], grad[b][j] += sum_s g[j][s]  M = 128 inputs, N <= 256 outputs, K = samples
//
// Both read their operands as plain float32 straight from where they live -- rows [row][sample] of the workspace and
// the staged parameters -- into a shared-memory ring: no separate re-tiling pass and no second copy of the activations
// exist.  Two loader variants (template parameter TMA):
//   * TMA (default): one thread issues tensor loads (cp.async.bulk.tensor.2d) that land in swizzled UMMA layouts -- K-major
//     operands with the 64-byte swizzle, LINEAR's sample-major activations as an MN-major operand on the 32-byte-atom
//     128-byte swizzle (details at tma_2d below); eight warps then only split;
//   * cp.async (fallback when a tensor map cannot describe an operand: base not 16-byte aligned, pitch not a multiple of
//     4 floats): eight loader warps copy 16-byte pieces (LINEAR's activations: 4-byte copies that transpose on the fly)
//     into the K-major no-swizzle layout and split what they loaded.
// The split: the tensor core reads the float32 bits as TF32 (it ignores the 13 low mantissa bits: hi = truncation, free)
// and lo = x - trunc(x) goes to a second ring; per 8-wide k step the issuing thread runs lo*hi + hi*lo + hi*hi
// (3xTF32, product error 2^-21).  The tensor core's accumulator add truncates, so only BLK = 256 k are accumulated
// inside it; four drain warps add every finished block into a MASTER accumulator in tensor memory with round-to-nearest
// adds (tcgen05.ld / FADD / tcgen05.st) -- N = 256 is the MMA shape that reaches the tensor pipe's rate (a 128-wide
// MMA costs the same 140-cycle floor, tools/ubench/f16_probe.cu), and 2 x 256 columns are all of tensor memory.
// Warp roles: 0..7 load / split, 8 issues the MMAs, 9..12 drain / epilogue (TMEM lanes 32 * (warp % 4)); TMA variant:
// + 13 tensor-load producer (raw ring), 14 producer of the pre-split weight lo part (lo ring, BLO).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vm.cuh"
#include "vm_plan.h"
#include "vm_tc.cuh"

namespace vmg {

using vmtc::fence_async_smem;
using vmtc::mbar_arrive;
using vmtc::mbar_init;
using vmtc::mbar_wait;
using vmtc::smem_u32;
using vmtc::tc_fence_after;
using vmtc::tc_fence_before;
using vmtc::umma_commit;

// k per stage, ring depths: bytes in flight are what this kernel lives on.  Measured on the gradient-phase products of a
// cfg4 train step (weight gradients / input gradients, ms): KS 16, RS 4, LS 3 (72 KB in flight): 209 / 202;
// KS 32, RS 2, LS 2 (full cache lines per row piece, but 48 KB in flight): 316 / 258.
constexpr int TM = 128, TN = 256, KS = 16;     // tile, k per stage
constexpr int BLK = 256 / KS;                  // stages per accumulation block inside the tensor core (256 k)
constexpr int RS = 6, LS = 2;                  // raw / lo ring depths
constexpr int KQ = KS / 4;                     // 16-byte pieces per row and stage
constexpr int A_BYTES = TM * KS * 4, B_BYTES = TN * KS * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int LOADERS = 256, THREADS = 416;
// TMA-fed variant.  A stage holds a raw slot from the tensor load's issue until its MMAs complete (load latency L ~ 1800
// cycles + split S + MMAs E) and a lo slot from the split's start until then (S + E); S is ~870 cycles of LDS / STS for 12
// pieces per thread plus ~700 .. 1400 for the proxy fence (fence.proxy.async after the lo pieces are written), E ~ 930
// (six 128 x 256 x 8 MMAs).  Throughput is min(raw slots / (L + S + E), lo slots / (S + E), sets / S, 1 / E) -- measured
// (cfg4 train step, cycles per stage on CTA 0 / weight-gradient ms / input-gradient ms):
//   raw 6, lo 2, 2 sets of 4 warps: 1750 / 145 / 154     raw 4, lo 4, 4 sets of 2 warps: 1840 / 154 / 158
//   raw 5, lo 3, 2 sets of 4 warps: 1310 / 112 / 130     raw 5, lo 3, 3 sets of 4 warps (576 threads, drain spills): 1660 / 138 / 160
constexpr int TRS = 5, TLS = 3, TSETS = 2, TSW = 8;   // raw / lo ring depths, splitter sets, splitter warps
constexpr int TMA_THREADS = (TSW + 7) * 32;            // + MMA issuer, four drain warps, two tensor-load producers (raw ring, lo ring)
constexpr size_t SMEM_BYTES = 1024 + (size_t)(RS + LS) * STAGE_BYTES + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "rings exceed the shared memory of an SM");

__device__ __forceinline__ uint32_t soff(int r, int k) {  // byte offset of element (r, k) in a [rows x KS] K-major tile
  return (uint32_t)((r >> 3) * (KQ * 128) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ uint64_t desc16(uint32_t saddr) {  // LBO 128 B, SBO = KQ core matrices of 4 k per 8-row group
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | ((uint64_t)(KQ * 8) << 32) | (1ull << 46);
}
// TMA-fed variant (TMA = true): one elected thread of a 14th warp issues tensor loads -- LINEAR_DW and the weight operand of
// LINEAR: box {16 k, rows} with the 64-byte swizzle (K-major, SBO 512 B, k step of 8 = +32 B); LINEAR's activations,
// sample-major in memory: four boxes {32 samples, 16 k} with the 128-byte swizzle on 32-byte atoms, the shared-memory form
// tcgen05 wants for an MN-major 32-bit operand (layout code 1, LBO 2048 B = next 32 samples, SBO 512 B = next 4 k, k step
// of 8 = +1024 B, instruction-descriptor bit 15) -- both validated on exact integer data by tools/ubench/tma_probe.cu.
// The eight loader warps then only split (lo = x - trunc(x) is elementwise: piece i of the stage's bytes -> piece i of the
// lo stage, whatever the swizzle).  Why: the cp.async path is bound by the LDGSTS issue rate (12 cp.async per thread and
// owned stage take 1.7 .. 2.4k cycles; the data itself arrives in time -- in-kernel counters, profiles/r2_summary.md).
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tm_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tm_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// piece q of a K-major [rows x 16] tile -> (row, k quarter) such that EIGHT CONSECUTIVE LANES hold the same k quarter of
// eight consecutive rows: their 16-byte shared-memory accesses are 128 contiguous bytes (one conflict-free wavefront).
// (row = q / 4, quarter = q % 4 -- contiguous along k -- puts four lanes on the same banks: every LDGSTS / LDS / STS of
// the loaders took four times its wavefronts and the MMA issuer waited for the loaders 62 % of its time.)
__device__ __forceinline__ void piece_of(int q, int& r, int& kq) {
  r = (q & 7) | ((q >> 5) << 3);
  kq = (q >> 3) & 3;
}

struct Args {
  // operands
  const float* a_base; long long a_pitch; int a_rows;  // LINEAR: in rows [k][sample] (+ sample offset), a_rows = K;  DW: in rows [i][sample], a_rows = valid i
  const float* b_base; long long b_pitch; int b_rows;  // B[n][k]: LINEAR: Wt rows (K-major); DW: g rows [j][sample]; b_rows = valid n of this tile set
  int K;                                               // LINEAR: n_in; DW: samples of the slab set (per CTA: [k_begin, k_end))
  int k_slab;                                          // DW: samples per CTA (multiple of 16); LINEAR: unused
  // LINEAR epilogue
  float* rows; long long C;                            // row workspace
  int out_row, pre_row, n_out, b_off, act, accumulate; // n_out: padded outputs written
  const float* par;
  int ab_on, ab_pre_row, ab_n, ab_act;                 // the OP_ACT_BWD that follows, fused: out[j] = act'(rows[ab_pre_row + j]) * out[j], j < ab_n
  // DW epilogue
  float* grad; int w_off, ld, gb_off, n_in;            // grad[w_off + i * ld + j], grad[gb_off + j]
  VmKRange kr;                                         // LINEAR: k-stage range per 128-output tile (structural zeros), by value
  unsigned long long skip;                             // DW: all-zero 128 x 128 tiles (bit i_tile * n_j128 + j128_tile)
  int n_j128;
  long long* dbg;                                      // SURJ_B200_FUSED_DBG=1: CTA (0,0,0)'s MMA issuer adds [10] cycles waiting for a split stage,
                                                       // [11] for the drain of a block, [12] its total cycles
};

// remote arrive on the barrier at the same shared-memory offset in CTA 0 of the cluster (the pair's leader)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope (arrivals from the peer CTA)
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// drain warps 9..12: add every finished block of the work accumulator into the master, then the epilogue from tensor memory.
// PAIR: blk_done lives in the leader CTA (8 arrivals: four warps of each CTA).
template <int MODE, bool PAIR>
__device__ __forceinline__ void drain_epilogue(const Args& g, int warp, int lane, int n_blocks, int n_stages, int n_tile, int n_valid, int n0,
                                               int m0, uint32_t t_work, uint32_t t_master, uint64_t* blk_full, uint64_t* blk_done) {
  const int quarter = warp & 3;
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
  for (int blk = 0; blk < n_blocks; ++blk) {
    mbar_wait(blk_full, (uint32_t)blk & 1u);
    tc_fence_after();
    if (n_blocks > 1) {  // master (+)= work, round-to-nearest
      for (int c0 = 0; c0 < n_tile; c0 += 32) {
        uint32_t w[32], m[32];
        tm_ld32(t_work + lane_addr + (uint32_t)c0, w);
        if (blk > 0) tm_ld32(t_master + lane_addr + (uint32_t)c0, m);
        tm_ld_wait();
        if (blk > 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) + __uint_as_float(m[i]));
        }
        tm_st32(t_master + lane_addr + (uint32_t)c0, w);
      }
      tm_st_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { if (PAIR) mbar_arrive_leader(blk_done); else mbar_arrive(blk_done); }
  }
  const uint32_t t_acc = (n_blocks > 1 ? t_master : t_work) + lane_addr;
  tc_fence_after();
  if (MODE == 0) {
    const long long col = (long long)m0 + quarter * 32 + lane;  // this thread's sample
    const int acc_row = g.pre_row >= 0 ? g.pre_row : g.out_row;
    const int n_write = min(TN, g.n_out - n0);                   // padded outputs of this tile
    if (g.ab_on) {
      // input-gradient product followed by the activation derivative (no bias, identity, no accumulate: vm_global.cu):
      // the 32 pre-activations of a column batch are requested together, before any of this batch's stores
      vm::with_act(g.ab_act, [&](auto tag) {
        constexpr int ACT = decltype(tag)::value;
        for (int c0 = 0; c0 < n_write; c0 += 32) {
          uint32_t w[32];
          float pre[32];
          if (n_stages > 0) tm_ld32(t_acc + (uint32_t)c0, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = n0 + c0 + i;
            pre[i] = (c0 + i < n_write && j < g.ab_n) ? g.rows[(long long)(g.ab_pre_row + j) * g.C + col] : 0.0f;
          }
          if (n_stages > 0) tm_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = n0 + c0 + i;
            if (c0 + i < n_write) {
              float x = n_stages > 0 ? __uint_as_float(w[i]) : 0.0f;
              if (j < g.ab_n) x = vm::act_bwd_fast(ACT, pre[i], x);
              g.rows[(long long)(g.out_row + j) * g.C + col] = x;
            }
          }
        }
      });
    } else {
      vm::with_act(g.act, [&](auto tag) {
        constexpr int ACT = decltype(tag)::value;
        for (int c0 = 0; c0 < n_write; c0 += 32) {
          uint32_t w[32];
          float prev[32];
          if (n_stages > 0) tm_ld32(t_acc + (uint32_t)c0, w);
          if (g.accumulate) {  // requested together, before this batch's stores (see vg_linear_pipe)
#pragma unroll
            for (int i = 0; i < 32; ++i) prev[i] = c0 + i < n_write ? g.rows[(long long)(acc_row + n0 + c0 + i) * g.C + col] : 0.0f;
          }
          if (n_stages > 0) tm_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = n0 + c0 + i;
            if (c0 + i < n_write) {
              float x = (n_stages > 0 ? __uint_as_float(w[i]) : 0.0f) + (g.b_off >= 0 ? __ldg(g.par + g.b_off + j) : 0.0f);
              if (g.accumulate) x += prev[i];
              if (g.pre_row >= 0) g.rows[(long long)(g.pre_row + j) * g.C + col] = x;
              g.rows[(long long)(g.out_row + j) * g.C + col] = vm::act_fwd(ACT, x);
            }
          }
        }
      });
    }
  } else if (n_stages > 0) {
    const int i = m0 + quarter * 32 + lane;  // TMEM lane = input feature
    for (int c0 = 0; c0 < n_valid; c0 += 32) {
      uint32_t w[32];
      tm_ld32(t_acc + (uint32_t)c0, w);
      tm_ld_wait();
      if (i < g.n_in) {
        float* dst = g.grad + g.w_off + (long long)i * g.ld + n0 + c0;
#pragma unroll
        for (int q = 0; q < 32; ++q)
          if (c0 + q < n_valid) atomicAdd(dst + q, __uint_as_float(w[q]));
      }
    }
  }
}

// MODE 0: LINEAR, grid (sample tiles, ceil(n_out / 256)).   MODE 1: LINEAR_DW, grid (ceil(n_in / 128), ceil(n_out / 256), slabs).
// TMA: tensor-load producer warp (THREADS + 32 threads, map_a / map_b from make_maps below); else cp.async loaders, maps unused.
// BLO (TMA, LINEAR): the B operand is a weight matrix whose lo part was computed once per call (vg_par_lo): a second
// producer thread tensor-loads it from map_blo straight into the lo ring, and the splitters only touch the activation tile.
template <int MODE, bool TMA, bool BLO = false>
__global__ void __launch_bounds__(TMA ? TMA_THREADS : THREADS, 1)
vg_gemm3(const __grid_constant__ Args g, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
         const __grid_constant__ CUtensorMap map_blo) {
  constexpr int RSx = TMA ? TRS : RS, LSx = TMA ? TLS : LS;   // ring depths of this variant
  constexpr int NSW = TMA ? TSW : 8;                          // warps 0 .. NSW-1 load / split, NSW issues the MMAs, NSW+1 .. NSW+4 drain, NSW+5 produces
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* raw = base;                                  // [RSx][A | B]
  unsigned char* lo = base + (size_t)RSx * STAGE_BYTES;        // [LSx][A | B]
  uint64_t* bars = (uint64_t*)(lo + (size_t)LSx * STAGE_BYTES);
  uint64_t* full = bars;             // [LSx]  4 arrivals (the warps of the set that owns the stage): split and visible
  uint64_t* lo_empty = full + LSx;    // [LSx]  MMA commit
  uint64_t* raw_empty = lo_empty + LSx;  // [RSx] MMA commit
  uint64_t* blk_full = raw_empty + RSx;  // [1]  MMA commit: a block's accumulator is complete
  uint64_t* blk_done = blk_full + 1;    // [1]  4 arrivals: the block has been added to the master
  uint64_t* raw_full = blk_done + 1;    // [RSx] TMA only: the stage's bytes have landed
  uint32_t* tmem_slot = (uint32_t*)(raw_full + RSx);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- this CTA's problem
  const int n0 = blockIdx.y * TN;                        // first output (B row) of the tile
  const int n_valid = min(TN, g.b_rows - n0);            // B rows that exist
  const int n_tile = (max(n_valid, 16) + 15) & ~15;      // MMA N
  int m0 = 0, k_begin = 0, k_end = g.K;
  if (MODE == 0) {
    m0 = blockIdx.x * TM;                                // first sample of the tile
    if (g.kr.enabled) {  // structural zeros: k stages that can contribute, per 128 outputs -> union over this tile's two halves
      int l = 1 << 30, h = 0;
      for (int half = 0; half < 2; ++half) {
        const int tl = 2 * (int)blockIdx.y + half;
        if (tl < 8 && half * 128 < n_valid && g.kr.hi[tl] > g.kr.lo[tl]) { l = min(l, (int)g.kr.lo[tl]); h = max(h, (int)g.kr.hi[tl]); }
      }
      k_begin = h > 0 ? l * 16 : 0;  // (the ranges are in 16-k blocks: vm_global.cu tc_info)
      k_end = h > 0 ? min(g.K, h * 16) : 0;
    }
  } else {
    m0 = blockIdx.x * TM;                                // first input feature of the tile
    k_begin = blockIdx.z * g.k_slab; k_end = min(g.K, k_begin + g.k_slab);
  }
  const int m_valid = MODE == 0 ? TM : min(TM, g.a_rows - m0);
  const int n_stages = k_end > k_begin ? (k_end - k_begin + KS - 1) / KS : 0;
  const int n_blocks = (n_stages + BLK - 1) / BLK;
  bool skip_tile = false;
  if (MODE == 1 && g.skip != 0ull) {  // both 128-column halves structurally zero: nothing but the bias gradient to do
    const int j128 = 2 * blockIdx.y;
    const bool z0 = (g.skip >> (blockIdx.x * g.n_j128 + j128)) & 1ull;
    const bool z1 = (j128 + 1 >= g.n_j128) || ((g.skip >> (blockIdx.x * g.n_j128 + j128 + 1)) & 1ull);
    skip_tile = z0 && z1;
  }

  if (tid == 0) {
    for (int s = 0; s < LSx; ++s) { mbar_init(full + s, TMA ? TSW / TSETS + (BLO ? 1 : 0) : LOADERS / 64); mbar_init(lo_empty + s, 1); }
    for (int s = 0; s < RSx; ++s) { mbar_init(raw_empty + s, 1); mbar_init(raw_full + s, 1); }
    mbar_init(blk_full, 1);
    mbar_init(blk_done, 4);
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_work = tmem, t_master = tmem + TN;

  if (TMA && warp < NSW) {
    // ------------------------------------------------------------------ splitters behind the tensor loads
    // TSETS sets of warps take the stages in turn; thread ts of the set splits pieces ts + SETT e of the stage's 1536
    // (consecutive lanes, consecutive 16 bytes: conflict-free; the first 512 pieces are A, row of a B piece = (piece - 512) / 4)
    constexpr int SETT = TSW * 32 / TSETS;             // threads of a set
    constexpr int PA = A_BYTES / 16 / SETT, PB = B_BYTES / 16 / SETT, HALF = (PA + PB) / 2;
    const int set = warp / (TSW / TSETS), ts = tid % SETT;
    float bsum[PB];
#pragma unroll
    for (int e = 0; e < PB; ++e) bsum[e] = 0.0f;
    auto low = [](const float4& x) {
      float4 l;
      l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
      l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
      l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
      l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
      return l;
    };
    if (!skip_tile) {
      const bool lrec = g.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
      long long lt[4] = {0, 0, 0, 0};
      for (int t = set; t < n_stages; t += TSETS) {
        long long c0 = lrec ? clock64() : 0, c1 = 0;
        if (lane == 0) {
          mbar_wait(raw_full + t % RSx, (uint32_t)(t / RSx) & 1u);
          c1 = lrec ? clock64() : 0;
          if (t >= LSx) mbar_wait(lo_empty + t % LSx, (uint32_t)(t / LSx - 1) & 1u);
        }
        __syncwarp();
        long long c2 = lrec ? clock64() : 0;
        const unsigned char* sr = raw + (size_t)(t % RSx) * STAGE_BYTES;
        unsigned char* sl = lo + (size_t)(t % LSx) * STAGE_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // two batches: all loads of a batch before its stores (independent loads), 48 registers
          float4 x[HALF];
#pragma unroll
          for (int q = 0; q < HALF; ++q) {
            const int e = h * HALF + q;
            if (e < PA || (!BLO && (ts >> 2) + (SETT / 4) * (e - PA) < n_tile)) x[q] = *reinterpret_cast<const float4*>(sr + 16 * (ts + SETT * e));
          }
#pragma unroll
          for (int q = 0; q < HALF; ++q) {
            const int e = h * HALF + q;
            if (e < PA || (!BLO && (ts >> 2) + (SETT / 4) * (e - PA) < n_tile)) {
              *reinterpret_cast<float4*>(sl + 16 * (ts + SETT * e)) = low(x[q]);
              if (MODE == 1 && e >= PA) bsum[e - PA] += (x[q].x + x[q].y) + (x[q].z + x[q].w);
            }
          }
        }
        long long c3 = lrec ? clock64() : 0;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full + t % LSx);
        if (lrec) { lt[0] += c1 - c0; lt[1] += c2 - c1; lt[2] += c3 - c2; lt[3] += clock64() - c3; }
      }
      if (lrec)
        for (int q = 0; q < 4; ++q) atomicAdd((unsigned long long*)g.dbg + 17 + q, (unsigned long long)lt[q]);
      if (MODE == 1 && blockIdx.x == 0 && g.gb_off >= 0) {  // bias gradient: the four pieces of a g row sit in four consecutive lanes
#pragma unroll
        for (int e = 0; e < PB; ++e) {
          float a = bsum[e];
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          const int r = (ts >> 2) + (SETT / 4) * e;
          if ((ts & 3) == 0 && r < n_valid) atomicAdd(g.grad + g.gb_off + n0 + r, a);
        }
      }
    } else if (MODE == 1 && blockIdx.x == 0 && g.gb_off >= 0) {
      for (int j = warp; j < n_valid; j += NSW) {
        const float* gr = g.b_base + (long long)(n0 + j) * g.b_pitch;
        float a = 0.0f;
        for (int s = k_begin + lane; s < k_end; s += 32) a += gr[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) atomicAdd(g.grad + g.gb_off + n0 + j, a);
      }
    }
  } else if (TMA && warp == NSW + 5) {
    if (lane == 0 && !skip_tile) {  // ------------------------------------------------------ tensor-load producer
      for (int t = 0; t < n_stages; ++t) {
        if (t >= RSx) mbar_wait(raw_empty + t % RSx, (uint32_t)(t / RSx - 1) & 1u);
        const int k0 = k_begin + t * KS;
        const uint32_t sa = smem_u32(raw + (size_t)(t % RSx) * STAGE_BYTES), sb = sa + A_BYTES;
        vmtc::mbar_expect_tx(raw_full + t % RSx, STAGE_BYTES);
        if (MODE == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_2d(sa + 2048u * c, &map_a, m0 + 32 * c, k0, raw_full + t % RSx);
        } else {
          tma_2d(sa, &map_a, k0, m0, raw_full + t % RSx);
        }
        tma_2d(sb, &map_b, k0, n0, raw_full + t % RSx);
      }
    }
  } else if (TMA && BLO && warp == NSW + 6) {
    if (lane == 0 && !skip_tile) {  // ------------------------------------------------------ lo-ring producer: the weights' lo part
      for (int t = 0; t < n_stages; ++t) {
        if (t >= LSx) mbar_wait(lo_empty + t % LSx, (uint32_t)(t / LSx - 1) & 1u);
        vmtc::mbar_expect_tx(full + t % LSx, B_BYTES);   // arrive + expect: the stage is full when this and the splitters' arrivals are in
        tma_2d(smem_u32(lo + (size_t)(t % LSx) * STAGE_BYTES) + A_BYTES, &map_blo, k_begin + t * KS, n0, full + t % LSx);
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ loaders: cp.async, then split what they loaded
    // A tile 128 x 16 floats = 512 pieces of 16 B (2 per thread); B tile 256 x 16 = 1024 pieces (4 per thread);
    // piece q -> row q / 4, k quarter q % 4 (K-major sources).  LINEAR's A is sample-major: element-wise 4-byte copies,
    // thread -> sample tid % 128, k quarters tid / 128 and tid / 128 + 2.
    // (Measured alternatives, weight-gradient / input-gradient products of a cfg4 train step in ms against 203 / 197 for
    // this version: two sets of four warps taking the stages in turn, one stage in flight per set: 286 / 230; lanes 0..3 /
    // 4..7 of every group of eight owning even / odd stages so that a warp instruction covers whole 128-byte lines, two
    // stages in flight per set: 348 / 208; 32-k stages with two raw slots: 316 / 258.  Every thread taking part in
    // every stage, with as many stages in flight as shared memory holds, wins.)
    // Two sets of four warps take the stages in turn (set = warp / 4 owns stages set, set + 2, ...): the per-stage chain
    // landed -> split -> proxy fence -> arrive -> refill is latency, not throughput, so two sets halve its cost per stage.
    // Each set keeps two stages in flight behind the one it is splitting (three of the six raw slots).
    const int set = warp >> 2, ts = tid & 127, wq = warp & 3;
    constexpr int PA = TM * KQ / (LOADERS / 2), PB = TN * KQ / (LOADERS / 2);  // pieces per thread and owned stage (A: 4, B: 8)
    float bsum[PB];  // DW: running sums of this thread's g rows (bias gradient)
#pragma unroll
    for (int e = 0; e < PB; ++e) bsum[e] = 0.0f;
    // Per-thread addressing, computed once: piece e of this thread is row r0 + 32 e, k quarter kq (piece_of(ts + 128 e)), so
    // shared-memory offsets are so0 + 2048 e and global addresses advance by 32 rows per piece and KS floats per stage.
    // (Computing row / quarter / validity / 64-bit address per piece and stage cost the loaders 1.7k cycles per owned
    // stage -- as much as the split -- and the MMA issuer waited for them 57 % of its time.)
    const int r0 = (ts & 7) | ((ts >> 5) << 3), kq0 = (ts >> 3) & 3;
    const uint32_t so0 = soff(r0, 4 * kq0);
    const float* pb0 = g.b_base + (long long)(n0 + r0) * g.b_pitch + k_begin + 4 * kq0;
    const long long pb_step = 32 * g.b_pitch;
    const float* pa0;      // MODE 1: like pb0.  MODE 0: row (lane / 8) of the stage, sample m0 + 32 wq + lane % 8
    long long pa_step;     // MODE 1: 32 rows.   MODE 0: 4 rows (one k quarter)
    uint32_t sa0 = 0;      // MODE 0: soff of this lane's element in copy 0
    if (MODE == 0) {
      pa0 = g.a_base + (long long)(k_begin + (lane >> 3)) * g.a_pitch + m0 + 32 * wq + (lane & 7);
      pa_step = 4 * g.a_pitch;
      sa0 = soff(32 * wq + (lane & 7), lane >> 3);
    } else {
      pa0 = g.a_base + (long long)(m0 + r0) * g.a_pitch + k_begin + 4 * kq0;
      pa_step = 32 * g.a_pitch;
    }
    auto issue = [&](int t) {
      const int k0 = k_begin + t * KS;
      const uint32_t sa = smem_u32(raw + (size_t)(t % RS) * STAGE_BYTES), sb = sa + A_BYTES;
      const bool whole = k0 + KS <= k_end;       // every k of the stage exists (all stages but a ragged last one)
      if (MODE == 0) {  // sample-major source, element-wise copies: a warp instruction moves 8 samples x the 4 k of one quarter
                        // (32 distinct banks; 4 sectors of 32 B in memory); step e -> samples 32 wq + 8 (e / 4) + lane % 8, k = 4 (e % 4) + lane / 8
        const float* pa = pa0 + (long long)(t * KS) * g.a_pitch;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const bool ok = whole || k0 + 4 * (e & 3) + (lane >> 3) < k_end;
          // soff(smp + 8 (e / 4), k + 4 (e % 4)) = sa0 + (e / 4) * KQ * 128 + (e % 4) * 128
          cp4(sa + sa0 + (uint32_t)((e >> 2) * (KQ * 128) + (e & 3) * 128), ok ? pa + (e & 3) * pa_step + 8 * (e >> 2) : g.a_base, ok ? 4 : 0);
        }
      } else {
        const float* pa = pa0 + t * KS;
#pragma unroll
        for (int e = 0; e < PA; ++e) {
          const int kk = k0 + 4 * kq0;
          const int nb = r0 + 32 * e < m_valid ? (whole ? 16 : (kk < k_end ? min(16, 4 * (k_end - kk)) : 0)) : 0;
          cp16(sa + so0 + 2048u * e, nb ? pa + e * pa_step : g.a_base, nb);
        }
      }
      const float* pb = pb0 + t * KS;
#pragma unroll
      for (int e = 0; e < PB; ++e) {
        const int kk = k0 + 4 * kq0;
        const int nb = r0 + 32 * e < n_valid ? (whole ? 16 : (kk < k_end ? min(16, 4 * (k_end - kk)) : 0)) : 0;
        if (r0 + 32 * e < n_tile) cp16(sb + so0 + 2048u * e, nb ? pb + e * pb_step : g.b_base, nb);
      }
    };
    // all of a stage's pieces are read before any lo piece is written: the loads are independent (the compiler cannot
    // prove the two rings disjoint and would otherwise order every load behind the previous piece's store)
    auto split = [&](int t) {
      unsigned char* sr = raw + (size_t)(t % RS) * STAGE_BYTES;
      unsigned char* sl = lo + (size_t)(t % LS) * STAGE_BYTES;
      auto low = [](const float4& x) {
        float4 l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
        l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
        l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        return l;
      };
      float4 xa[PA], xb[PB];
      uint32_t oa[PA];
      if (MODE == 0) __syncwarp();  // the 4 k of a piece came from 4 lanes of this warp
#pragma unroll
      for (int e = 0; e < PA; ++e) {
        if (MODE == 0) {  // the warp splits the 128 pieces it copied
          const int pi = lane + 32 * e, c = 16 * wq + (pi >> 3);
          oa[e] = soff(((c >> 2) << 3) | (pi & 7), 4 * (c & 3));
        } else {
          oa[e] = so0 + 2048u * e;
        }
        xa[e] = *reinterpret_cast<const float4*>(sr + oa[e]);
      }
#pragma unroll
      for (int e = 0; e < PB; ++e)
        if (r0 + 32 * e < n_tile) xb[e] = *reinterpret_cast<const float4*>(sr + A_BYTES + so0 + 2048u * e);
#pragma unroll
      for (int e = 0; e < PA; ++e) *reinterpret_cast<float4*>(sl + oa[e]) = low(xa[e]);
#pragma unroll
      for (int e = 0; e < PB; ++e)
        if (r0 + 32 * e < n_tile) {
          *reinterpret_cast<float4*>(sl + A_BYTES + so0 + 2048u * e) = low(xb[e]);
          if (MODE == 1) bsum[e] += (xb[e].x + xb[e].y) + (xb[e].z + xb[e].w);
        }
    };
    if (!skip_tile) {
      // per set: the stage being split plus two in flight; the refill of a slot is issued at the END of an iteration (it
      // waits for the MMAs of this set's previous stage, which have had two stage times to finish)
      for (int t = set; t < set + 4; t += 2) {
        if (t < n_stages) issue(t);
        cp_commit();
      }
      const bool lrec = g.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
      long long lt[5] = {0, 0, 0, 0, 0};
      for (int t = set; t < n_stages; t += 2) {
        long long c0 = lrec ? clock64() : 0;
        cp_wait<1>();                                        // this thread's pieces of stage t have landed
        long long c1 = lrec ? clock64() : 0;
        if (t >= LS) {                                       // one lane polls, the warp follows (128 spinning threads cost
          if (lane == 0) mbar_wait(lo_empty + t % LS, (uint32_t)(t / LS - 1) & 1u);  // shared-memory bandwidth the MMAs need)
          __syncwarp();
        }
        long long c2 = lrec ? clock64() : 0;
        split(t);
        fence_async_smem();
        __syncwarp();                                        // one arrival per warp: 128 arrivals per stage on one shared-memory
        if (lane == 0) mbar_arrive(full + t % LS);           // word serialise (they were most of the loaders' per-stage chain)
        long long c3 = lrec ? clock64() : 0;
        const int tn = t + 4;                                // into the slot this set's stage t - 2 used
        if (tn < n_stages) {
          if (tn >= RS) {
            if (lane == 0) mbar_wait(raw_empty + tn % RS, (uint32_t)(tn / RS - 1) & 1u);
            __syncwarp();
          }
          long long c4 = lrec ? clock64() : 0;
          issue(tn);
          if (lrec) { lt[3] += c4 - c3; lt[4] += clock64() - c4; }
        }
        cp_commit();
        if (lrec) { lt[0] += c1 - c0; lt[1] += c2 - c1; lt[2] += c3 - c2; }
      }
      if (lrec)
        for (int q = 0; q < 5; ++q) atomicAdd((unsigned long long*)g.dbg + 17 + q, (unsigned long long)lt[q]);
    } else if (MODE == 1 && blockIdx.x == 0 && g.gb_off >= 0) {
      // masked tile that still owes the bias gradient of its columns: plain column sums
      for (int j = warp; j < n_valid; j += 8) {
        const float* gr = g.b_base + (long long)(n0 + j) * g.b_pitch;
        float a = 0.0f;
        for (int s = k_begin + lane; s < k_end; s += 32) a += gr[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) atomicAdd(g.grad + g.gb_off + n0 + j, a);
      }
    }
    if (MODE == 1 && !skip_tile && blockIdx.x == 0 && g.gb_off >= 0) {  // bias gradient: a row's four k quarters sit in lanes l, l + 8, l + 16, l + 24
#pragma unroll
      for (int e = 0; e < PB; ++e) {
        float a = bsum[e];
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 16);
        if (kq0 == 0 && r0 + 32 * e < n_valid) atomicAdd(g.grad + g.gb_off + n0 + r0 + 32 * e, a);
      }
    }
  } else if (warp == NSW) {
    if (lane == 0 && !skip_tile) {  // ---------------------------------------------------- MMA issuer
      const bool rec = g.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      long long w_full = 0, w_blk = 0;
      const long long t_start = rec ? clock64() : 0;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TMA && MODE == 0) ? (1u << 15) : 0u) | ((uint32_t)(n_tile >> 3) << 17) |
                             ((uint32_t)(TM >> 4) << 24);
      for (int t = 0; t < n_stages; ++t) {
        const int blk = t / BLK;
        long long c0 = rec ? clock64() : 0;
        if (t % BLK == 0 && blk > 0) mbar_wait(blk_done, (uint32_t)(blk - 1) & 1u);  // the work accumulator has been drained
        long long c1 = rec ? clock64() : 0;
        mbar_wait(full + t % LSx, (uint32_t)(t / LSx) & 1u);
        if (rec) { w_blk += c1 - c0; w_full += clock64() - c1; }
        tc_fence_after();
        const uint32_t sr = smem_u32(raw + (size_t)(t % RSx) * STAGE_BYTES), sl = smem_u32(lo + (size_t)(t % LSx) * STAGE_BYTES);
        uint64_t da_hi, db_hi, da_lo, db_lo, adv_a, adv_b;  // adv: start-address step (in 16 B) of a k step of 8
        if (TMA) {
          mbar_wait(raw_full + t % RSx, (uint32_t)(t / RSx) & 1u);  // (already complete: the splitters waited for it)
          if (MODE == 0) { da_hi = desc_sw(sr, 2048, 512, 1); da_lo = desc_sw(sl, 2048, 512, 1); adv_a = 1024 >> 4; }
          else { da_hi = desc_sw(sr, 16, 512, 4); da_lo = desc_sw(sl, 16, 512, 4); adv_a = 32 >> 4; }
          db_hi = desc_sw(sr + A_BYTES, 16, 512, 4); db_lo = desc_sw(sl + A_BYTES, 16, 512, 4); adv_b = 32 >> 4;
        } else {
          da_hi = desc16(sr); db_hi = desc16(sr + A_BYTES); da_lo = desc16(sl); db_lo = desc16(sl + A_BYTES);
          adv_a = adv_b = 16;  // two core matrices = 256 B
        }
#pragma unroll
        for (int ks = 0; ks < KS / 8; ++ks) {
          const uint64_t aa = adv_a * ks, ab = adv_b * ks;
          vmtc::mma_tf32(t_work, da_lo + aa, db_hi + ab, idesc, (t % BLK == 0 && ks == 0) ? 0u : 1u);
          vmtc::mma_tf32(t_work, da_hi + aa, db_lo + ab, idesc, 1u);
          vmtc::mma_tf32(t_work, da_hi + aa, db_hi + ab, idesc, 1u);
        }
        umma_commit(raw_empty + t % RSx);
        umma_commit(lo_empty + t % LSx);
        if (t % BLK == BLK - 1 || t == n_stages - 1) umma_commit(blk_full);
      }
      if (rec) {
        atomicAdd((unsigned long long*)g.dbg + 10, (unsigned long long)w_full);
        atomicAdd((unsigned long long*)g.dbg + 11, (unsigned long long)w_blk);
        atomicAdd((unsigned long long*)g.dbg + 12, (unsigned long long)(clock64() - t_start));
        atomicAdd((unsigned long long*)g.dbg + 13, (unsigned long long)n_stages);
      }
    }
  } else if (warp <= NSW + 4 && !skip_tile) {
    drain_epilogue<MODE, false>(g, warp, lane, n_blocks, n_stages, n_tile, n_valid, n0, m0, t_work, t_master, blk_full, blk_done);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster -- two neighbouring sample tiles (LINEAR) / input-feature tiles
// (LINEAR_DW) -- run ONE 256 x 256 x 8 MMA per step.  Each CTA loads and splits its own 128 A rows and only HALF of the B
// tile (the tensor core of each SM reads the other half from its peer), so per SM and stage the loads drop from 24 to
// 16 KB, the split from 48 to 32 KB of shared-memory traffic and the MMA operand reads from 72 to 48 KB: the one-CTA
// kernel above is bound by exactly that shared-memory bandwidth (6 MMAs of a stage alone need ~1060 cycles of it, the
// split and the fills the rest of the measured 1750).  The leader (cluster rank 0) issues the MMAs and owns the barriers
// the splitters / drain warps of both CTAs arrive on (full, blk_done); MMA completion is multicast to both CTAs' copies
// of raw_empty / lo_empty / blk_full.  Tensor loads only (there is no cp.async pair kernel).
constexpr int PRS = 7, PLS = 5;                      // raw / lo ring depths of the pair kernel (16 KB stages)
constexpr int PB_BYTES = (TN / 2) * KS * 4, PSTAGE = A_BYTES + PB_BYTES;
static_assert(1024 + (size_t)(PRS + PLS) * PSTAGE + 256 <= SMEM_BYTES, "pair rings exceed the allocation");

__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS + 32, 1)
vg_gemm3_pair(const __grid_constant__ Args g, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* raw = base;                                  // [PRS][A | B half]
  unsigned char* lo = base + (size_t)PRS * PSTAGE;            // [PLS][A | B half]
  uint64_t* bars = (uint64_t*)(lo + (size_t)PLS * PSTAGE);
  uint64_t* full = bars;                 // [PLS]  leader's copy: 8 arrivals (four warps of the owning set in each CTA)
  uint64_t* lo_empty = full + PLS;        // [PLS]  MMA commit, multicast
  uint64_t* raw_empty = lo_empty + PLS;   // [PRS] MMA commit, multicast
  uint64_t* blk_full = raw_empty + PRS;  // [1]   MMA commit, multicast
  uint64_t* blk_done = blk_full + 1;     // [1]   leader's copy: 8 arrivals
  uint64_t* raw_full = blk_done + 1;     // [PRS] this CTA's tensor loads have landed
  uint32_t* tmem_slot = (uint32_t*)(raw_full + PRS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

  const int n0 = blockIdx.y * TN;
  const int n_valid = min(TN, g.b_rows - n0);
  const int n_tile = (max(n_valid, 16) + 15) & ~15;      // MMA N; this CTA holds rows [rank * n_tile / 2, + n_tile / 2) of the B tile
  const int n_half = n_tile >> 1, b_row0 = n0 + (int)rank * n_half;
  const int m0 = blockIdx.x * TM;                        // first sample (LINEAR) / input feature (DW) of THIS CTA's half of M = 256
  int k_begin = 0, k_end = g.K;
  if (MODE == 0) {
    if (g.kr.enabled) {
      int l = 1 << 30, h = 0;
      for (int half = 0; half < 2; ++half) {
        const int tl = 2 * (int)blockIdx.y + half;
        if (tl < 8 && half * 128 < n_valid && g.kr.hi[tl] > g.kr.lo[tl]) { l = min(l, (int)g.kr.lo[tl]); h = max(h, (int)g.kr.hi[tl]); }
      }
      k_begin = h > 0 ? l * 16 : 0;
      k_end = h > 0 ? min(g.K, h * 16) : 0;
    }
  } else {
    k_begin = blockIdx.z * g.k_slab; k_end = min(g.K, k_begin + g.k_slab);
  }
  const int n_stages = k_end > k_begin ? (k_end - k_begin + KS - 1) / KS : 0;
  const int n_blocks = (n_stages + BLK - 1) / BLK;

  if (tid == 0) {
    for (int s = 0; s < PLS; ++s) { mbar_init(full + s, 8); mbar_init(lo_empty + s, 1); }
    for (int s = 0; s < PRS; ++s) { mbar_init(raw_empty + s, 1); mbar_init(raw_full + s, 1); }
    mbar_init(blk_full, 1);
    mbar_init(blk_done, 8);
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_work = tmem, t_master = tmem + TN;

  if (warp < 8) {
    // ------------------------------------------------------------------ splitters: 1024 pieces per stage, 8 per thread of the owning set
    const int set = warp >> 2, ts = tid & 127;
    constexpr int PA = A_BYTES / 16 / 128, PB = PB_BYTES / 16 / 128;  // 4, 4
    float bsum[PB];
#pragma unroll
    for (int e = 0; e < PB; ++e) bsum[e] = 0.0f;
    auto low = [](const float4& x) {
      float4 l;
      l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
      l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
      l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
      l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
      return l;
    };
    for (int t = set; t < n_stages; t += 2) {
      if (lane == 0) {
        mbar_wait(raw_full + t % PRS, (uint32_t)(t / PRS) & 1u);
        if (t >= PLS) mbar_wait(lo_empty + t % PLS, (uint32_t)(t / PLS - 1) & 1u);
      }
      __syncwarp();
      const unsigned char* sr = raw + (size_t)(t % PRS) * PSTAGE;
      unsigned char* sl = lo + (size_t)(t % PLS) * PSTAGE;
      float4 x[PA + PB];
#pragma unroll
      for (int e = 0; e < PA + PB; ++e)
        if (e < PA || (ts >> 2) + 32 * (e - PA) < n_half) x[e] = *reinterpret_cast<const float4*>(sr + 16 * (ts + 128 * e));
#pragma unroll
      for (int e = 0; e < PA + PB; ++e)
        if (e < PA || (ts >> 2) + 32 * (e - PA) < n_half) {
          *reinterpret_cast<float4*>(sl + 16 * (ts + 128 * e)) = low(x[e]);
          if (MODE == 1 && e >= PA) bsum[e - PA] += (x[e].x + x[e].y) + (x[e].z + x[e].w);
        }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(full + t % PLS);
    }
    if (MODE == 1 && blockIdx.x < 2 && g.gb_off >= 0) {  // bias gradient from the first pair of every column of tiles: each CTA its half of the g rows
#pragma unroll
      for (int e = 0; e < PB; ++e) {
        float a = bsum[e];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        const int r = (ts >> 2) + 32 * e;
        if ((ts & 3) == 0 && r < n_half && b_row0 + r < g.b_rows) atomicAdd(g.grad + g.gb_off + b_row0 + r, a);
      }
    }
  } else if (warp == 13) {
    if (lane == 0) {  // -------------------------------------------------------------------- tensor-load producer (each CTA its own operands)
      for (int t = 0; t < n_stages; ++t) {
        if (t >= PRS) mbar_wait(raw_empty + t % PRS, (uint32_t)(t / PRS - 1) & 1u);
        const int k0 = k_begin + t * KS;
        const uint32_t sa = smem_u32(raw + (size_t)(t % PRS) * PSTAGE), sb = sa + A_BYTES;
        vmtc::mbar_expect_tx(raw_full + t % PRS, PSTAGE);
        if (MODE == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_2d(sa + 2048u * c, &map_a, m0 + 32 * c, k0, raw_full + t % PRS);
        } else {
          tma_2d(sa, &map_a, k0, m0, raw_full + t % PRS);
        }
        tma_2d(sb, &map_b, k0, b_row0, raw_full + t % PRS);
      }
    }
  } else if (warp == 8) {
    if (lane == 0 && rank == 0) {  // ------------------------------------------------------- MMA issuer: the leader, for both SMs
      const bool rec = g.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      long long w_full = 0, w_blk = 0;
      const long long t_start = rec ? clock64() : 0;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (MODE == 0 ? (1u << 15) : 0u) | ((uint32_t)(n_tile >> 3) << 17) |
                             ((uint32_t)((2 * TM) >> 4) << 24);
      for (int t = 0; t < n_stages; ++t) {
        const int blk = t / BLK;
        long long c0 = rec ? clock64() : 0;
        if (t % BLK == 0 && blk > 0) mbar_wait_cluster(blk_done, (uint32_t)(blk - 1) & 1u);
        long long c1 = rec ? clock64() : 0;
        mbar_wait_cluster(full + t % PLS, (uint32_t)(t / PLS) & 1u);
        if (rec) { w_blk += c1 - c0; w_full += clock64() - c1; }
        tc_fence_after();
        const uint32_t sr = smem_u32(raw + (size_t)(t % PRS) * PSTAGE), sl = smem_u32(lo + (size_t)(t % PLS) * PSTAGE);
        uint64_t da_hi, da_lo, adv_a;
        if (MODE == 0) { da_hi = desc_sw(sr, 2048, 512, 1); da_lo = desc_sw(sl, 2048, 512, 1); adv_a = 1024 >> 4; }
        else { da_hi = desc_sw(sr, 16, 512, 4); da_lo = desc_sw(sl, 16, 512, 4); adv_a = 32 >> 4; }
        const uint64_t db_hi = desc_sw(sr + A_BYTES, 16, 512, 4), db_lo = desc_sw(sl + A_BYTES, 16, 512, 4);
#pragma unroll
        for (int ks = 0; ks < KS / 8; ++ks) {
          const uint64_t aa = adv_a * ks, ab = (uint64_t)(2 * ks);
          mma_tf32_pair(t_work, da_lo + aa, db_hi + ab, idesc, (t % BLK == 0 && ks == 0) ? 0u : 1u);
          mma_tf32_pair(t_work, da_hi + aa, db_lo + ab, idesc, 1u);
          mma_tf32_pair(t_work, da_hi + aa, db_hi + ab, idesc, 1u);
        }
        umma_commit_pair(raw_empty + t % PRS);
        umma_commit_pair(lo_empty + t % PLS);
        if (t % BLK == BLK - 1 || t == n_stages - 1) umma_commit_pair(blk_full);
      }
      if (rec) {
        atomicAdd((unsigned long long*)g.dbg + 10, (unsigned long long)w_full);
        atomicAdd((unsigned long long*)g.dbg + 11, (unsigned long long)w_blk);
        atomicAdd((unsigned long long*)g.dbg + 12, (unsigned long long)(clock64() - t_start));
        atomicAdd((unsigned long long*)g.dbg + 13, (unsigned long long)n_stages);
      }
    }
  } else if (warp <= 12) {
    drain_epilogue<MODE, true>(g, warp, lane, n_blocks, n_stages, n_tile, n_valid, n0, m0, t_work, t_master, blk_full, blk_done);
  }
  tc_fence_before();
  cluster_sync_all();   // nobody leaves (or frees tensor memory) while the peer may still read this CTA's shared memory / arrive here
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// ---- host: tensor maps of the TMA-fed variant (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda)
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmapEncodeFn tmap_encode_fn() {
  static TmapEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (TmapEncodeFn)f;
  }();
  return fn;
}
// float32 matrix [outer][inner] with row pitch `pitch` floats; false if TMA cannot describe it (alignment) -> caller uses cp.async
inline bool make_tmap(CUtensorMap* m, const float* base, long long inner, long long outer, long long pitch, int box_inner, int box_outer,
                      CUtensorMapSwizzle sw) {
  TmapEncodeFn enc = tmap_encode_fn();
  if (!enc || ((uintptr_t)base & 15u) || (pitch & 3) || inner <= 0 || outer <= 0) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer}, strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer}, es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the lo image of a LINEAR weight operand: same geometry as map_b on the lo copy of the staged parameters
inline bool make_blo_map(const Args& a, const float* par, const float* par_lo, CUtensorMap* m) {
  return make_tmap(m, par_lo + (a.b_base - par), a.K, a.b_rows, a.b_pitch, KS, TN, CU_TENSOR_MAP_SWIZZLE_64B);
}
// lo[i] = w[i] - trunc_tf32(w[i]) over the staged parameters, once per backward call
__global__ void vg_par_lo(const float* __restrict__ par, float* __restrict__ lo, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float w = par[i];
    lo[i] = w - __uint_as_float(__float_as_uint(w) & 0xffffe000u);
  }
}
// (The CTA-pair kernel is kept for the record and behind SURJ_B200_GEMM3_TMA bits 2 / 3: see vm_global.cu.)
// maps of a launch: MODE 0 (LINEAR): A = in rows [K][samples] read in boxes {32 samples, 16 k}; MODE 1 (LINEAR_DW): A = in rows
// [n_in][samples] in boxes {16 samples, 128 rows}.  B: rows [b_rows][K] in boxes {16 k, 256 rows}.  Out-of-bounds rows / k read as zeros.
// pair: the CTA-pair kernel, whose CTAs load half a B tile each (boxes of 128 rows).
inline bool make_maps(int mode, const Args& a, long long samples, bool pair, CUtensorMap* ma, CUtensorMap* mb) {
  const int bn = pair ? TN / 2 : TN;
  if (mode == 0)
    return make_tmap(ma, a.a_base, samples, a.a_rows, a.a_pitch, 32, KS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
           make_tmap(mb, a.b_base, a.K, a.b_rows, a.b_pitch, KS, bn, CU_TENSOR_MAP_SWIZZLE_64B);
  return make_tmap(ma, a.a_base, a.K, a.a_rows, a.a_pitch, KS, TM, CU_TENSOR_MAP_SWIZZLE_64B) &&
         make_tmap(mb, a.b_base, a.K, a.b_rows, a.b_pitch, KS, bn, CU_TENSOR_MAP_SWIZZLE_64B);
}

}  // namespace vmg

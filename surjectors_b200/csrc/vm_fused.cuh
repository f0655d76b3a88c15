// Fused conditioner + rational-quadratic-spline kernel of the per-op executor (sm_100a): ONE launch per coupling layer
// does what the reference does in masked_coupling.py:73-85 with a make_mlp conditioner (conditioners/mlp.py:24-29) and
// distrax.RationalQuadraticSpline (rq_masked_coupling.py:59-61):
//
//     h_1 = act(m W_0 + b_0), ..., theta = h_L W_L + b_L        [B, n_t, 3K+1]   (never leaves the SM)
//     z_d, ld_d = spline(theta[:, d, :])(y_d),  lp += sum_d ld_d
//
// and, in training mode, the spline's VJP instead (theta's cotangent goes to the row workspace for the weight-gradient
// and input-gradient products that follow; hidden pre / post activations are kept for them).
//
// A CTA owns a tile of 128 samples and is persistent over tiles.  Data flow per tile:
//   * the 128 x K activation operand (conditioner input, then h_1, h_2) is STATIONARY in shared memory as two fp16
//     planes (hi, lo) scaled by a power of two per sample row, in the K-major no-swizzle UMMA layout;
//   * the weights stream through a ring of shared-memory slots (cp.async.bulk + mbarrier), pre-split into fp16 hi / lo
//     planes scaled by a power of two per matrix and pre-tiled ONCE per call in exactly the order the ring consumes;
//   * one thread issues tcgen05.mma.kind::f16: per 16-wide k step lo*hi + hi*lo + hi*hi into an FP32 accumulator in
//     tensor memory -- a 22-bit-mantissa product ("fp16x3", tools/ubench/f16_probe.cu: max error 6e-7 of sum |terms| at
//     K = 256, the FP32 FMA chain is at 2e-7) at twice the TF32 rate and half its shared-memory bytes, which is what
//     lets the activation operand stay resident;
//   * theta of ONE transformed dim (3K+1 columns, padded to 16) is one accumulator; two accumulators alternate, so
//     the spline of dim d (8 epilogue warps in two groups, a thread per sample, theta read straight from tensor memory
//     in 16-column chunks) runs while the tensor core computes dim d + 1.
// Warp roles: 0 = weight producer, 1 = MMA issuer, 2..9 = epilogue (group g = (warp - 2) / 4 owns accumulator g).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "vm.cuh"
#include "vm_ops.cuh"
#include "vm_rqs.cuh"
#include "vm_tc.cuh"

namespace vmf {

using vmtc::fence_async_smem;
using vmtc::mbar_arrive;
using vmtc::mbar_expect_tx;
using vmtc::mbar_init;
using vmtc::mbar_wait;
using vmtc::smem_u32;
using vmtc::tc_fence_after;
using vmtc::tc_fence_before;
using vmtc::umma_commit;

constexpr int TM = 128;          // samples per tile = TMEM lanes
constexpr int KMAX = 256;        // widest activation operand (conditioner input / hidden layer)
constexpr int NMAX = 256;        // widest accumulator
constexpr int KS = 16;           // k per ring stage = k of one kind::f16 MMA
constexpr int SLOTS = 5;
constexpr int SLOT_BYTES = NMAX * KS * 2 * 2;   // hi + lo planes of a 256 x 16 stage
constexpr int A_PLANE = TM * KMAX * 2;          // bytes of one activation plane
constexpr int MAX_HIDDEN = 3;
constexpr int THREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr int BIAS_STRIDE = NMAX + 32;          // floats per staged bias slab (chunk reads run up to 16 past 3K+1)
constexpr size_t SMEM_BYTES = 1024 + 2 * A_PLANE + SLOTS * SLOT_BYTES + 8192;

struct Desc {
  int in_row, n_in;                         // conditioner input rows [n_in][C]
  int n_hidden;                             // 1..3 hidden layers
  int h_n[MAX_HIDDEN], h_boff[MAX_HIDDEN], h_act[MAX_HIDDEN], h_post[MAX_HIDDEN], h_pre[MAX_HIDDEN];  // post / pre: -1 = do not keep
  int k_pad[MAX_HIDDEN + 1];                // K of every matrix, padded to 16
  int n_t, P, N;                            // transformed dims, 3K+1, P padded to 16
  int b_off;                                // bias of the last layer (compact [n_t * P]) or -1
  int K_bins, inner;                        // spline bins, vm::INNER_RQS / INNER_RQS_INV
  float lo, hi;
  int ein_row, eout_row, tdims, skip_lp;    // event rows in / out, itab offset of the transformed dims
  int mode;                                 // 0: value + log-det; 1: VJP (theta cotangent -> th_row)
  int th_row, gout_row, gin_row;
  long long img_off[MAX_HIDDEN + 1];        // byte offset of every matrix's staged image
  int scale_slot[MAX_HIDDEN + 1];           // index of every matrix's max |w| (float bits) in the scale array
};

// ------------------------------------------------------------------------------------------ weight images
// max |w| per matrix (float bits, atomicMax on the unsigned pattern), over the [out][in] copy of the staged weights
struct ImgMat {
  int wt_off, ldt, n_rows, n_k;      // W^T[n_rows][ldt], n_k valid k per row
  int rows_per_block, n_blocks, N;   // block j covers rows [j * rows_per_block, +rows_per_block), padded to N rows in the image
  int k_pad, scale_slot;
  long long img_off;
};

constexpr int IMG_BATCH = 32;
struct ImgBatch { ImgMat m[IMG_BATCH]; };  // passed by value: no device-side table to allocate or upload

__global__ void __launch_bounds__(256)
vf_wmax(const float* __restrict__ par, const __grid_constant__ ImgBatch mats, unsigned* __restrict__ scales) {
  const ImgMat& m = mats.m[blockIdx.y];
  float mx = 0.0f;
  const long long total = (long long)m.n_rows * m.n_k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / m.n_k), k = (int)(i % m.n_k);
    mx = fmaxf(mx, fabsf(par[m.wt_off + (size_t)r * m.ldt + k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0f) atomicMax(scales + m.scale_slot, __float_as_uint(mx));
}

// exponent e with |x| = f * 2^e, f in [0.5, 1)  (0 / inf / nan -> 15: scale 1)
__device__ __forceinline__ int exp_of(float amax) {
  const unsigned u = __float_as_uint(amax) & 0x7fffffffu;
  if (u == 0u || u >= 0x7f800000u) return 15;
  int e = (int)(u >> 23) - 126;
  if ((u >> 23) == 0u) e = -126 - (__clz(u) - 9);  // subnormal
  return max(e, -100);
}
__device__ __forceinline__ float pow2f(int e) {  // 2^e for -126 <= e <= 127
  return __uint_as_float((unsigned)(e + 127) << 23);
}
__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// image of one matrix: [block][k stage][plane hi | lo][N x 16 halfs in the K-major no-swizzle layout (SBO 256 B, LBO 128 B)]
// one thread = one (block, row, k stage): 16 consecutive k of one output row
__global__ void __launch_bounds__(256)
vf_image(const float* __restrict__ par, const __grid_constant__ ImgBatch mats, const unsigned* __restrict__ scales,
         unsigned char* __restrict__ img) {
  const ImgMat& m = mats.m[blockIdx.y];
  const int nks = m.k_pad / KS;
  const long long total = (long long)m.n_blocks * m.N * nks;
  const float sc = pow2f(15 - exp_of(__uint_as_float(scales[m.scale_slot])));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ks = (int)(i % nks);
    const int n = (int)((i / nks) % m.N);
    const int blk = (int)(i / ((long long)nks * m.N));
    const int r = blk * m.rows_per_block + n;
    const bool row_ok = n < m.rows_per_block && r < m.n_rows;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float x[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = ks * KS + 2 * q + e;
        x[e] = (row_ok && k < m.n_k) ? par[m.wt_off + (size_t)r * m.ldt + k] * sc : 0.0f;
      }
      __half h0, l0, h1, l1;
      split_h(x[0], h0, l0);
      split_h(x[1], h1, l1);
      hi[q] = pack2(h0, h1);
      lo[q] = pack2(l0, l1);
    }
    unsigned char* stage = img + m.img_off + ((size_t)blk * nks + ks) * ((size_t)m.N * 64);
    unsigned char* ph = stage + (size_t)(n >> 3) * 256 + (size_t)(n & 7) * 16;
    unsigned char* pl = ph + (size_t)m.N * 32;
    *reinterpret_cast<uint4*>(ph) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(ph + 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    *reinterpret_cast<uint4*>(pl) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(pl + 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  }
}

// ------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc_f16(int n) {  // D = f32, A = B = f16, both K-major, M = 128
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- the spline of one transformed dim with theta in a tensor-memory accumulator (one thread = one sample = one TMEM
// lane; columns can only be addressed warp-uniformly, so everything is a sweep over 16-column chunks).
//   pass 1  theta = D * scale + bias written back in place (all 3K+1 columns); maxima of the width / height logits
//   pass 2  e = exp(theta - max) written back in place, normalisers            -> no later pass evaluates an exponential
//   walk    knots by a Kahan-compensated running sum of the bin sizes (a position error is amplified by 1 / bin size);
//           the search runs while "v < next knot" has not fired, the other side's prefix advances with it
//   slopes  predicated sweep for the two slope parameters of the bin
// The arithmetic is that of vm::rqs_locate (vm_rqs.cuh) except that exp is ex2.approx with the argument's rounding error
// folded back in (expf-grade, 6 instructions); value, log-det and the scalar part of the VJP are the shared
// vm::rqs_eval / vm::rqs_bwd_core.
struct Chunk { uint32_t r[16]; };
__device__ __forceinline__ void tm_issue(uint32_t taddr, Chunk& c) { tmem_ld16_issue(taddr, c.r); }
// the wait names the destination registers as read-write operands: nothing that uses them can be scheduled above it
__device__ __forceinline__ void tm_wait(Chunk& c) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(c.r[0]), "+r"(c.r[1]), "+r"(c.r[2]), "+r"(c.r[3]), "+r"(c.r[4]), "+r"(c.r[5]), "+r"(c.r[6]), "+r"(c.r[7]),
                 "+r"(c.r[8]), "+r"(c.r[9]), "+r"(c.r[10]), "+r"(c.r[11]), "+r"(c.r[12]), "+r"(c.r[13]), "+r"(c.r[14]), "+r"(c.r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tm_wait(Chunk& a, Chunk& b) {
  tm_wait(a);
  asm volatile(""
               : "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]), "+r"(b.r[6]), "+r"(b.r[7]),
                 "+r"(b.r[8]), "+r"(b.r[9]), "+r"(b.r[10]), "+r"(b.r[11]), "+r"(b.r[12]), "+r"(b.r[13]), "+r"(b.r[14]), "+r"(b.r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tm_store(uint32_t taddr, const Chunk& c) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
               "r"(c.r[0]), "r"(c.r[1]), "r"(c.r[2]), "r"(c.r[3]), "r"(c.r[4]), "r"(c.r[5]), "r"(c.r[6]), "r"(c.r[7]), "r"(c.r[8]),
               "r"(c.r[9]), "r"(c.r[10]), "r"(c.r[11]), "r"(c.r[12]), "r"(c.r[13]), "r"(c.r[14]), "r"(c.r[15])
               : "memory");
}
__device__ __forceinline__ void tm_store_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fast_exp(float x) {  // 2^t, t = x log2(e), times (1 + r ln 2), r = exact rounding error of t
  const float t = x * 1.4426950216293334961f;
  float r = fmaf(x, 1.4426950216293334961f, -t);
  r = fmaf(x, 1.925963033500011079e-08f, r);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return e * fmaf(r, 0.6931471805599453f, 1.0f);
}
#define CF(c, i) __uint_as_float((c).r[i])
#define CSET(c, i, x) (c).r[i] = __float_as_uint(x)

// activation of a hidden unit: swish / sigmoid through ex2.approx + rcp.approx (a few ulp on a quantity that only shapes
// hidden units -- the same choice as the small-network kernels, smallnet.cuh); everything else as vm::act_fwd
template <int ACT>
__device__ __forceinline__ float hidden_act(float x) {
  if (ACT == vm::ACT_SWISH || ACT == vm::ACT_SIGMOID) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return ACT == vm::ACT_SWISH ? x * r : r;
  }
  return vm::act_fwd(ACT, x);
}

struct SplineTm {
  uint32_t taddr;      // accumulator base + this warp's lane quarter
  float scale;         // 2^(row exponent + matrix exponent - 30)
  uint32_t bias;       // shared-memory address of the dim's bias slab (zeros when the layer has none)
  int K;
  float lo, hi;
};
__device__ __forceinline__ float lds1(uint32_t a) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}

// One SIDE of the spline (the K width logits or the K height logits) lives in REGISTERS from one tensor-memory read on
// (NCH chunks of 16; a round trip to tensor memory costs hundreds of cycles while the tensor core is writing the other
// accumulator): theta -> maximum -> exponentials in place -> normaliser -> walk.  FULL: K == 16 NCH (no masking).
// The side the input is searched on (heights for the spline inverse, widths for the forward map) goes first and yields
// the bin; the other side then takes its prefix up to that bin.  Knots are running sums of the bin sizes in 32-bit FIXED
// POINT (2^-s of the range with (hi - lo) 2^s < 2^30: resolution 7e-9 for [-3, 3], finer than a float32 ulp of the knot,
// exact adds -- the float path needs a Kahan-compensated sum for the same accuracy, four dependent adds per bin): one
// FFMA, one F2I and one IADD per bin, and every per-bin decision is an independent predicate on the running sum (the
// sum is monotone), so the walk has no serial select chain.
struct SideOut {
  float m, z;            // maximum of the logits, normaliser
  float p0, p1;          // knots either side of the bin
  float size;            // bin size straight from the softmax (no knot cancellation)
  float p_bin, p_cum;    // softmax mass of the bin and before it
};

template <int NCH, bool FULL>
__device__ __forceinline__ void side_load(const SplineTm& s, int col0, Chunk (&e)[NCH], SideOut& o) {
  const int K = s.K;
#pragma unroll
  for (int c = 0; c < NCH; ++c) tm_issue(s.taddr + (uint32_t)(col0 + 16 * c), e[c]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) tm_wait(e[c]);
  float m = -INFINITY;
  const uint32_t ba = s.bias + 4u * (uint32_t)col0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (FULL) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bv = lds4(ba + (uint32_t)(64 * c + 16 * q));
        const float t0 = fmaf(CF(e[c], 4 * q), s.scale, bv.x), t1 = fmaf(CF(e[c], 4 * q + 1), s.scale, bv.y);
        const float t2 = fmaf(CF(e[c], 4 * q + 2), s.scale, bv.z), t3 = fmaf(CF(e[c], 4 * q + 3), s.scale, bv.w);
        m = fmaxf(fmaxf(m, fmaxf(t0, t1)), fmaxf(t2, t3));
        CSET(e[c], 4 * q, t0); CSET(e[c], 4 * q + 1, t1); CSET(e[c], 4 * q + 2, t2); CSET(e[c], 4 * q + 3, t3);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (16 * c + i < K) {
          const float t = fmaf(CF(e[c], i), s.scale, lds1(ba + 4u * (uint32_t)(16 * c + i)));
          m = fmaxf(m, t);
          CSET(e[c], i, t);
        }
    }
  }
  float z0 = 0.0f, z1 = 0.0f;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (FULL || 16 * c + i < K) {
        const float x = fast_exp(CF(e[c], i) - m);
        z0 += x;
        CSET(e[c], i, x);
      }
      if (FULL || 16 * c + i + 1 < K) {
        const float x = fast_exp(CF(e[c], i + 1) - m);
        z1 += x;
        CSET(e[c], i + 1, x);
      }
    }
  o.m = m; o.z = z0 + z1;
}

// fixed-point scale of the knots: S = 2^s with (hi - lo) S in [2^29, 2^30)
__device__ __forceinline__ float knot_scale(float range) { return pow2f(30 - exp_of(range)); }

// SEARCH: bin = number of upper knots <= v (the last bin takes what is left); else: prefix up to the given bin
template <int NCH, bool FULL, bool SEARCH>
__device__ __forceinline__ int side_walk(const SplineTm& s, const Chunk (&e)[NCH], float vs, int bin_in, SideOut& o) {
  const int K = s.K;
  const float lo = s.lo, hi = s.hi;
  const float span = (hi - lo) - (float)K * VM_RQS_MIN_BIN, iz = 1.0f / o.z;
  const float S = knot_scale(hi - lo), iS = 1.0f / S;
  const float cS = span * iz * S, mS = VM_RQS_MIN_BIN * S;
  const int vI = SEARCH ? __float2int_rd(fminf(fmaxf((vs - lo) * S, 0.0f), 1073741824.0f)) : 0;
  int cs = 0, pk = 0, p1 = 0, bin = 0;
  float cum = 0.0f, eb = 0.0f;
  bool prev_below = true;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = 16 * c + i;
      if (FULL || k < K) {
        const float ek = CF(e[c], i);
        cs += __float2int_rn(fmaf(ek, cS, mS));
        const bool below = SEARCH ? (cs <= vI && k != K - 1) : (k < bin_in);
        const bool first = prev_below && !below;
        cum += below ? ek : 0.0f;
        pk = below ? cs : pk;
        eb = first ? ek : eb;
        p1 = first ? cs : p1;
        if (SEARCH) bin += below ? 1 : 0;
        prev_below = below;
      }
    }
  if (!SEARCH) bin = bin_in;
  const bool last = bin == K - 1;
  const float p = eb * iz;
  o.p0 = fmaf((float)pk, iS, lo);
  o.p1 = last ? hi : fmaf((float)p1, iS, lo);
  o.size = last ? hi - o.p0 : fmaf(p, span, VM_RQS_MIN_BIN);
  o.p_bin = p;
  o.p_cum = cum * iz;
  return bin;
}

template <int NCH>
__device__ __forceinline__ void side_store(const SplineTm& s, int col0, const Chunk (&e)[NCH]) {  // exp(theta - max) back in place
#pragma unroll
  for (int c = 0; c < NCH; ++c) tm_store(s.taddr + (uint32_t)(col0 + 16 * c), e[c]);
}

// KEEP: write the exponentials back to the accumulator for the VJP's sweep (a side's chunks may run past K into the
// next region, whose entries are stored back unchanged; each side is stored before the next one is loaded)
template <int NCH, bool FULL, bool KEEP>
__device__ __forceinline__ void tm_locate(const SplineTm& s, float v, bool forward, vm::RqsBin& b) {
  const int K = s.K;
  const float lo = s.lo, hi = s.hi;
  const bool below = v <= lo, above = !below && v >= hi;
  const float vs = (v == v) ? v : lo;  // NaN: first bin, like distrax
  SideOut s1, s2;                      // search side, other side
  int bin;
  {
    Chunk e[NCH];
    const int col0 = forward ? 0 : K;
    side_load<NCH, FULL>(s, col0, e, s1);
    bin = side_walk<NCH, FULL, true>(s, e, vs, 0, s1);
    if (KEEP) { side_store<NCH>(s, col0, e); tm_store_wait(); }
  }
  {
    Chunk e[NCH];
    const int col0 = forward ? K : 0;
    side_load<NCH, FULL>(s, col0, e, s2);
    side_walk<NCH, FULL, false>(s, e, vs, bin, s2);
    if (KEEP) { side_store<NCH>(s, col0, e); tm_store_wait(); }
  }
  const SideOut& sw = forward ? s1 : s2;
  const SideOut& sh = forward ? s2 : s1;
  b.mw = sw.m; b.mh = sh.m; b.zw = sw.z; b.zh = sh.z;
  if (below || above) {
    b.bin = below ? -1 : K;
    b.x0 = b.x1 = b.y0 = b.y1 = below ? lo : hi;
    b.bw = b.bh = 1.0f;
    b.pw_cum = b.pw_bin = b.ph_cum = b.ph_bin = 0.0f;
  } else {
    b.bin = bin;
    b.x0 = sw.p0; b.x1 = sw.p1; b.y0 = sh.p0; b.y1 = sh.p1; b.bw = sw.size; b.bh = sh.size;
    b.pw_cum = sw.p_cum; b.pw_bin = sw.p_bin; b.ph_cum = sh.p_cum; b.ph_bin = sh.p_bin;
  }
  // ---- the two slope parameters: predicated sweep over the K + 1 of them, the next chunk in flight behind the current
  const int s0 = below ? 0 : (above ? K : bin), s1i = (below || above) ? s0 : bin + 1;
  float u0 = 0.0f, u1 = 0.0f;
  const uint32_t ba = s.bias + 8u * (uint32_t)K;
  Chunk ua, ub;
  tm_issue(s.taddr + (uint32_t)(2 * K), ua);
  for (int k0 = 0; k0 <= K; k0 += 32) {
    tm_wait(ua);
    if (k0 + 16 <= K) tm_issue(s.taddr + (uint32_t)(2 * K + k0 + 16), ub);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (k0 + i == s0) u0 = fmaf(CF(ua, i), s.scale, lds1(ba + 4u * (uint32_t)(k0 + i)));
      if (k0 + i == s1i) u1 = fmaf(CF(ua, i), s.scale, lds1(ba + 4u * (uint32_t)(k0 + i)));
    }
    if (k0 + 16 <= K) {
      tm_wait(ub);
      if (k0 + 32 <= K) tm_issue(s.taddr + (uint32_t)(2 * K + k0 + 32), ua);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (k0 + 16 + i == s0) u0 = fmaf(CF(ub, i), s.scale, lds1(ba + 4u * (uint32_t)(k0 + 16 + i)));
        if (k0 + 16 + i == s1i) u1 = fmaf(CF(ub, i), s.scale, lds1(ba + 4u * (uint32_t)(k0 + 16 + i)));
      }
    }
  }
  b.u0 = u0; b.u1 = u1;
  b.d0 = vm::rqs_slope(u0);
  b.d1 = vm::rqs_slope(u1);
}

// value + log-det of one (sample, dim)
template <int NCH, bool FULL>
__device__ __forceinline__ void tm_fwd(const SplineTm& s, float v, bool forward, float& out, float& ld) {
  vm::RqsBin b;
  tm_locate<NCH, FULL, false>(s, v, forward, b);
  vm::rqs_eval(b, s.K, s.lo, s.hi, v, forward, out, ld);
}

// VJP: cotangent of theta -> rows [th_row + c][sample] (p = &rows[th_row * C + sample]); returns the cotangent of v.
// After tm_locate<KEEP> the accumulator holds exp(theta - max) for widths / heights: the sweep is a multiply per entry.
// release(): called as soon as the last tensor-memory read of this dim has completed -- the accumulator goes back to the
// MMA issuer while the second side's cotangents and the slope rows are still being computed and stored (the VJP holds an
// accumulator for its whole epilogue otherwise: per pair of dims the issuer idled for the length of one epilogue).
template <int NCH, bool FULL, class Release>
__device__ __forceinline__ float tm_bwd(const SplineTm& s, float* __restrict__ p, size_t C, float v, bool forward, float g_out, float g_ld,
                                        long long* tt, Release&& release) {
  const long long c0 = tt ? clock64() : 0;
  vm::RqsBin b;
  tm_locate<NCH, FULL, true>(s, v, forward, b);
  const long long c1 = tt ? clock64() : 0;
  vm::RqsGrad G;
  vm::rqs_bwd_core(b, s.K, s.lo, s.hi, v, forward, g_out, g_ld, G);
  const long long c2 = tt ? clock64() : 0;
  const int K = s.K;
  const bool tail = G.ks >= 0;
  const int kb = b.bin;
  const float span = (s.hi - s.lo) - (float)K * VM_RQS_MIN_BIN;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const float cz = tail ? 0.0f : span / (side ? b.zh : b.zw);
    const float A = side ? G.Ah : G.Aw, Bc = side ? G.Bh : G.Bw, dot = side ? G.doth : G.dotw;
    const float cA = cz * (A - dot), cB = cz * (Bc - dot), c0 = -cz * dot;  // coefficient of exp(theta_k - max) below / at / above the bin
    Chunk e[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) tm_issue(s.taddr + (uint32_t)(side * K + 16 * c), e[c]);
#pragma unroll
    for (int c = 0; c < NCH; ++c) tm_wait(e[c]);
    if (side == 1) release();
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = 16 * c + i;
        if (FULL || k < K) {
          p[(size_t)(side * K + k) * C] = CF(e[c], i) * (k < kb ? cA : (k == kb ? cB : c0));
        }
      }
  }
  // the K + 1 slope parameters: all but two (one in a tail) have zero cotangent -- zero the rows, then patch (same thread,
  // same addresses: program order)
  {
    float* q = p + (size_t)(2 * K) * C;
#pragma unroll 8
    for (int k = 0; k <= K; ++k) q[(size_t)k * C] = 0.0f;
    if (tail) {
      if (G.ks >= 2 * K && G.ks <= 3 * K) p[(size_t)G.ks * C] = G.g_tail;
    } else {
      q[(size_t)kb * C] = G.g_u0;
      q[(size_t)(kb + 1) * C] = G.g_u1;
    }
  }
  if (tt) { tt[0] += c1 - c0; tt[1] += c2 - c1; tt[2] += clock64() - c2; }
  return G.g_v;
}

// runs f(NCH, FULL) for the spline's bin count: K = 64 (BASELINE cfg4) gets unmasked code, every other K <= 64 the masked
// 4-chunk code, 64 < K <= 80 the masked 5-chunk code
template <class F>
__device__ __forceinline__ void with_chunks(int K, F&& f) {
  if (K == 64) f(std::integral_constant<int, 4>{}, std::true_type{});
  else if (K <= 64) f(std::integral_constant<int, 4>{}, std::false_type{});
  else f(std::integral_constant<int, 5>{}, std::false_type{});
}

// ------------------------------------------------------------------------------------------ the kernel
// rows: [n_rows][C] row workspace of the chunk; n_tiles = C / 128; n_valid = samples of the chunk that exist (the rest
// of the last tile computes on zeros and writes nothing that is read).
template <int MODE>  // 0: value + log-det, 1: VJP -- two kernels, so that each one's dim loop stays compact
__global__ void __launch_bounds__(THREADS, 1)
vg_fused_rqs(const __grid_constant__ Desc d, float* __restrict__ rows, int C, int n_tiles, const float* __restrict__ par, const int32_t* __restrict__ itab,
             const unsigned char* __restrict__ img, const unsigned* __restrict__ scales, long long* __restrict__ dbg) {
  // dbg (SURJ_B200_FUSED_DBG=1, else nullptr): CTA 0 adds the cycles its roles spend waiting -- [0] producer on a free ring
  // slot, [1] MMA issuer on weights, [2] on a free accumulator, [3] on the activation operand, [4] epilogue warp 2 on a
  // finished accumulator (dims), [5] the CTA's total cycles, [6] epilogue warp 2 in the hidden-layer phases
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* a_hi = base;
  unsigned char* a_lo = base + A_PLANE;
  unsigned char* ring = base + 2 * A_PLANE;
  unsigned char* misc = ring + SLOTS * SLOT_BYTES;
  uint64_t* full = (uint64_t*)misc;               // [SLOTS]
  uint64_t* empty = full + SLOTS;                 // [SLOTS]
  uint64_t* d_full = empty + SLOTS;               // [2]
  uint64_t* d_empty = d_full + 2;                 // [2]
  uint64_t* a_ready = d_empty + 2;                // [1]
  uint32_t* tmem_slot = (uint32_t*)(a_ready + 1);
  float* s_max = (float*)(misc + 256);            // [2][128]
  float* s_lp = s_max + 2 * TM;                   // [128]
  float* s_bias = s_lp + TM;                      // [2 groups][2][BIAS_STRIDE]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_jobs = d.n_hidden + d.n_t;          // accumulator jobs per tile
  if (tid == 0) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(d_full + b, 1); mbar_init(d_empty + b, 4); }
    mbar_init(a_ready, EPI_THREADS);
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
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool rec = dbg != nullptr && blockIdx.x == 0;
  const long long t_start = rec ? clock64() : 0;
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;
  long long tt[3] = {0, 0, 0};
#define TIMED(acc, stmt) do { if (rec) { const long long t_ = clock64(); stmt; acc += clock64() - t_; } else { stmt; } } while (0)

  if (warp == 0) {
    if (lane == 0) {  // ---------------------------------------------------------------- weight producer
      long long it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int job = 0; job < n_jobs; ++job) {
          const int l = job < d.n_hidden ? job : d.n_hidden;
          const int n = job < d.n_hidden ? d.h_n[job] : d.N;
          const int nks = d.k_pad[l] / KS;
          const uint32_t bytes = (uint32_t)n * 64u;
          const unsigned char* src = img + d.img_off[l] + (job < d.n_hidden ? 0 : (size_t)(job - d.n_hidden) * nks * bytes);
          for (int ks = 0; ks < nks; ++ks, ++it) {
            const int s = (int)(it % SLOTS);
            if (it >= SLOTS) TIMED(w0, mbar_wait(empty + s, (uint32_t)((it / SLOTS) - 1) & 1u));
            mbar_expect_tx(full + s, bytes);
            bulk_g2s(ring + (size_t)s * SLOT_BYTES, src + (size_t)ks * bytes, bytes, full + s);
          }
        }
      }
      if (rec) atomicAdd((unsigned long long*)dbg + 0, (unsigned long long)w0);
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------------------------------------------------------- MMA issuer
      long long it = 0, job_no = 0;
      uint32_t a_phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int job = 0; job < n_jobs; ++job, ++job_no) {
          const int l = job < d.n_hidden ? job : d.n_hidden;
          const int n = job < d.n_hidden ? d.h_n[job] : d.N;
          const int nks = d.k_pad[l] / KS;
          const int buf = (int)(job_no & 1);
          if (job <= d.n_hidden) {  // a new activation operand: conditioner input, h_1, ...
            TIMED(w2, mbar_wait(a_ready, a_phase));
            a_phase ^= 1u;
          }
          if (job_no >= 2) TIMED(w1, mbar_wait(d_empty + buf, (uint32_t)((job_no >> 1) - 1) & 1u));
          tc_fence_after();
          const uint32_t dcol = tmem + (uint32_t)(buf * NMAX);
          const uint32_t id = idesc_f16(n);
          for (int ks = 0; ks < nks; ++ks, ++it) {
            const int s = (int)(it % SLOTS);
            TIMED(w0, mbar_wait(full + s, (uint32_t)(it / SLOTS) & 1u));
            tc_fence_after();
            const uint32_t sb = smem_u32(ring + (size_t)s * SLOT_BYTES);
            const uint64_t db_hi = desc_of(sb, 128, 256), db_lo = desc_of(sb + (uint32_t)n * 32u, 128, 256);
            const uint64_t da_hi = desc_of(smem_u32(a_hi) + (uint32_t)ks * 256u, 128, 4096);
            const uint64_t da_lo = desc_of(smem_u32(a_lo) + (uint32_t)ks * 256u, 128, 4096);
            mma_f16(dcol, da_lo, db_hi, id, ks > 0 ? 1u : 0u);
            mma_f16(dcol, da_hi, db_lo, id, 1u);
            mma_f16(dcol, da_hi, db_hi, id, 1u);
            umma_commit(empty + s);
          }
          umma_commit(d_full + buf);
        }
      }
      if (rec) {
        atomicAdd((unsigned long long*)dbg + 1, (unsigned long long)w0);
        atomicAdd((unsigned long long*)dbg + 2, (unsigned long long)w1);
        atomicAdd((unsigned long long*)dbg + 3, (unsigned long long)w2);
      }
    }
  } else {  // ------------------------------------------------------------------------------ epilogue warps 2..9
    const int ew = warp - 2, grp = ew >> 2;
    const int quarter = warp & 3;                    // the TMEM lanes a warp may touch: 32 * (warp id % 4)
    const int r = quarter * 32 + lane;               // sample row of this thread inside the tile
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const bool forward_spline = d.inner == vm::INNER_RQS_INV;
    uint32_t full_phase[2] = {0u, 0u};
    long long job_no = 0;
    unsigned char* my_a_hi = a_hi + (size_t)(r >> 3) * 4096 + (size_t)(r & 7) * 16;
    unsigned char* my_a_lo = a_lo + (size_t)(r >> 3) * 4096 + (size_t)(r & 7) * 16;
    auto wscale = [&](int l) { return pow2f(exp_of(__uint_as_float(__ldg(scales + d.scale_slot[l]))) - 15); };

    for (int t = 0; t < my_tiles; ++t) {
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const size_t col = (size_t)tile * TM + r;      // this thread's sample column in the row workspace
      // ---- conditioner input -> activation operand (this group's half of the k range)
      const long long t_hid = rec ? clock64() : 0;
      int e_row;
      {
        const int k_lo = grp * (KMAX / 2), k_hi = min(d.k_pad[0], k_lo + KMAX / 2);
        float mx = 0.0f;
        for (int k = k_lo; k < min(k_hi, d.n_in); ++k) mx = fmaxf(mx, fabsf(rows[(size_t)(d.in_row + k) * C + col]));
        s_max[grp * TM + r] = mx;
        bar_sync(1, EPI_THREADS);
        mx = fmaxf(s_max[r], s_max[TM + r]);
        e_row = exp_of(mx);
        const float sc = pow2f(15 - e_row);
        for (int k8 = k_lo; k8 < k_hi; k8 += 8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = k8 + 2 * q;
            const float x0 = k < d.n_in ? rows[(size_t)(d.in_row + k) * C + col] * sc : 0.0f;
            const float x1 = k + 1 < d.n_in ? rows[(size_t)(d.in_row + k + 1) * C + col] * sc : 0.0f;
            __half h0, l0, h1, l1;
            split_h(x0, h0, l0);
            split_h(x1, h1, l1);
            hi[q] = pack2(h0, h1);
            lo[q] = pack2(l0, l1);
          }
          *reinterpret_cast<uint4*>(my_a_hi + (size_t)(k8 >> 3) * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(my_a_lo + (size_t)(k8 >> 3) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_async_smem();
        mbar_arrive(a_ready);
      }
      // ---- hidden layers: accumulator -> bias, activation -> next activation operand
      for (int l = 0; l < d.n_hidden; ++l, ++job_no) {
        const int buf = (int)(job_no & 1);
        const int n = d.h_n[l];
        mbar_wait(d_full + buf, full_phase[buf]);
        full_phase[buf] ^= 1u;
        tc_fence_after();
        const float out_scale = pow2f(e_row - 15) * wscale(l);
        const uint32_t dcol = tmem + lane_addr + (uint32_t)(buf * NMAX);
        const int n_lo = grp * (NMAX / 2), n_hi = min(n, n_lo + NMAX / 2);
        const int act = d.h_act[l], boff = d.h_boff[l];
        // sweep 1: the largest |pre-activation| of the row.  |act(x)| <= |x| (+ 1 for sigmoid / softplus), so it bounds the
        // row of activations and fixes the row's power-of-two scale without evaluating the activation twice.
        float mx = 0.0f;
        for (int c0 = n_lo; c0 < n_hi; c0 += 16) {
          uint32_t v[16];
          tmem_ld16_issue(dcol + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bv = boff >= 0 ? __ldg(reinterpret_cast<const float4*>(par + boff + c0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(fmaf(__uint_as_float(v[4 * q]), out_scale, bv.x)), fabsf(fmaf(__uint_as_float(v[4 * q + 1]), out_scale, bv.y))),
                                 fmaxf(fabsf(fmaf(__uint_as_float(v[4 * q + 2]), out_scale, bv.z)), fabsf(fmaf(__uint_as_float(v[4 * q + 3]), out_scale, bv.w)))));
          }
        }
        s_max[grp * TM + r] = mx;
        bar_sync(1, EPI_THREADS);
        mx = fmaxf(s_max[r], s_max[TM + r]) + ((act == vm::ACT_SIGMOID || act == vm::ACT_SOFTPLUS) ? 1.0f : 0.0f);
        e_row = exp_of(mx);
        const float sc = pow2f(15 - e_row);
        float* post = d.h_post[l] >= 0 ? rows + (size_t)d.h_post[l] * C + col : nullptr;
        float* pre = d.h_pre[l] >= 0 ? rows + (size_t)d.h_pre[l] * C + col : nullptr;
        vm::with_act(act, [&](auto tag) {  // sweep 2, one compact copy per activation
          constexpr int ACT = decltype(tag)::value;
          for (int c0 = n_lo; c0 < n_hi; c0 += 16) {
            uint32_t v[16];
            tmem_ld16_issue(dcol + (uint32_t)c0, v);
            tmem_ld_wait();
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bv = boff >= 0 ? __ldg(reinterpret_cast<const float4*>(par + boff + c0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
              float a[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = c0 + 4 * q + e;
                const float x = fmaf(__uint_as_float(v[4 * q + e]), out_scale, bb[e]);
                a[e] = hidden_act<ACT>(x);
                if (pre != nullptr) pre[(size_t)c * C] = x;
                if (post != nullptr) post[(size_t)c * C] = a[e];
              }
              __half h0, l0, h1, l1;
              split_h(a[0] * sc, h0, l0);
              split_h(a[1] * sc, h1, l1);
              hi[2 * q] = pack2(h0, h1);
              lo[2 * q] = pack2(l0, l1);
              split_h(a[2] * sc, h0, l0);
              split_h(a[3] * sc, h1, l1);
              hi[2 * q + 1] = pack2(h0, h1);
              lo[2 * q + 1] = pack2(l0, l1);
            }
            *reinterpret_cast<uint4*>(my_a_hi + (size_t)(c0 >> 3) * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(my_a_hi + (size_t)((c0 >> 3) + 1) * 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(my_a_lo + (size_t)(c0 >> 3) * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(my_a_lo + (size_t)((c0 >> 3) + 1) * 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
        });
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(a_ready);
        bar_sync(1, EPI_THREADS);                     // every epilogue warp has finished reading the accumulator
        if (grp == buf && lane == 0) mbar_arrive(d_empty + buf);
      }
      if (rec) w1 += clock64() - t_hid;
      // ---- one accumulator per transformed dim: the spline (or its VJP) straight from tensor memory
      const float th_scale = pow2f(e_row - 15) * wscale(d.n_hidden);
      float lp_acc = 0.0f;
      const float g_lp = MODE ? rows[(size_t)vm::ROW_GLP * C + col] : 0.0f;
      for (int j = 0; j < d.n_t; ++j, ++job_no) {
        const int buf = (int)(job_no & 1);
        if (buf != grp) {                             // the other group's accumulator: only keep the phase bookkeeping
          full_phase[buf] ^= 1u;
          continue;
        }
        const int par_slot = (int)((job_no >> 1) & 1);
        float* bias = s_bias + (size_t)(grp * 2 + par_slot) * BIAS_STRIDE;
        const int gt = tid - 64 - grp * 128;          // 0..127 inside the group
        const long long t_a = rec ? clock64() : 0;
        for (int c = gt; c < d.N + 16; c += 128) bias[c] = (d.b_off >= 0 && c < d.P) ? __ldg(par + d.b_off + (size_t)j * d.P + c) : 0.0f;
        bar_sync(2 + grp, 128);
        if (rec) w2 += clock64() - t_a;
        TIMED(w0, mbar_wait(d_full + buf, full_phase[buf]));
        full_phase[buf] ^= 1u;
        tc_fence_after();
        const long long t_c = rec ? clock64() : 0;
        const int dim = itab[d.tdims + j];
        const float v = rows[(size_t)(d.ein_row + dim) * C + col];
        if (rec && v == 123.456f) w2 += 1;  // (forces the load to complete inside the timed segment)
        const long long t_d = rec ? clock64() : 0;
        const SplineTm th{tmem + lane_addr + (uint32_t)(buf * NMAX), th_scale, smem_u32(bias), d.K_bins, d.lo, d.hi};
        if (MODE == 0) {
          float out = 0.0f, ld = 0.0f;
          with_chunks(d.K_bins, [&](auto nch, auto full) { tm_fwd<decltype(nch)::value, decltype(full)::value>(th, v, forward_spline, out, ld); });
          rows[(size_t)(d.eout_row + dim) * C + col] = out;
          if (!d.skip_lp) lp_acc += ld;
        } else {
          const float g_out = rows[(size_t)(d.gout_row + dim) * C + col];
          float* gp = rows + (size_t)(d.th_row + j * d.P) * C + col;
          float g_v = 0.0f;
          with_chunks(d.K_bins, [&](auto nch, auto full) {
            g_v = tm_bwd<decltype(nch)::value, decltype(full)::value>(th, gp, (size_t)C, v, forward_spline, g_out, g_lp, rec ? tt : nullptr, [&] {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(d_empty + buf);
            });
          });
          rows[(size_t)(d.gin_row + dim) * C + col] = g_v;
        }
        if (rec) { w3 += clock64() - t_d; w4 += t_d - t_c; }
        if (MODE == 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty + buf);
        }
      }
      // ---- log-det of the tile: both groups' halves -> the lp row
      if (MODE == 0) {
        if (grp == 1) s_lp[r] = lp_acc;
        bar_sync(1, EPI_THREADS);
        if (grp == 0 && !d.skip_lp) rows[(size_t)vm::ROW_LP * C + col] += lp_acc + s_lp[r];
      } else {
        bar_sync(1, EPI_THREADS);
      }
      // a group that skipped the last jobs has not observed their completion: the barrier above is passed only after
      // the group that ran the last job saw its accumulator complete, i.e. after every MMA of this tile has read the
      // activation operand -- it may now be overwritten by the next tile's input.
    }
  }
  if (rec && tid == 64) {
    atomicAdd((unsigned long long*)dbg + 4, (unsigned long long)w0);
    atomicAdd((unsigned long long*)dbg + 5, (unsigned long long)(clock64() - t_start));
    atomicAdd((unsigned long long*)dbg + 6, (unsigned long long)w1);
    atomicAdd((unsigned long long*)dbg + 7, (unsigned long long)w2);
    atomicAdd((unsigned long long*)dbg + 8, (unsigned long long)w3);
    atomicAdd((unsigned long long*)dbg + 9, (unsigned long long)w4);
    atomicAdd((unsigned long long*)dbg + 14, (unsigned long long)tt[0]);
    atomicAdd((unsigned long long*)dbg + 15, (unsigned long long)tt[1]);
    atomicAdd((unsigned long long*)dbg + 16, (unsigned long long)tt[2]);
  }
#undef TIMED
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace vmf

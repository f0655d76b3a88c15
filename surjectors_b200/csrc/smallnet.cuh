// Fused whole-chain log_prob kernels for small conditioner networks (sm_100a).
//
// One thread owns one sample.  Every width, parameter offset, coupling mask,
// MADE mask and permutation is a compile-time constant of the chain type, so
// activations live in registers, structurally-zero products (masked MADE
// weights, conditioner rows for transformed dims, conditioner columns for
// pass-through dims, triangular halves of L/U) are never issued, and the
// parameter blob is read from shared memory with warp-uniform (broadcast)
// LDS.128.  The whole Chain -- conditioners, transforms, log-det, decoder
// log-density, base log-prob -- is one pass over the batch.
//
// Reference semantics implemented here (file:line under /root/reference):
//   Chain fold                         surjectors/_src/surjectors/chain.py:25-45
//   MaskedCoupling inverse             surjectors/_src/bijectors/masked_coupling.py:73-85
//   MaskedAutoregressive inverse       surjectors/_src/bijectors/masked_autoregressive.py:73-79
//   MADE / MaskedLinear                surjectors/_src/conditioners/nn/made.py:69-91, masked_linear.py:35-58
//   make_mlp (hk.nets.MLP)             surjectors/_src/conditioners/mlp.py:7-29
//   LULinear                           surjectors/_src/bijectors/lu_linear.py:68-90
//   Permutation                        surjectors/_src/bijectors/permutation.py:30-37
//   Slice                              surjectors/_src/surjectors/slice.py:58-68
//   MaskedCouplingInferenceFunnel      surjectors/_src/surjectors/masked_coupling_inference_funnel.py:72-91
//   MaskedAutoregressiveInferenceFunnel surjectors/_src/surjectors/masked_autoregressive_inference_funnel.py:77-90
//   base log_prob + lc                 surjectors/_src/distributions/transformed_distribution.py:87-93
// The VJP follows SURVEY.md appendix C (the reference relies on JAX autodiff).
#pragma once
#include <cuda_runtime.h>

namespace surj {

typedef unsigned long long u64;

enum : int { ACT_IDENTITY = 0, ACT_RELU = 1, ACT_SWISH = 2, ACT_TANH = 3, ACT_SIGMOID = 4, ACT_SOFTPLUS = 5,
             ACT_GELU = 6, ACT_ELU = 7, ACT_LEAKY_RELU = 8 };
enum : int { INNER_AFFINE = 0, INNER_AFFINE_INV = 1, INNER_RQS = 2, INNER_RQS_INV = 3 };
enum : int { LAYOUT_SPLIT = 0, LAYOUT_INTERLEAVED = 1 };

#define SURJ_DEV __device__ __forceinline__
#define SURJ_HALF_LOG_2PI 0.91893853320467274178f

// ---------------------------------------------------------------- value types
// f2 is a 64-bit register pair worked on by PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2, which also
// take a broadcast scalar or an immediate as an operand; each lane is a correctly rounded FP32 operation and the pair
// issues at the full FP32 rate in ONE issue slot (tools/ubench/ffma2.cu).  Used two ways: the matvecs below pack two
// adjacent OUTPUTS of one sample (one broadcast activation x two adjacent weights), and the compact kernel
// (compact.cuh, packed.cuh) packs two SAMPLES per thread.
struct f2 {
  u64 v;
  f2() = default;
  SURJ_DEV f2(float s) { asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(s)); }
  SURJ_DEV f2(float lo, float hi) { asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); }
  SURJ_DEV float lo() const { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
  SURJ_DEV float hi() const { return __uint_as_float((unsigned)(v >> 32)); }
};
SURJ_DEV f2 operator+(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SURJ_DEV f2 operator-(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SURJ_DEV f2 operator*(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SURJ_DEV f2 operator-(f2 a) { return f2(0.0f) - a; }
SURJ_DEV f2 operator/(f2 a, f2 b) { return f2(a.lo() / b.lo(), a.hi() / b.hi()); }
SURJ_DEV f2& operator+=(f2& a, f2 b) { a = a + b; return a; }
using ::fmaf;  // keep the scalar overload visible next to the packed one
SURJ_DEV f2 fmaf(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}

SURJ_DEV float hsum(float a) { return a; }
SURJ_DEV float hsum(f2 a) { return a.lo() + a.hi(); }

// ------------------------------------------------------------------ scalar math
// exp feeds log-dets and Gaussian residuals directly: ex2.approx (2 ulp) with the argument's rounding error folded back
// in, i.e. expf-grade accuracy in 6 instructions.  The sigmoid of swish / sigmoid activations only shapes hidden units:
// ex2.approx + approximate reciprocal (a few ulp) -- measured to leave the log-prob error at the 1e-7 level while
// removing a third of the straight-line instruction stream.
SURJ_DEV float fast_exp(float x) {  // 2^t, t = x log2(e), times (1 + r ln 2) with r the exactly computed rounding error of t
  const float t = x * 1.4426950216293334961f;
  float r = fmaf(x, 1.4426950216293334961f, -t);
  r = fmaf(x, 1.925963033500011079e-08f, r);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return e * fmaf(r, 0.6931471805599453f, 1.0f);
}
SURJ_DEV float fast_sigmoid(float x) {  // 1 / (1 + 2^(-x log2 e)): FMUL, MUFU.EX2, FADD, MUFU.RCP (2^t = inf gives exactly 0)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
SURJ_DEV float softplus_f(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }

template <int ACT>
SURJ_DEV float act_fwd(float a) {
  if constexpr (ACT == ACT_IDENTITY) return a;
  else if constexpr (ACT == ACT_RELU) return fmaxf(a, 0.0f);
  else if constexpr (ACT == ACT_SWISH) return a * fast_sigmoid(a);
  else if constexpr (ACT == ACT_TANH) return tanhf(a);
  else if constexpr (ACT == ACT_SIGMOID) return fast_sigmoid(a);
  else if constexpr (ACT == ACT_SOFTPLUS) return softplus_f(a);
  else if constexpr (ACT == ACT_GELU) {
    const float u = 0.7978845608028654f * (a + 0.044715f * a * a * a);
    return 0.5f * a * (1.0f + tanhf(u));
  } else if constexpr (ACT == ACT_ELU) return a > 0.0f ? a : expm1f(a);
  else return a > 0.0f ? a : 0.01f * a;
}

// g * act'(a) given the pre-activation a
template <int ACT>
SURJ_DEV float act_bwd(float a, float g) {
  if constexpr (ACT == ACT_IDENTITY) return g;
  else if constexpr (ACT == ACT_RELU) return a > 0.0f ? g : 0.0f;
  else if constexpr (ACT == ACT_SWISH) {
    const float s = fast_sigmoid(a);
    return g * (s * (1.0f + a * (1.0f - s)));
  } else if constexpr (ACT == ACT_TANH) {
    const float t = tanhf(a);
    return g * (1.0f - t * t);
  } else if constexpr (ACT == ACT_SIGMOID) {
    const float s = fast_sigmoid(a);
    return g * s * (1.0f - s);
  } else if constexpr (ACT == ACT_SOFTPLUS) return g * fast_sigmoid(a);
  else if constexpr (ACT == ACT_GELU) {
    const float c = 0.7978845608028654f;
    const float u = c * (a + 0.044715f * a * a * a);
    const float t = tanhf(u);
    const float du = c * (1.0f + 3.0f * 0.044715f * a * a);
    return g * (0.5f * (1.0f + t) + 0.5f * a * (1.0f - t * t) * du);
  } else if constexpr (ACT == ACT_ELU) return a > 0.0f ? g : g * expf(a);
  else return a > 0.0f ? g : 0.01f * g;
}

// --------------------------------------------------------------- mask predicates
// on(i, j): is weight (row i = input, col j = output) structurally non-zero / needed?
struct PAll {
  static __host__ __device__ constexpr bool on(int, int) { return true; }
};
// rows / cols selected by 64-bit masks (bit set = active)
template <u64 ROWS, u64 COLS>
struct PRowCol {
  static __host__ __device__ constexpr bool on(int i, int j) { return ((ROWS >> i) & 1ull) && ((COLS >> j) & 1ull); }
};

// MADE masks in closed form (TFP _make_dense_autoregressive_masks, input_order='left-to-right',
// hidden_degrees='equal'; called by the reference at conditioners/nn/made.py:54-56):
//   input degree of y_i            = i + 1
//   degree of unit u of a hidden layer with U units and floor M = max(M, ceil((u+1)(D-1)/(U+1)))
//   hidden mask  [in, out]: deg_in(i) <= deg_out(j);  output mask: deg_in(i) < (j / P) + 1
template <int D>
struct DegInput {
  static __host__ __device__ constexpr int of(int i) { return i + 1; }
};
template <int D, int U, int M>
struct DegHidden {
  static __host__ __device__ constexpr int of(int u) {
    const int c = ((u + 1) * (D - 1) + U) / (U + 1);
    return c > M ? c : M;
  }
};
template <class DI, class DO>
struct PMadeHidden {
  static __host__ __device__ constexpr bool on(int i, int j) { return DI::of(i) <= DO::of(j); }
};
template <class DI, int P>
struct PMadeOut {
  static __host__ __device__ constexpr bool on(int i, int j) { return DI::of(i) < (j / P) + 1; }
};

// One Linear layer: out = act(in @ W + b).  W is [IN, OUT] row-major at float offset WOFF of
// the parameter blob, b at BOFF (BOFF < 0: no bias).  P prunes structural zeros.
template <int IN_, int OUT_, int ACT_, int WOFF_, int BOFF_, class P_ = PAll>
struct Lin {
  static constexpr int IN = IN_, OUT = OUT_, ACT = ACT_, WOFF = WOFF_, BOFF = BOFF_;
  using P = P_;
  static __host__ __device__ constexpr bool col_on(int j) {
    for (int i = 0; i < IN; ++i)
      if (P::on(i, j)) return true;
    return false;
  }
  static __host__ __device__ constexpr bool row_on(int i) {
    for (int j = 0; j < OUT; ++j)
      if (P::on(i, j)) return true;
    return false;
  }
};

// The three matvecs issue packed FFMA2 over PAIRS OF ADJACENT OUTPUTS (j, j + 1) of the thread's one sample: the
// activation is the broadcast operand, the two weights are adjacent in the row-major blob (they arrive as halves of the
// same LDS.128), the accumulator pair is two adjacent registers whose halves the scalar code before / after reads and
// writes directly.  The generated kernels are bound by instruction delivery, so halving the FMA
// instructions at an unchanged number of warps is worth ~20 % of the stream.  Pairs of which only one weight is
// structurally non-zero (P::on) fall back to a scalar FFMA on that half.

// pre[j] = b[j] + sum_i in[i] W[i][j]    (columns with no active row still get their bias:
// they are real outputs of the reference network unless the caller never reads them)
template <class L, bool BIAS>
SURJ_DEV void lin_mac(const float* __restrict__ sp, const float (&in)[L::IN], float (&pre)[L::OUT]) {
  constexpr int NPAIR = L::OUT / 2;
  f2 acc[NPAIR > 0 ? NPAIR : 1];
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) {
    if (BIAS) acc[p] = (L::BOFF >= 0) ? f2(sp[(L::BOFF >= 0 ? L::BOFF : 0) + 2 * p], sp[(L::BOFF >= 0 ? L::BOFF : 0) + 2 * p + 1]) : f2(0.0f);
    else acc[p] = f2(pre[2 * p], pre[2 * p + 1]);
  }
  float tail = 0.0f;  // last column of an odd-width layer
  if (L::OUT & 1) tail = BIAS ? ((L::BOFF >= 0) ? sp[(L::BOFF >= 0 ? L::BOFF : 0) + L::OUT - 1] : 0.0f) : pre[L::OUT - 1];
#pragma unroll
  for (int i = 0; i < L::IN; ++i) {
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
      const int o = L::WOFF + i * L::OUT + 2 * p;
      const bool lo_on = L::P::on(i, 2 * p), hi_on = L::P::on(i, 2 * p + 1);
      if (lo_on && hi_on) acc[p] = fmaf(f2(in[i]), f2(sp[o], sp[o + 1]), acc[p]);
      else if (lo_on) acc[p] = f2(fmaf(in[i], sp[o], acc[p].lo()), acc[p].hi());
      else if (hi_on) acc[p] = f2(acc[p].lo(), fmaf(in[i], sp[o + 1], acc[p].hi()));
    }
    if ((L::OUT & 1) && L::P::on(i, L::OUT - 1)) tail = fmaf(in[i], sp[L::WOFF + i * L::OUT + L::OUT - 1], tail);
  }
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) { pre[2 * p] = acc[p].lo(); pre[2 * p + 1] = acc[p].hi(); }
  if (L::OUT & 1) pre[L::OUT - 1] = tail;
}
template <class L>
SURJ_DEV void lin_pre(const float* __restrict__ sp, const float (&in)[L::IN], float (&pre)[L::OUT]) { lin_mac<L, true>(sp, in, pre); }
// accumulate only: pre[j] += sum_i in[i] W[i][j]
template <class L>
SURJ_DEV void lin_acc(const float* __restrict__ sp, const float (&in)[L::IN], float (&pre)[L::OUT]) { lin_mac<L, false>(sp, in, pre); }

// g_in[i] (+)= sum_j W[i][j] g_pre[j]
template <class L, bool ACCUM>
SURJ_DEV void lin_bwd_in(const float* __restrict__ sp, const float (&g_pre)[L::OUT], float (&g_in)[L::IN]) {
  constexpr int NPAIR = L::OUT / 2;
#pragma unroll
  for (int i = 0; i < L::IN; ++i) {
    float a = ACCUM ? g_in[i] : 0.0f;
    f2 a2 = f2(0.0f);
    bool any_pair = false;
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
      const int o = L::WOFF + i * L::OUT + 2 * p;
      const bool lo_on = L::P::on(i, 2 * p), hi_on = L::P::on(i, 2 * p + 1);
      if (lo_on && hi_on) { a2 = fmaf(f2(sp[o], sp[o + 1]), f2(g_pre[2 * p], g_pre[2 * p + 1]), a2); any_pair = true; }
      else if (lo_on) a = fmaf(sp[o], g_pre[2 * p], a);
      else if (hi_on) a = fmaf(sp[o + 1], g_pre[2 * p + 1], a);
    }
    if ((L::OUT & 1) && L::P::on(i, L::OUT - 1)) a = fmaf(sp[L::WOFF + i * L::OUT + L::OUT - 1], g_pre[L::OUT - 1], a);
    g_in[i] = any_pair ? a + hsum(a2) : a;
  }
}

// ------------------------------------------------------------------------- sinks
// Parameter-gradient accumulation.  dW[i][j] = sum over the block's samples of in[i]*g[j] is a
// contraction over the batch, so it is done block-cooperatively: every thread parks its in[] and
// g[] in shared memory ([feature][sample], padded), then the block runs a register-tiled outer
// product whose K dimension is the sample index and adds the result into the block's private
// gradient copy in shared memory.
struct NoSink {
  template <class L>
  SURJ_DEV void linear(const float (&)[L::IN], const float (&)[L::OUT]) {}
  template <int N, int LOFF, int UOFF, int DOFF>
  SURJ_DEV void lu(const float*, const float*, const float (&)[N], const float (&)[N], const float (&)[N],
                   const float (&)[N], float) {}
};

template <int NT>
struct BlockSink {
  static constexpr int LDT = NT + 4;  // row stride (floats): 16B-aligned rows, bank offset 4 per row
  float* sg;     // [n_params] block-private gradient accumulator (shared)
  float* tile;   // [(IN+1+OUT) * LDT] staging tile (shared)
  int tid;

  // generic: contraction of A[NA][NT] (rows 0..NA-1) with G[NG][NT] over the sample index;
  // result (a, g) is added to sg[off(a, g)].  Work item = a 4x4 tile of (a, g) pairs and one of
  // KS slices of the sample range; the KS lanes of a tile are adjacent in the warp, combine their
  // partial sums with shuffles, and lane 0 of the group -- the only thread of the block that ever
  // touches those addresses -- adds them into sg without atomics (so the result is deterministic).
  template <int NA, int NG, class Off>
  SURJ_DEV void contract(Off off) {
    constexpr int TA = 4, TG = 4;
    constexpr int NTA = (NA + TA - 1) / TA, NTG = (NG + TG - 1) / TG, NE = NTA * NTG;
    constexpr int KS_RAW = NE >= NT ? 1 : NT / NE;
    constexpr int KS = KS_RAW >= 8 ? 8 : KS_RAW >= 4 ? 4 : KS_RAW >= 2 ? 2 : 1;
    constexpr int KLEN = NT / KS;
    constexpr int ITEMS = NE * KS;
    constexpr int ROUNDS = (ITEMS + NT - 1) / NT;
    static_assert(KLEN % 4 == 0, "K slice must be a multiple of 4");
    const float* A = tile;
    const float* G = tile + NA * LDT;
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) {
      const int w = r * NT + tid;
      const bool live = w < ITEMS;  // ITEMS is a multiple of KS: a group of KS lanes is live or dead as a whole
      const int wc = live ? w : 0;
      const int e = wc / KS, ks = wc % KS;
      const int ta = e % NTA, tg = e / NTA;
      float acc[TA][TG];
#pragma unroll
      for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int g = 0; g < TG; ++g) acc[a][g] = 0.0f;
      int ra[TA], rg[TG];
#pragma unroll
      for (int a = 0; a < TA; ++a) ra[a] = min(ta + a * NTA, NA - 1);
#pragma unroll
      for (int g = 0; g < TG; ++g) rg[g] = min(tg + g * NTG, NG - 1);
      const int k0 = ks * KLEN;
#pragma unroll 2
      for (int k = k0; k < k0 + KLEN; k += 4) {
        float4 av[TA], gv[TG];
#pragma unroll
        for (int a = 0; a < TA; ++a) av[a] = *reinterpret_cast<const float4*>(A + ra[a] * LDT + k);
#pragma unroll
        for (int g = 0; g < TG; ++g) gv[g] = *reinterpret_cast<const float4*>(G + rg[g] * LDT + k);
#pragma unroll
        for (int a = 0; a < TA; ++a)
#pragma unroll
          for (int g = 0; g < TG; ++g) {
            acc[a][g] = fmaf(av[a].x, gv[g].x, acc[a][g]);
            acc[a][g] = fmaf(av[a].y, gv[g].y, acc[a][g]);
            acc[a][g] = fmaf(av[a].z, gv[g].z, acc[a][g]);
            acc[a][g] = fmaf(av[a].w, gv[g].w, acc[a][g]);
          }
      }
#pragma unroll
      for (int o = KS / 2; o > 0; o >>= 1)
#pragma unroll
        for (int a = 0; a < TA; ++a)
#pragma unroll
          for (int g = 0; g < TG; ++g) acc[a][g] += __shfl_xor_sync(0xffffffffu, acc[a][g], o);
      if (live && ks == 0) {
#pragma unroll
        for (int a = 0; a < TA; ++a)
#pragma unroll
          for (int g = 0; g < TG; ++g) {
            const int ia = ta + a * NTA, ig = tg + g * NTG;
            if (ia < NA && ig < NG) {
              const int o = off(ia, ig);
              if (o >= 0) sg[o] += acc[a][g];
            }
          }
      }
    }
  }

  // dW += in^T g ; db += sum g   (one extra A row of ones carries the bias)
  template <class L>
  SURJ_DEV void linear(const float (&in)[L::IN], const float (&g)[L::OUT]) {
    float* A = tile;
    float* G = tile + (L::IN + 1) * LDT;
#pragma unroll
    for (int i = 0; i < L::IN; ++i) A[i * LDT + tid] = L::row_on(i) ? in[i] : 0.0f;
    A[L::IN * LDT + tid] = 1.0f;
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) G[j * LDT + tid] = g[j];
    __syncthreads();
    contract<L::IN + 1, L::OUT>([](int i, int j) -> int {
      if (i == L::IN) return L::BOFF >= 0 ? L::BOFF + j : -1;
      return L::P::on(i, j) ? L::WOFF + i * L::OUT + j : -1;  // masked weights get exactly 0 grad
    });
    __syncthreads();
  }

  // LULinear parameter grads: gL = t^T g_z (strict lower), gU = y^T g_t
  // (strict upper), g d_raw = (y_i g_t_i + g_lc / u_i) * sigmoid(d_raw_i)
  template <int N, int LOFF, int UOFF, int DOFF>
  SURJ_DEV void lu(const float* __restrict__ sp, const float* scr, const float (&y)[N], const float (&t)[N],
                   const float (&g_z)[N], const float (&g_t)[N], float g_lc) {
    float* A = tile;
    float* G = tile + N * LDT;
#pragma unroll
    for (int i = 0; i < N; ++i) { A[i * LDT + tid] = t[i]; G[i * LDT + tid] = g_z[i]; }
    __syncthreads();
    contract<N, N>([](int r, int c) -> int { return r > c ? LOFF + r * (r - 1) / 2 + c : -1; });
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) { A[i * LDT + tid] = y[i]; G[i * LDT + tid] = g_t[i]; }
    __syncthreads();
    contract<N, N>([](int r, int c) -> int { return r < c ? UOFF + r * (2 * N - r - 1) / 2 + (c - r - 1) : -1; });
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i)
      A[i * LDT + tid] = (y[i] * g_t[i] + g_lc / scr[i]) * fast_sigmoid(sp[DOFF + i]);
    G[tid] = 1.0f;
    __syncthreads();
    contract<N, 1>([](int i, int) -> int { return DOFF + i; });
    __syncthreads();
  }
};

// --------------------------------------------------------------------------- MLP
template <class... Ls>
struct Mlp;
template <class L0, class... Ls>
struct Mlp<L0, Ls...> {
  using First = L0;
  using Rest = Mlp<Ls...>;
  static constexpr int N = 1 + sizeof...(Ls);
  static constexpr int IN = L0::IN;
  static constexpr int OUT = Rest::N == 0 ? L0::OUT : Rest::OUT;
  static constexpr int MAXW = (L0::IN + 1 + L0::OUT) > Rest::MAXW ? (L0::IN + 1 + L0::OUT) : Rest::MAXW;
};
template <>
struct Mlp<> {
  static constexpr int N = 0, IN = 0, OUT = 0, MAXW = 0;
};

template <class M>
SURJ_DEV void mlp_fwd(const float* __restrict__ sp, const float (&in)[M::IN], float (&out)[M::OUT]) {
  using L = typename M::First;
  float pre[L::OUT];
  lin_pre<L>(sp, in, pre);
  if constexpr (M::N == 1) {
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) out[j] = act_fwd<L::ACT>(pre[j]);
  } else {
    float h[L::OUT];
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) h[j] = act_fwd<L::ACT>(pre[j]);
    mlp_fwd<typename M::Rest>(sp, h, out);
  }
}

// Recompute-forward + backward through an MLP.  g_out is the cotangent of the MLP output;
// g_in receives the cotangent of the input (overwritten).  Parameter grads go to the sink.
template <class M, class Sink>
SURJ_DEV void mlp_bwd(const float* __restrict__ sp, const float (&in)[M::IN], const float (&g_out)[M::OUT],
                      float (&g_in)[M::IN], Sink& sink) {
  using L = typename M::First;
  float pre[L::OUT];
  lin_pre<L>(sp, in, pre);
  float g_pre[L::OUT];
  if constexpr (M::N == 1) {
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) g_pre[j] = act_bwd<L::ACT>(pre[j], g_out[j]);
  } else {
    float h[L::OUT], g_h[L::OUT];
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) h[j] = act_fwd<L::ACT>(pre[j]);
    mlp_bwd<typename M::Rest, Sink>(sp, h, g_out, g_h, sink);
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) g_pre[j] = act_bwd<L::ACT>(pre[j], g_h[j]);
  }
  sink.template linear<L>(in, g_pre);
  lin_bwd_in<L, false>(sp, g_pre, g_in);
}

// Forward through the MLP keeping every layer's pre-activations alive on the recursion stack, then
// `gfun(out, g_out)` turns the network output into its cotangent, then backward.  One forward
// instead of the two that mlp_fwd + mlp_bwd spend (shorter instruction stream: the generated
// kernels are instruction-fetch bound).
template <class M, class Sink, class G>
SURJ_DEV void mlp_fwd_bwd(const float* __restrict__ sp, const float (&in)[M::IN], G& gfun, float (&g_in)[M::IN],
                          Sink& sink) {
  using L = typename M::First;
  float pre[L::OUT];
  lin_pre<L>(sp, in, pre);
  float g_pre[L::OUT];
  if constexpr (M::N == 1) {
    float out[L::OUT], g_out[L::OUT];
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) out[j] = act_fwd<L::ACT>(pre[j]);
    gfun(out, g_out);
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) g_pre[j] = act_bwd<L::ACT>(pre[j], g_out[j]);
  } else {
    float h[L::OUT], g_h[L::OUT];
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) h[j] = act_fwd<L::ACT>(pre[j]);
    mlp_fwd_bwd<typename M::Rest, Sink, G>(sp, h, gfun, g_h, sink);
#pragma unroll
    for (int j = 0; j < L::OUT; ++j) g_pre[j] = act_bwd<L::ACT>(pre[j], g_h[j]);
  }
  sink.template linear<L>(in, g_pre);
  lin_bwd_in<L, false>(sp, g_pre, g_in);
}

// -------------------------------------------------------------------------- MADE
// First masked layer takes y (masked weights) plus an optional dense context Linear on c;
// the remaining masked layers are an Mlp whose Lin predicates carry the MADE masks.
// L0: Lin<D, H0, ACT, ...mask...>; CTX: Lin<DC, H0, IDENTITY, ...> or void; REST: Mlp<...> ending
// in the [H, D*P] output layer.
template <class L0_, class CTX_, class REST_>
struct Made {
  using L0 = L0_;
  using CTX = CTX_;
  using REST = REST_;
  static constexpr int D = L0::IN, H0 = L0::OUT, OUT = REST::OUT;
  static constexpr int MAXW0 = (L0::IN + 1 + L0::OUT);
  static constexpr int MAXW = MAXW0 > REST::MAXW ? MAXW0 : REST::MAXW;
};
struct NoCtx {
  static constexpr int IN = 0, OUT = 0;
};

template <class MD, int DC>
SURJ_DEV void made_fwd(const float* __restrict__ sp, const float (&y)[MD::D], const float (&c)[DC > 0 ? DC : 1],
                       float (&out)[MD::OUT]) {
  using L0 = typename MD::L0;
  float pre[L0::OUT];
  lin_pre<L0>(sp, y, pre);
  if constexpr (DC > 0) {
    using C = typename MD::CTX;
    float cp[C::OUT];
    lin_pre<C>(sp, c, cp);
#pragma unroll
    for (int j = 0; j < L0::OUT; ++j) pre[j] += cp[j];
  }
  float h[L0::OUT];
#pragma unroll
  for (int j = 0; j < L0::OUT; ++j) h[j] = act_fwd<L0::ACT>(pre[j]);
  mlp_fwd<typename MD::REST>(sp, h, out);
}

template <class MD, int DC, class Sink>
SURJ_DEV void made_bwd(const float* __restrict__ sp, const float (&y)[MD::D], const float (&c)[DC > 0 ? DC : 1],
                       const float (&g_out)[MD::OUT], float (&g_y)[MD::D], float (&g_c)[DC > 0 ? DC : 1], Sink& sink) {
  using L0 = typename MD::L0;
  float pre[L0::OUT];
  lin_pre<L0>(sp, y, pre);
  if constexpr (DC > 0) {
    using C = typename MD::CTX;
    float cp[C::OUT];
    lin_pre<C>(sp, c, cp);
#pragma unroll
    for (int j = 0; j < L0::OUT; ++j) pre[j] += cp[j];
  }
  float h[L0::OUT], g_h[L0::OUT], g_pre[L0::OUT];
#pragma unroll
  for (int j = 0; j < L0::OUT; ++j) h[j] = act_fwd<L0::ACT>(pre[j]);
  mlp_bwd<typename MD::REST, Sink>(sp, h, g_out, g_h, sink);
#pragma unroll
  for (int j = 0; j < L0::OUT; ++j) g_pre[j] = act_bwd<L0::ACT>(pre[j], g_h[j]);
  sink.template linear<L0>(y, g_pre);
  lin_bwd_in<L0, false>(sp, g_pre, g_y);
  if constexpr (DC > 0) {
    using C = typename MD::CTX;
    sink.template linear<C>(c, g_pre);
    lin_bwd_in<C, false>(sp, g_pre, g_c);
  }
}

template <class MD, int DC, class Sink, class G>
SURJ_DEV void made_fwd_bwd(const float* __restrict__ sp, const float (&y)[MD::D], const float (&c)[DC > 0 ? DC : 1],
                           G& gfun, float (&g_y)[MD::D], float (&g_c)[DC > 0 ? DC : 1], Sink& sink) {
  using L0 = typename MD::L0;
  float pre[L0::OUT];
  lin_pre<L0>(sp, y, pre);
  if constexpr (DC > 0) {
    using C = typename MD::CTX;
    float cp[C::OUT];
    lin_pre<C>(sp, c, cp);
#pragma unroll
    for (int j = 0; j < L0::OUT; ++j) pre[j] += cp[j];
  }
  float h[L0::OUT], g_h[L0::OUT], g_pre[L0::OUT];
#pragma unroll
  for (int j = 0; j < L0::OUT; ++j) h[j] = act_fwd<L0::ACT>(pre[j]);
  mlp_fwd_bwd<typename MD::REST, Sink, G>(sp, h, gfun, g_h, sink);
#pragma unroll
  for (int j = 0; j < L0::OUT; ++j) g_pre[j] = act_bwd<L0::ACT>(pre[j], g_h[j]);
  sink.template linear<L0>(y, g_pre);
  lin_bwd_in<L0, false>(sp, g_pre, g_y);
  if constexpr (DC > 0) {
    using C = typename MD::CTX;
    sink.template linear<C>(c, g_pre);
    lin_bwd_in<C, false>(sp, g_pre, g_c);
  }
}

// --------------------------------------------------------- affine inner bijector
// SURVEY appendix A.3 / C.  theta index of (shift, log_scale) for dim d among D dims:
template <int LAYOUT, int D>
struct ThetaIdx {
  static __host__ __device__ constexpr int shift(int d) { return LAYOUT == LAYOUT_SPLIT ? d : 2 * d; }
  static __host__ __device__ constexpr int lscale(int d) { return LAYOUT == LAYOUT_SPLIT ? D + d : 2 * d + 1; }
};

template <int INNER>
SURJ_DEV void affine_inv(float y, float m, float s, float& z, float& ld) {
  if constexpr (INNER == INNER_AFFINE) {  // ScalarAffine(m, exp(s)).inverse: z=(y-m)e^{-s}, ld=-s
    z = (y - m) * fast_exp(-s);
    ld = -s;
  } else {  // Inverse(ScalarAffine).inverse = forward: z = e^{s} y + m, ld = +s
    z = fmaf(fast_exp(s), y, m);
    ld = s;
  }
}
// cotangents: given g_z, g_ld -> g_y, g_m, g_s
template <int INNER>
SURJ_DEV void affine_inv_bwd(float y, float m, float s, float g_z, float g_ld, float& g_y, float& g_m, float& g_s) {
  if constexpr (INNER == INNER_AFFINE) {
    const float e = fast_exp(-s);
    const float z = (y - m) * e;
    g_y = g_z * e;
    g_m = -g_y;
    g_s = -g_z * z - g_ld;
  } else {
    const float e = fast_exp(s);
    g_y = g_z * e;
    g_m = g_z;
    g_s = g_z * e * y + g_ld;
  }
}

// Independent(Normal(mu, exp(ls))).log_prob(v) summed over NV dims; theta = [mu | ls] (split)
template <int NV>
SURJ_DEV float normal_lp(const float (&v)[NV], const float (&th)[2 * NV]) {
  float lp = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float r = (v[j] - th[j]) * fast_exp(-th[NV + j]);
    lp += -0.5f * r * r - th[NV + j] - SURJ_HALF_LOG_2PI;
  }
  return lp;
}
// g = cotangent of the log-prob.  g_th: cotangent of theta; g_v: cotangent of v
template <int NV>
SURJ_DEV void normal_lp_bwd(const float (&v)[NV], const float (&th)[2 * NV], float g, float (&g_th)[2 * NV],
                            float (&g_v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float e = fast_exp(-th[NV + j]);
    const float r = (v[j] - th[j]) * e;
    g_th[j] = g * r * e;
    g_th[NV + j] = g * (r * r - 1.0f);
    g_v[j] = -g * r * e;
  }
}

// ------------------------------------------------------------------------ layers
// Every layer type exposes:
//   DIN, DOUT, MAXW (max staging rows the sink needs), NSCR (floats of per-block scratch)
//   prologue(sp, scr, tid, nt)         block-wide precompute into scratch
//   inverse(sp, scr, in, x, out, lc)   z and likelihood contribution (lc is accumulated)
//   backward(sp, scr, in, x, g_out, g_lc, g_in, g_x, sink)
//       g_in is overwritten with the cotangent of `in`; g_x is accumulated.
// DX is the chain-level conditioning width; USEX says whether this layer consumes x.

template <int N>
struct XArr {
  static constexpr int n = N > 0 ? N : 1;
};

// MaskedCoupling with an affine inner bijector (masked_coupling.py:73-85).
// MASK bit d set = dimension d passes through and feeds the conditioner.
template <int D, int DX, bool USEX, u64 MASK, int INNER, int LAYOUT, class COND>
struct Coupling {
  static constexpr int DIN = D, DOUT = D, MAXW = COND::MAXW, NSCR = 0;
  static constexpr int CIN = D + (USEX ? DX : 0);
  static_assert(COND::IN == CIN, "conditioner input width");
  static_assert(COND::OUT == 2 * D, "conditioner output width");
  using TI = ThetaIdx<LAYOUT, D>;
  static __host__ __device__ constexpr bool pass(int d) { return (MASK >> d) & 1ull; }

  SURJ_DEV static void prologue(const float*, float*, int, int) {}

  SURJ_DEV static void cond_in(const float (&in)[D], const float (&x)[XArr<DX>::n], float (&m)[CIN]) {
#pragma unroll
    for (int d = 0; d < D; ++d) m[d] = pass(d) ? in[d] : 0.0f;
    if constexpr (USEX) {
#pragma unroll
      for (int k = 0; k < DX; ++k) m[D + k] = x[k];
    }
  }

  SURJ_DEV static void inverse(const float* __restrict__ sp, const float*, const float (&in)[D],
                               const float (&x)[XArr<DX>::n], float (&out)[D], float& lc) {
    float m[CIN], th[2 * D];
    cond_in(in, x, m);
    mlp_fwd<COND>(sp, m, th);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (pass(d)) out[d] = in[d];
      else {
        float ld;
        affine_inv<INNER>(in[d], th[TI::shift(d)], th[TI::lscale(d)], out[d], ld);
        lc += ld;
      }
    }
  }

  template <class Sink>
  SURJ_DEV static void backward(const float* __restrict__ sp, const float*, const float (&in)[D],
                                const float (&x)[XArr<DX>::n], const float (&g_out)[D], float g_lc, float (&g_in)[D],
                                float (&g_x)[XArr<DX>::n], Sink& sink) {
    float m[CIN], g_m[CIN];
    cond_in(in, x, m);
    auto gfun = [&](const float (&th)[2 * D], float (&g_th)[2 * D]) {
#pragma unroll
      for (int j = 0; j < 2 * D; ++j) g_th[j] = 0.0f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (pass(d)) g_in[d] = g_out[d];
        else
          affine_inv_bwd<INNER>(in[d], th[TI::shift(d)], th[TI::lscale(d)], g_out[d], g_lc, g_in[d],
                                g_th[TI::shift(d)], g_th[TI::lscale(d)]);
      }
    };
    mlp_fwd_bwd<COND, Sink>(sp, m, gfun, g_m, sink);
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (pass(d)) g_in[d] += g_m[d];
    if constexpr (USEX) {
#pragma unroll
      for (int k = 0; k < DX; ++k) g_x[k] += g_m[D + k];
    }
  }
};

// Decoder tail shared by Slice and the funnels: lc += Independent(Normal(dec([z, x]))).log_prob(v)
template <int K, int NV, int DX, bool USEX, class DEC>
struct DecoderTail {
  static constexpr int CIN = K + (USEX ? DX : 0);
  static_assert(DEC::IN == CIN, "decoder input width");
  static_assert(DEC::OUT == 2 * NV, "decoder output width");
  SURJ_DEV static void cat(const float (&z)[K], const float (&x)[XArr<DX>::n], float (&c)[CIN]) {
#pragma unroll
    for (int d = 0; d < K; ++d) c[d] = z[d];
    if constexpr (USEX) {
#pragma unroll
      for (int k = 0; k < DX; ++k) c[K + k] = x[k];
    }
  }
  SURJ_DEV static float lp(const float* __restrict__ sp, const float (&z)[K], const float (&x)[XArr<DX>::n],
                           const float (&v)[NV]) {
    float c[CIN], th[2 * NV];
    cat(z, x, c);
    mlp_fwd<DEC>(sp, c, th);
    return normal_lp<NV>(v, th);
  }
  // g_z += d lp / d z * g ; g_v = ... ; g_x += ...
  template <class Sink>
  SURJ_DEV static void bwd(const float* __restrict__ sp, const float (&z)[K], const float (&x)[XArr<DX>::n],
                           const float (&v)[NV], float g, float (&g_z)[K], float (&g_v)[NV],
                           float (&g_x)[XArr<DX>::n], Sink& sink) {
    float c[CIN], g_c[CIN];
    cat(z, x, c);
    auto gfun = [&](const float (&th)[2 * NV], float (&g_th)[2 * NV]) { normal_lp_bwd<NV>(v, th, g, g_th, g_v); };
    mlp_fwd_bwd<DEC, Sink>(sp, c, gfun, g_c, sink);
#pragma unroll
    for (int d = 0; d < K; ++d) g_z[d] += g_c[d];
    if constexpr (USEX) {
#pragma unroll
      for (int k = 0; k < DX; ++k) g_x[k] += g_c[K + k];
    }
  }
};

// Slice (slice.py:58-68)
template <int D, int K, int DX, bool USEX, class DEC>
struct SliceL {
  static constexpr int DIN = D, DOUT = K, MAXW = DEC::MAXW, NSCR = 0;
  using Tail = DecoderTail<K, D - K, DX, USEX, DEC>;
  SURJ_DEV static void prologue(const float*, float*, int, int) {}
  SURJ_DEV static void inverse(const float* __restrict__ sp, const float*, const float (&in)[D],
                               const float (&x)[XArr<DX>::n], float (&out)[K], float& lc) {
    float v[D - K];
#pragma unroll
    for (int d = 0; d < K; ++d) out[d] = in[d];
#pragma unroll
    for (int j = 0; j < D - K; ++j) v[j] = in[K + j];
    lc += Tail::lp(sp, out, x, v);
  }
  template <class Sink>
  SURJ_DEV static void backward(const float* __restrict__ sp, const float*, const float (&in)[D],
                                const float (&x)[XArr<DX>::n], const float (&g_out)[K], float g_lc, float (&g_in)[D],
                                float (&g_x)[XArr<DX>::n], Sink& sink) {
    float z[K], v[D - K], g_z[K], g_v[D - K];
#pragma unroll
    for (int d = 0; d < K; ++d) { z[d] = in[d]; g_z[d] = g_out[d]; }
#pragma unroll
    for (int j = 0; j < D - K; ++j) v[j] = in[K + j];
    Tail::bwd(sp, z, x, v, g_lc, g_z, g_v, g_x, sink);
#pragma unroll
    for (int d = 0; d < K; ++d) g_in[d] = g_z[d];
#pragma unroll
    for (int j = 0; j < D - K; ++j) g_in[K + j] = g_v[j];
  }
};

// MaskedCouplingInferenceFunnel with affine inner bijector
// (masked_coupling_inference_funnel.py:72-91): mask = arange(D) >= K, i.e. dims [0,K) are
// transformed conditioned on dims [K,D) (+x); then the decoder scores y[K:] given [z, x].
template <int D, int K, int DX, bool USEX, int INNER, int LAYOUT, class COND, class DEC>
struct CouplingFunnel {
  static constexpr u64 MASK = (D >= 64 ? ~0ull : ((1ull << D) - 1ull)) & ~((1ull << K) - 1ull);
  using Cpl = Coupling<D, DX, USEX, MASK, INNER, LAYOUT, COND>;
  using Tail = DecoderTail<K, D - K, DX, USEX, DEC>;
  static constexpr int DIN = D, DOUT = K, NSCR = 0;
  static constexpr int MAXW = COND::MAXW > DEC::MAXW ? COND::MAXW : DEC::MAXW;
  SURJ_DEV static void prologue(const float*, float*, int, int) {}
  SURJ_DEV static void inverse(const float* __restrict__ sp, const float* scr, const float (&in)[D],
                               const float (&x)[XArr<DX>::n], float (&out)[K], float& lc) {
    float faux[D], v[D - K];
    Cpl::inverse(sp, scr, in, x, faux, lc);
#pragma unroll
    for (int d = 0; d < K; ++d) out[d] = faux[d];
#pragma unroll
    for (int j = 0; j < D - K; ++j) v[j] = in[K + j];
    lc += Tail::lp(sp, out, x, v);
  }
  template <class Sink>
  SURJ_DEV static void backward(const float* __restrict__ sp, const float* scr, const float (&in)[D],
                                const float (&x)[XArr<DX>::n], const float (&g_out)[K], float g_lc, float (&g_in)[D],
                                float (&g_x)[XArr<DX>::n], Sink& sink) {
    using TI = typename Cpl::TI;
    float m[Cpl::CIN], g_m[Cpl::CIN], v[D - K], g_v[D - K];
    Cpl::cond_in(in, x, m);
#pragma unroll
    for (int j = 0; j < D - K; ++j) v[j] = in[K + j];
    // one conditioner forward: theta -> z -> decoder forward + backward -> cotangent of theta
    auto gfun = [&](const float (&th)[2 * D], float (&g_th)[2 * D]) {
      float z[K], g_z[K];
#pragma unroll
      for (int d = 0; d < K; ++d) {
        float ld;
        affine_inv<INNER>(in[d], th[TI::shift(d)], th[TI::lscale(d)], z[d], ld);
        g_z[d] = g_out[d];
      }
      Tail::bwd(sp, z, x, v, g_lc, g_z, g_v, g_x, sink);
#pragma unroll
      for (int j = 0; j < 2 * D; ++j) g_th[j] = 0.0f;
#pragma unroll
      for (int d = 0; d < K; ++d)
        affine_inv_bwd<INNER>(in[d], th[TI::shift(d)], th[TI::lscale(d)], g_z[d], g_lc, g_in[d], g_th[TI::shift(d)],
                              g_th[TI::lscale(d)]);
    };
    mlp_fwd_bwd<COND, Sink>(sp, m, gfun, g_m, sink);
#pragma unroll
    for (int j = 0; j < D - K; ++j) g_in[K + j] = g_m[K + j] + g_v[j];  // dims >= K pass through and feed the conditioner
    if constexpr (USEX) {
#pragma unroll
      for (int k = 0; k < DX; ++k) g_x[k] += g_m[D + k];
    }
  }
};

// MaskedAutoregressive with affine inner bijector, MADE output interleaved [D, 2]
// (masked_autoregressive.py:73-79).  DC = context width (0 = unconditional).
template <int D, int DX, bool USEX, int INNER, class MADE_>
struct Maf {
  static constexpr int DIN = D, DOUT = D, MAXW = MADE_::MAXW, NSCR = 0;
  static constexpr int DC = USEX ? DX : 0;
  static_assert(MADE_::OUT == 2 * D, "MADE output width");
  SURJ_DEV static void prologue(const float*, float*, int, int) {}
  SURJ_DEV static void ctx(const float (&x)[XArr<DX>::n], float (&c)[XArr<DC>::n]) {
    if constexpr (DC > 0) {
#pragma unroll
      for (int k = 0; k < DX; ++k) c[k] = x[k];
    } else c[0] = 0.0f;
  }
  SURJ_DEV static void inverse(const float* __restrict__ sp, const float*, const float (&in)[D],
                               const float (&x)[XArr<DX>::n], float (&out)[D], float& lc) {
    float c[XArr<DC>::n], th[2 * D];
    ctx(x, c);
    made_fwd<MADE_, DC>(sp, in, c, th);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float ld;
      affine_inv<INNER>(in[d], th[2 * d], th[2 * d + 1], out[d], ld);
      lc += ld;
    }
  }
  template <class Sink>
  SURJ_DEV static void backward(const float* __restrict__ sp, const float*, const float (&in)[D],
                                const float (&x)[XArr<DX>::n], const float (&g_out)[D], float g_lc, float (&g_in)[D],
                                float (&g_x)[XArr<DX>::n], Sink& sink) {
    float c[XArr<DC>::n], g_y[D], g_c[XArr<DC>::n];
    ctx(x, c);
    auto gfun = [&](const float (&th)[2 * D], float (&g_th)[2 * D]) {
#pragma unroll
      for (int d = 0; d < D; ++d)
        affine_inv_bwd<INNER>(in[d], th[2 * d], th[2 * d + 1], g_out[d], g_lc, g_in[d], g_th[2 * d], g_th[2 * d + 1]);
    };
    made_fwd_bwd<MADE_, DC, Sink>(sp, in, c, gfun, g_y, g_c, sink);
#pragma unroll
    for (int d = 0; d < D; ++d) g_in[d] += g_y[d];
    if constexpr (DC > 0) {
#pragma unroll
      for (int k = 0; k < DX; ++k) g_x[k] += g_c[k];
    }
  }
};

// MaskedAutoregressiveInferenceFunnel (masked_autoregressive_inference_funnel.py:77-90):
// MAF on y[:K] with context [y[K:], x], then decoder on [z, x] scoring y[K:].
template <int D, int K, int DX, bool USEX, int INNER, class MADE_, class DEC>
struct MafFunnel {
  static constexpr int DIN = D, DOUT = K, NSCR = 0;
  static constexpr int DC = (D - K) + (USEX ? DX : 0);
  static constexpr int MAXW = MADE_::MAXW > DEC::MAXW ? MADE_::MAXW : DEC::MAXW;
  static_assert(MADE_::OUT == 2 * K && MADE_::D == K, "MADE shape");
  using Tail = DecoderTail<K, D - K, DX, USEX, DEC>;
  SURJ_DEV static void prologue(const float*, float*, int, int) {}
  SURJ_DEV static void split(const float (&in)[D], const float (&x)[XArr<DX>::n], float (&yp)[K], float (&v)[D - K],
                             float (&c)[DC]) {
#pragma unroll
    for (int d = 0; d < K; ++d) yp[d] = in[d];
#pragma unroll
    for (int j = 0; j < D - K; ++j) { v[j] = in[K + j]; c[j] = in[K + j]; }
    if constexpr (USEX) {
#pragma unroll
      for (int k = 0; k < DX; ++k) c[D - K + k] = x[k];
    }
  }
  SURJ_DEV static void inverse(const float* __restrict__ sp, const float*, const float (&in)[D],
                               const float (&x)[XArr<DX>::n], float (&out)[K], float& lc) {
    float yp[K], v[D - K], c[DC], th[2 * K];
    split(in, x, yp, v, c);
    made_fwd<MADE_, DC>(sp, yp, c, th);
#pragma unroll
    for (int d = 0; d < K; ++d) {
      float ld;
      affine_inv<INNER>(yp[d], th[2 * d], th[2 * d + 1], out[d], ld);
      lc += ld;
    }
    lc += Tail::lp(sp, out, x, v);
  }
  template <class Sink>
  SURJ_DEV static void backward(const float* __restrict__ sp, const float*, const float (&in)[D],
                                const float (&x)[XArr<DX>::n], const float (&g_out)[K], float g_lc, float (&g_in)[D],
                                float (&g_x)[XArr<DX>::n], Sink& sink) {
    float yp[K], v[D - K], c[DC], g_v[D - K], g_yp[K], g_y2[K], g_c[DC];
    split(in, x, yp, v, c);
    // one MADE forward: its output feeds the transform, the decoder runs forward + backward in the
    // middle, and the cotangent of theta flows straight back
    auto gfun = [&](const float (&th)[2 * K], float (&g_th)[2 * K]) {
      float z[K], g_z[K];
#pragma unroll
      for (int d = 0; d < K; ++d) {
        float ld;
        affine_inv<INNER>(yp[d], th[2 * d], th[2 * d + 1], z[d], ld);
        g_z[d] = g_out[d];
      }
      Tail::bwd(sp, z, x, v, g_lc, g_z, g_v, g_x, sink);
#pragma unroll
      for (int d = 0; d < K; ++d)
        affine_inv_bwd<INNER>(yp[d], th[2 * d], th[2 * d + 1], g_z[d], g_lc, g_yp[d], g_th[2 * d], g_th[2 * d + 1]);
    };
    made_fwd_bwd<MADE_, DC, Sink>(sp, yp, c, gfun, g_y2, g_c, sink);
#pragma unroll
    for (int d = 0; d < K; ++d) g_in[d] = g_yp[d] + g_y2[d];
#pragma unroll
    for (int j = 0; j < D - K; ++j) g_in[K + j] = g_v[j] + g_c[j];
    if constexpr (USEX) {
#pragma unroll
      for (int k = 0; k < DX; ++k) g_x[k] += g_c[D - K + k];
    }
  }
};

// LULinear (lu_linear.py:68-90): z = (y @ U) @ L, U upper with diag softplus(d)+1e-4, L unit
// lower; lc = sum log diag (same for every sample).  Entries are stored in row-major
// tril_indices(n,-1) / triu_indices(n,1) order.  scratch: [0,N) = diag(U), [N] = lc.
template <int N, int DX, int LOFF, int UOFF, int DOFF>
struct LuL {
  static constexpr int DIN = N, DOUT = N, MAXW = 2 * N + 1, NSCR = N + 1;
  static __host__ __device__ constexpr int lidx(int r, int c) { return LOFF + r * (r - 1) / 2 + c; }  // r > c
  static __host__ __device__ constexpr int uidx(int r, int c) {                                       // r < c
    return UOFF + r * (2 * N - r - 1) / 2 + (c - r - 1);
  }
  SURJ_DEV static void prologue(const float* __restrict__ sp, float* scr, int tid, int nt) {
    for (int i = tid; i < N; i += nt) scr[i] = softplus_f(sp[DOFF + i]) + 1e-4f;
    __syncthreads();
    if (tid == 0) {
      float s = 0.0f;
      for (int i = 0; i < N; ++i) s += logf(scr[i]);
      scr[N] = s;
    }
  }
  SURJ_DEV static void yu(const float* __restrict__ sp, const float* scr, const float (&in)[N], float (&t)[N]) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
      float a = in[c] * scr[c];
#pragma unroll
      for (int r = 0; r < c; ++r) a = fmaf(in[r], sp[uidx(r, c)], a);
      t[c] = a;
    }
  }
  SURJ_DEV static void inverse(const float* __restrict__ sp, const float* scr, const float (&in)[N],
                               const float (&)[XArr<DX>::n], float (&out)[N], float& lc) {
    float t[N];
    yu(sp, scr, in, t);
#pragma unroll
    for (int c = 0; c < N; ++c) {
      float a = t[c];
#pragma unroll
      for (int r = c + 1; r < N; ++r) a = fmaf(t[r], sp[lidx(r, c)], a);
      out[c] = a;
    }
    lc += scr[N];
  }
  template <class Sink>
  SURJ_DEV static void backward(const float* __restrict__ sp, const float* scr, const float (&in)[N],
                                const float (&)[XArr<DX>::n], const float (&g_out)[N], float g_lc, float (&g_in)[N],
                                float (&)[XArr<DX>::n], Sink& sink) {
    float t[N], g_t[N];
    yu(sp, scr, in, t);
    // g_t = g_z L^T : g_t[r] = g_z[r] + sum_{c<r} L[r][c] g_z[c]
#pragma unroll
    for (int r = 0; r < N; ++r) {
      float a = g_out[r];
#pragma unroll
      for (int c = 0; c < r; ++c) a = fmaf(sp[lidx(r, c)], g_out[c], a);
      g_t[r] = a;
    }
    // g_y = g_t U^T : g_y[r] = u_rr g_t[r] + sum_{c>r} U[r][c] g_t[c]
#pragma unroll
    for (int r = 0; r < N; ++r) {
      float a = scr[r] * g_t[r];
#pragma unroll
      for (int c = r + 1; c < N; ++c) a = fmaf(sp[uidx(r, c)], g_t[c], a);
      g_in[r] = a;
    }
    sink.template lu<N, LOFF, UOFF, DOFF>(sp, scr, in, t, g_out, g_t, g_lc);
  }
};

// Permutation (permutation.py:30-37): z = y[..., inv] with inv = scatter(perm -> arange).
// INV... is that inverse index vector, spelled out as template arguments so every register
// index is a compile-time constant.
template <int DX, int... INV>
struct PermL {
  static constexpr int D = sizeof...(INV);
  static constexpr int DIN = D, DOUT = D, MAXW = 0, NSCR = 0;
  SURJ_DEV static void prologue(const float*, float*, int, int) {}
  template <int POS, int FIRST, int... REST>
  SURJ_DEV static void gather(const float (&in)[D], float (&out)[D]) {
    out[POS] = in[FIRST];
    if constexpr (sizeof...(REST) > 0) gather<POS + 1, REST...>(in, out);
  }
  template <int POS, int FIRST, int... REST>
  SURJ_DEV static void scatter(const float (&g_out)[D], float (&g_in)[D]) {
    g_in[FIRST] = g_out[POS];
    if constexpr (sizeof...(REST) > 0) scatter<POS + 1, REST...>(g_out, g_in);
  }
  SURJ_DEV static void inverse(const float*, const float*, const float (&in)[D], const float (&)[XArr<DX>::n],
                               float (&out)[D], float&) {
    gather<0, INV...>(in, out);
  }
  template <class Sink>
  SURJ_DEV static void backward(const float*, const float*, const float (&)[D], const float (&)[XArr<DX>::n],
                                const float (&g_out)[D], float, float (&g_in)[D], float (&)[XArr<DX>::n], Sink&) {
    scatter<0, INV...>(g_out, g_in);
  }
};

}  // namespace surj

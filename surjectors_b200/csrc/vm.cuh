// Fused whole-chain log_prob / VJP kernel for small conditioner networks (sm_100a), second
// generation: a compact interpreter instead of per-chain straight-line code.
//
// Why: the first generation unrolled every matvec of a chain into one instruction stream
// (10k+ SASS instructions, 167 KB for BASELINE cfg2).  ncu showed it fetch-bound: the stream
// does not fit the 32 KB L1.5 instruction cache and `stalled_no_instruction` was the top stall.
// Here the hot loops are a handful of rolled matvec / contraction loops (a few KB of code) driven
// by a per-chain *program* the host compiles at plan time, so one kernel serves every chain whose
// parameters fit in shared memory -- no code generation, no registry.
//
// Data layout inside a CTA (NT threads, S samples per thread, T = NT * S samples per tile):
//   rows   [n_rows][T]   every per-sample quantity (event dims, hidden units, thetas, cotangents)
//                        is a "row"; a thread only ever touches its own S columns, so all traffic
//                        on rows is bank-conflict-free and needs no barrier -- except the weight-
//                        gradient contraction, which is CTA-cooperative over the sample axis.
//   par    [n_par]       parameters, *compacted and padded* by a gather table built at plan time:
//                        conditioner rows / columns that can never matter (transformed dims of a
//                        coupling input, pass-through dims of its output) are dropped, MADE masks
//                        are applied (w * mask, exactly the reference's masked_linear.py:49), L / U
//                        of LULinear are expanded to dense triangular matrices, every matrix has
//                        its leading dimension padded to a multiple of 8 and is also stored
//                        transposed for the input-gradient matvec.
//   grad   [n_prim]      (backward) the CTA's parameter-gradient accumulator, same indexing as the
//                        primary (non-transposed) part of `par`.
//
// Reference semantics implemented by the ops (file:line under /root/reference):
//   OP_LINEAR        hk.Linear / MaskedLinear        conditioners/mlp.py:24-29, nn/masked_linear.py:35-58
//   OP_AFFINE        distrax.ScalarAffine (+Inverse) bijectors/affine_masked_coupling.py:46-49,
//                                                    masked_coupling.py:78-84, masked_autoregressive.py:73-79
//   OP_RQS           distrax.RationalQuadraticSpline bijectors/rq_masked_coupling.py:59-61
//   OP_NORMAL_LP     Independent(Normal).log_prob    surjectors/slice.py:66-68, funnels :88-91
//   OP_BASE          base log_prob                   distributions/transformed_distribution.py:91-93
//   OP_GATHER        Permutation / concat            bijectors/permutation.py:30-37, masked_coupling.py:75-77
//   the *_BWD ops    SURVEY.md appendix C (the reference relies on jax.value_and_grad)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace vm {

enum : int32_t {
  OP_END = 0,
  OP_LOAD,          // a0 dst_row, a1 n, a2 src (0 y, 1 x)
  OP_STORE,         // a0 src_row, a1 n, a2 dst (0 z_out, 1 g_y, 2 g_x)
  OP_GATHER,        // a0 dst_row, a1 n, a2 itab : row[dst+k] = row[itab[k]]
  OP_SCATTER_ADD,   // a0 src_row, a1 n, a2 itab : row[itab[k]] += row[src+k]
  OP_COPY,          // a0 dst_row, a1 src_row, a2 n
  OP_ADD,           // a0 dst_row, a1 src_row, a2 n : dst += src
  OP_ZERO,          // a0 dst_row, a1 n
  OP_LINEAR,        // a0 in_row, a1 n_in, a2 w_off, a3 ld, a4 out_row, a5 n_out_pad, a6 b_off, a7 act,
                    // a8 accumulate, a9 pre_row (-1 none)
  OP_AFFINE,        // a0 in_row, a1 out_row, a2 n_t, a3 itab(dims), a4 shift_row, a5 shift_stride, a6 ls_row,
                    // a7 ls_stride, a8 inner, a9 skip_lp (1: do not add the log-det to lp)
  OP_RQS,           // a0 in_row, a1 out_row, a2 n_t, a3 itab(dims), a4 theta_row, a5 K, a6 inner, a7 skip_lp ; f0 lo, f1 hi
  OP_NORMAL_LP,     // a0 v_row, a1 n_v, a2 mu_row, a3 ls_row
  OP_ADD_SCALAR,    // a0 scalar slot
  OP_BASE,          // a0 z_row, a1 n
  OP_STORE_LP,
  OP_BASE_BWD,      // a0 z_row, a1 n, a2 g_row
  OP_NORMAL_LP_BWD, // a0 v_row, a1 n_v, a2 mu_row, a3 ls_row (overwritten by g_mu, g_ls), a4 gv_row
  OP_AFFINE_BWD,    // a0 in_row, a1 n_t, a2 itab, a3 shift_row, a4 shift_stride, a5 ls_row, a6 ls_stride, a7 inner,
                    // a8 gout_row, a9 gin_row  (thetas overwritten by their cotangents)
  OP_RQS_BWD,       // a0 in_row, a1 n_t, a2 itab, a3 theta_row, a4 K, a5 inner, a6 gout_row, a7 gin_row ; f0 lo, f1 hi
  OP_ACT_BWD,       // a0 g_row (in place), a1 pre_row, a2 n, a3 act
  OP_LINEAR_DW,     // a0 in_row, a1 n_in, a2 g_row, a3 n_out_pad, a4 w_off, a5 ld, a6 b_off
                    // (the input-gradient matvec g_in = g W^T is an OP_LINEAR on the transposed copy of W)
  OP_SCALAR_BWD,    // a0 scalar slot : slot += sum over the tile of g_lp
  OP_NORMAL_SAMPLE  // a0 out_row, a1 n_v, a2 mu_row, a3 ls_row, a4 eps_col : out = mu + exp(ls) * eps[:, eps_col + j];
                    // lp += log N(out; mu, exp(ls))   (decoder sample_and_log_prob, slice.py:74-76)
};

enum : int { ACT_IDENTITY = 0, ACT_RELU = 1, ACT_SWISH = 2, ACT_TANH = 3, ACT_SIGMOID = 4, ACT_SOFTPLUS = 5,
             ACT_GELU = 6, ACT_ELU = 7, ACT_LEAKY_RELU = 8 };
enum : int { INNER_AFFINE = 0, INNER_AFFINE_INV = 1, INNER_RQS = 2, INNER_RQS_INV = 3 };

struct Op {  // 64 bytes
  int32_t code;
  int32_t a[13];
  float f0, f1;
};

// stage codes of the parameter gather table
enum : int32_t { STAGE_ZERO = -1, STAGE_ONE = -2, STAGE_SOFTPLUS_BASE = -3 };  // <= -3: softplus(params[-(c+3)]) + 1e-4

struct ScalarJob {  // prologue: scalars[slot] = sum_i log(par[off + i * stride]), i < n   (LULinear log-det)
  int32_t slot, off, stride, n;
};

struct Program {            // device-resident (plan-owned) tables
  const Op* ops;            int32_t n_ops;                  // the program this launch runs (log_prob, its VJP, or sampling direction)
  const int32_t* stage;     int32_t n_par, n_prim;          // gather table [n_par]
  const int32_t* grad_dst;  const int32_t* grad_aux;        // [n_prim] write-back tables
  const int32_t* itab;      int32_t n_itab;
  const ScalarJob* sjobs;   int32_t n_sjobs, n_scalars;
  int32_t n_rows, d_in, d_x, d_out, n_params;
};

#define VM_DEV __device__ __forceinline__
#define VM_HALF_LOG_2PI 0.91893853320467274178f

VM_DEV float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
VM_DEV float softplus_f(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }

VM_DEV float act_fwd(int act, float a) {
  switch (act) {
    case ACT_IDENTITY: return a;
    case ACT_RELU: return fmaxf(a, 0.0f);
    case ACT_SWISH: return a * sigmoid_f(a);
    case ACT_TANH: return tanhf(a);
    case ACT_SIGMOID: return sigmoid_f(a);
    case ACT_SOFTPLUS: return softplus_f(a);
    case ACT_GELU: {
      const float u = 0.7978845608028654f * (a + 0.044715f * a * a * a);
      return 0.5f * a * (1.0f + tanhf(u));
    }
    case ACT_ELU: return a > 0.0f ? a : expm1f(a);
    default: return a > 0.0f ? a : 0.01f * a;
  }
}
VM_DEV float act_bwd(int act, float a, float g) {
  switch (act) {
    case ACT_IDENTITY: return g;
    case ACT_RELU: return a > 0.0f ? g : 0.0f;
    case ACT_SWISH: {
      const float s = sigmoid_f(a);
      return g * (s * (1.0f + a * (1.0f - s)));
    }
    case ACT_TANH: {
      const float t = tanhf(a);
      return g * (1.0f - t * t);
    }
    case ACT_SIGMOID: {
      const float s = sigmoid_f(a);
      return g * s * (1.0f - s);
    }
    case ACT_SOFTPLUS: return g * sigmoid_f(a);
    case ACT_GELU: {
      const float c = 0.7978845608028654f;
      const float u = c * (a + 0.044715f * a * a * a);
      const float t = tanhf(u);
      const float du = c * (1.0f + 3.0f * 0.044715f * a * a);
      return g * (0.5f * (1.0f + t) + 0.5f * a * (1.0f - t * t) * du);
    }
    case ACT_ELU: return a > 0.0f ? g : g * expf(a);
    default: return a > 0.0f ? g : 0.01f * g;
  }
}

// act_bwd with the sigmoid of swish / sigmoid / softplus units from ex2.approx + rcp.approx (relative error ~2^-21 on a factor
// of a cotangent; the gradient bound is 1e-4 of the tensor scale).  The per-op executor's ACT_BWD sweep is otherwise
// issue-bound, not HBM-bound: accurate expf + IEEE division are ~30 of its ~40 instructions per element.
VM_DEV float act_bwd_fast(int act, float a, float g) {
  if (act == ACT_SWISH || act == ACT_SIGMOID || act == ACT_SOFTPLUS) {
    float e, s;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f + e));
    if (act == ACT_SWISH) return g * (s * (1.0f + a * (1.0f - s)));
    if (act == ACT_SIGMOID) return g * s * (1.0f - s);
    return g * s;
  }
  return act_bwd(act, a, g);
}

// Runs f(std::integral_constant<int, ACT>{}) for the runtime activation code `act`.  Epilogues that apply an activation to
// an unrolled register tile call this ONCE around the loop: each activation gets its own compact copy of the loop,
// instead of the whole switch (tanhf, expm1f, ... inlined) being repeated per element -- that repetition made
// vg_linear_tc 285 KB of code, most of it executed once per CTA at the cold instruction-fetch rate.
template <class F>
VM_DEV void with_act(int act, F&& f) {
  switch (act) {
    case ACT_IDENTITY: f(std::integral_constant<int, ACT_IDENTITY>{}); break;
    case ACT_RELU: f(std::integral_constant<int, ACT_RELU>{}); break;
    case ACT_SWISH: f(std::integral_constant<int, ACT_SWISH>{}); break;
    case ACT_TANH: f(std::integral_constant<int, ACT_TANH>{}); break;
    case ACT_SIGMOID: f(std::integral_constant<int, ACT_SIGMOID>{}); break;
    case ACT_SOFTPLUS: f(std::integral_constant<int, ACT_SOFTPLUS>{}); break;
    case ACT_GELU: f(std::integral_constant<int, ACT_GELU>{}); break;
    case ACT_ELU: f(std::integral_constant<int, ACT_ELU>{}); break;
    default: f(std::integral_constant<int, ACT_LEAKY_RELU>{}); break;
  }
}

// parameter staging (gather table) and the inverse map for gradients
VM_DEV float stage_value(int32_t sidx, const float* __restrict__ params) {
  if (sidx >= 0) return __ldg(params + sidx);
  if (sidx == STAGE_ZERO) return 0.0f;
  if (sidx == STAGE_ONE) return 1.0f;
  return softplus_f(__ldg(params + (-(sidx + 3)))) + 1e-4f;  // lu_linear.py:76: softplus(diag) + 1e-4
}
// g = gradient w.r.t. the staged value `staged`; writes the gradient of the raw parameter it came from
VM_DEV void writeback_grad(int32_t dst, int32_t aux, float g, float staged, const float* scal_g,
                           const float* __restrict__ params, float* __restrict__ out) {
  if (dst >= 0) out[dst] = g;
  else if (dst <= STAGE_SOFTPLUS_BASE) {  // LULinear diagonal: u = softplus(raw) + 1e-4, lc = sum log u
    const int p = -(dst + 3);
    out[p] = (g + scal_g[aux] / staged) * sigmoid_f(__ldg(params + p));
  }
}

}  // namespace vm

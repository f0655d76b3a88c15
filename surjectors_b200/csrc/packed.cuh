// Packed two-sample arithmetic for sm_100a: a thread that owns TWO samples keeps every per-sample quantity as a
// 64-bit register pair and works on it with Blackwell's packed FP32 instructions.  Used by the compact fused kernel
// (compact.cuh).  tools/ubench/ffma2.cu measured FFMA2 at the full FP32 rate (73.9 TFLOP/s) with half the issue slots.
#pragma once
#include "smallnet.cuh"

namespace surj {

// per-lane map of a scalar function (transcendentals, selects: MUFU and FSEL have no packed form)
template <class F>
SURJ_DEV float vmap(float a, F f) { return f(a); }
template <class F>
SURJ_DEV f2 vmap(f2 a, F f) { return f2(f(a.lo()), f(a.hi())); }
template <class F>
SURJ_DEV float vmap(float a, float b, F f) { return f(a, b); }
template <class F>
SURJ_DEV f2 vmap(f2 a, f2 b, F f) { return f2(f(a.lo(), b.lo()), f(a.hi(), b.hi())); }


// both lanes of the approximate sigmoid used for hidden activations: 1 / (1 + 2^(-x log2 e)), multiplies packed
SURJ_DEV f2 fast_sigmoid(f2 x) {
  const f2 t = x * f2(-1.4426950408889634f);
  float e0, e1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t.lo()));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t.hi()));
  const f2 d = f2(e0, e1) + f2(1.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.lo()));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.hi()));
  return f2(r0, r1);
}

// two-sample forms: the sigmoid / tanh / select run per lane, the polynomial part of the common
// activations (swish, sigmoid, tanh) stays packed
template <int ACT>
SURJ_DEV f2 act_fwd(f2 a) {
  if constexpr (ACT == ACT_IDENTITY) return a;
  else if constexpr (ACT == ACT_SWISH) return a * fast_sigmoid(a);
  else return vmap(a, [](float x) { return act_fwd<ACT>(x); });
}
template <int ACT>
SURJ_DEV f2 act_bwd(f2 a, f2 g) {
  if constexpr (ACT == ACT_IDENTITY) return g;
  else if constexpr (ACT == ACT_SWISH) {
    const f2 s = fast_sigmoid(a);
    return g * (s * fmaf(a, f2(1.0f) - s, f2(1.0f)));
  } else if constexpr (ACT == ACT_SIGMOID) {
    const f2 s = fast_sigmoid(a);
    return g * s * (f2(1.0f) - s);
  } else return vmap(a, g, [](float x, float y) { return act_bwd<ACT>(x, y); });
}

// ---- a thread's two samples are rows s, s + 1 of the batch, `n` of them (0..2) inside it
SURJ_DEV void load_scalar(const float* __restrict__ p, long long s, int n, float dflt, f2& v) {
  v = f2(n > 0 ? __ldg(p + s) : dflt, n > 1 ? __ldg(p + s + 1) : dflt);
}
SURJ_DEV void store_scalar(float* __restrict__ p, long long s, int n, f2 v) {
  if (n > 1) *reinterpret_cast<float2*>(p + s) = make_float2(v.lo(), v.hi());  // s is even: 8-byte aligned
  else if (n > 0) p[s] = v.lo();
}
SURJ_DEV void load_elem(const float* __restrict__ p, long long s, int n, int stride, int d, f2& v) {
  v = f2(n > 0 ? __ldg(p + s * stride + d) : 0.0f, n > 1 ? __ldg(p + (s + 1) * stride + d) : 0.0f);
}
SURJ_DEV f2 keep_valid(f2 v, int n) { return f2(n > 0 ? v.lo() : 0.0f, n > 1 ? v.hi() : 0.0f); }

}  // namespace surj

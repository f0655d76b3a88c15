// distrax.RationalQuadraticSpline as scalar device code: value, log-det and the VJP.
//
// The reference reaches it through rq_masked_coupling.py:59-61 and
// rq_masked_coupling_inference_funnel.py:82-84 with distrax's defaults
// (boundary_slopes='unconstrained', min_bin_size = min_knot_slope = 1e-4); distrax itself is not
// vendored in the reference, its algorithm is restated in SURVEY.md appendix A.3 and oracle/distrax.py:
//   widths  w = softmax(uw) * (range - K * 1e-4) + 1e-4,  x-knots = lo + cumsum(w) (last forced to hi)
//   heights likewise from uh,  knot slopes d = softplus(us + log(e^{1 - 1e-4} - 1)) + 1e-4
//   bin k with pos[k] <= v < pos[k+1]; v <= lo / v >= hi: linear tails with the first / last slope
// log_prob through a plain spline bijector uses the spline INVERSE (quadratic solve); an
// Inverse-wrapped one uses the forward map.
//
// theta is read / its cotangent written through an accessor `th(k)` -> float&, k in [0, 3K+1), so
// the same code serves shared-memory rows (fused kernel) and global-memory rows (per-op executor).
// Nothing of size K is ever materialised: knots are accumulated while walking the bins.
#pragma once
#include "vm.cuh"

namespace vm {

#define VM_RQS_MIN_BIN 1e-4f
#define VM_RQS_MIN_SLOPE 1e-4f
#define VM_RQS_SLOPE_OFFSET 0.5411666523f  // log(exp(1 - 1e-4) - 1)

struct RqsBin {
  int bin;                        // selected bin, -1 = below range, K = above range
  float x0, x1, y0, y1, d0, d1;   // knots / slopes of the bin (tails: d0 = d1 = the boundary slope)
  float bw, bh;                   // width / height of the bin, taken from the softmax directly (no knot cancellation)
  float mw, mh, zw, zh;           // softmax max and normaliser of widths / heights
  float pw_cum, pw_bin, ph_cum, ph_bin;  // softmax mass before the bin and of the bin
  float u0, u1;                   // raw slope parameters behind d0 / d1 (tails: u0 = the boundary slope's)
};

VM_DEV float rqs_softplus(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }
VM_DEV float rqs_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
VM_DEV float rqs_slope(float u) { return rqs_softplus(u + VM_RQS_SLOPE_OFFSET) + VM_RQS_MIN_SLOPE; }

template <class Th>
VM_DEV void rqs_locate(const Th& th, int K, float lo, float hi, float v, bool forward, RqsBin& b) {
  float mw = -INFINITY, mh = -INFINITY;
  for (int k = 0; k < K; ++k) {
    mw = fmaxf(mw, th(k));
    mh = fmaxf(mh, th(K + k));
  }
  float zw = 0.0f, zh = 0.0f;
  for (int k = 0; k < K; ++k) {
    zw += expf(th(k) - mw);
    zh += expf(th(K + k) - mh);
  }
  b.mw = mw; b.mh = mh; b.zw = zw; b.zh = zh;
  b.pw_cum = b.pw_bin = b.ph_cum = b.ph_bin = 0.0f;
  if (v <= lo) {  // distrax: below_range = v <= pos[0]
    b.bin = -1;
    b.u0 = b.u1 = th(2 * K);
    b.d0 = b.d1 = rqs_slope(b.u0);
    b.x0 = b.x1 = b.y0 = b.y1 = lo;
    b.bw = b.bh = 1.0f;
    return;
  }
  if (v >= hi) {
    b.bin = K;
    b.u0 = b.u1 = th(3 * K);
    b.d0 = b.d1 = rqs_slope(b.u0);
    b.x0 = b.x1 = b.y0 = b.y1 = hi;
    b.bw = b.bh = 1.0f;
    return;
  }
  const float span = (hi - lo) - (float)K * VM_RQS_MIN_BIN;
  const float iw = 1.0f / zw, ih = 1.0f / zh;
  // knot positions = lo + running sum of the bin sizes; the sum is Kahan-compensated because a
  // position error is amplified by 1 / bin_width in (v - x_k) / bw
  float xk = lo, yk = lo, cw = 0.0f, ch = 0.0f, ew = 0.0f, eh = 0.0f, cpw = 0.0f, cph = 0.0f;
  int bin = -2;
  for (int k = 0; k < K; ++k) {
    const float pw = expf(th(k) - mw) * iw, ph = expf(th(K + k) - mh) * ih;
    const float wk = pw * span + VM_RQS_MIN_BIN, hk = ph * span + VM_RQS_MIN_BIN;
    {
      const float yv = wk - ew, t = cw + yv;
      ew = (t - cw) - yv;
      cw = t;
    }
    {
      const float yv = hk - eh, t = ch + yv;
      eh = (t - ch) - yv;
      ch = t;
    }
    const float xn = (k == K - 1) ? hi : lo + cw;
    const float yn = (k == K - 1) ? hi : lo + ch;
    const float p0 = forward ? xk : yk, p1 = forward ? xn : yn;
    if (bin == -2 && ((v >= p0 && v < p1) || (k == 0 && !(v == v)))) {  // NaN: first bin, like distrax
      bin = k;
      b.x0 = xk; b.x1 = xn; b.y0 = yk; b.y1 = yn;
      b.bw = (k == K - 1) ? hi - xk : wk;
      b.bh = (k == K - 1) ? hi - yk : hk;
      b.pw_cum = cpw; b.pw_bin = pw; b.ph_cum = cph; b.ph_bin = ph;
    }
    if (k == 0 && bin == -2) {  // remember bin 0 as the "no bin matched" fallback of distrax
      b.x0 = xk; b.x1 = xn; b.y0 = yk; b.y1 = yn;
      b.bw = (K == 1) ? hi - xk : wk;
      b.bh = (K == 1) ? hi - yk : hk;
      b.pw_bin = pw; b.ph_bin = ph;
    }
    xk = xn; yk = yn; cpw += pw; cph += ph;
  }
  if (bin == -2) bin = 0;
  b.bin = bin;
  b.u0 = th(2 * K + bin);
  b.u1 = th(2 * K + bin + 1);
  b.d0 = rqs_slope(b.u0);
  b.d1 = rqs_slope(b.u1);
}

// value and log-det of the located bin.  forward: y(x), ld = log dy/dx; else x(y), ld = -log dy/dx.
VM_DEV void rqs_eval(const RqsBin& b, int K, float lo, float hi, float v, bool forward, float& out, float& ld) {
  if (b.bin < 0 || b.bin >= K) {
    const float e = b.bin < 0 ? lo : hi;
    if (forward) { out = (v - e) * b.d0 + e; ld = logf(b.d0); }
    else { out = (v - e) / b.d0 + e; ld = -logf(b.d0); }
    return;
  }
  const float bw = b.bw, bh = b.bh;
  const float s = bh / bw;
  const float t = b.d1 + b.d0 - 2.0f * s;
  float z;
  if (forward) {
    z = fminf(fmaxf((v - b.x0) / bw, 0.0f), 1.0f);
  } else {
    const float w = fminf(fmaxf((v - b.y0) / bh, 0.0f), 1.0f);
    const float c = -s * w, bq = b.d0 - t * w, a = s - bq;
    z = -2.0f * c / (bq + sqrtf(bq * bq - 4.0f * a * c));
    z = fminf(fmaxf(z, 0.0f), 1.0f);
  }
  const float sq = z * z, z1 = z - sq, om = (1.0f - z) * (1.0f - z);
  const float den = s + t * z1;
  const float l = 2.0f * logf(s) + logf(b.d1 * sq + 2.0f * s * z1 + b.d0 * om) - 2.0f * logf(den);
  if (forward) {
    out = b.y0 + bh * (s * sq + b.d0 * z1) / den;
    ld = l;
  } else {
    out = bw * z + b.x0;
    ld = -l;
  }
}

// Scalar part of the VJP: cotangents of (out, ld) = (g_out, g_ld) -> g_v and the handful of coefficients from which
// the cotangent of every theta follows (rqs_bwd below sweeps them over an accessor; the fused conditioner + spline
// kernel sweeps them over tensor-memory chunks).
struct RqsGrad {
  float g_v;                           // cotangent of the transformed input
  int ks;                              // tails: index of the one theta with a non-zero cotangent (-1: inside the range)
  float g_tail;                        // tails: that cotangent
  float Aw, Bw, dotw, Ah, Bh, doth;    // widths / heights: g_theta[k] = span * p_k * ((k < bin ? A : k == bin ? B : 0) - dot)
  float g_u0, g_u1;                    // cotangents of the raw slope parameters of knots bin, bin + 1
};

VM_DEV void rqs_bwd_core(const RqsBin& b, int K, float lo, float hi, float v, bool forward, float g_out, float g_ld, RqsGrad& G) {
  float g_v, gx0 = 0.0f, gx1 = 0.0f, gy0 = 0.0f, gy1 = 0.0f, gd0 = 0.0f, gd1 = 0.0f;
  G.ks = -1;
  G.g_tail = G.Aw = G.Bw = G.dotw = G.Ah = G.Bh = G.doth = G.g_u0 = G.g_u1 = 0.0f;  // (a tail leaves the in-range coefficients unused: defined, so that a caller may multiply them by zero)
  if (b.bin < 0 || b.bin >= K) {
    const float e = b.bin < 0 ? lo : hi;
    float gd;
    if (forward) { g_v = g_out * b.d0; gd = g_out * (v - e) + g_ld / b.d0; }
    else { g_v = g_out / b.d0; gd = -g_out * (v - e) / (b.d0 * b.d0) - g_ld / b.d0; }
    G.ks = b.bin < 0 ? 2 * K : 3 * K;
    G.g_tail = gd * rqs_sigmoid(b.u0 + VM_RQS_SLOPE_OFFSET);
    G.g_v = g_v;
    return;
  }
  const float bw = b.bw, bh = b.bh;
  const float s = bh / bw;
  const float t = b.d1 + b.d0 - 2.0f * s;
  float gs, gt, gbw, gbh;
  if (forward) {
    const float zr = (v - b.x0) / bw;
    const bool clipped = !(zr > 0.0f && zr < 1.0f);
    const float z = fminf(fmaxf(zr, 0.0f), 1.0f);
    const float sq = z * z, z1 = z - sq, omz = 1.0f - z, om = omz * omz;
    const float nn = s * sq + b.d0 * z1, den = s + t * z1, num = b.d1 * sq + 2.0f * s * z1 + b.d0 * om;
    const float gl = g_ld;
    gs = 2.0f * gl / s;
    const float gnum = gl / num;
    float gden = -2.0f * gl / den;
    gy0 = g_out;
    gbh = g_out * nn / den;
    const float gnn = g_out * bh / den;
    gden -= g_out * bh * nn / (den * den);
    gs += gnn * sq;
    float gq1 = gnn * s;
    gd0 = gnn * z1;
    float gz1 = gnn * b.d0;
    gs += gden;
    gt = gden * z1;
    gz1 += gden * t;
    gd1 = gnum * sq;
    gq1 += gnum * b.d1;
    gs += gnum * 2.0f * z1;
    gz1 += gnum * 2.0f * s;
    gd0 += gnum * om;
    const float gom = gnum * b.d0;
    float gz = gq1 * 2.0f * z + gz1 * (1.0f - 2.0f * z) - gom * 2.0f * omz;
    if (clipped) gz = 0.0f;
    g_v = gz / bw;
    gx0 = -gz / bw;
    gbw = -gz * z / bw;
  } else {
    const float wr = (v - b.y0) / bh;
    const bool wclip = !(wr > 0.0f && wr < 1.0f);
    const float wq = fminf(fmaxf(wr, 0.0f), 1.0f);
    const float c = -s * wq, bq = b.d0 - t * wq, a = s - bq;
    const float r = sqrtf(bq * bq - 4.0f * a * c), den0 = bq + r;
    const float zr = -2.0f * c / den0;
    const bool zclip = !(zr > 0.0f && zr < 1.0f);
    const float z = fminf(fmaxf(zr, 0.0f), 1.0f);
    const float sq = z * z, z1 = z - sq, omz = 1.0f - z, om = omz * omz;
    const float den = s + t * z1, num = b.d1 * sq + 2.0f * s * z1 + b.d0 * om;
    const float gl = -g_ld;
    gs = 2.0f * gl / s;
    const float gnum = gl / num, gden = -2.0f * gl / den;
    gs += gden;
    gt = gden * z1;
    float gz1 = gden * t;
    gd1 = gnum * sq;
    gs += gnum * 2.0f * z1;
    gz1 += gnum * 2.0f * s;
    gd0 = gnum * om;
    const float gq1 = gnum * b.d1, gom = gnum * b.d0;
    float gz = gq1 * 2.0f * z + gz1 * (1.0f - 2.0f * z) - gom * 2.0f * omz;
    gx0 = g_out;
    gbw = g_out * z;
    gz += g_out * bw;
    if (zclip) gz = 0.0f;
    float gc = gz * (-2.0f / den0);
    const float gden0 = -gz * zr / den0;
    float gbq = gden0;
    const float gdisc = gden0 / (2.0f * r);
    gbq += gdisc * 2.0f * bq;
    const float ga = gdisc * (-4.0f * c);
    gc += gdisc * (-4.0f * a);
    gs += ga;
    gbq -= ga;
    gd0 += gbq;
    gt += gbq * (-wq);
    float gwq = gbq * (-t);
    gs += gc * (-wq);
    gwq += gc * (-s);
    if (wclip) gwq = 0.0f;
    g_v = gwq / bh;
    gy0 = -gwq / bh;
    gbh = -gwq * wq / bh;
  }
  gd0 += gt;
  gd1 += gt;
  gs -= 2.0f * gt;
  gbh += gs / bw;
  gbw -= gs * s / bw;
  gy1 = gbh;
  gy0 -= gbh;
  gx1 = gbw;
  gx0 -= gbw;
  // knots -> widths / heights -> softmax logits.  Knot k (1 <= k <= K-1) = lo + sum_{m<k} w_m;
  // knots 0 and K are the constants lo / hi.
  const int kb = b.bin;
  const float e1 = (kb + 1 <= K - 1) ? 1.0f : 0.0f;
  G.Aw = (kb >= 1 ? gx0 : 0.0f) + gx1 * e1; G.Bw = gx1 * e1;
  G.Ah = (kb >= 1 ? gy0 : 0.0f) + gy1 * e1; G.Bh = gy1 * e1;
  G.dotw = G.Aw * b.pw_cum + G.Bw * b.pw_bin; G.doth = G.Ah * b.ph_cum + G.Bh * b.ph_bin;
  G.g_u0 = gd0 * rqs_sigmoid(b.u0 + VM_RQS_SLOPE_OFFSET);
  G.g_u1 = gd1 * rqs_sigmoid(b.u1 + VM_RQS_SLOPE_OFFSET);
  G.g_v = g_v;
}

// Cotangents of (out, ld) = (g_out, g_ld) -> g_v and, written over theta in place, g_theta.
template <class Th>
VM_DEV float rqs_bwd(Th& th, int K, float lo, float hi, float v, bool forward, float g_out, float g_ld) {
  RqsBin b;
  rqs_locate(th, K, lo, hi, v, forward, b);
  RqsGrad G;
  rqs_bwd_core(b, K, lo, hi, v, forward, g_out, g_ld, G);
  if (G.ks >= 0) {
    for (int k = 0; k < 3 * K + 1; ++k) th(k) = 0.0f;
    th(G.ks) = G.g_tail;
    return G.g_v;
  }
  const int kb = b.bin;
  const float span = (hi - lo) - (float)K * VM_RQS_MIN_BIN;
  const float iw = 1.0f / b.zw, ih = 1.0f / b.zh;
  for (int k = 0; k < K; ++k) {
    const float pw = expf(th(k) - b.mw) * iw, ph = expf(th(K + k) - b.mh) * ih;
    const float sw = k < kb ? G.Aw : (k == kb ? G.Bw : 0.0f), sh = k < kb ? G.Ah : (k == kb ? G.Bh : 0.0f);
    th(k) = span * pw * (sw - G.dotw);
    th(K + k) = span * ph * (sh - G.doth);
  }
  for (int k = 0; k <= K; ++k) th(2 * K + k) = 0.0f;
  th(2 * K + kb) = G.g_u0;
  th(2 * K + kb + 1) = G.g_u1;
  return G.g_v;
}


}  // namespace vm

// Compact fused kernel for chains of affine MaskedCoupling / MaskedCouplingInferenceFunnel layers with small
// two-hidden-layer MLP conditioners (BASELINE cfg2 and its family), sm_100a.
//
// Why a third fused kernel.  The generated straight-line kernels (chain.cuh) execute ~10k instructions once per
// warp, and tools/ubench/icache.cu measured what that costs on a B200: code that does not loop inside ~32 KB is
// issued at <= 0.16 instructions/clk per warp and <= 2.3 per SM however many warps are resident (3.55 per SM for a
// 24-32 KB loop).  So this kernel is LOOPS: one body per layer kind, driven by a small layer table, about 3k
// instructions in total (instruction-cache hit rate 98 %), and it needs no build-time code generation.  What it
// pays for that is shared memory: ~106 rows x 448 samples leave room for one 224-thread CTA (7 warps) per SM, and
// at that occupancy it runs the cfg2 train step in 67 us against 52 us for the generated kernel (DESIGN.md
// section 2), so AUTO uses it for chains of its family that have no generated kernel.
//
//   * a CTA owns a tile of T = 2 * NT samples; every thread owns TWO consecutive samples as packed f32x2 pairs
//     (FFMA2 / FADD2 / FMUL2, smallnet.cuh / packed.cuh) -- one warp-uniform LDS.128 of weights feeds four FFMA2 = 8 FMAs;
//   * wide per-sample vectors (layer inputs, conditioner / decoder outputs, cotangents) live in shared memory as
//     rows [feature][sample] (a thread only touches its own 8-byte column: conflict-free, no barrier);
//   * the NARROW hidden layers (width H <= 16, a template constant) live in registers, so the three matvecs of a
//     conditioner are: runtime loop over inputs -> H accumulators; unrolled H x H; runtime loop over outputs;
//   * parameters are staged once per CTA into a compacted layout (only rows / columns that can matter, the last
//     layer transposed) through a gather table built at plan time; the CTA's gradient accumulator uses the same
//     layout and is scattered back into its partial row at the end;
//   * weight gradients are the one CTA-cooperative step: operands are parked as rows and contracted over the
//     sample axis with 2x4 register tiles and packed FFMA2 over sample pairs (no atomics, deterministic);
//   * the backward pass recomputes each layer's conditioner instead of storing hidden activations, which keeps
//     ~100 floats per sample in shared memory so that 448 samples (one CTA per SM at batch 65 536) are resident.
//
// Reference semantics (file:line under /root/reference): masked_coupling.py:73-85,
// masked_coupling_inference_funnel.py:72-91, conditioners/mlp.py:7-29, transformed_distribution.py:87-93;
// the VJP follows SURVEY.md appendix C.
#pragma once
#include "packed.cuh"

namespace cc {

using surj::f2;
using surj::u64;

constexpr int MAX_LAYERS = 12;
constexpr int ROW_ONES = 0;

struct Mlp {     // offsets into the staged parameter block (floats)
  int w1, b1, w2, b2, w3t, b3;
  int n_in;      // inputs (rows listed in itab at in_idx)
  int in_idx;    // itab offset: absolute row index of every input
  int n_out;     // outputs, padded to a multiple of 4 (pad rows have zero weights)
  int gin_idx;   // itab offset: absolute row (in the cotangent block) that receives d/d input k (accumulated)
};
struct Layer {
  int d_in, d_out;    // event widths
  int n_t;            // transformed dims
  int n_p;            // pass-through dims
  int n_v;            // funnel: dims scored by the decoder (0: plain coupling)
  int inner;          // INNER_AFFINE / INNER_AFFINE_INV
  int u_in, u_out;    // first row of the layer's input / output block
  int t_idx;          // itab offset: transformed dims (index inside the event)
  int p_idx;          // itab offset: pass-through dims (index inside the event)
  Mlp cond, dec;
};
struct Program {
  int n_layers, d_in, d_out, n_staged, n_params, n_itab;
  int row_th, row_g, row_park, n_rows_fwd, n_rows_bwd;  // first rows of the theta / cotangent / parking blocks
  const int* stage_src;  // [n_staged] parameter index of every staged value (-1: zero)
  const int* itab;       // [n_itab]
  Layer layers[MAX_LAYERS];
};
struct Io {
  const float* y; const float* params; const float* g_lp; const float* g_z;
  float g_scale;
  float* lp_out; float* z_out; float* g_y; float* partials;
  int npad, with_base, backward;
  long long B;
  unsigned* tickets; int n_tickets;
};

#define CC_DEV __device__ __forceinline__

CC_DEV f2 ld2(const float* p) { f2 r; r.v = *reinterpret_cast<const u64*>(p); return r; }
CC_DEV void st2(float* p, f2 v) { *reinterpret_cast<u64*>(p) = v.v; }

// One weight-gradient contraction (built once per CTA from the layer table, see build_job).
struct Job {
  int a_list;            // itab offset of the A-row list (float offsets of the rows), na - 1 entries; the last A row is the ones row
  int na, ng;            // operand rows
  int g_off;             // float offset of G row 0
  int dst_w, sa, sgs;    // sg[dst_w + a * sa + g * sgs] += sum_s A[a][s] G[g][s]
  int dst_b;             // sg[dst_b + g] += sum_s G[g][s]   (the ones row)
  int nta, inv_nta;      // 2-row A tiles; inv_nta = ceil(65536 / nta): e / nta == (e * inv_nta) >> 16 for e < 4096
  int log2ks, items;     // K slices per (a, g) tile and work items (= tiles << log2ks)
};

struct Tile {        // a thread's view of the CTA's shared memory
  float* rows;       // rows[r * ldt + column]
  float* my;         // rows + 2 * tid: this thread's pair of every row
  const float* par;  // staged parameters
  float* sg;         // gradient accumulator (staged layout; backward only)
  const int* itab;   // index lists, already multiplied by ldt (float offsets of rows)
  const Job* jobs;
  int ldt, T, tid, nt;
  CC_DEV float* mine(int r) const { return my + r * ldt; }
  CC_DEV float* at(int off) const { return my + off; }
};

// exp of both lanes: 2^t with t = x log2(e) evaluated by ex2.approx, times (1 + r ln 2) where r is the rounding
// error of t computed exactly with FMAs -- the accuracy of expf (this feeds log-dets and Gaussian residuals), with
// the multiplies packed.
CC_DEV f2 exp2v(f2 x) {
  const f2 t = x * f2(1.4426950216293334961f);
  f2 r = fmaf(x, f2(1.4426950216293334961f), -t);
  r = fmaf(x, f2(1.925963033500011079e-08f), r);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t.lo()));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t.hi()));
  return f2(e0, e1) * fmaf(r, f2(0.6931471805599453f), f2(1.0f));
}

// Activation with the one extra value its derivative wants: the sigmoid for swish / sigmoid (so the backward pass does
// not evaluate it again), the pre-activation otherwise.
template <int ACT>
CC_DEV f2 act_fwd_aux(f2 p, f2& aux) {
  if constexpr (ACT == surj::ACT_SWISH) { aux = surj::fast_sigmoid(p); return p * aux; }
  else if constexpr (ACT == surj::ACT_SIGMOID) { aux = surj::fast_sigmoid(p); return aux; }
  else { aux = p; return surj::act_fwd<ACT>(p); }
}
template <int ACT>
CC_DEV f2 act_bwd_aux(f2 h, f2 aux, f2 g) {
  if constexpr (ACT == surj::ACT_SWISH) return g * fmaf(h, f2(1.0f) - aux, aux);  // s + a s (1 - s) with a s = h
  else if constexpr (ACT == surj::ACT_SIGMOID) return g * aux * (f2(1.0f) - aux);
  else return surj::act_bwd<ACT>(aux, g);
}

// ---- hidden layers forward: p1 = b1 + in W1, h1 = act(p1), p2 = b2 + h1 W2, h2 = act(p2); (p1, p2) come back as the
// activations' aux values (see act_fwd_aux)
template <int H, int ACT>
CC_DEV void hidden_fwd(const Tile& t, const Mlp& m, f2 (&p1)[H], f2 (&h1)[H], f2 (&p2)[H], f2 (&h2)[H]) {
#pragma unroll
  for (int j = 0; j < H; ++j) p1[j] = f2(t.par[m.b1 + j]);
  const int* idx = t.itab + m.in_idx;
  const float4* w = reinterpret_cast<const float4*>(t.par + m.w1);
#pragma unroll 2
  for (int k = 0; k < m.n_in; ++k) {
    const f2 a = ld2(t.at(idx[k]));
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 wv = w[k * (H / 4) + q];
      p1[4 * q + 0] = fmaf(a, wv.x, p1[4 * q + 0]);
      p1[4 * q + 1] = fmaf(a, wv.y, p1[4 * q + 1]);
      p1[4 * q + 2] = fmaf(a, wv.z, p1[4 * q + 2]);
      p1[4 * q + 3] = fmaf(a, wv.w, p1[4 * q + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < H; ++j) {
    h1[j] = act_fwd_aux<ACT>(p1[j], p1[j]);
    p2[j] = f2(t.par[m.b2 + j]);
  }
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float4* w2 = reinterpret_cast<const float4*>(t.par + m.w2 + i * H);
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 wv = w2[q];
      p2[4 * q + 0] = fmaf(h1[i], wv.x, p2[4 * q + 0]);
      p2[4 * q + 1] = fmaf(h1[i], wv.y, p2[4 * q + 1]);
      p2[4 * q + 2] = fmaf(h1[i], wv.z, p2[4 * q + 2]);
      p2[4 * q + 3] = fmaf(h1[i], wv.w, p2[4 * q + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < H; ++j) h2[j] = act_fwd_aux<ACT>(p2[j], p2[j]);
}

// ---- output layer forward: row[out0 + o] = b3[o] + sum_i h2[i] W3T[o][i], four outputs per step
template <int H>
CC_DEV void out_fwd(const Tile& t, const Mlp& m, const f2 (&h2)[H], int out0) {
  float* dst = t.mine(out0);
  const float4* w = reinterpret_cast<const float4*>(t.par + m.w3t);
#pragma unroll 1
  for (int o = 0; o < m.n_out; o += 4) {
    const float4 b = *reinterpret_cast<const float4*>(t.par + m.b3 + o);
    f2 acc[4] = {f2(b.x), f2(b.y), f2(b.z), f2(b.w)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int r = 0; r < H / 4; ++r) {
        const float4 wv = w[(o + q) * (H / 4) + r];
        acc[q] = fmaf(h2[4 * r + 0], wv.x, acc[q]);
        acc[q] = fmaf(h2[4 * r + 1], wv.y, acc[q]);
        acc[q] = fmaf(h2[4 * r + 2], wv.z, acc[q]);
        acc[q] = fmaf(h2[4 * r + 3], wv.w, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) st2(dst + (o + q) * t.ldt, acc[q]);
  }
}

// ---- weight-gradient contraction over the tile's samples (CTA-cooperative; callers bracket it with barriers).
// Work item = a 2x4 tile of (a, g) pairs x one of 2^log2ks slices of the sample axis; the lanes of an item are
// adjacent and combine with shuffles, the first of them -- the only thread of the CTA that touches those addresses --
// adds the sums into sg.  The products run as packed FFMA2 over even / odd samples (an LDS.128 delivers two aligned
// pairs of each operand).  Warps without a live item skip the job.
// Not inlined: one copy of this loop nest serves every weight matrix.
static __device__ __noinline__ void contract_job(const Job* jp, float* rows, float* sg, const int* itab, int T, int tid, int nt) {
  constexpr int TA = 2, TG = 4;
  const Job j = *jp;
  const int KS = 1 << j.log2ks, KLEN = T >> j.log2ks;
  const int* alist = itab + j.a_list;
#pragma unroll 1
  for (int w0 = 0; w0 < j.items; w0 += nt) {
    if (w0 + (tid & ~31) >= j.items) continue;  // the whole warp is idle in this round
    const int w = w0 + tid;
    const bool live = w < j.items;
    const int wc = live ? w : 0;
    const int e = wc >> j.log2ks, ks = wc & (KS - 1);
    const int tg = (e * j.inv_nta) >> 16, ta = e - tg * j.nta;
    const int ntg = (j.ng + TG - 1) / TG;
    const float* ap[TA];
    const float* gp[TG];
    int ia[TA], ig[TG];
#pragma unroll
    for (int a = 0; a < TA; ++a) {
      ia[a] = ta + a * j.nta;
      const int c = min(ia[a], j.na - 1);
      ap[a] = rows + (c < j.na - 1 ? alist[c] : 0) + ks * KLEN;  // row 0 is the ones row
    }
#pragma unroll
    for (int g = 0; g < TG; ++g) {
      ig[g] = tg + g * ntg;
      gp[g] = rows + j.g_off + min(ig[g], j.ng - 1) * (T + 4) + ks * KLEN;
    }
    f2 acc2[TA][TG];
#pragma unroll
    for (int a = 0; a < TA; ++a)
#pragma unroll
      for (int g = 0; g < TG; ++g) acc2[a][g] = f2(0.0f);
#pragma unroll 2
    for (int k = 0; k < KLEN; k += 4) {
      ulonglong2 av[TA], gv[TG];
#pragma unroll
      for (int a = 0; a < TA; ++a) av[a] = *reinterpret_cast<const ulonglong2*>(ap[a] + k);
#pragma unroll
      for (int g = 0; g < TG; ++g) gv[g] = *reinterpret_cast<const ulonglong2*>(gp[g] + k);
#pragma unroll
      for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int g = 0; g < TG; ++g) {
          f2 ax, ay, gx, gy;
          ax.v = av[a].x; ay.v = av[a].y; gx.v = gv[g].x; gy.v = gv[g].y;
          acc2[a][g] = fmaf(ax, gx, acc2[a][g]);
          acc2[a][g] = fmaf(ay, gy, acc2[a][g]);
        }
    }
    float acc[TA][TG];
#pragma unroll
    for (int a = 0; a < TA; ++a)
#pragma unroll
      for (int g = 0; g < TG; ++g) acc[a][g] = surj::hsum(acc2[a][g]);
#pragma unroll 1
    for (int o = KS >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int g = 0; g < TG; ++g) acc[a][g] += __shfl_xor_sync(0xffffffffu, acc[a][g], o);
    if (live && ks == 0) {
#pragma unroll
      for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int g = 0; g < TG; ++g) {
          if (ia[a] < j.na && ig[g] < j.ng) {
            const int o = ia[a] < j.na - 1 ? j.dst_w + ia[a] * j.sa + ig[g] * j.sgs : j.dst_b + ig[g];
            sg[o] += acc[a][g];
          }
        }
    }
  }
}

CC_DEV void contract(const Tile& t, int job) { contract_job(t.jobs + job, t.rows, t.sg, t.itab, t.T, t.tid, t.nt); }

// Jobs of layer l: 6 * l + {0, 1, 2} = decoder dW3, dW2, dW1; + {3, 4, 5} = conditioner dW3, dW2, dW1.
// dW3 = h2^T g_out (stored transposed [out][H]; h2 parked), dW2 = h1^T g_p2 (both parked), dW1 = in^T g_p1 (g_p1 parked
// in the output block).
CC_DEV Job build_job(const Mlp& m, int which, int H, int row_th, int row_park, int hm, int park_list, int ldt, int nt) {
  Job j;
  if (which == 0) { j.a_list = park_list; j.na = H + 1; j.ng = m.n_out; j.g_off = row_th * ldt; j.dst_w = m.w3t; j.sa = 1; j.sgs = H; j.dst_b = m.b3; }
  else if (which == 1) { j.a_list = park_list; j.na = H + 1; j.ng = H; j.g_off = (row_park + hm) * ldt; j.dst_w = m.w2; j.sa = H; j.sgs = 1; j.dst_b = m.b2; }
  else { j.a_list = m.in_idx; j.na = m.n_in + 1; j.ng = H; j.g_off = row_th * ldt; j.dst_w = m.w1; j.sa = H; j.sgs = 1; j.dst_b = m.b1; }
  j.nta = (j.na + 1) / 2;
  j.inv_nta = (65536 + j.nta - 1) / j.nta;
  const int ne = j.nta * ((j.ng + 3) / 4);
  j.log2ks = 0;
  while (j.log2ks < 4 && (ne << (j.log2ks + 1)) <= nt) ++j.log2ks;
  j.items = ne << j.log2ks;
  return j;
}

// ---- backward through one MLP.  On entry row[g0 + o] holds the cotangent of output o (pad rows: zero); the
// cotangent of input k is ADDED to the row at float offset itab[gin_idx + k].  Parameter gradients go to t.sg.
// park_a / park_g: first rows of two parking blocks; job0: index of this MLP's first contraction job.
template <int H, int ACT>
CC_DEV void mlp_bwd(const Tile& t, const Mlp& m, const f2 (&p1)[H], const f2 (&h1)[H], const f2 (&p2)[H],
                    const f2 (&h2)[H], int g0, int park_a, int park_g, int job0) {
  float* pa = t.mine(park_a);
  float* pg = t.mine(park_g);
  float* go = t.mine(g0);
#pragma unroll
  for (int i = 0; i < H; ++i) st2(pa + i * t.ldt, h2[i]);
  __syncthreads();
  contract(t, job0);
  // g_h2[i] = sum_o W3T[o][i] g_out[o]
  f2 gh[H];
#pragma unroll
  for (int i = 0; i < H; ++i) gh[i] = f2(0.0f);
  const float4* w3 = reinterpret_cast<const float4*>(t.par + m.w3t);
#pragma unroll 2
  for (int o = 0; o < m.n_out; ++o) {
    const f2 g = ld2(go + o * t.ldt);
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 wv = w3[o * (H / 4) + q];
      gh[4 * q + 0] = fmaf(g, wv.x, gh[4 * q + 0]);
      gh[4 * q + 1] = fmaf(g, wv.y, gh[4 * q + 1]);
      gh[4 * q + 2] = fmaf(g, wv.z, gh[4 * q + 2]);
      gh[4 * q + 3] = fmaf(g, wv.w, gh[4 * q + 3]);
    }
  }
  f2 gp2[H], gp1[H];
#pragma unroll
  for (int j = 0; j < H; ++j) gp2[j] = act_bwd_aux<ACT>(h2[j], p2[j], gh[j]);
  // g_h1[i] = sum_j W2[i][j] g_p2[j]
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float4* w = reinterpret_cast<const float4*>(t.par + m.w2 + i * H);
    f2 a = f2(0.0f);
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 wv = w[q];
      a = fmaf(gp2[4 * q + 0], wv.x, a);
      a = fmaf(gp2[4 * q + 1], wv.y, a);
      a = fmaf(gp2[4 * q + 2], wv.z, a);
      a = fmaf(gp2[4 * q + 3], wv.w, a);
    }
    gp1[i] = act_bwd_aux<ACT>(h1[i], p1[i], a);
  }
  __syncthreads();  // every thread is done with the dW3 operands: the parking rows and the output block are free
#pragma unroll
  for (int i = 0; i < H; ++i) {
    st2(pa + i * t.ldt, h1[i]);
    st2(pg + i * t.ldt, gp2[i]);
    st2(go + i * t.ldt, gp1[i]);
  }
  __syncthreads();
  contract(t, job0 + 1);
  contract(t, job0 + 2);
  // g_in[k] += sum_j W1[k][j] g_p1[j]
  const int* gidx = t.itab + m.gin_idx;
  const float4* w1 = reinterpret_cast<const float4*>(t.par + m.w1);
#pragma unroll 2
  for (int k = 0; k < m.n_in; ++k) {
    f2 a0 = f2(0.0f), a1 = f2(0.0f);
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 wv = w1[k * (H / 4) + q];
      a0 = fmaf(gp1[4 * q + 0], wv.x, a0);
      a1 = fmaf(gp1[4 * q + 1], wv.y, a1);
      a0 = fmaf(gp1[4 * q + 2], wv.z, a0);
      a1 = fmaf(gp1[4 * q + 3], wv.w, a1);
    }
    float* dst = t.at(gidx[k]);
    st2(dst, ld2(dst) + (a0 + a1));
  }
  __syncthreads();  // the contractions are done reading the parked operands
}

// ---- one layer, forward (log-density) or backward (VJP) phase.  Both phases run the same conditioner / decoder
// forward code (the backward phase recomputes instead of storing hidden activations), so the two are ONE body with
// the phase as a runtime flag: the instruction footprint is what this kernel is about.
//   forward : conditioner -> affine transform of the transformed dims -> (funnel) decoder -> log N(y[K:]; mu, sigma)
//   backward: (funnel) decoder forward + VJP -> conditioner forward -> affine VJP -> conditioner VJP
// The cotangent block (row_g ...) holds d loss / d (layer output) on entry to the backward phase and
// d loss / d (layer input) on exit; g_lp is d loss / d lp of the thread's samples; lc accumulates the log-density.
template <int HC, int HD, int ACT>
CC_DEV void layer_step(const Tile& t, const Layer& L, int job0, bool bwd, int row_th, int row_g, int row_park, f2 g_lp, f2& lc) {
  constexpr int HM = HC > HD ? HC : HD;
  float* uin = t.mine(L.u_in);
  float* uout = t.mine(L.u_out);
  float* th = t.mine(row_th);
  float* gr = t.mine(row_g);
#pragma unroll 1
  for (int sub = 0; sub < 2; ++sub) {
    if ((sub == 0) == bwd) {  // decoder part: second in the forward phase, first in the backward phase
      if (L.n_v == 0) {       // plain coupling: nothing to score; forward copies the pass-through dims
        if (!bwd) {
          const int* pidx = t.itab + L.p_idx;
#pragma unroll 1
          for (int k = 0; k < L.n_p; ++k) st2(uout + pidx[k], ld2(uin + pidx[k]));
        }
        continue;
      }
      f2 p1[HD], h1[HD], p2[HD], h2[HD];
      hidden_fwd<HD, ACT>(t, L.dec, p1, h1, p2, h2);
      out_fwd<HD>(t, L.dec, h2, row_th);
      // lc += Independent(Normal(mu, exp(ls))).log_prob(y[K:]) with (mu, ls) = decoder(z)
      const float* vv = uin + L.d_out * t.ldt;
      float* gv = gr + L.d_out * t.ldt;
      float* ths = th + L.n_v * t.ldt;
#pragma unroll 1
      for (int j = 0; j < L.n_v; ++j) {
        const int o = j * t.ldt;
        const f2 v = ld2(vv + o), mu = ld2(th + o), ls = ld2(ths + o);
        const f2 e = exp2v(-ls);
        const f2 r = (v - mu) * e;
        if (!bwd) {
          lc = lc + (f2(-0.5f) * r * r - ls - f2(SURJ_HALF_LOG_2PI));
        } else {
          const f2 gre = g_lp * r * e;
          st2(th + o, gre);                            // d / d mu
          st2(ths + o, g_lp * (r * r - f2(1.0f)));     // d / d log_sigma
          st2(gv + o, -gre);                           // d / d v: the scored dims are inputs K.. of the layer
        }
      }
      if (bwd) mlp_bwd<HD, ACT>(t, L.dec, p1, h1, p2, h2, row_th, row_park, row_park + HM, job0);
    } else {  // conditioner part
      f2 p1[HC], h1[HC], p2[HC], h2[HC];
      hidden_fwd<HC, ACT>(t, L.cond, p1, h1, p2, h2);
      out_fwd<HC>(t, L.cond, h2, row_th);
      const int* tidx = t.itab + L.t_idx;
      float* ths = th + L.n_t * t.ldt;
      const bool direct = L.inner == surj::INNER_AFFINE;  // z = (y - m) e^{-s}, ld = -s ; else z = e^{s} y + m, ld = +s
#pragma unroll 1
      for (int k = 0; k < L.n_t; ++k) {
        const int d = tidx[k], o = k * t.ldt;
        const f2 u = ld2(uin + d), sh = ld2(th + o), ls = ld2(ths + o);
        const f2 e = exp2v(direct ? -ls : ls);
        if (!bwd) {
          st2(uout + d, direct ? (u - sh) * e : fmaf(e, u, sh));
          lc = direct ? lc - ls : lc + ls;
        } else {
          const f2 gz = ld2(gr + d);
          const f2 gu = gz * e;
          st2(th + o, direct ? -gu : gz);                                             // d / d shift
          st2(ths + o, direct ? -(gu * (u - sh)) - g_lp : fmaf(gu, u, g_lp));         // d / d log_scale
          st2(gr + d, gu);
        }
      }
      if (bwd) mlp_bwd<HC, ACT>(t, L.cond, p1, h1, p2, h2, row_th, row_park, row_park + HM, job0 + 3);
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernel
// Shared memory: staged parameters | gradient accumulator (backward) | index lists | contraction jobs (backward) | rows.
template <int HC, int HD, int ACT>
__global__ void __launch_bounds__(256, 1) chain_kernel(const __grid_constant__ Program P, const __grid_constant__ Io io) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int HM = HC > HD ? HC : HD;
  const int tid = threadIdx.x, nt = blockDim.x, T = 2 * nt, ldt = T + 4;
  const int backward = io.backward;
  if (backward && blockIdx.x == 0)  // arm the tickets of the gradient-reduction kernel that follows on the stream
    for (int i = tid; i < io.n_tickets; i += nt) io.tickets[i] = 0u;
  const int ns4 = (P.n_staged + 3) & ~3;
  float* par = reinterpret_cast<float*>(smem_raw);
  float* sg = par + ns4;
  int* itab = reinterpret_cast<int*>(sg + (backward ? ns4 + 4 : 0));
  Job* jobs = reinterpret_cast<Job*>(itab + ((P.n_itab + 3) & ~3));
  float* rows = reinterpret_cast<float*>(jobs + (backward ? 6 * P.n_layers : 0));
#pragma unroll 8
  for (int i = tid; i < P.n_staged; i += nt) {  // independent gathers: keep several in flight
    const int s = __ldg(P.stage_src + i);
    par[i] = s >= 0 ? __ldg(io.params + s) : 0.0f;
  }
  if (backward) {
#pragma unroll 1
    for (int i = tid; i < ns4 + 4; i += nt) sg[i] = 0.0f;
  }
#pragma unroll 1
  for (int i = tid; i < P.n_itab; i += nt) itab[i] = P.itab[i] * ldt;  // every list entry is a row (or a dim = row delta)
#pragma unroll 1
  for (int i = tid; i < T; i += nt) rows[i] = 1.0f;                    // row 0: the ones row
  const int park_list = P.n_itab - HM;  // the plan appends the parking-row list to itab
  if (backward) {
#pragma unroll 1
    for (int i = tid; i < 6 * P.n_layers; i += nt) {
      const Layer& L = P.layers[i / 6];
      const int which = i % 6;
      jobs[i] = which < 3 ? build_job(L.dec, which, HD, P.row_th, P.row_park, HM, park_list, ldt, nt)
                          : build_job(L.cond, which - 3, HC, P.row_th, P.row_park, HM, park_list, ldt, nt);
    }
  }
  __syncthreads();

  Tile t;
  t.rows = rows; t.my = rows + 2 * tid; t.par = par; t.sg = sg; t.itab = itab; t.jobs = jobs;
  t.ldt = ldt; t.T = T; t.tid = tid; t.nt = nt;
  f2 lp_total = f2(0.0f);
  const long long n_tiles = (io.B + T - 1) / T;
  const int zrow = P.layers[P.n_layers - 1].u_out;
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long s = tile * T + 2 * tid;
    const int n = s >= io.B ? 0 : (int)(io.B - s < 2 ? io.B - s : 2);
    // ---- the thread's two rows of y (2 * d_in consecutive floats) -> first input block
    {
      float* u0 = t.mine(P.layers[0].u_in);
      const float* y0 = io.y + s * P.d_in;
      if (n == 2 && (P.d_in & 3) == 0) {
        const float4* r0 = reinterpret_cast<const float4*>(y0);
        const float4* r1 = reinterpret_cast<const float4*>(y0 + P.d_in);
#pragma unroll 1
        for (int q = 0; q < P.d_in / 4; ++q) {
          const float4 a = __ldg(r0 + q), b = __ldg(r1 + q);
          float* u = u0 + 4 * q * ldt;
          st2(u, f2(a.x, b.x)); st2(u + ldt, f2(a.y, b.y)); st2(u + 2 * ldt, f2(a.z, b.z)); st2(u + 3 * ldt, f2(a.w, b.w));
        }
      } else {
#pragma unroll 1
        for (int d = 0; d < P.d_in; ++d) {
          const float a = n > 0 ? __ldg(y0 + d) : 0.0f, b = n > 1 ? __ldg(y0 + P.d_in + d) : 0.0f;
          st2(u0 + d * ldt, f2(a, b));
        }
      }
    }
    f2 lp = f2(0.0f), g_lp = f2(0.0f);
    const int n_steps = backward ? 2 * P.n_layers : P.n_layers;
#pragma unroll 1
    for (int step = 0; step < n_steps; ++step) {
      const bool bwd = step >= P.n_layers;
      if (step == P.n_layers) {
        // ---- end of the forward phase (backward launches): cotangent of z = the caller's g_z plus the base density's -g_lp z
        if (io.g_lp != nullptr) surj::load_scalar(io.g_lp, s, n, 0.0f, g_lp);
        else g_lp = surj::keep_valid(f2(io.g_scale), n);
#pragma unroll 1
        for (int d = 0; d < P.d_out; ++d) {
          f2 g = f2(0.0f);
          if (io.g_z != nullptr) surj::load_elem(io.g_z, s, n, P.d_out, d, g);
          if (io.with_base) g = g - g_lp * ld2(t.mine(zrow + d));
          st2(t.mine(P.row_g + d), g);
        }
      }
      const int l = bwd ? 2 * P.n_layers - 1 - step : step;
      layer_step<HC, HD, ACT>(t, P.layers[l], 6 * l, bwd, P.row_th, P.row_g, P.row_park, g_lp, lp);
    }
    if (io.with_base) {
#pragma unroll 1
      for (int d = 0; d < P.d_out; ++d) {
        const f2 z = ld2(t.mine(zrow + d));
        lp = lp + (f2(-0.5f) * z * z - f2(SURJ_HALF_LOG_2PI));
      }
    }
    if (io.lp_out != nullptr) surj::store_scalar(io.lp_out, s, n, lp);
    if (io.z_out != nullptr) {
#pragma unroll 1
      for (int d = 0; d < P.d_out; ++d) {
        const f2 z = ld2(t.mine(zrow + d));
        if (n > 0) io.z_out[s * P.d_out + d] = z.lo();
        if (n > 1) io.z_out[(s + 1) * P.d_out + d] = z.hi();
      }
    }
    if (!backward) continue;
    lp_total += surj::keep_valid(lp, n);
    if (io.g_y != nullptr) {
#pragma unroll 1
      for (int d = 0; d < P.d_in; ++d) {
        const f2 g = ld2(t.mine(P.row_g + d));
        if (n > 0) io.g_y[s * P.d_in + d] = g.lo();
        if (n > 1) io.g_y[(s + 1) * P.d_in + d] = g.hi();
      }
    }
  }
  if (!backward) return;
  // ---- this CTA's partial gradient row: [n_params] gradients (scattered back through the gather table), then sum(lp)
  float lp_sum = surj::hsum(lp_total);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lp_sum += __shfl_xor_sync(0xffffffffu, lp_sum, o);
  __syncthreads();
  if ((tid & 31) == 0) atomicAdd(sg + ns4, lp_sum);
  float* out = io.partials + (size_t)blockIdx.x * io.npad;
#pragma unroll 1
  for (int i = tid; i < io.npad; i += nt) out[i] = 0.0f;
  __syncthreads();
#pragma unroll 1
  for (int i = tid; i < P.n_staged; i += nt) {
    const int s = P.stage_src[i];
    if (s >= 0) out[s] = sg[i];
  }
  if (tid == 0) out[P.n_params] = sg[ns4];
}

}  // namespace cc

// Host side of the fused small-network path: compiles a chain (SurjLayerDesc[]) into the
// program the interpreter kernel runs (vm.cuh / vm_kernels.cuh), uploads its tables once per
// plan and launches it.  See vm.cuh for the data layout.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "plan.h"
#include "vm_kernels.cuh"
#include "vm_plan.h"

namespace {

using vm::Op;

inline int pad8(int n) { return (n + 7) & ~7; }
inline int pad4(int n) { return (n + 3) & ~3; }

struct Mat {      // one compacted matrix (+ bias) inside the staged parameter area
  int n_in = 0, n_out = 0;  // compact sizes
  int w_off = 0, ld = 0;    // W[n_in][ld]
  int wt_off = 0, ldt = 0;  // W^T[n_out][ldt]
  int b_off = -1;           // b[ld]
};

struct NetRows {  // rows of one Linear layer during a (training-mode) forward
  int in_row, n_in, pre_row, post_row, act;
  Mat m;
};

struct Builder {
  const SurjPlan* plan;
  std::vector<int32_t> stage, grad_dst, grad_aux, itab;
  std::vector<vm::ScalarJob> sjobs;
  int n_scalars = 0;
  int rows_next = vm::ROW_FIRST_FREE;  // rows 0..2: ones, lp, g_lp (vm_ops.cuh)
  int scratch_base = 0, scratch_top = 0, scratch_max = 0;
  std::string why;

  int alloc_rows(int n) {
    const int r = rows_next;
    rows_next += n;
    return r;
  }
  int scratch(int n) {
    const int r = scratch_top;
    scratch_top += n;
    scratch_max = std::max(scratch_max, scratch_top);
    return r;
  }
  void scratch_reset() { scratch_top = scratch_base; }
  int tab(const std::vector<int>& v) {
    const int off = (int)itab.size();
    for (int x : v) itab.push_back(x);
    return off;
  }
  void push(int32_t code, int32_t dst, int32_t aux = -1) {
    stage.push_back(code);
    grad_dst.push_back(dst);
    grad_aux.push_back(aux);
  }
  // src(i, j) -> {stage code, grad dst, grad aux}; bias(j) -> param index or -1 for "no bias vector"
  struct Ent { int32_t code, dst, aux; };
  Mat add_matrix(int n_in, int n_out, const std::function<Ent(int, int)>& src, const std::function<int(int)>* bias) {
    Mat m;
    m.n_in = n_in; m.n_out = n_out; m.ld = pad8(n_out); m.ldt = pad8(n_in);
    while (stage.size() & 3) push(vm::STAGE_ZERO, -1);
    m.w_off = (int)stage.size();
    for (int i = 0; i < n_in; ++i)
      for (int j = 0; j < m.ld; ++j) {
        if (j < n_out) { const Ent e = src(i, j); push(e.code, e.dst, e.aux); }
        else push(vm::STAGE_ZERO, -1);
      }
    if (bias) {
      m.b_off = (int)stage.size();
      for (int j = 0; j < m.ld; ++j) {
        if (j < n_out) { const int p = (*bias)(j); push(p, p); }
        else push(vm::STAGE_ZERO, -1);
      }
    }
    m.wt_off = (int)stage.size();
    for (int j = 0; j < n_out; ++j)
      for (int i = 0; i < m.ldt; ++i) {
        if (i < n_in) push(src(i, j).code, -1);
        else push(vm::STAGE_ZERO, -1);
      }
    return m;
  }
};

Op mk(int code, std::initializer_list<int> a, float f0 = 0.0f, float f1 = 0.0f) {
  Op o;
  std::memset(&o, 0, sizeof o);
  o.code = code;
  int k = 0;
  for (int v : a) o.a[k++] = v;
  o.f0 = f0; o.f1 = f1;
  return o;
}
// out = act(b + in W): a10 / a11 locate the [out][in] copy of the same weights (what the tensor-core
// kernel stages as its K-major B operand), a12 = number of real (unpadded) outputs
Op lin_fwd(const Mat& m, int in_row, int out_row, int act, int accumulate, int pre_row) {
  return mk(vm::OP_LINEAR, {in_row, m.n_in, m.w_off, m.ld, out_row, m.ld, m.b_off, act, accumulate, pre_row, m.wt_off,
                            m.ldt, m.n_out});
}
// g_in = g W^T: an OP_LINEAR on the transposed copy (whose [out][in] copy is the primary matrix)
Op lin_bwd_in(const Mat& m, int g_row, int dst_row, int accumulate) {
  return mk(vm::OP_LINEAR, {g_row, m.n_out, m.wt_off, m.ldt, dst_row, m.ldt, -1, vm::ACT_IDENTITY, accumulate, -1, m.w_off,
                            m.ld, m.n_in});
}

// ----- MLP (hk.nets.MLP / Linear stack) with optional row / column compaction of its first / last layer
std::vector<Mat> build_mlp(Builder& B, const SurjMlpDesc& d, const std::vector<int>* rows, const std::vector<int>* cols) {
  std::vector<Mat> out;
  for (int l = 0; l < d.n_layers; ++l) {
    const int in_w = d.widths[l], out_w = d.widths[l + 1];
    std::vector<int> ri, ci;
    if (l == 0 && rows) ri = *rows; else { ri.resize(in_w); for (int i = 0; i < in_w; ++i) ri[i] = i; }
    if (l == d.n_layers - 1 && cols) ci = *cols; else { ci.resize(out_w); for (int j = 0; j < out_w; ++j) ci[j] = j; }
    const int64_t w0 = d.w_off[l], b0 = d.b_off[l];
    auto src = [&](int i, int j) -> Builder::Ent {
      const int p = (int)(w0 + (int64_t)ri[i] * out_w + ci[j]);
      return {p, p, -1};
    };
    std::function<int(int)> bias = [&](int j) -> int { return (int)(b0 + ci[j]); };
    out.push_back(B.add_matrix((int)ri.size(), (int)ci.size(), src, b0 >= 0 ? &bias : nullptr));
  }
  return out;
}

// forward ops of a Linear stack; input rows must already be filled.  Returns the rows of every layer.
std::vector<NetRows> emit_mlp_fwd(Builder& B, std::vector<Op>& ops, const std::vector<Mat>& mats, const int32_t* acts,
                                  int in_row, bool train) {
  std::vector<NetRows> rows;
  int cur = in_row;
  for (size_t l = 0; l < mats.size(); ++l) {
    const Mat& m = mats[l];
    NetRows r;
    r.in_row = cur; r.n_in = m.n_in; r.act = acts[l]; r.m = m;
    r.post_row = B.scratch(m.ld);
    r.pre_row = (train && r.act != vm::ACT_IDENTITY) ? B.scratch(m.ld) : -1;
    ops.push_back(lin_fwd(m, cur, r.post_row, r.act, 0, r.pre_row));
    rows.push_back(r);
    cur = r.post_row;
  }
  return rows;
}

// backward through a Linear stack whose last layer's post rows hold the cotangent of its output.
// The cotangent of each layer's input overwrites that input's rows (which must span pad8(n_in) rows),
// except for layer 0 when gin0_row >= 0 (then it is accumulated there) or need_gin0 is false.
void emit_mlp_bwd(std::vector<Op>& ops, const std::vector<NetRows>& rows, bool need_gin0, int gin0_row = -1) {
  for (int l = (int)rows.size() - 1; l >= 0; --l) {
    const NetRows& r = rows[l];
    if (r.act != vm::ACT_IDENTITY) ops.push_back(mk(vm::OP_ACT_BWD, {r.post_row, r.pre_row, r.m.n_out, r.act}));
    ops.push_back(mk(vm::OP_LINEAR_DW, {r.in_row, r.n_in, r.post_row, r.m.n_out, r.m.w_off, r.m.ld, r.m.b_off}));
    if (l > 0 || need_gin0) {
      const bool acc = (l == 0 && gin0_row >= 0);
      ops.push_back(lin_bwd_in(r.m, r.post_row, acc ? gin0_row : r.in_row, acc ? 1 : 0));
    }
  }
}

std::vector<int> iota(int n, int start = 0) {
  std::vector<int> v(n);
  for (int i = 0; i < n; ++i) v[i] = start + i;
  return v;
}

struct MadeMats {
  Mat l0, ctx;
  bool has_ctx = false;
  std::vector<Mat> rest;
  std::vector<int32_t> rest_acts;
};

MadeMats build_made(Builder& B, const SurjMadeDesc& d) {
  MadeMats mm;
  std::vector<int32_t> hidden(d.hidden, d.hidden + d.n_hidden);
  const int64_t total = surj_made_masks(d.n_params, d.input_size, hidden.data(), d.n_hidden, nullptr);
  std::vector<uint8_t> masks((size_t)total);
  surj_made_masks(d.n_params, d.input_size, hidden.data(), d.n_hidden, masks.data());
  int64_t moff = 0;
  int prev = d.input_size;
  for (int l = 0; l <= d.n_hidden; ++l) {
    const int cur = l < d.n_hidden ? d.hidden[l] : d.input_size * d.n_params;
    const int64_t w0 = d.w_off[l], b0 = d.b_off[l];
    const uint8_t* mk_ = masks.data() + moff;
    auto src = [&](int i, int j) -> Builder::Ent {
      const int p = (int)(w0 + (int64_t)i * cur + j);
      if (mk_[(size_t)i * cur + j]) return {p, p, -1};
      return {vm::STAGE_ZERO, -1, -1};  // w * mask: masked weights are exact zeros with exact-zero gradients
    };
    std::function<int(int)> bias = [&](int j) -> int { return (int)(b0 + j); };
    Mat m = B.add_matrix(prev, cur, src, &bias);
    if (l == 0) mm.l0 = m;
    else { mm.rest.push_back(m); mm.rest_acts.push_back(l < d.n_hidden ? d.act : vm::ACT_IDENTITY); }
    moff += (int64_t)prev * cur;
    prev = cur;
  }
  if (d.ctx_dim > 0) {
    mm.has_ctx = true;
    const int64_t w0 = d.ctx_w_off, b0 = d.ctx_b_off;
    const int h0 = d.hidden[0];
    auto src = [&](int i, int j) -> Builder::Ent {
      const int p = (int)(w0 + (int64_t)i * h0 + j);
      return {p, p, -1};
    };
    std::function<int(int)> bias = [&](int j) -> int { return (int)(b0 + j); };
    mm.ctx = B.add_matrix(d.ctx_dim, h0, src, &bias);
  }
  return mm;
}

struct LayerCode {
  std::vector<Op> fwd;    // inference-mode forward
  std::vector<Op> dec;    // funnels: decoder forward + backward, run FIRST in the backward program -- it only needs the layer's
                          // kept output rows, so its scratch rows are released before the conditioner's are allocated
  std::vector<Op> train;  // training-mode forward recompute (keeps pre-activations)
  std::vector<Op> bwd;    // backward, assumes `train` just ran with the same scratch layout
  std::vector<Op> gen;    // sampling direction: forward_and_likelihood_contribution (E[l+1] -> E[l])
  bool gen_ok = false;
};

}  // namespace

static size_t vm_smem_bytes(const VmPlan& v, int prog, int T) {
  const bool backward = prog == VM_PROG_BWD;
  const size_t n_ops = prog == VM_PROG_BWD ? v.ops_bwd.size() : (prog == VM_PROG_GEN ? v.ops_gen.size() : v.ops_fwd.size());
  size_t b = n_ops * sizeof(Op);
  b += (size_t)pad4((int)v.itab.size()) * 4;
  b += (size_t)pad4((int)v.stage.size()) * 4;
  b += 2 * (size_t)pad4(v.n_scalars) * 4;
  if (backward) b += (size_t)pad4((int)v.stage.size() + 1) * 4;
  b += (size_t)v.n_rows * T * 4;
  return b;
}

// Returns nullptr (and leaves a reason in *why) when the chain does not fit the fused path.
VmPlan* vm_compile(const SurjPlan* plan, std::string* why) {
  Builder B;
  B.plan = plan;
  const int L = (int)plan->layers.size();
  const int dx = plan->dim_x;
  auto fail = [&](const std::string& s) -> VmPlan* { if (why) *why = s; return nullptr; };

  // ---- rows: ones | x | events E_0..E_L | gx | ga | gb | scratch
  int max_d = plan->dim_in;
  for (const SurjLayerDesc& Ld : plan->layers) max_d = std::max(max_d, std::max(Ld.dim_in, Ld.dim_out));
  const int X = B.alloc_rows(pad8(std::max(dx, 1)));
  std::vector<int> E(L + 1);
  E[0] = B.alloc_rows(pad8(plan->dim_in));
  for (int l = 0; l < L; ++l) {
    const SurjLayerDesc& Ld = plan->layers[l];
    E[l + 1] = (Ld.kind == SURJ_LAYER_SLICE) ? E[l] : B.alloc_rows(pad8(Ld.dim_out));
  }
  const int GX = B.alloc_rows(pad8(std::max(dx, 1)));
  const int G[2] = {B.alloc_rows(pad8(max_d)), B.alloc_rows(pad8(max_d))};
  B.scratch_base = B.scratch_top = B.rows_next;

  std::vector<LayerCode> code(L);
  int n_eps = 0;  // noise columns are assigned in sampling order (last layer first)
  std::vector<int> eps_col(L, -1);
  for (int l = L - 1; l >= 0; --l) {
    const SurjLayerDesc& Ld = plan->layers[l];
    if (Ld.kind == SURJ_LAYER_SLICE || Ld.kind == SURJ_LAYER_COUPLING_FUNNEL) {
      eps_col[l] = n_eps;
      n_eps += Ld.dim_in - Ld.dim_out;
    }
  }
  for (int l = 0; l < L; ++l) {
    const SurjLayerDesc& Ld = plan->layers[l];
    LayerCode& C = code[l];
    const int D = Ld.dim_in, K = Ld.dim_out;
    const int dxu = Ld.use_x ? dx : 0;
    const int Ein = E[l], Eout = E[l + 1];
    const int Gout = G[(L - (l + 1)) & 1], Gin = G[(L - l) & 1];
    const bool rqs = Ld.inner == SURJ_INNER_RQS || Ld.inner == SURJ_INNER_RQS_INV;
    const int Kb = Ld.n_bins;
    const int P = rqs ? 3 * Kb + 1 : 2;
    const float lo = Ld.range_min, hi = Ld.range_max;
    // inner-bijector ops on transformed dims `dims` (itab offset), thetas at row th
    auto inner_fwd = [&](std::vector<Op>& ops, int n_t, int dims, int th, bool interleaved) {
      if (rqs) ops.push_back(mk(vm::OP_RQS, {Ein, Eout, n_t, dims, th, Kb, Ld.inner}, lo, hi));
      else if (interleaved) ops.push_back(mk(vm::OP_AFFINE, {Ein, Eout, n_t, dims, th, 2, th + 1, 2, Ld.inner}));
      else ops.push_back(mk(vm::OP_AFFINE, {Ein, Eout, n_t, dims, th, 1, th + n_t, 1, Ld.inner}));
    };
    auto inner_bwd = [&](std::vector<Op>& ops, int n_t, int dims, int th, bool interleaved) {
      if (rqs) ops.push_back(mk(vm::OP_RQS_BWD, {Ein, n_t, dims, th, Kb, Ld.inner, Gout, Gin}, lo, hi));
      else if (interleaved) ops.push_back(mk(vm::OP_AFFINE_BWD, {Ein, n_t, dims, th, 2, th + 1, 2, Ld.inner, Gout, Gin}));
      else ops.push_back(mk(vm::OP_AFFINE_BWD, {Ein, n_t, dims, th, 1, th + n_t, 1, Ld.inner, Gout, Gin}));
    };

    auto xrows = [&](std::vector<int>& v) { for (int k = 0; k < dxu; ++k) v.push_back(X + k); };

    // decoder shared by SLICE / funnels: c = [z, x] -> Linear stack -> (mu | log_sigma)
    std::vector<Mat> dec_mats;
    const bool has_dec = Ld.kind == SURJ_LAYER_SLICE || Ld.kind == SURJ_LAYER_COUPLING_FUNNEL ||
                         Ld.kind == SURJ_LAYER_MAF_FUNNEL || Ld.kind == SURJ_LAYER_MLP_FUNNEL;
    if (has_dec) dec_mats = build_mlp(B, Ld.decoder, nullptr, nullptr);
    const int nv = D - K;
    auto emit_dec_fwd = [&](std::vector<Op>& ops, bool train, std::vector<NetRows>& rows, int& c_row) {
      const int dec_x = Ld.kind == SURJ_LAYER_MLP_FUNNEL ? 0 : dxu;  // mlp.py:53: decoder(z), no x
      const int cin = K + dec_x;
      c_row = B.scratch(pad8(cin));
      std::vector<int> src = iota(K, Eout);
      if (dec_x) xrows(src);
      ops.push_back(mk(vm::OP_GATHER, {c_row, cin, B.tab(src)}));
      rows = emit_mlp_fwd(B, ops, dec_mats, Ld.decoder.act, c_row, train);
    };

    switch (Ld.kind) {
      case SURJ_LAYER_COUPLING:
      case SURJ_LAYER_COUPLING_FUNNEL: {
        const bool funnel = Ld.kind == SURJ_LAYER_COUPLING_FUNNEL;
        std::vector<int> pass, tdim;
        for (int d = 0; d < D; ++d) {
          const bool p = funnel ? (d >= K) : (Ld.mask[d] != 0);
          (p ? pass : tdim).push_back(d);
        }
        const int n_pass = (int)pass.size(), n_t = (int)tdim.size();
        if (n_t == 0) {  // nothing is transformed: identity with zero log-det
          if (funnel) return fail("funnel without transformed dims");
          C.fwd.push_back(mk(vm::OP_COPY, {Eout, Ein, D}));
          C.bwd.push_back(mk(vm::OP_COPY, {Gin, Gout, D}));
          break;
        }
        std::vector<int> rsel = pass;
        for (int k = 0; k < dxu; ++k) rsel.push_back(D + k);
        // compact theta: affine -> [shift(n_t) | log_scale(n_t)]; spline -> [n_t][3K+1]
        std::vector<int> csel;
        if (rqs) {
          for (int k = 0; k < n_t; ++k)
            for (int p = 0; p < P; ++p) csel.push_back(tdim[k] * P + p);
        } else {
          for (int k = 0; k < n_t; ++k) csel.push_back(Ld.layout == SURJ_LAYOUT_SPLIT ? tdim[k] : 2 * tdim[k]);
          for (int k = 0; k < n_t; ++k) csel.push_back(Ld.layout == SURJ_LAYOUT_SPLIT ? D + tdim[k] : 2 * tdim[k] + 1);
        }
        const int cin = (int)rsel.size();
        if (cin == 0) return fail("conditioner without inputs");
        std::vector<Mat> mats = build_mlp(B, Ld.conditioner, &rsel, &csel);
        std::vector<int> msrc;
        for (int d : pass) msrc.push_back(Ein + d);
        xrows(msrc);
        const int t_dims = B.tab(tdim);
        auto emit_fwd = [&](std::vector<Op>& ops, bool train, std::vector<NetRows>& rows, int& m_row) {
          m_row = B.scratch(pad8(cin));
          ops.push_back(mk(vm::OP_GATHER, {m_row, cin, B.tab(msrc)}));
          rows = emit_mlp_fwd(B, ops, mats, Ld.conditioner.act, m_row, train);
        };
        {  // inference forward
          B.scratch_reset();
          std::vector<NetRows> rows, drows;
          int m_row, c_row;
          if (!funnel && n_pass > 0) C.fwd.push_back(mk(vm::OP_COPY, {Eout, Ein, D}));
          emit_fwd(C.fwd, false, rows, m_row);
          const int th = rows.back().post_row;
          inner_fwd(C.fwd, n_t, t_dims, th, false);
          if (funnel) {
            emit_dec_fwd(C.fwd, false, drows, c_row);
            const int thd = drows.back().post_row;
            C.fwd.push_back(mk(vm::OP_NORMAL_LP, {Ein + K, nv, thd, thd + nv}));
          }
        }
        if (funnel) {  // sampling direction (masked_coupling_inference_funnel.py:93-106): y- ~ decoder([z, x]), then the
                       // coupling forward of z conditioned on the sampled y- (+ x)
          B.scratch_reset();
          std::vector<NetRows> drows;
          int c_row;
          emit_dec_fwd(C.gen, false, drows, c_row);
          const int thd = drows.back().post_row;
          C.gen.push_back(mk(vm::OP_NORMAL_SAMPLE, {Ein + K, nv, thd, thd + nv, eps_col[l]}));
          const int m_row = B.scratch(pad8(cin));
          C.gen.push_back(mk(vm::OP_GATHER, {m_row, cin, B.tab(msrc)}));  // [y[K:], x]: rows of E[l] just written
          std::vector<NetRows> rows = emit_mlp_fwd(B, C.gen, mats, Ld.conditioner.act, m_row, false);
          const int th = rows.back().post_row;
          if (rqs) C.gen.push_back(mk(vm::OP_RQS, {Eout, Ein, n_t, t_dims, th, Kb, Ld.inner ^ 1, 0}, lo, hi));
          else C.gen.push_back(mk(vm::OP_AFFINE, {Eout, Ein, n_t, t_dims, th, 1, th + n_t, 1, Ld.inner ^ 1, 0}));
          C.gen_ok = true;
        }
        if (!funnel) {  // sampling direction (masked_coupling.py:59-71): same conditioner on the pass-through dims of z,
                        // inner bijector in the other direction (inner ^ 1 swaps X <-> Inverse(X))
          B.scratch_reset();
          std::vector<int> gsrc;
          for (int d : pass) gsrc.push_back(Eout + d);
          xrows(gsrc);
          const int m_row = B.scratch(pad8(cin));
          if (n_pass > 0) C.gen.push_back(mk(vm::OP_COPY, {Ein, Eout, D}));
          C.gen.push_back(mk(vm::OP_GATHER, {m_row, cin, B.tab(gsrc)}));
          std::vector<NetRows> rows = emit_mlp_fwd(B, C.gen, mats, Ld.conditioner.act, m_row, false);
          const int th = rows.back().post_row;
          if (rqs) C.gen.push_back(mk(vm::OP_RQS, {Eout, Ein, n_t, t_dims, th, Kb, Ld.inner ^ 1, 0}, lo, hi));
          else C.gen.push_back(mk(vm::OP_AFFINE, {Eout, Ein, n_t, t_dims, th, 1, th + n_t, 1, Ld.inner ^ 1, 0}));
          C.gen_ok = true;
        }
        {  // training forward + backward
          B.scratch_reset();
          std::vector<NetRows> rows, drows;
          int m_row, c_row = -1;
          if (funnel) {
            emit_dec_fwd(C.dec, true, drows, c_row);
            const int thd = drows.back().post_row;
            C.dec.push_back(mk(vm::OP_NORMAL_LP_BWD, {Ein + K, nv, thd, thd + nv, Gin + K}));
            emit_mlp_bwd(C.dec, drows, true);
            C.dec.push_back(mk(vm::OP_ADD, {Gout, c_row, K}));
            if (dxu) C.dec.push_back(mk(vm::OP_ADD, {GX, c_row + K, dxu}));
            B.scratch_reset();  // the decoder is done: the conditioner reuses its rows
          } else {
            // (emitted ahead of the training-mode recompute, not between it and the inner bijector's VJP: the per-op
            // executor fuses [conditioner products, spline VJP] into one launch when they are adjacent)
            C.dec.push_back(mk(vm::OP_COPY, {Gin, Gout, D}));
          }
          emit_fwd(C.train, true, rows, m_row);
          const int th = rows.back().post_row;
          inner_bwd(C.bwd, n_t, t_dims, th, false);
          emit_mlp_bwd(C.bwd, rows, true);
          if (n_pass > 0) {
            std::vector<int> dst;
            for (int d : pass) dst.push_back(Gin + d);
            C.bwd.push_back(mk(vm::OP_SCATTER_ADD, {m_row, n_pass, B.tab(dst)}));
          }
          if (dxu) C.bwd.push_back(mk(vm::OP_ADD, {GX, m_row + n_pass, dxu}));
        }
        break;
      }
      case SURJ_LAYER_MAF:
      case SURJ_LAYER_MAF_FUNNEL: {
        const bool funnel = Ld.kind == SURJ_LAYER_MAF_FUNNEL;
        const int Dm = funnel ? K : D;  // width the MADE transforms
        const int ctxw = funnel ? (D - K + dxu) : dxu;
        if (Ld.made.n_params != P) return fail("MADE n_params does not match the inner bijector");
        MadeMats mm = build_made(B, Ld.made);
        std::vector<int> csrc;
        if (funnel) for (int j = 0; j < D - K; ++j) csrc.push_back(Ein + K + j);
        xrows(csrc);
        const int t_dims = B.tab(iota(Dm));
        const int act = Ld.made.act;
        struct MadeRows { int cc_row, p0_row, h0_row; std::vector<NetRows> rest; };
        auto emit_fwd = [&](std::vector<Op>& ops, bool train, MadeRows& R) {
          R.cc_row = -1;
          R.h0_row = B.scratch(mm.l0.ld);
          const bool need_p0 = mm.has_ctx || (train && act != vm::ACT_IDENTITY);
          R.p0_row = need_p0 ? B.scratch(mm.l0.ld) : -1;
          if (mm.has_ctx) {
            R.cc_row = B.scratch(pad8(ctxw));
            ops.push_back(mk(vm::OP_GATHER, {R.cc_row, ctxw, B.tab(csrc)}));
            ops.push_back(lin_fwd(mm.l0, Ein, R.p0_row, vm::ACT_IDENTITY, 0, -1));
            ops.push_back(lin_fwd(mm.ctx, R.cc_row, R.h0_row, act, 1, R.p0_row));
          } else {
            ops.push_back(lin_fwd(mm.l0, Ein, R.h0_row, act, 0, R.p0_row));
          }
          R.rest = emit_mlp_fwd(B, ops, mm.rest, mm.rest_acts.data(), R.h0_row, train);
        };
        std::vector<Mat> dmats_unused;
        {
          B.scratch_reset();
          MadeRows R;
          emit_fwd(C.fwd, false, R);
          const int th = R.rest.back().post_row;
          inner_fwd(C.fwd, Dm, t_dims, th, true);
          if (funnel) {
            std::vector<NetRows> drows;
            int c_row;
            emit_dec_fwd(C.fwd, false, drows, c_row);
            const int thd = drows.back().post_row;
            C.fwd.push_back(mk(vm::OP_NORMAL_LP, {Ein + K, nv, thd, thd + nv}));
          }
        }
        if (!funnel) {  // sampling direction (masked_autoregressive.py:63-71): y = 0, then D passes of
                        // y = inner(MADE(y, x)).forward(z); the log-det of the last pass is the contribution
          B.scratch_reset();
          std::vector<Op> one;
          {
            // the MADE reads the current y estimate (rows Ein) instead of the layer input
            MadeRows R;
            R.cc_row = -1;
            R.h0_row = B.scratch(mm.l0.ld);
            R.p0_row = mm.has_ctx ? B.scratch(mm.l0.ld) : -1;
            if (mm.has_ctx) {
              R.cc_row = B.scratch(pad8(ctxw));
              one.push_back(mk(vm::OP_GATHER, {R.cc_row, ctxw, B.tab(csrc)}));
              one.push_back(lin_fwd(mm.l0, Ein, R.p0_row, vm::ACT_IDENTITY, 0, -1));
              one.push_back(lin_fwd(mm.ctx, R.cc_row, R.h0_row, act, 1, R.p0_row));
            } else {
              one.push_back(lin_fwd(mm.l0, Ein, R.h0_row, act, 0, -1));
            }
            R.rest = emit_mlp_fwd(B, one, mm.rest, mm.rest_acts.data(), R.h0_row, false);
            const int th = R.rest.back().post_row;
            if (rqs) one.push_back(mk(vm::OP_RQS, {Eout, Ein, Dm, t_dims, th, Kb, Ld.inner ^ 1, 1}, lo, hi));
            else one.push_back(mk(vm::OP_AFFINE, {Eout, Ein, Dm, t_dims, th, 2, th + 1, 2, Ld.inner ^ 1, 1}));
          }
          C.gen.push_back(mk(vm::OP_ZERO, {Ein, Dm}));
          for (int it = 0; it < Dm; ++it) {
            C.gen.insert(C.gen.end(), one.begin(), one.end());
            if (it == Dm - 1) C.gen.back().a[rqs ? 7 : 9] = 0;  // only the last pass contributes its log-det
          }
          C.gen_ok = true;
        }
        {
          B.scratch_reset();
          MadeRows R;
          if (funnel) {
            std::vector<NetRows> drows;
            int c_row;
            emit_dec_fwd(C.dec, true, drows, c_row);
            const int thd = drows.back().post_row;
            C.dec.push_back(mk(vm::OP_NORMAL_LP_BWD, {Ein + K, nv, thd, thd + nv, Gin + K}));
            emit_mlp_bwd(C.dec, drows, true);
            C.dec.push_back(mk(vm::OP_ADD, {Gout, c_row, K}));
            if (dxu) C.dec.push_back(mk(vm::OP_ADD, {GX, c_row + K, dxu}));
            B.scratch_reset();  // the decoder is done: the MADE reuses its rows
          }
          emit_fwd(C.train, true, R);
          const int th = R.rest.back().post_row;
          inner_bwd(C.bwd, Dm, t_dims, th, true);
          emit_mlp_bwd(C.bwd, R.rest, true);  // cotangent of h0 lands in the h0 rows
          if (act != vm::ACT_IDENTITY) C.bwd.push_back(mk(vm::OP_ACT_BWD, {R.h0_row, R.p0_row, mm.l0.n_out, act}));
          C.bwd.push_back(mk(vm::OP_LINEAR_DW, {Ein, Dm, R.h0_row, mm.l0.n_out, mm.l0.w_off, mm.l0.ld, mm.l0.b_off}));
          // g_y += g_pre0 W0^T, accumulated straight into the cotangent rows AFFINE_BWD just wrote
          C.bwd.push_back(lin_bwd_in(mm.l0, R.h0_row, Gin, 1));
          if (mm.has_ctx) {
            C.bwd.push_back(mk(vm::OP_LINEAR_DW, {R.cc_row, ctxw, R.h0_row, mm.ctx.n_out, mm.ctx.w_off, mm.ctx.ld, mm.ctx.b_off}));
            C.bwd.push_back(lin_bwd_in(mm.ctx, R.h0_row, R.cc_row, 0));
            if (funnel) {
              C.bwd.push_back(mk(vm::OP_ADD, {Gin + K, R.cc_row, D - K}));
              if (dxu) C.bwd.push_back(mk(vm::OP_ADD, {GX, R.cc_row + (D - K), dxu}));
            } else if (dxu) {
              C.bwd.push_back(mk(vm::OP_ADD, {GX, R.cc_row, dxu}));
            }
          }
        }
        break;
      }
      case SURJ_LAYER_SLICE: {
        {
          B.scratch_reset();
          std::vector<NetRows> drows;
          int c_row;
          emit_dec_fwd(C.fwd, false, drows, c_row);
          const int thd = drows.back().post_row;
          C.fwd.push_back(mk(vm::OP_NORMAL_LP, {Ein + K, nv, thd, thd + nv}));
        }
        {
          B.scratch_reset();
          std::vector<NetRows> drows;
          int c_row;
          emit_dec_fwd(C.train, true, drows, c_row);
          const int thd = drows.back().post_row;
          C.bwd.push_back(mk(vm::OP_NORMAL_LP_BWD, {Ein + K, nv, thd, thd + nv, Gin + K}));
          emit_mlp_bwd(C.bwd, drows, true);
          C.bwd.push_back(mk(vm::OP_COPY, {Gin, Gout, K}));
          C.bwd.push_back(mk(vm::OP_ADD, {Gin, c_row, K}));
          if (dxu) C.bwd.push_back(mk(vm::OP_ADD, {GX, c_row + K, dxu}));
        }
        {  // sampling direction (slice.py:70-78): y = [z, y-] with y- ~ decoder([z, x]); contribution = log p(y- | z)
          B.scratch_reset();
          std::vector<NetRows> drows;
          int c_row;
          emit_dec_fwd(C.gen, false, drows, c_row);  // z already sits in the first K rows (E[l+1] aliases E[l])
          const int thd = drows.back().post_row;
          C.gen.push_back(mk(vm::OP_NORMAL_SAMPLE, {Ein + K, nv, thd, thd + nv, eps_col[l]}));
          C.gen_ok = true;
        }
        break;
      }
      case SURJ_LAYER_MLP_FUNNEL: {
        // mlp.py:47-56: z = LULinear(y+) + Linear(y-); lc = decoder(z).log_prob(y-) + log|det|
        const int n = K;
        const int slot = B.n_scalars++;
        const int64_t lo_ = Ld.lu_lower_off, up_ = Ld.lu_upper_off, dg_ = Ld.lu_diag_off;
        auto usrc = [&](int r, int c) -> Builder::Ent {
          if (r < c) { const int p = (int)(up_ + (int64_t)r * (2 * n - r - 1) / 2 + (c - r - 1)); return {p, p, -1}; }
          if (r == c) { const int code = -((int)dg_ + r) - 3; return {code, code, slot}; }
          return {vm::STAGE_ZERO, -1, -1};
        };
        auto lsrc = [&](int r, int c) -> Builder::Ent {
          if (r > c) { const int p = (int)(lo_ + (int64_t)r * (r - 1) / 2 + c); return {p, p, -1}; }
          if (r == c) return {vm::STAGE_ONE, -1, -1};
          return {vm::STAGE_ZERO, -1, -1};
        };
        const Mat U = B.add_matrix(n, n, usrc, nullptr);
        const Mat Lm = B.add_matrix(n, n, lsrc, nullptr);
        B.sjobs.push_back({slot, U.w_off, U.ld + 1, n});
        const Mat Wp = build_mlp(B, Ld.conditioner, nullptr, nullptr)[0];
        {
          B.scratch_reset();
          const int t = B.scratch(pad8(n));
          C.fwd.push_back(lin_fwd(U, Ein, t, vm::ACT_IDENTITY, 0, -1));
          C.fwd.push_back(lin_fwd(Lm, t, Eout, vm::ACT_IDENTITY, 0, -1));
          C.fwd.push_back(lin_fwd(Wp, Ein + K, Eout, vm::ACT_IDENTITY, 1, -1));
          C.fwd.push_back(mk(vm::OP_ADD_SCALAR, {slot}));
          std::vector<NetRows> drows;
          int c_row;
          emit_dec_fwd(C.fwd, false, drows, c_row);
          const int thd = drows.back().post_row;
          C.fwd.push_back(mk(vm::OP_NORMAL_LP, {Ein + K, nv, thd, thd + nv}));
        }
        {
          B.scratch_reset();
          const int t = B.scratch(pad8(n)), gt = B.scratch(pad8(n)), gv = B.scratch(pad8(nv)), gw = B.scratch(pad8(nv));
          C.train.push_back(lin_fwd(U, Ein, t, vm::ACT_IDENTITY, 0, -1));
          std::vector<NetRows> drows;
          int c_row;
          emit_dec_fwd(C.train, true, drows, c_row);
          const int thd = drows.back().post_row;
          C.bwd.push_back(mk(vm::OP_NORMAL_LP_BWD, {Ein + K, nv, thd, thd + nv, gv}));
          emit_mlp_bwd(C.bwd, drows, true);
          C.bwd.push_back(mk(vm::OP_ADD, {Gout, c_row, K}));                                        // g_z += decoder input gradient
          C.bwd.push_back(mk(vm::OP_LINEAR_DW, {Ein + K, nv, Gout, Wp.n_out, Wp.w_off, Wp.ld, Wp.b_off}));  // g_W' = y-^T g_z, g_b
          C.bwd.push_back(lin_bwd_in(Wp, Gout, gw, 0));                                             // g_z W'^T
          C.bwd.push_back(lin_bwd_in(Lm, Gout, gt, 0));                                             // g_t = g_z L^T
          C.bwd.push_back(mk(vm::OP_LINEAR_DW, {t, n, Gout, n, Lm.w_off, Lm.ld, -1}));
          C.bwd.push_back(lin_bwd_in(U, gt, Gin, 0));                                               // g_y+ = g_t U^T
          C.bwd.push_back(mk(vm::OP_LINEAR_DW, {Ein, n, gt, n, U.w_off, U.ld, -1}));
          C.bwd.push_back(mk(vm::OP_SCALAR_BWD, {slot}));
          C.bwd.push_back(mk(vm::OP_COPY, {Gin + K, gv, nv}));                                      // g_y- = decoder score + W' path
          C.bwd.push_back(mk(vm::OP_ADD, {Gin + K, gw, nv}));
        }
        break;
      }
      case SURJ_LAYER_LU: {
        // lu_linear.py:68-90: z = (y @ U) @ L, U upper with diag softplus(d) + 1e-4, L unit lower,
        // entries in row-major tril_indices(n, -1) / triu_indices(n, 1) order
        const int n = D;
        const int slot = B.n_scalars++;
        const int64_t lo = Ld.lu_lower_off, up = Ld.lu_upper_off, dg = Ld.lu_diag_off;
        auto usrc = [&](int r, int c) -> Builder::Ent {
          if (r < c) { const int p = (int)(up + (int64_t)r * (2 * n - r - 1) / 2 + (c - r - 1)); return {p, p, -1}; }
          if (r == c) { const int code = -((int)dg + r) - 3; return {code, code, slot}; }
          return {vm::STAGE_ZERO, -1, -1};
        };
        auto lsrc = [&](int r, int c) -> Builder::Ent {
          if (r > c) { const int p = (int)(lo + (int64_t)r * (r - 1) / 2 + c); return {p, p, -1}; }
          if (r == c) return {vm::STAGE_ONE, -1, -1};
          return {vm::STAGE_ZERO, -1, -1};
        };
        const Mat U = B.add_matrix(n, n, usrc, nullptr);
        const Mat Lm = B.add_matrix(n, n, lsrc, nullptr);
        B.sjobs.push_back({slot, U.w_off, U.ld + 1, n});
        {
          B.scratch_reset();
          const int t = B.scratch(pad8(n));
          C.fwd.push_back(lin_fwd(U, Ein, t, vm::ACT_IDENTITY, 0, -1));
          C.fwd.push_back(lin_fwd(Lm, t, Eout, vm::ACT_IDENTITY, 0, -1));
          C.fwd.push_back(mk(vm::OP_ADD_SCALAR, {slot}));
        }
        {
          B.scratch_reset();
          const int t = B.scratch(pad8(n)), gt = B.scratch(pad8(n));
          C.train.push_back(lin_fwd(U, Ein, t, vm::ACT_IDENTITY, 0, -1));
          C.bwd.push_back(lin_bwd_in(Lm, Gout, gt, 0));                                                    // g_t = g_z L^T
          C.bwd.push_back(mk(vm::OP_LINEAR_DW, {t, n, Gout, n, Lm.w_off, Lm.ld, -1}));                     // g_L = t^T g_z
          C.bwd.push_back(lin_bwd_in(U, gt, Gin, 0));                                                      // g_y = g_t U^T
          C.bwd.push_back(mk(vm::OP_LINEAR_DW, {Ein, n, gt, n, U.w_off, U.ld, -1}));                       // g_U = y^T g_t
          C.bwd.push_back(mk(vm::OP_SCALAR_BWD, {slot}));
        }
        break;
      }
      case SURJ_LAYER_PERMUTATION: {
        std::vector<int32_t> inv(D);
        surj_permutation_inverse(Ld.permutation, D, inv.data());
        std::vector<int> f(D), b(D);
        for (int d = 0; d < D; ++d) { f[d] = Ein + inv[d]; b[d] = Gout + Ld.permutation[d]; }
        C.fwd.push_back(mk(vm::OP_GATHER, {Eout, D, B.tab(f)}));   // z = y[..., inv]  (permutation.py:37)
        C.bwd.push_back(mk(vm::OP_GATHER, {Gin, D, B.tab(b)}));    // g_y = g_z[..., perm]
        std::vector<int> gsel(D);
        for (int d = 0; d < D; ++d) gsel[d] = Eout + Ld.permutation[d];
        C.gen.push_back(mk(vm::OP_GATHER, {Ein, D, B.tab(gsel)}));  // y = z[..., perm]  (permutation.py:27-28)
        C.gen_ok = true;
        break;
      }
      default:
        return fail("unknown layer kind");
    }
  }

  VmPlan* v = new VmPlan();
  const int dz = plan->dim_out;
  auto head = [&](std::vector<Op>& ops) {
    ops.push_back(mk(vm::OP_LOAD, {E[0], plan->dim_in, 0}));
    if (dx > 0) ops.push_back(mk(vm::OP_LOAD, {X, dx, 1}));
    for (int l = 0; l < L; ++l) ops.insert(ops.end(), code[l].fwd.begin(), code[l].fwd.end());
    ops.push_back(mk(vm::OP_BASE, {E[L], dz}));
    ops.push_back(mk(vm::OP_STORE, {E[L], dz, 0}));
    ops.push_back(mk(vm::OP_STORE_LP, {}));
  };
  head(v->ops_fwd);
  head(v->ops_bwd);
  v->ops_bwd.push_back(mk(vm::OP_BASE_BWD, {E[L], dz, G[0]}));
  if (dx > 0) v->ops_bwd.push_back(mk(vm::OP_ZERO, {GX, dx}));
  for (int l = L - 1; l >= 0; --l) {
    v->ops_bwd.insert(v->ops_bwd.end(), code[l].dec.begin(), code[l].dec.end());
    v->ops_bwd.insert(v->ops_bwd.end(), code[l].train.begin(), code[l].train.end());
    v->ops_bwd.insert(v->ops_bwd.end(), code[l].bwd.begin(), code[l].bwd.end());
  }
  v->ops_bwd.push_back(mk(vm::OP_STORE, {G[L & 1], plan->dim_in, 1}));
  if (dx > 0) v->ops_bwd.push_back(mk(vm::OP_STORE, {GX, dx, 2}));
  // sampling direction: latent z -> y through the layers in reverse (chain.py:47-54); only when every layer
  // has a forward (LULinear has none upstream, lu_linear.py:92-93; surjections need a decoder sample)
  bool gen_ok = true;
  for (int l = 0; l < L; ++l)
    if (!code[l].gen_ok) {
      gen_ok = false;
      char buf[96];
      snprintf(buf, sizeof buf, "layer %d (kind %d) has no forward / sampling direction here", l, plan->layers[l].kind);
      v->gen_why = buf;
      break;
    }
  if (gen_ok) {
    v->ops_gen.push_back(mk(vm::OP_LOAD, {E[L], dz, 0}));
    if (dx > 0) v->ops_gen.push_back(mk(vm::OP_LOAD, {X, dx, 1}));
    for (int l = L - 1; l >= 0; --l) v->ops_gen.insert(v->ops_gen.end(), code[l].gen.begin(), code[l].gen.end());
    v->ops_gen.push_back(mk(vm::OP_STORE, {E[0], plan->dim_in, 0}));
    v->ops_gen.push_back(mk(vm::OP_STORE_LP, {}));
  }

  while (B.stage.size() & 3) B.push(vm::STAGE_ZERO, -1);
  v->stage = B.stage; v->grad_dst = B.grad_dst; v->grad_aux = B.grad_aux; v->itab = B.itab; v->sjobs = B.sjobs;
  v->n_scalars = B.n_scalars;
  v->n_eps = n_eps;
  v->n_rows = B.scratch_base + (B.scratch_max - B.scratch_base);
  v->d_in = plan->dim_in; v->d_x = dx; v->d_out = dz; v->n_params = (int)plan->n_params;

  // ---- fused-kernel launch shape: parameters + one tile of rows must fit shared memory, with >= 2 CTAs per SM
  // when possible; chains that do not fit run op by op out of global memory (vm_global.cu)
  const size_t cap = 227 * 1024, want = 110 * 1024;
  const int cands[][2] = {{64, 1}, {32, 1}};
  for (int ci = 0; ci < 2 && !v->fused_ok; ++ci) {
    const size_t need = vm_smem_bytes(*v, VM_PROG_BWD, cands[ci][0] * cands[ci][1]);
    if (need <= want || (ci == 1 && need <= cap)) {
      v->nt = cands[ci][0]; v->s = cands[ci][1]; v->fused_ok = true;
    }
  }
  if (!v->fused_ok) v->fused_why = "parameters + activations of one tile exceed the shared memory of an SM";
  v->smem_fwd = vm_smem_bytes(*v, VM_PROG_FWD, v->nt * v->s);
  v->smem_bwd = vm_smem_bytes(*v, VM_PROG_BWD, v->nt * v->s);
  v->smem_gen = vm_smem_bytes(*v, VM_PROG_GEN, v->nt * v->s);
  return v;
}

void vm_destroy(VmPlan* v) {
  if (!v) return;
  if (v->d_blob) cudaFree(v->d_blob);
  delete v;
}

typedef void (*VmKernel)(vm::Program, vm::Io, const float*, float*, int, long long, unsigned*, int);
static VmKernel vm_kernel_for(int nt, int s) {
  if (nt == 64 && s == 1) return vm::vm_kernel<64, 1>;
  if (nt == 32 && s == 1) return vm::vm_kernel<32, 1>;
  return nullptr;
}

int vm_prepare(VmPlan* v) {
  if (v->uploaded) return SURJ_OK;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return surj_cuda_fail(e, "cudaGetDevice");
  cudaDeviceGetAttribute(&v->n_sm, cudaDevAttrMultiProcessorCount, dev);
  auto a16 = [](size_t n) { return (n + 15) & ~(size_t)15; };
  const size_t b_fwd = a16(v->ops_fwd.size() * sizeof(Op)), b_bwd = a16(v->ops_bwd.size() * sizeof(Op));
  const size_t b_gen = a16(std::max<size_t>(v->ops_gen.size(), 1) * sizeof(Op));
  const size_t b_stage = a16(v->stage.size() * 4), b_itab = a16(std::max<size_t>(v->itab.size(), 1) * 4);
  const size_t b_sj = a16(std::max<size_t>(v->sjobs.size(), 1) * sizeof(vm::ScalarJob));
  const size_t total = b_fwd + b_bwd + b_gen + 3 * b_stage + b_itab + b_sj;
  std::vector<unsigned char> host(total, 0);
  size_t o = 0;
  auto put = [&](const void* p, size_t n, size_t slot) { if (n) std::memcpy(host.data() + o, p, n); const size_t r = o; o += slot; return r; };
  const size_t o_fwd = put(v->ops_fwd.data(), v->ops_fwd.size() * sizeof(Op), b_fwd);
  const size_t o_bwd = put(v->ops_bwd.data(), v->ops_bwd.size() * sizeof(Op), b_bwd);
  const size_t o_gen = put(v->ops_gen.data(), v->ops_gen.size() * sizeof(Op), b_gen);
  const size_t o_stage = put(v->stage.data(), v->stage.size() * 4, b_stage);
  const size_t o_gd = put(v->grad_dst.data(), v->grad_dst.size() * 4, b_stage);
  const size_t o_ga = put(v->grad_aux.data(), v->grad_aux.size() * 4, b_stage);
  const size_t o_it = put(v->itab.data(), v->itab.size() * 4, b_itab);
  const size_t o_sj = put(v->sjobs.data(), v->sjobs.size() * sizeof(vm::ScalarJob), b_sj);
  e = cudaMalloc(&v->d_blob, total);
  if (e != cudaSuccess) return surj_cuda_fail(e, "cudaMalloc(plan tables)");
  e = cudaMemcpy(v->d_blob, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return surj_cuda_fail(e, "cudaMemcpy(plan tables)");
  unsigned char* d = (unsigned char*)v->d_blob;
  v->d_ops_fwd = (const Op*)(d + o_fwd);
  v->d_ops_bwd = (const Op*)(d + o_bwd);
  v->d_ops_gen = (const Op*)(d + o_gen);
  vm::Program& P = v->prog;
  P.ops = nullptr;
  P.n_ops = 0;
  P.stage = (const int32_t*)(d + o_stage); P.n_par = (int)v->stage.size(); P.n_prim = (int)v->stage.size();
  P.grad_dst = (const int32_t*)(d + o_gd); P.grad_aux = (const int32_t*)(d + o_ga);
  P.itab = (const int32_t*)(d + o_it); P.n_itab = (int)v->itab.size();
  P.sjobs = (const vm::ScalarJob*)(d + o_sj); P.n_sjobs = (int)v->sjobs.size(); P.n_scalars = v->n_scalars;
  P.n_rows = v->n_rows; P.d_in = v->d_in; P.d_x = v->d_x; P.d_out = v->d_out; P.n_params = v->n_params;
  if (!v->fused_ok) {
    v->uploaded = true;
    return SURJ_OK;
  }
  VmKernel k = vm_kernel_for(v->nt, v->s);
  // the attribute is per kernel, not per plan: always raise it to the device limit so plans with
  // different footprints can share the kernel
  int optin = 227 * 1024;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if ((size_t)optin < std::max(std::max(v->smem_fwd, v->smem_bwd), v->smem_gen)) {
    surj_set_error("fused kernel needs more shared memory than this device offers");
    return SURJ_E_CUDA;
  }
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
  if (e != cudaSuccess) return surj_cuda_fail(e, "cudaFuncSetAttribute(vm_kernel)");
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v->occ_fwd, k, v->nt, v->smem_fwd);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v->occ_bwd, k, v->nt, v->smem_bwd);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v->occ_gen, k, v->nt, v->smem_gen);
  if (e != cudaSuccess) return surj_cuda_fail(e, "occupancy(vm_kernel)");
  if (v->occ_fwd < 1 || v->occ_bwd < 1) {
    surj_set_error("vm_kernel does not fit on an SM");
    return SURJ_E_CUDA;
  }
  v->uploaded = true;
  return SURJ_OK;
}

int vm_grid(const VmPlan* v, int64_t B, int prog) {
  const int T = v->nt * v->s;
  const int64_t tiles = (B + T - 1) / T;
  const int64_t cap = (int64_t)v->n_sm * (prog == VM_PROG_BWD ? v->occ_bwd : (prog == VM_PROG_GEN ? v->occ_gen : v->occ_fwd));
  return (int)std::min<int64_t>(tiles, cap);
}

int vm_max_grid(const VmPlan* v, int64_t B) {  // upper bound usable before a device is known
  const int T = v->nt * v->s;
  const int64_t tiles = (B + T - 1) / T;
  return (int)std::min<int64_t>(tiles, 148 * 16);
}

int vm_launch(VmPlan* v, cudaStream_t st, int prog, int with_base, const float* y, const float* x,
              const float* params, const float* g_lp, float g_scale, const float* g_z, float* z_out, float* lp_out,
              float* g_y, float* g_x, float* partials, int npad, int64_t B, int grid, unsigned* tickets,
              int n_tickets, const float* eps) {
  const bool backward = prog == VM_PROG_BWD;
  vm::Program P = v->prog;
  P.ops = backward ? v->d_ops_bwd : (prog == VM_PROG_GEN ? v->d_ops_gen : v->d_ops_fwd);
  P.n_ops = (int)(backward ? v->ops_bwd.size() : (prog == VM_PROG_GEN ? v->ops_gen.size() : v->ops_fwd.size()));
  vm::Io io;
  io.y = y; io.x = x; io.g_lp = g_lp; io.g_z = g_z; io.z_out = z_out; io.lp_out = lp_out; io.g_y = g_y; io.g_x = g_x;
  io.g_scale = g_scale; io.with_base = with_base; io.backward = backward ? 1 : 0;
  io.eps = eps; io.n_eps = v->n_eps;
  VmKernel k = vm_kernel_for(v->nt, v->s);
  k<<<grid, v->nt, backward ? v->smem_bwd : (prog == VM_PROG_GEN ? v->smem_gen : v->smem_fwd), st>>>(P, io, params, partials, npad, (long long)B, tickets,
                                                                n_tickets);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return surj_cuda_fail(e, "vm_kernel launch");
  return SURJ_OK;
}

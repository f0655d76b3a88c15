// Host side of the compact fused kernel (compact.cuh): decides whether a chain is in its family, builds the
// staged-parameter gather table, the row map and the layer table, picks the CTA size, and launches.
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

#include "compact.cuh"
#include "compact.h"
#include "compact_variants.h"
#include "plan.h"

namespace {

using cc::Layer;
using cc::Mlp;
using cc::Program;

using cc::Kernel;
using cc::Variant;

// the kernel instantiations live in compact_k{1,2,3}.cu (one activation each) so that they compile in parallel
const Variant* find_variant(int hc, int hd, int act) {
  const Variant* tabs[] = {cc::variants_relu(), cc::variants_swish(), cc::variants_tanh()};
  for (const Variant* t : tabs)
    for (int i = 0; i < cc::N_VARIANTS; ++i)
      if (t[i].hc == hc && t[i].act == act && (hd == 0 || t[i].hd == hd)) return t + i;
  return nullptr;
}

int pad4(int n) { return (n + 3) & ~3; }

// hk.nets.MLP with exactly two equal hidden layers, the same activation on both, a linear output layer, biases everywhere
bool mlp_shape(const SurjMlpDesc& m, int* h, int* act, std::string* why, const char* what) {
  if (m.n_layers != 3 || m.widths[1] != m.widths[2] || m.act[0] != m.act[1] || m.act[2] != SURJ_ACT_IDENTITY) {
    *why = std::string(what) + " is not an MLP with two equal hidden layers and a linear output";
    return false;
  }
  for (int l = 0; l < 3; ++l)
    if (m.b_off[l] < 0) {
      *why = std::string(what) + " has a layer without bias";
      return false;
    }
  *h = m.widths[1];
  *act = m.act[0];
  return true;
}

}  // namespace

struct CcPlan {
  Program prog;  // device pointers filled by cc_prepare
  std::vector<int> stage_src, itab;
  int hc = 0, hd = 0, act = 0, n_sm = 148, max_smem = 0;
  Kernel fn = nullptr;
  int* d_stage = nullptr;
  int* d_itab = nullptr;
  bool ready = false;
};

CcPlan* cc_compile(const SurjPlan* plan, std::string* why) {
  const int n = (int)plan->layers.size();
  if (n > cc::MAX_LAYERS) { *why = "more layers than the compact kernel's table holds"; return nullptr; }
  if (plan->dim_x != 0) { *why = "conditional chains are not in the compact kernel's family"; return nullptr; }
  int hc = 0, hd = 0, act = -1;
  for (const SurjLayerDesc& L : plan->layers) {
    if (L.kind != SURJ_LAYER_COUPLING && L.kind != SURJ_LAYER_COUPLING_FUNNEL) { *why = "layer kind outside {MaskedCoupling, MaskedCouplingInferenceFunnel}"; return nullptr; }
    if (L.inner != SURJ_INNER_AFFINE && L.inner != SURJ_INNER_AFFINE_INV) { *why = "inner bijector is not affine"; return nullptr; }
    if (L.dim_in > 64) { *why = "event wider than 64"; return nullptr; }
    int h, a;
    if (!mlp_shape(L.conditioner, &h, &a, why, "conditioner")) return nullptr;
    if ((hc && h != hc) || (act >= 0 && a != act)) { *why = "conditioners differ in hidden width or activation"; return nullptr; }
    hc = h; act = a;
    if (L.kind == SURJ_LAYER_COUPLING_FUNNEL) {
      if (!mlp_shape(L.decoder, &h, &a, why, "decoder")) return nullptr;
      if ((hd && h != hd) || a != act) { *why = "decoders differ in hidden width or activation"; return nullptr; }
      hd = h;
    }
  }
  const Variant* var = find_variant(hc, hd, act);
  if (!var) { *why = "no compact kernel instantiated for these hidden widths / activation"; return nullptr; }

  CcPlan* c = new CcPlan();
  c->hc = var->hc; c->hd = var->hd; c->act = act; c->fn = var->fn;
  Program& P = c->prog;
  P = Program{};
  P.n_layers = n; P.d_in = plan->dim_in; P.d_out = plan->dim_out; P.n_params = (int)plan->n_params;
  std::vector<int>& src = c->stage_src;
  std::vector<int>& itab = c->itab;
  auto stage_block = [&](int count) { int off = (int)src.size(); src.resize(off + pad4(count), -1); return off; };
  // rows: 0 = ones, then one block per layer input and the final output
  int row = 1;
  std::vector<int> u_row(n + 1);
  u_row[0] = row; row += plan->dim_in;
  for (int l = 0; l < n; ++l) { u_row[l + 1] = row; row += plan->layers[l].dim_out; }
  int th_rows = std::max(c->hc, c->hd), d_max = plan->dim_in;
  // first pass over the layers: staged parameters, index lists (rows of the cotangent block are patched below)
  struct Pending { int layer; bool dec; int n; std::vector<int> rel; };
  std::vector<Pending> gin;
  for (int l = 0; l < n; ++l) {
    const SurjLayerDesc& D = plan->layers[l];
    Layer& L = P.layers[l];
    L = Layer{};
    L.d_in = D.dim_in; L.d_out = D.dim_out; L.inner = D.inner; L.u_in = u_row[l]; L.u_out = u_row[l + 1];
    d_max = std::max(d_max, D.dim_in);
    std::vector<int> tdims, pdims;
    const bool funnel = D.kind == SURJ_LAYER_COUPLING_FUNNEL;
    for (int d = 0; d < D.dim_in; ++d) {
      const bool pass = funnel ? d >= D.dim_out : D.mask[d] != 0;  // funnel mask: arange(D) >= n_keep
      (pass ? pdims : tdims).push_back(d);
    }
    L.n_t = (int)tdims.size(); L.n_p = (int)pdims.size(); L.n_v = funnel ? D.dim_in - D.dim_out : 0;
    L.t_idx = (int)itab.size(); itab.insert(itab.end(), tdims.begin(), tdims.end());
    L.p_idx = (int)itab.size(); itab.insert(itab.end(), pdims.begin(), pdims.end());
    auto stage_mlp = [&](const SurjMlpDesc& m, int H, const std::vector<int>& in_dims, const std::vector<int>& out_cols, Mlp& M) {
      const int wout = m.widths[3];
      M.n_in = (int)in_dims.size();
      M.w1 = stage_block(M.n_in * H);
      for (int k = 0; k < M.n_in; ++k)
        for (int j = 0; j < H; ++j) src[M.w1 + k * H + j] = (int)m.w_off[0] + in_dims[k] * H + j;
      M.b1 = stage_block(H);
      for (int j = 0; j < H; ++j) src[M.b1 + j] = (int)m.b_off[0] + j;
      M.w2 = stage_block(H * H);
      for (int i = 0; i < H * H; ++i) src[M.w2 + i] = (int)m.w_off[1] + i;
      M.b2 = stage_block(H);
      for (int j = 0; j < H; ++j) src[M.b2 + j] = (int)m.b_off[1] + j;
      M.n_out = pad4((int)out_cols.size());
      M.w3t = stage_block(M.n_out * H);
      M.b3 = stage_block(M.n_out);
      for (size_t o = 0; o < out_cols.size(); ++o) {
        for (int i = 0; i < H; ++i) src[M.w3t + (int)o * H + i] = (int)m.w_off[2] + i * wout + out_cols[o];
        src[M.b3 + (int)o] = (int)m.b_off[2] + out_cols[o];
      }
    };
    // conditioner: inputs = pass-through dims (the masked input is zero elsewhere, masked_coupling.py:75-77);
    // outputs needed = (shift, log_scale) of the transformed dims in bijector_fn's layout
    std::vector<int> cols;
    for (int d : tdims) cols.push_back(D.layout == SURJ_LAYOUT_SPLIT ? d : 2 * d);
    for (int d : tdims) cols.push_back(D.layout == SURJ_LAYOUT_SPLIT ? D.dim_in + d : 2 * d + 1);
    stage_mlp(D.conditioner, c->hc, pdims, cols, L.cond);
    L.cond.in_idx = (int)itab.size();
    for (int d : pdims) itab.push_back(L.u_in + d);
    gin.push_back({l, false, L.n_p, pdims});
    th_rows = std::max(th_rows, L.cond.n_out);
    if (funnel) {
      std::vector<int> in_dims(D.dim_out), dcols(2 * L.n_v);
      for (int k = 0; k < D.dim_out; ++k) in_dims[k] = k;
      for (int j = 0; j < 2 * L.n_v; ++j) dcols[j] = j;  // jnp.split(decoder(z), 2, -1): [mu | log_sigma]
      stage_mlp(D.decoder, c->hd, in_dims, dcols, L.dec);
      L.dec.in_idx = (int)itab.size();
      for (int k = 0; k < D.dim_out; ++k) itab.push_back(L.u_out + k);
      gin.push_back({l, true, D.dim_out, in_dims});
      th_rows = std::max(th_rows, L.dec.n_out);
    }
  }
  P.n_staged = (int)src.size();
  P.row_th = row; row += th_rows;
  P.n_rows_fwd = row;
  P.row_g = row; row += d_max;
  P.row_park = row; row += 2 * std::max(c->hc, c->hd);
  P.n_rows_bwd = row;
  for (const Pending& g : gin) {
    Mlp& M = g.dec ? P.layers[g.layer].dec : P.layers[g.layer].cond;
    M.gin_idx = (int)itab.size();
    for (int k = 0; k < g.n; ++k) itab.push_back(P.row_g + g.rel[k]);
  }
  for (int i = 0; i < std::max(c->hc, c->hd); ++i) itab.push_back(P.row_park + i);  // must stay last (compact.cuh: park_list)
  P.n_itab = (int)itab.size();
  return c;
}

void cc_destroy(CcPlan* c) {
  if (!c) return;
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->d_itab) cudaFree(c->d_itab);
  delete c;
}

static size_t cc_smem(const CcPlan* c, int nt, bool backward) {
  const Program& P = c->prog;
  const size_t ns4 = pad4(P.n_staged);
  const size_t rows = backward ? P.n_rows_bwd : P.n_rows_fwd;
  const size_t jobs = backward ? 6 * (size_t)P.n_layers * (sizeof(cc::Job) / 4) : 0;
  return (ns4 + (backward ? ns4 + 4 : 0) + pad4(P.n_itab) + jobs + rows * (2 * (size_t)nt + 4)) * 4;
}

int cc_prepare(CcPlan* c) {
  if (c->ready) return SURJ_OK;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_stage, c->stage_src.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_itab, c->itab.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(c->d_stage, c->stage_src.data(), c->stage_src.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(c->d_itab, c->itab.data(), c->itab.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return surj_cuda_fail(e, "compact kernel setup");
  c->prog.stage_src = c->d_stage;
  c->prog.itab = c->d_itab;
  if (cc_smem(c, 32, true) > (size_t)c->max_smem) {
    surj_set_error("compact kernel: one 64-sample tile does not fit in shared memory");
    return SURJ_E_UNSUPPORTED;
  }
  c->ready = true;
  return SURJ_OK;
}

// Threads per CTA (a multiple of 32, each thread owns two samples) and CTAs: the largest tile that fits in shared
// memory, shrunk so that the passes over the batch are evenly filled (one CTA per SM is resident).
void cc_shape(const CcPlan* c, int64_t B, bool backward, int* nt_out, int* grid_out) {
  int nt_max = 256;
  while (nt_max > 32 && cc_smem(c, nt_max, backward) > (size_t)c->max_smem) nt_max -= 32;
  const int64_t pairs = (B + 1) / 2;
  const int64_t passes = (pairs + (int64_t)nt_max * c->n_sm - 1) / ((int64_t)nt_max * c->n_sm);
  int64_t nt = (pairs + passes * c->n_sm - 1) / (passes * c->n_sm);
  nt = std::min<int64_t>(nt_max, std::max<int64_t>(32, (nt + 31) / 32 * 32));
  const int64_t tiles = (B + 2 * nt - 1) / (2 * nt);
  *nt_out = (int)nt;
  *grid_out = (int)std::min<int64_t>(tiles, c->n_sm);
}

int cc_max_grid(const CcPlan* c) { return c->n_sm; }

int cc_launch(CcPlan* c, cudaStream_t st, int backward, int with_base, const float* y, const float* params, const float* g_lp,
              float g_scale, const float* g_z, float* z_out, float* lp_out, float* g_y, float* partials, int npad, int64_t B,
              unsigned* tickets, int n_tickets) {
  int nt, grid;
  cc_shape(c, B, backward != 0, &nt, &grid);
  cc::Io io;
  io.y = y; io.params = params; io.g_lp = g_lp; io.g_z = g_z; io.g_scale = g_scale;
  io.lp_out = lp_out; io.z_out = z_out; io.g_y = g_y; io.partials = partials;
  io.npad = npad; io.with_base = with_base; io.backward = backward; io.B = B; io.tickets = tickets; io.n_tickets = n_tickets;
  c->fn<<<grid, nt, cc_smem(c, nt, backward != 0), st>>>(c->prog, io);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return surj_cuda_fail(e, "compact chain kernel launch");
  return SURJ_OK;
}

int cc_grid(const CcPlan* c, int64_t B, bool backward) {
  int nt, grid;
  cc_shape(c, B, backward, &nt, &grid);
  return grid;
}

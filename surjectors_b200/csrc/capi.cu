// C ABI (include/surj_b200.h): plan construction/validation, dispatch to the fused whole-chain
// kernels (chain.cuh, registered per chain) or the per-layer path (layered.cu), gradient
// reduction and the fused Adam update.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>
#include <string>

#include "compact.h"
#include "plan.h"
#include "vm_plan.h"

// ------------------------------------------------------------------------ errors
static thread_local std::string g_err;
void surj_set_error(const std::string& msg) { g_err = msg; }
int surj_cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return SURJ_E_CUDA;
}
extern "C" const char* surj_last_error(void) { return g_err.c_str(); }
extern "C" int32_t surj_abi_version(void) { return SURJ_ABI_VERSION; }

// ---------------------------------------------------------------------- registry
namespace {
struct RegEntry {
  SmallKernels k;
  std::string name;
};
std::map<unsigned long long, RegEntry>& registry() {
  static std::map<unsigned long long, RegEntry> r;
  return r;
}
std::mutex& reg_mutex() {
  static std::mutex m;
  return m;
}
}  // namespace

void surj_register_small(unsigned long long key, SmallKernels k, const char* name) {
  std::lock_guard<std::mutex> g(reg_mutex());
  RegEntry e;
  e.k = k;
  e.name = name;
  registry()[key] = e;
}
const SmallKernels* surj_find_small(unsigned long long key) {
  std::lock_guard<std::mutex> g(reg_mutex());
  auto it = registry().find(key);
  return it == registry().end() ? nullptr : &it->second.k;
}
int surj_registry_count() { return (int)registry().size(); }
extern "C" int32_t surj_registry_size(void) { return surj_registry_count(); }

// ----------------------------------------------------------------------- helpers
extern "C" int64_t surj_made_masks(int32_t n_params, int32_t input_size, const int32_t* hidden, int32_t n_hidden,
                                   uint8_t* out) {
  // TFP _make_dense_autoregressive_masks (input_order='left-to-right', hidden_degrees='equal'),
  // called by the reference at surjectors/_src/conditioners/nn/made.py:54-56.
  if (input_size <= 0 || n_params <= 0 || n_hidden < 0 || n_hidden > SURJ_MAX_NET_LAYERS) return SURJ_E_INVALID;
  std::vector<std::vector<int>> deg;
  deg.emplace_back();
  for (int d = 1; d <= input_size; ++d) deg[0].push_back(d);
  for (int l = 0; l < n_hidden; ++l) {
    const int units = hidden[l];
    if (units <= 0) return SURJ_E_INVALID;
    int mn = deg.back()[0];
    for (int v : deg.back()) mn = v < mn ? v : mn;
    mn = mn < input_size - 1 ? mn : input_size - 1;
    std::vector<int> d(units);
    for (int u = 1; u <= units; ++u) {
      // ceil(u * (D - 1) / (units + 1)) evaluated like numpy: float64 division then ceil
      const int c = (int)std::ceil((double)(u * (input_size - 1)) / (double)(units + 1));
      d[u - 1] = c > mn ? c : mn;
    }
    deg.push_back(d);
  }
  int64_t total = 0;
  for (int l = 0; l <= n_hidden; ++l) {
    const int rows = (int)deg[l].size();
    const int cols = l < n_hidden ? (int)deg[l + 1].size() : input_size * n_params;
    if (out) {
      for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
          uint8_t m;
          if (l < n_hidden) m = deg[l][i] <= deg[l + 1][j];
          else m = deg[l][i] < deg[0][j / n_params];
          out[total + (int64_t)i * cols + j] = m;
        }
    }
    total += (int64_t)rows * cols;
  }
  return total;
}

extern "C" int surj_permutation_inverse(const int32_t* permutation, int32_t n, int32_t* inverse_out) {
  // permutation.py:31-36: zeros(size, int).at[perm].set(arange(size))
  if (!permutation || !inverse_out || n <= 0) return SURJ_E_INVALID;
  for (int i = 0; i < n; ++i) inverse_out[i] = 0;
  for (int i = 0; i < n; ++i) {
    if (permutation[i] < 0 || permutation[i] >= n) {
      surj_set_error("permutation entry out of range");
      return SURJ_E_INVALID;
    }
    inverse_out[permutation[i]] = i;
  }
  return SURJ_OK;
}

// ------------------------------------------------------------------- validation
static int check_mlp(const SurjMlpDesc& m, int in_w, int out_w, int64_t np, const char* what) {
  char buf[256];
  if (m.n_layers <= 0 || m.n_layers > SURJ_MAX_NET_LAYERS) {
    snprintf(buf, sizeof buf, "%s: n_layers=%d out of range", what, m.n_layers);
    surj_set_error(buf);
    return SURJ_E_INVALID;
  }
  if (m.widths[0] != in_w || m.widths[m.n_layers] != out_w) {
    snprintf(buf, sizeof buf, "%s: expected %d -> %d, descriptor says %d -> %d", what, in_w, out_w, m.widths[0],
             m.widths[m.n_layers]);
    surj_set_error(buf);
    return SURJ_E_INVALID;
  }
  for (int l = 0; l < m.n_layers; ++l) {
    const int64_t wsz = (int64_t)m.widths[l] * m.widths[l + 1];
    if (m.widths[l] <= 0 || m.widths[l + 1] <= 0 || m.w_off[l] < 0 || m.w_off[l] + wsz > np ||
        (m.b_off[l] >= 0 && m.b_off[l] + m.widths[l + 1] > np) || m.act[l] < 0 || m.act[l] > SURJ_ACT_LEAKY_RELU) {
      snprintf(buf, sizeof buf, "%s: layer %d has bad width / offset / activation", what, l);
      surj_set_error(buf);
      return SURJ_E_INVALID;
    }
  }
  return SURJ_OK;
}

static int check_made(const SurjMadeDesc& m, int in_w, int n_par, int ctx, int64_t np, const char* what) {
  char buf[256];
  if (m.input_size != in_w || m.n_params != n_par || m.ctx_dim != ctx || m.n_hidden < 1 ||
      m.n_hidden > SURJ_MAX_NET_LAYERS || m.act < 0 || m.act > SURJ_ACT_LEAKY_RELU) {
    snprintf(buf, sizeof buf, "%s: MADE(input=%d, params=%d, ctx=%d, hidden layers=%d) does not fit layer (%d, %d, %d)",
             what, m.input_size, m.n_params, m.ctx_dim, m.n_hidden, in_w, n_par, ctx);
    surj_set_error(buf);
    return SURJ_E_INVALID;
  }
  int prev = in_w;
  for (int l = 0; l <= m.n_hidden; ++l) {
    const int cur = l < m.n_hidden ? m.hidden[l] : in_w * n_par;
    if (cur <= 0 || m.w_off[l] < 0 || m.w_off[l] + (int64_t)prev * cur > np || m.b_off[l] < 0 ||
        m.b_off[l] + cur > np) {
      snprintf(buf, sizeof buf, "%s: masked layer %d has bad width / offset", what, l);
      surj_set_error(buf);
      return SURJ_E_INVALID;
    }
    prev = cur;
  }
  if (ctx > 0 && (m.ctx_w_off < 0 || m.ctx_w_off + (int64_t)ctx * m.hidden[0] > np || m.ctx_b_off < 0 ||
                  m.ctx_b_off + m.hidden[0] > np)) {
    surj_set_error(std::string(what) + ": context Linear offsets out of range");
    return SURJ_E_INVALID;
  }
  return SURJ_OK;
}

static int inner_params(const SurjLayerDesc& L) {
  if (L.inner == SURJ_INNER_AFFINE || L.inner == SURJ_INNER_AFFINE_INV) return 2;
  return 3 * L.n_bins + 1;
}

static int validate_layer(const SurjLayerDesc& L, int k, int d_in, int dim_x, int64_t np) {
  char what[64];
  snprintf(what, sizeof what, "layer %d", k);
  char buf[256];
  if (L.dim_in != d_in) {
    snprintf(buf, sizeof buf, "%s: dim_in=%d but the chain delivers %d", what, L.dim_in, d_in);
    surj_set_error(buf);
    return SURJ_E_INVALID;
  }
  if (L.use_x && dim_x <= 0) {
    surj_set_error(std::string(what) + ": use_x set but the chain has no x");
    return SURJ_E_INVALID;
  }
  const int dx = L.use_x ? dim_x : 0;
  const bool funnel = L.kind == SURJ_LAYER_COUPLING_FUNNEL || L.kind == SURJ_LAYER_MAF_FUNNEL ||
                      L.kind == SURJ_LAYER_SLICE || L.kind == SURJ_LAYER_MLP_FUNNEL;
  if (funnel) {
    if (L.dim_out <= 0 || L.dim_out >= L.dim_in) {
      snprintf(buf, sizeof buf, "%s: n_keep=%d must be in (0, %d)", what, L.dim_out, L.dim_in);
      surj_set_error(buf);
      return SURJ_E_INVALID;
    }
    // the MLP funnel's decoder sees z only (mlp.py:53), the others [z, x]
    int r = check_mlp(L.decoder, L.dim_out + (L.kind == SURJ_LAYER_MLP_FUNNEL ? 0 : dx), 2 * (L.dim_in - L.dim_out), np,
                      "decoder");
    if (r) return r;
  } else if (L.dim_out != L.dim_in) {
    surj_set_error(std::string(what) + ": bijector must keep the width");
    return SURJ_E_INVALID;
  }
  const bool has_inner = L.kind == SURJ_LAYER_COUPLING || L.kind == SURJ_LAYER_COUPLING_FUNNEL ||
                         L.kind == SURJ_LAYER_MAF || L.kind == SURJ_LAYER_MAF_FUNNEL;
  if (has_inner) {
    if (L.inner < SURJ_INNER_AFFINE || L.inner > SURJ_INNER_RQS_INV) {
      surj_set_error(std::string(what) + ": unknown inner bijector");
      return SURJ_E_UNSUPPORTED;
    }
    if (L.inner >= SURJ_INNER_RQS) {
      if (L.n_bins < 1 || !(L.range_max > L.range_min) || L.n_bins * 1e-4f > L.range_max - L.range_min) {
        surj_set_error(std::string(what) + ": bad spline bins / range");
        return SURJ_E_INVALID;
      }
    }
  }
  switch (L.kind) {
    case SURJ_LAYER_COUPLING: {
      if (!L.mask) {
        surj_set_error(std::string(what) + ": coupling needs a mask");
        return SURJ_E_INVALID;
      }
      return check_mlp(L.conditioner, L.dim_in + dx, L.dim_in * inner_params(L), np, "conditioner");
    }
    case SURJ_LAYER_COUPLING_FUNNEL:
      return check_mlp(L.conditioner, L.dim_in + dx, L.dim_in * inner_params(L), np, "conditioner");
    case SURJ_LAYER_MAF:
      return check_made(L.made, L.dim_in, inner_params(L), dx, np, "MADE");
    case SURJ_LAYER_MAF_FUNNEL:
      return check_made(L.made, L.dim_out, inner_params(L), L.dim_in - L.dim_out + dx, np, "MADE");
    case SURJ_LAYER_SLICE:
      return SURJ_OK;
    case SURJ_LAYER_MLP_FUNNEL: {
      const int64_t n = L.dim_out, tri = n * (n - 1) / 2;
      if (L.lu_lower_off < 0 || L.lu_lower_off + tri > np || L.lu_upper_off < 0 || L.lu_upper_off + tri > np ||
          L.lu_diag_off < 0 || L.lu_diag_off + n > np) {
        surj_set_error(std::string(what) + ": LULinear offsets out of range");
        return SURJ_E_INVALID;
      }
      if (L.conditioner.n_layers != 1) {
        surj_set_error(std::string(what) + ": the y- path of an MLP funnel is ONE hk.Linear");
        return SURJ_E_INVALID;
      }
      return check_mlp(L.conditioner, L.dim_in - L.dim_out, L.dim_out, np, "w_prime");
    }
    case SURJ_LAYER_LU: {
      const int64_t n = L.dim_in, tri = n * (n - 1) / 2;
      if (L.lu_lower_off < 0 || L.lu_lower_off + tri > np || L.lu_upper_off < 0 || L.lu_upper_off + tri > np ||
          L.lu_diag_off < 0 || L.lu_diag_off + n > np) {
        surj_set_error(std::string(what) + ": LULinear offsets out of range");
        return SURJ_E_INVALID;
      }
      return SURJ_OK;
    }
    case SURJ_LAYER_PERMUTATION: {
      if (!L.permutation) {
        surj_set_error(std::string(what) + ": permutation missing");
        return SURJ_E_INVALID;
      }
      std::vector<char> seen(L.dim_in, 0);
      for (int d = 0; d < L.dim_in; ++d) {
        const int p = L.permutation[d];
        if (p < 0 || p >= L.dim_in || seen[p]) {
          surj_set_error(std::string(what) + ": not a permutation");
          return SURJ_E_INVALID;
        }
        seen[p] = 1;
      }
      return SURJ_OK;
    }
    default:
      surj_set_error(std::string(what) + ": unknown layer kind");
      return SURJ_E_UNSUPPORTED;
  }
}

// -------------------------------------------------------------------------- plan
extern "C" int surj_plan_create(const SurjLayerDesc* layers, int32_t n_layers, int32_t dim_in, int32_t dim_x,
                                int64_t n_params, int32_t mode, int32_t base, SurjPlan** out) {
  if (!layers || !out || n_layers <= 0 || dim_in <= 0 || dim_x < 0 || n_params < 0 ||
      (base != SURJ_BASE_STANDARD_NORMAL && base != SURJ_BASE_NONE)) {
    surj_set_error("surj_plan_create: bad arguments");
    return SURJ_E_INVALID;
  }
  int d = dim_in;
  for (int k = 0; k < n_layers; ++k) {
    int r = validate_layer(layers[k], k, d, dim_x, n_params);
    if (r) return r;
    d = layers[k].dim_out;
  }
  SurjPlan* p = new SurjPlan();
  p->layers.assign(layers, layers + n_layers);
  p->masks.resize(n_layers);
  p->perms.resize(n_layers);
  for (int k = 0; k < n_layers; ++k) {
    SurjLayerDesc& L = p->layers[k];
    if (L.mask) {
      p->masks[k].assign(L.mask, L.mask + L.dim_in);
      L.mask = p->masks[k].data();
    }
    if (L.permutation) {
      p->perms[k].assign(L.permutation, L.permutation + L.dim_in);
      L.permutation = p->perms[k].data();
    }
  }
  p->dim_in = dim_in;
  p->dim_x = dim_x;
  p->dim_out = d;
  p->n_params = n_params;
  p->mode = mode;
  p->base = base;
  surj_chain_key(p->layers.data(), n_layers, dim_in, dim_x, n_params, &p->key);
  // Every chain compiles to one program (vm_plan.cu) with two executors; two more kernels cover special families.
  // AUTO picks, in order: a generated straight-line kernel compiled in for exactly this chain, the compact looped kernel
  // (affine coupling / coupling-funnel chains with small two-hidden-layer MLPs, no code generation needed), the fused
  // interpreter kernel (parameters + a tile of activations fit in shared memory), the per-op executor (any width).
  std::string why;
  p->vmp = vm_compile(p, &why);
  if (mode == SURJ_MODE_AUTO || mode == SURJ_MODE_FUSED_GEN1) p->fused = surj_find_small(p->key);
  std::string cc_why;
  if (mode == SURJ_MODE_AUTO || mode == SURJ_MODE_FUSED_COMPACT) p->ccp = cc_compile(p, &cc_why);
  if (p->fused) p->exec = 1;
  else if (p->ccp) p->exec = 4;
  else if (p->vmp && p->vmp->fused_ok && (mode == SURJ_MODE_AUTO || mode == SURJ_MODE_FUSED)) p->exec = 2;
  else if (p->vmp && (mode == SURJ_MODE_AUTO || mode == SURJ_MODE_LAYERED)) p->exec = 3;
  if (p->exec == 0) {
    char buf[400];
    if (!p->vmp) snprintf(buf, sizeof buf, "chain is not supported: %s", why.c_str());
    else if (mode == SURJ_MODE_FUSED_COMPACT) snprintf(buf, sizeof buf, "chain cannot run on the compact fused kernel: %s", cc_why.c_str());
    else if (mode == SURJ_MODE_FUSED_GEN1)
      snprintf(buf, sizeof buf, "no generated kernel compiled in for chain key 0x%016llx (%d registered)",
               (unsigned long long)p->key, surj_registry_count());
    else snprintf(buf, sizeof buf, "chain cannot run on the fused kernel: %s", p->vmp->fused_why.c_str());
    surj_set_error(buf);
    vm_destroy(p->vmp);
    delete p;
    return SURJ_E_UNSUPPORTED;
  }
  *out = p;
  return SURJ_OK;
}

extern "C" void surj_plan_destroy(SurjPlan* plan) {
  if (!plan) return;
  cc_destroy(plan->ccp);
  vm_destroy(plan->vmp);
  delete plan;
}
extern "C" uint64_t surj_plan_key(const SurjPlan* plan) { return plan ? plan->key : 0; }
extern "C" int32_t surj_plan_is_fused(const SurjPlan* plan) {
  if (!plan) return 0;
  return plan->exec == 3 ? 0 : (plan->exec == 4 ? 3 : plan->exec);
}
extern "C" int32_t surj_plan_dim_out(const SurjPlan* plan) { return plan ? plan->dim_out : -1; }

// A plan is bound to the device that is current at its first launch: its staging tables are allocated there and
// its occupancy facts are that device's.  Calls from any other device are refused instead of dereferencing another
// GPU's tables (one plan per device: the host layer keys its plan cache by device).
static int ensure_device(const SurjPlan* p) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return surj_cuda_fail(e, "cudaGetDevice");
  if (p->dev_ready) {
    if (dev != p->device) {
      surj_set_error("plan is bound to CUDA device " + std::to_string(p->device) + " but device " + std::to_string(dev) +
                     " is current: create one plan per device");
      return SURJ_E_INVALID;
    }
    return SURJ_OK;
  }
  p->device = dev;
  e = cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return surj_cuda_fail(e, "cudaDeviceGetAttribute");
  if (p->exec == 2 || p->exec == 3) {
    int r = vm_prepare(p->vmp);
    if (r) return r;
  }
  if (p->exec == 4) {
    int r = cc_prepare(p->ccp);
    if (r) return r;
  }
  if (p->exec == 1) {
    e = p->fused->prepare();
    if (e != cudaSuccess) return surj_cuda_fail(e, "cudaFuncSetAttribute(fused chain)");
    e = p->fused->occupancy(&p->fwd_blocks, &p->bwd_blocks);
    if (e != cudaSuccess) return surj_cuda_fail(e, "occupancy(fused chain)");
    if (p->fwd_blocks < 1 || p->bwd_blocks < 1) {
      surj_set_error("fused chain kernel does not fit on an SM");
      return SURJ_E_CUDA;
    }
  }
  p->dev_ready = true;
  return SURJ_OK;
}

static int fused_grid(const SurjPlan* p, int64_t B, bool backward) {
  if (p->exec == 2) return vm_grid(p->vmp, B, backward ? VM_PROG_BWD : VM_PROG_FWD);
  if (p->exec == 4) return cc_grid(p->ccp, B, backward);
  const int nt = p->fused->nt;
  const int64_t tiles = (B + nt - 1) / nt;
  const int64_t cap = (int64_t)p->n_sm * (backward ? p->bwd_blocks : p->fwd_blocks);
  return (int)(tiles < cap ? tiles : cap);
}
static int fused_npad(const SurjPlan* p) { return (int)((p->n_params + 1 + 3) / 4 * 4); }
static size_t fused_bwd_workspace(int64_t rows, int npad, int np);

extern "C" size_t surj_plan_workspace_bytes(const SurjPlan* plan, int64_t batch, int32_t backward) {
  if (!plan || batch <= 0) return 0;
  if (plan->exec == 1 || plan->exec == 2 || plan->exec == 4) {
    if (!backward) return 0;
    int64_t rows;
    if (ensure_device(plan) == SURJ_OK) rows = fused_grid(plan, batch, true);
    else if (plan->exec == 2) rows = vm_max_grid(plan->vmp, batch);
    else if (plan->exec == 4) rows = cc_max_grid(plan->ccp);
    else {  // no device yet: an upper bound of one partial row per possible CTA
      const int64_t tiles = (batch + plan->fused->nt - 1) / plan->fused->nt;
      rows = tiles < 4096 ? tiles : 4096;
    }
    return fused_bwd_workspace(rows, fused_npad(plan), (int)plan->n_params);
  }
  return vm_global_workspace_bytes(plan->vmp, batch, backward != 0);
}

extern "C" int32_t surj_plan_sampling_recursions(const SurjPlan* plan) {
  if (!plan || !plan->vmp || plan->exec != 3) return 0;
  return vm_global_maf_groups(plan->vmp);
}

extern "C" int32_t surj_plan_launches(const SurjPlan* plan, int32_t backward) {
  if (!plan) return 0;
  if (plan->exec == 1 || plan->exec == 2 || plan->exec == 4) return backward ? 2 : 1;
  return vm_global_launches(plan->vmp, 1, backward != 0);
}

static int check_io(const SurjPlan* plan, const void* y, const void* x, const void* params, int64_t B) {
  if (!plan || !y || !params || B <= 0 || (plan->dim_x > 0 && !x)) {
    surj_set_error("null plan / buffer, non-positive batch, or x missing for a conditional chain");
    return SURJ_E_INVALID;
  }
  if (((uintptr_t)y | (uintptr_t)x | (uintptr_t)params) & 15u) {
    surj_set_error("device pointers must be 16-byte aligned");
    return SURJ_E_INVALID;
  }
  return SURJ_OK;
}

extern "C" int surj_chain_logprob_fwd(const SurjPlan* plan, SurjStream stream, const float* y, const float* x,
                                      const float* params, float* z_out, float* lp_out, void* workspace,
                                      size_t workspace_bytes, int64_t batch) {
  int r = check_io(plan, y, x, params, batch);
  if (r) return r;
  if (!lp_out) {
    surj_set_error("lp_out is required");
    return SURJ_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->exec == 4) {
    r = ensure_device(plan);
    if (r) return r;
    return cc_launch(plan->ccp, st, 0, plan->base == SURJ_BASE_STANDARD_NORMAL, y, params, nullptr, 0.0f, nullptr, z_out, lp_out,
                     nullptr, nullptr, 0, batch, nullptr, 0);
  }
  if (plan->exec == 2) {
    r = ensure_device(plan);
    if (r) return r;
    return vm_launch(plan->vmp, st, VM_PROG_FWD, plan->base == SURJ_BASE_STANDARD_NORMAL, y, x, params, nullptr, 0.0f, nullptr,
                     z_out, lp_out, nullptr, nullptr, nullptr, 0, batch, fused_grid(plan, batch, false), nullptr, 0);
  }
  if (plan->exec == 1) {
    r = ensure_device(plan);
    if (r) return r;
    cudaError_t e = plan->fused->fwd(fused_grid(plan, batch, false), st, plan->base == SURJ_BASE_STANDARD_NORMAL, y, x, params,
                                     z_out, lp_out, batch);
    if (e != cudaSuccess) return surj_cuda_fail(e, "chain_fwd_kernel launch");
    return SURJ_OK;
  }
  r = ensure_device(plan);
  if (r) return r;
  return vm_global_run(plan->vmp, st, VM_PROG_FWD, plan->base == SURJ_BASE_STANDARD_NORMAL, y, x, params, nullptr, 0.0f, nullptr,
                       z_out, lp_out, nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, batch);
}

// ---- gradient reduction (+ optional fused Adam) for the fused executors
// Every CTA of the backward kernel leaves one partial row [npad] (column np = its sum of lp).
// Level 1: grid (32-column blocks) x (RSL row slices); each CTA sums its rows in a fixed order and
// stores one slice row.  Level 2: the last CTA of a column block to arrive (ticket) adds the RSL
// slice rows in a fixed order -- the result is deterministic -- writes g_params / loss and, when asked,
// applies the optax.adam update in the same pass.  The last CTA overall advances the step counter.
// Tickets are zeroed by the backward kernel that precedes this one on the stream.
struct AdamArgs {
  float* params; float* mu; float* nu; int32_t* count;
  float lr, b1, b2, eps;
};
// Data-parallel exchange over NVLink peer memory (SurjDp): every rank owns a mailbox box[2][world][npad] of
// 8-byte slots (float value | step tag << 32) that its peers write into directly.
struct DpArgs {
  unsigned long long* const* peer_box;  // [world] device pointers to the peers' mailboxes (nullptr: single GPU)
  unsigned long long* my_box;
  int rank, world;
  int* err;                     // device flag, sticky: 1 once a peer did not arrive in time -- every later step is a no-op
  int* err_host;                // host-mapped mirror of the flag (the host polls it without synchronising)
  long long timeout_cycles;     // 0: wait for ever (what NCCL does)
};
constexpr int RSL = 8;

__device__ __forceinline__ void adam_update(const AdamArgs& ad, int col, float g, int t) {
  const float c1 = 1.0f / (1.0f - powf(ad.b1, (float)t)), c2 = 1.0f / (1.0f - powf(ad.b2, (float)t));
  const float m = ad.b1 * ad.mu[col] + (1.0f - ad.b1) * g;
  const float v = ad.b2 * ad.nu[col] + (1.0f - ad.b2) * g * g;
  ad.mu[col] = m;
  ad.nu[col] = v;
  ad.params[col] -= ad.lr * (m * c1) / (sqrtf(v * c2) + ad.eps);
}

__global__ void __launch_bounds__(256)
surj_reduce_partials_kernel(const float* __restrict__ partials, int nrows, int npad, int np,
                            float* __restrict__ slices, unsigned* __restrict__ tickets,
                            float* __restrict__ g_params, float* __restrict__ loss_out, float loss_scale, AdamArgs ad,
                            DpArgs dp) {
  __shared__ float red[8][33];
  __shared__ unsigned s_ticket;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int per = (nrows + RSL - 1) / RSL;
  const int r0 = blockIdx.y * per, r1 = min(nrows, r0 + per);
  const bool multi = dp.peer_box != nullptr;
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
  if (col <= np) {
    int r = r0 + w;
    for (; r + 24 < r1; r += 32) {
      a0 += partials[(size_t)r * npad + col];
      a1 += partials[(size_t)(r + 8) * npad + col];
      a2 += partials[(size_t)(r + 16) * npad + col];
      a3 += partials[(size_t)(r + 24) * npad + col];
    }
    for (; r < r1; r += 8) a0 += partials[(size_t)r * npad + col];
  }
  red[w][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (w == 0) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][lane];
    if (col <= np) slices[(size_t)blockIdx.y * npad + col] = s;
    __threadfence();
    if (lane == 0) s_ticket = atomicAdd(tickets + blockIdx.x, 1u);
    __syncwarp();
    if (s_ticket == RSL - 1) {  // last slice of this column block: finish the columns
      __threadfence();
      const bool valid = col <= np;  // the last column block is partly empty; all 32 lanes stay converged
      float g = 0.0f;
      if (valid) {
#pragma unroll
        for (int k = 0; k < RSL; ++k) g += __ldcg(slices + (size_t)k * npad + col);
      }
      bool dp_fail = false;
      if (multi) {
        // A step whose exchange timed out must not touch the replicated state: it would add a stale mailbox value and
        // let the replicas drift apart silently.  The flag is sticky, the update and the step counter are skipped for
        // this and every later step, and the host raises on its next poll (TrainStep.check / surj_dp_error).
        const bool dead = *reinterpret_cast<volatile int*>(dp.err) != 0;
        bool late = false;
        // Data parallel: this warp finishes its 32 columns across ALL ranks.  My shard's sums go straight into every
        // rank's mailbox (NVLink peer stores); value and step tag travel in ONE aligned 8-byte store, so a value is
        // usable the moment it is visible (no fence + separate flag round trip).  The peers' contributions are summed
        // in rank order (bit-identical on every rank) and Adam is applied.  The mailbox is double-buffered by step
        // parity: a rank can be at most one step ahead of the slowest.
        const unsigned t = (unsigned)(*ad.count + 1);
        const int buf = (int)(t & 1u);
        if (valid && !dead) {
          const size_t slot = ((size_t)buf * dp.world + dp.rank) * npad + col;
          const unsigned long long packed = ((unsigned long long)t << 32) | (unsigned long long)__float_as_uint(g);
          for (int p = 0; p < dp.world; ++p) reinterpret_cast<volatile unsigned long long*>(dp.peer_box[p])[slot] = packed;
          g = 0.0f;
          for (int r = 0; r < dp.world; ++r) {
            const volatile unsigned long long* src =
              reinterpret_cast<const volatile unsigned long long*>(dp.my_box) + ((size_t)buf * dp.world + r) * npad + col;
            unsigned long long v = *src;
            const long long t0 = clock64();
            while ((unsigned)(v >> 32) != t) {
              if (dp.timeout_cycles > 0 && clock64() - t0 > dp.timeout_cycles) {  // a peer is gone: report, do not hang
                late = true;
                break;
              }
              v = *src;
            }
            if (late) break;
            g += __uint_as_float((unsigned)v);
          }
        }
        __syncwarp();
        dp_fail = dead || __any_sync(0xffffffffu, late);
        if (dp_fail && lane == 0 && !dead) {
          *reinterpret_cast<volatile int*>(dp.err) = 1;
          *reinterpret_cast<volatile int*>(dp.err_host) = 1;
          __threadfence_system();
        }
      }
      if (valid && !dp_fail) {
        if (col < np) {
          g_params[col] = g;
          if (ad.params != nullptr) adam_update(ad, col, g, *ad.count + 1);  // optax.adam
        } else if (loss_out) {
          *loss_out = loss_scale * g;
        }
      }
      __threadfence();
      if (lane == 0) {
        const unsigned done = atomicAdd(tickets + gridDim.x, 1u);
        // every column block has read count; a failed exchange (any column block of this step) leaves it where it was
        if (done == gridDim.x - 1 && ad.params != nullptr && !(multi && *reinterpret_cast<volatile int*>(dp.err) != 0))
          *ad.count += 1;
      }
    }
  }
}

static size_t fused_bwd_workspace(int64_t rows, int npad, int np) {
  const size_t ncb = (size_t)(np + 1 + 31) / 32;
  return ((size_t)rows + RSL) * npad * sizeof(float) + (ncb + 1 + 3) / 4 * 4 * sizeof(unsigned);
}

static int fused_backward(const SurjPlan* plan, cudaStream_t st, const float* y, const float* x, const float* params,
                          const float* g_lp, float g_scale, const float* g_z, float* g_params, float* loss_out,
                          float* lp_out, float* g_y, float* g_x, void* workspace, size_t workspace_bytes, int64_t batch,
                          AdamArgs ad, DpArgs dp) {
  int r = ensure_device(plan);
  if (r) return r;
  const int grid = fused_grid(plan, batch, true);
  const int npad = fused_npad(plan);
  const int np = (int)plan->n_params;
  if (!workspace || workspace_bytes < fused_bwd_workspace(grid, npad, np) || ((uintptr_t)workspace & 15u)) {
    surj_set_error("workspace too small or misaligned for the backward pass");
    return SURJ_E_WORKSPACE;
  }
  float* partials = (float*)workspace;
  float* slices = partials + (size_t)grid * npad;
  unsigned* tickets = (unsigned*)(slices + (size_t)RSL * npad);
  const int ncb = (np + 1 + 31) / 32;
  const int with_base = plan->base == SURJ_BASE_STANDARD_NORMAL;
  if (plan->exec == 2) {
    r = vm_launch(plan->vmp, st, VM_PROG_BWD, with_base, y, x, params, g_lp, g_scale, g_z, nullptr, lp_out, g_y, g_x, partials,
                  npad, batch, grid, tickets, ncb + 1);
    if (r) return r;
  } else if (plan->exec == 4) {
    r = cc_launch(plan->ccp, st, 1, with_base, y, params, g_lp, g_scale, g_z, nullptr, lp_out, g_y, partials, npad, batch, tickets,
                  ncb + 1);
    if (r) return r;
  } else {
    cudaError_t e = plan->fused->bwd(grid, st, with_base, y, x, params, g_lp, g_scale, g_z, partials, npad, lp_out,
                                     nullptr, g_y, g_x, batch, tickets, ncb + 1);
    if (e != cudaSuccess) return surj_cuda_fail(e, "chain_bwd_kernel launch");
  }
  surj_reduce_partials_kernel<<<dim3(ncb, RSL), 256, 0, st>>>(partials, grid, npad, np, slices, tickets, g_params,
                                                             loss_out, g_lp ? 1.0f : g_scale, ad, dp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return surj_cuda_fail(e, "reduce_partials launch");
  return SURJ_OK;
}

extern "C" int surj_chain_logprob_bwd(const SurjPlan* plan, SurjStream stream, const float* y, const float* x,
                                      const float* params, const float* g_lp, float g_scale, const float* g_z,
                                      float* g_params, float* loss_out, float* lp_out, float* g_y, float* g_x,
                                      void* workspace, size_t workspace_bytes, int64_t batch) {
  int r = check_io(plan, y, x, params, batch);
  if (r) return r;
  if (!g_params) {
    surj_set_error("g_params is required");
    return SURJ_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->exec == 1 || plan->exec == 2 || plan->exec == 4)
    return fused_backward(plan, st, y, x, params, g_lp, g_scale, g_z, g_params, loss_out, lp_out, g_y, g_x, workspace,
                          workspace_bytes, batch, AdamArgs{nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0}, DpArgs{});
  r = ensure_device(plan);
  if (r) return r;
  return vm_global_run(plan->vmp, st, VM_PROG_BWD, plan->base == SURJ_BASE_STANDARD_NORMAL, y, x, params, g_lp, g_scale, g_z,
                       nullptr, lp_out, g_y, g_x, g_params, loss_out, workspace, workspace_bytes, batch);
}

// -------------------------------------------------------------------------- Adam
// optax.adam: mu = b1 mu + (1-b1) g; nu = b2 nu + (1-b2) g^2; bias correction with t = count+1;
// p -= lr * mu_hat / (sqrt(nu_hat) + eps).  (examples/coupling_inference_surjection.py:78,88-89)
__global__ void __launch_bounds__(256)
surj_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu, float* __restrict__ nu,
                 const int32_t* __restrict__ count, long long n, float lr, float b1, float b2, float eps,
                 float gscale) {
  const int t = *count + 1;
  const float c1 = 1.0f / (1.0f - powf(b1, (float)t));
  const float c2 = 1.0f / (1.0f - powf(b2, (float)t));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float m = b1 * mu[i] + (1.0f - b1) * gi;
    const float v = b2 * nu[i] + (1.0f - b2) * gi * gi;
    mu[i] = m;
    nu[i] = v;
    p[i] -= lr * (m * c1) / (sqrtf(v * c2) + eps);
  }
}
// separate tiny kernel so every thread of the update sees the same count
__global__ void surj_count_inc_kernel(int32_t* count) { *count += 1; }

extern "C" int surj_adam_step(SurjStream stream, float* params, const float* grads, float* mu, float* nu,
                              int32_t* count, int64_t n, float lr, float b1, float b2, float eps, float grad_scale) {
  if (!params || !grads || !mu || !nu || !count || n <= 0) {
    surj_set_error("surj_adam_step: bad arguments");
    return SURJ_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  surj_adam_kernel<<<blocks, 256, 0, st>>>(params, grads, mu, nu, count, n, lr, b1, b2, eps, grad_scale);
  surj_count_inc_kernel<<<1, 1, 0, st>>>(count);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return surj_cuda_fail(e, "adam launch");
  return SURJ_OK;
}

// The whole jitted `step` of the reference's examples (examples/coupling_inference_surjection.py:81-90):
// loss = -sum(lp) (g_scale = -1) or -mean, value_and_grad, optax.adam update -- for the fused executors
// two launches (backward kernel; gradient reduction with the Adam update fused into its last stage).
extern "C" int surj_train_step(const SurjPlan* plan, SurjStream stream, const float* y, const float* x, float* params,
                               float g_scale, float* g_params, float* mu, float* nu, int32_t* count, float lr, float b1,
                               float b2, float eps, float* loss_out, void* workspace, size_t workspace_bytes,
                               int64_t batch) {
  int r = check_io(plan, y, x, params, batch);
  if (r) return r;
  if (!g_params || !mu || !nu || !count) {
    surj_set_error("surj_train_step: g_params, mu, nu and count are required");
    return SURJ_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->exec == 1 || plan->exec == 2 || plan->exec == 4)
    return fused_backward(plan, st, y, x, params, nullptr, g_scale, nullptr, g_params, loss_out, nullptr, nullptr,
                          nullptr, workspace, workspace_bytes, batch, AdamArgs{params, mu, nu, count, lr, b1, b2, eps}, DpArgs{});
  r = surj_chain_logprob_bwd(plan, stream, y, x, params, nullptr, g_scale, nullptr, g_params, loss_out, nullptr, nullptr,
                             nullptr, workspace, workspace_bytes, batch);
  if (r) return r;
  return surj_adam_step(stream, params, g_params, mu, nu, count, plan->n_params, lr, b1, b2, eps, 1.0f);
}

// Sampling direction: Surjector.forward_and_likelihood_contribution through the Chain in reverse
// (surjectors/_src/surjectors/chain.py:47-54).  Always runs on the chain program (the generated kernels
// only cover log_prob and its VJP): fused interpreter kernel when it fits, else the per-op executor.
extern "C" int32_t surj_plan_noise_dim(const SurjPlan* plan) { return plan && plan->vmp ? plan->vmp->n_eps : 0; }

extern "C" int surj_chain_forward(const SurjPlan* plan, SurjStream stream, const float* z, const float* x,
                                  const float* eps, const float* params, float* y_out, float* ld_out, void* workspace,
                                  size_t workspace_bytes, int64_t batch) {
  int r = check_io(plan, z, x, params, batch);
  if (r) return r;
  if (!y_out || !ld_out) {
    surj_set_error("y_out and ld_out are required");
    return SURJ_E_INVALID;
  }
  if (!plan->vmp || plan->vmp->ops_gen.empty()) {
    surj_set_error(plan->vmp ? plan->vmp->gen_why : std::string("chain has no program"));
    return SURJ_E_UNSUPPORTED;
  }
  if (plan->vmp->n_eps > 0 && !eps) {
    surj_set_error("this chain samples decoder distributions: eps[B, surj_plan_noise_dim] is required");
    return SURJ_E_INVALID;
  }
  r = vm_prepare(plan->vmp);
  if (r) return r;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->vmp->fused_ok && plan->mode != SURJ_MODE_LAYERED)
    return vm_launch(plan->vmp, st, VM_PROG_GEN, 0, z, x, params, nullptr, 0.0f, nullptr, y_out, ld_out, nullptr, nullptr,
                     nullptr, 0, batch, vm_grid(plan->vmp, batch, VM_PROG_GEN), nullptr, 0, eps);
  return vm_global_run(plan->vmp, st, VM_PROG_GEN, 0, z, x, params, nullptr, 0.0f, nullptr, y_out, ld_out, nullptr, nullptr,
                       nullptr, nullptr, workspace, workspace_bytes, batch, eps);
}

extern "C" size_t surj_plan_forward_workspace_bytes(const SurjPlan* plan, int64_t batch) {
  if (!plan || !plan->vmp || batch <= 0) return 0;
  if (plan->vmp->fused_ok && plan->mode != SURJ_MODE_LAYERED) return 0;
  return vm_global_workspace_bytes(plan->vmp, batch, false);
}

// ---------------------------------------------------------- data-parallel train step
// Replicated parameters, sharded batch: the flat gradient (+ loss) of every rank must be summed before the
// Adam update.  For the KB-sized gradients of the small-network configs an NCCL
// all-reduce (launch + protocol latency) costs as much as a fifth of the step, so the exchange is fused into
// the gradient-reduction kernel: each rank stores its reduced gradient, tagged with the step number, directly into every
// peer's mailbox (NVLink / NVSwitch peer memory), waits until the peers' tagged values have landed in its own, sums them in
// rank order (bit-identical on all ranks) and applies Adam -- backward kernel + ONE kernel, no NCCL call in the step.
typedef unsigned long long SurjSlot;  // float value | (step tag << 32)
struct SurjDp {
  int rank = 0, world = 1, npad = 0;
  SurjSlot* box = nullptr;       // [2][world][npad]
  std::vector<void*> opened;     // peers' mappings (cudaIpcOpenMemHandle)
  SurjSlot** d_peer_box = nullptr;
  int* d_err = nullptr;          // sticky device flag
  int* h_err = nullptr;          // pinned, host-mapped mirror (cudaHostAllocMapped)
  int* h_err_dev = nullptr;      // device pointer of the mirror
  long long timeout_cycles = 0;  // 0 = wait for ever
  int device = 0;
};

static long long dp_cycles_of_ms(int device, long long ms) {
  if (ms <= 0) return 0;
  int khz = 1965000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  return ms * (long long)khz;
}

extern "C" int surj_dp_create(int32_t rank, int32_t world, int64_t n_params, SurjDp** out, unsigned char* handle_out) {
  if (!out || !handle_out || world < 1 || rank < 0 || rank >= world || n_params <= 0) {
    surj_set_error("surj_dp_create: bad arguments");
    return SURJ_E_INVALID;
  }
  SurjDp* d = new SurjDp();
  d->rank = rank; d->world = world;
  d->npad = (int)((n_params + 1 + 3) / 4 * 4);
  const size_t bytes = (size_t)2 * world * d->npad * sizeof(SurjSlot);
  cudaError_t e = cudaMalloc((void**)&d->box, bytes);
  if (e == cudaSuccess) e = cudaMemset(d->box, 0, bytes);  // tag 0: no step has been delivered
  if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_err, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(d->d_err, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&d->h_err, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) { *d->h_err = 0; e = cudaHostGetDevicePointer((void**)&d->h_err_dev, d->h_err, 0); }
  if (e == cudaSuccess) e = cudaGetDevice(&d->device);
  if (e != cudaSuccess) { delete d; return surj_cuda_fail(e, "surj_dp_create"); }
  {  // default 30 s; SURJ_B200_DP_TIMEOUT_MS=0 waits for ever like a NCCL collective
    const char* env = getenv("SURJ_B200_DP_TIMEOUT_MS");
    d->timeout_cycles = dp_cycles_of_ms(d->device, env ? atoll(env) : 30000ll);
  }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, d->box);
  if (e != cudaSuccess) { delete d; return surj_cuda_fail(e, "cudaIpcGetMemHandle"); }
  static_assert(sizeof(cudaIpcMemHandle_t) == SURJ_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof h);
  *out = d;
  return SURJ_OK;
}

extern "C" int surj_dp_connect(SurjDp* d, const unsigned char* handles) {
  if (!d || !handles) { surj_set_error("surj_dp_connect: bad arguments"); return SURJ_E_INVALID; }
  std::vector<SurjSlot*> boxes(d->world);
  for (int r = 0; r < d->world; ++r) {
    if (r == d->rank) { boxes[r] = d->box; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * SURJ_IPC_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return surj_cuda_fail(e, "cudaIpcOpenMemHandle (peer mailbox)");
    d->opened.push_back(p);
    boxes[r] = (SurjSlot*)p;
  }
  cudaError_t e = cudaMalloc((void**)&d->d_peer_box, sizeof(SurjSlot*) * d->world);
  if (e == cudaSuccess) e = cudaMemcpy(d->d_peer_box, boxes.data(), sizeof(SurjSlot*) * d->world, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return surj_cuda_fail(e, "surj_dp_connect");
  return SURJ_OK;
}

extern "C" void surj_dp_destroy(SurjDp* d) {
  if (!d) return;
  for (void* p : d->opened) cudaIpcCloseMemHandle(p);
  if (d->d_peer_box) cudaFree(d->d_peer_box);
  if (d->d_err) cudaFree(d->d_err);
  if (d->h_err) cudaFreeHost(d->h_err);
  if (d->box) cudaFree(d->box);
  delete d;
}

extern "C" int32_t surj_dp_error(const SurjDp* d) {  // synchronises: for tests / end of run, not the hot path
  int v = 0;
  if (d && d->d_err) cudaMemcpy(&v, d->d_err, sizeof v, cudaMemcpyDeviceToHost);
  return v;
}

// Non-blocking: reads the host-mapped mirror of the flag.  Cheap enough to poll before every step.
extern "C" int32_t surj_dp_error_nosync(const SurjDp* d) {
  return (d && d->h_err) ? *reinterpret_cast<volatile int*>(d->h_err) : 0;
}

extern "C" int surj_dp_set_timeout_ms(SurjDp* d, int64_t ms) {
  if (!d || ms < 0) { surj_set_error("surj_dp_set_timeout_ms: bad arguments"); return SURJ_E_INVALID; }
  d->timeout_cycles = dp_cycles_of_ms(d->device, (long long)ms);
  return SURJ_OK;
}

extern "C" int surj_train_step_dp(const SurjPlan* plan, SurjDp* d, SurjStream stream, const float* y, const float* x,
                                  float* params, float g_scale, float* g_params, float* mu, float* nu, int32_t* count,
                                  float lr, float b1, float b2, float eps, float* loss_out, void* workspace,
                                  size_t workspace_bytes, int64_t batch) {
  int r = check_io(plan, y, x, params, batch);
  if (r) return r;
  if (!d || !d->d_peer_box || !g_params || !mu || !nu || !count) {
    surj_set_error("surj_train_step_dp: connect the SurjDp first; g_params, mu, nu and count are required");
    return SURJ_E_INVALID;
  }
  if (plan->exec != 1 && plan->exec != 2 && plan->exec != 4) {
    surj_set_error("the fused peer-memory exchange serves the fused executors; use an all-reduce for wide networks");
    return SURJ_E_UNSUPPORTED;
  }
  if (d->npad != fused_npad(plan)) {
    surj_set_error("SurjDp was created for a different parameter count");
    return SURJ_E_INVALID;
  }
  DpArgs dp{d->d_peer_box, d->box, d->rank, d->world, d->d_err, d->h_err_dev, d->timeout_cycles};
  return fused_backward(plan, (cudaStream_t)stream, y, x, params, nullptr, g_scale, nullptr, g_params, loss_out, nullptr,
                        nullptr, nullptr, workspace, workspace_bytes, batch, AdamArgs{params, mu, nu, count, lr, b1, b2, eps},
                        dp);
}

// ------------------------------------------------------------- FP32 peak probe
__global__ void __launch_bounds__(256) surj_fp32_peak_kernel(float* __restrict__ sink, int iters) {
  float a[16];
  const float m = 1.0f + 1e-7f * (float)(threadIdx.x & 3), c = 1e-9f * (float)blockIdx.x;
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = (float)(threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], m, c);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 123.456f) sink[0] = s;  // keeps the loop alive without a store on the hot path
}

extern "C" int surj_bench_fp32_peak(SurjStream stream, int32_t iters, float* sink, double* flops_out) {
  if (!sink || iters <= 0) {
    surj_set_error("surj_bench_fp32_peak: bad arguments");
    return SURJ_E_INVALID;
  }
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = n_sm * 8;
  surj_fp32_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, iters);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return surj_cuda_fail(e, "fp32 peak probe launch");
  if (flops_out) *flops_out = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
  return SURJ_OK;
}

// Canonical 64-bit key of a chain structure: FNV-1a over an explicit int64 serialisation of the
// descriptors (no struct padding, pointers replaced by the data they point to).  Built twice:
// into libsurj_b200.so, and into the tiny host-only _surjkey.so the build-time code generator
// uses to name the fused kernels it emits.
#include <cstring>
#include <vector>

#include "../../include/surj_b200.h"

namespace {
struct Ser {
  std::vector<int64_t> v;
  void i(int64_t x) { v.push_back(x); }
  void f(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    v.push_back((int64_t)u);
  }
  void mlp(const SurjMlpDesc& m) {
    i(m.n_layers);
    for (int l = 0; l <= m.n_layers && m.n_layers > 0; ++l) i(m.widths[l]);
    for (int l = 0; l < m.n_layers; ++l) { i(m.act[l]); i(m.w_off[l]); i(m.b_off[l]); }
  }
  void made(const SurjMadeDesc& m) {
    i(m.input_size);
    if (m.input_size <= 0) return;
    i(m.n_params); i(m.n_hidden); i(m.act); i(m.ctx_dim);
    for (int l = 0; l < m.n_hidden; ++l) i(m.hidden[l]);
    for (int l = 0; l <= m.n_hidden; ++l) { i(m.w_off[l]); i(m.b_off[l]); }
    if (m.ctx_dim > 0) { i(m.ctx_w_off); i(m.ctx_b_off); }
  }
};
}  // namespace

extern "C" int surj_chain_key(const SurjLayerDesc* layers, int32_t n_layers, int32_t dim_in, int32_t dim_x,
                              int64_t n_params, uint64_t* key_out) {
  if (!layers || !key_out || n_layers <= 0) return SURJ_E_INVALID;
  Ser s;
  s.i(SURJ_ABI_VERSION); s.i(dim_in); s.i(dim_x); s.i(n_params); s.i(n_layers);
  for (int k = 0; k < n_layers; ++k) {
    const SurjLayerDesc& L = layers[k];
    s.i(L.kind); s.i(L.dim_in); s.i(L.dim_out); s.i(L.use_x); s.i(L.inner); s.i(L.layout); s.i(L.n_bins);
    s.f(L.range_min); s.f(L.range_max);
    if (L.mask) { s.i(L.dim_in); for (int d = 0; d < L.dim_in; ++d) s.i(L.mask[d] ? 1 : 0); } else s.i(0);
    if (L.permutation) { s.i(L.dim_in); for (int d = 0; d < L.dim_in; ++d) s.i(L.permutation[d]); } else s.i(0);
    s.mlp(L.conditioner); s.made(L.made); s.mlp(L.decoder);
    s.i(L.lu_lower_off); s.i(L.lu_upper_off); s.i(L.lu_diag_off);
  }
  uint64_t h = 1469598103934665603ull;
  const unsigned char* p = reinterpret_cast<const unsigned char*>(s.v.data());
  for (size_t n = 0; n < s.v.size() * sizeof(int64_t); ++n) { h ^= p[n]; h *= 1099511628211ull; }
  *key_out = h;
  return SURJ_OK;
}

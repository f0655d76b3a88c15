// Compact fused kernel, tanh instantiations (see compact_variants.h).
#include "compact_variants.h"

namespace cc {
const Variant* variants_tanh() {
  static const Variant v[N_VARIANTS] = CC_VARIANTS_FOR(surj::ACT_TANH);
  return v;
}
}  // namespace cc

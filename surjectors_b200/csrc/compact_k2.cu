// Compact fused kernel, swish instantiations (see compact_variants.h).
#include "compact_variants.h"

namespace cc {
const Variant* variants_swish() {
  static const Variant v[N_VARIANTS] = CC_VARIANTS_FOR(surj::ACT_SWISH);
  return v;
}
}  // namespace cc

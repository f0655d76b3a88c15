// Compact fused kernel, relu instantiations (see compact_variants.h).
#include "compact_variants.h"

namespace cc {
const Variant* variants_relu() {
  static const Variant v[N_VARIANTS] = CC_VARIANTS_FOR(surj::ACT_RELU);
  return v;
}
}  // namespace cc

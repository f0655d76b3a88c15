// Generic per-layer path (any width).  Placeholder until the wide-network kernels land.
#include "plan.h"

int surj_layered_create(SurjPlan*) {
  surj_set_error("per-layer path not built yet; only chains with a fused kernel compiled in are supported");
  return SURJ_E_UNSUPPORTED;
}
void surj_layered_destroy(SurjPlan*) {}
size_t surj_layered_workspace_bytes(const SurjPlan*, int64_t, int) { return 0; }
int surj_layered_launches(const SurjPlan*, int) { return 0; }
int surj_layered_fwd(const SurjPlan*, cudaStream_t, const float*, const float*, const float*, float*, float*, void*,
                     size_t, int64_t) { return SURJ_E_UNSUPPORTED; }
int surj_layered_bwd(const SurjPlan*, cudaStream_t, const float*, const float*, const float*, const float*, float,
                     const float*, float*, float*, float*, float*, float*, void*, size_t, int64_t) { return SURJ_E_UNSUPPORTED; }

// Internal plan object shared by the C ABI (capi.cu) and the per-layer path (layered.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/surj_b200.h"
#include "registry.h"

struct SurjPlan {
  std::vector<SurjLayerDesc> layers;           // mask / permutation pointers point into the vectors below
  std::vector<std::vector<uint8_t>> masks;
  std::vector<std::vector<int32_t>> perms;
  int32_t dim_in = 0, dim_x = 0, dim_out = 0;
  int64_t n_params = 0;
  int32_t mode = 0;
  int32_t base = 0;  // SurjBase
  uint64_t key = 0;
  const SmallKernels* fused = nullptr;
  // lazily filled device facts
  mutable bool dev_ready = false;
  mutable int n_sm = 148, fwd_blocks = 4, bwd_blocks = 2;
  // per-layer path (layered.cu)
  struct Layered* layered = nullptr;
};

void surj_set_error(const std::string& msg);
int surj_cuda_fail(cudaError_t e, const char* what);

// layered.cu
int surj_layered_create(SurjPlan* plan);
void surj_layered_destroy(SurjPlan* plan);
size_t surj_layered_workspace_bytes(const SurjPlan* plan, int64_t batch, int backward);
int surj_layered_launches(const SurjPlan* plan, int backward);
int surj_layered_fwd(const SurjPlan* plan, cudaStream_t st, const float* y, const float* x, const float* params,
                     float* z_out, float* lp_out, void* ws, size_t ws_bytes, int64_t B);
int surj_layered_bwd(const SurjPlan* plan, cudaStream_t st, const float* y, const float* x, const float* params,
                     const float* g_lp, float g_scale, const float* g_z, float* g_params, float* loss_out,
                     float* lp_out, float* g_y, float* g_x, void* ws, size_t ws_bytes, int64_t B);

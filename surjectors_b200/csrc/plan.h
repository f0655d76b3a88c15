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
  const SmallKernels* fused = nullptr;  // generation-1 fused kernel (per-chain generated code), if registered
  struct VmPlan* vmp = nullptr;         // generation-2 fused kernel (interpreter), if the chain fits
  struct CcPlan* ccp = nullptr;         // compact looped kernel for affine coupling / funnel chains (compact.cuh)
  // lazily filled device facts
  mutable bool dev_ready = false;
  mutable int device = -1;  // the device the plan was first used on (ensure_device)
  mutable int n_sm = 148, fwd_blocks = 4, bwd_blocks = 2;
  int exec = 0;  // 1 generated straight-line kernel, 2 fused interpreter kernel, 3 per-op executor, 4 compact looped kernel
};

void surj_set_error(const std::string& msg);
int surj_cuda_fail(cudaError_t e, const char* what);

// vm_plan.cu -- fused small-network interpreter path
struct VmPlan* vm_compile(const SurjPlan* plan, std::string* why);
void vm_destroy(struct VmPlan* v);
int vm_prepare(struct VmPlan* v);
int vm_grid(const struct VmPlan* v, int64_t B, int prog);
int vm_max_grid(const struct VmPlan* v, int64_t B);
int vm_launch(struct VmPlan* v, cudaStream_t st, int prog, int with_base, const float* y, const float* x,
              const float* params, const float* g_lp, float g_scale, const float* g_z, float* z_out, float* lp_out,
              float* g_y, float* g_x, float* partials, int npad, int64_t B, int grid, unsigned* tickets,
              int n_tickets, const float* eps = nullptr);

// Compact fused kernel for affine coupling / coupling-funnel chains (compact.cuh): host interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

struct SurjPlan;
struct CcPlan;

CcPlan* cc_compile(const SurjPlan* plan, std::string* why);  // nullptr (+ why): the chain is outside the kernel's family
void cc_destroy(CcPlan* c);
int cc_prepare(CcPlan* c);
int cc_grid(const CcPlan* c, int64_t B, bool backward);  // CTAs (= partial gradient rows) a launch uses; needs cc_prepare
int cc_max_grid(const CcPlan* c);
int cc_launch(CcPlan* c, cudaStream_t st, int backward, int with_base, const float* y, const float* params, const float* g_lp,
              float g_scale, const float* g_z, float* z_out, float* lp_out, float* g_y, float* partials, int npad, int64_t B,
              unsigned* tickets, int n_tickets);

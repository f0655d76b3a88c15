// Registry of fused whole-chain kernels compiled into the library, keyed by the canonical
// chain key (key.cpp).  Each generated translation unit registers itself at load time.
#pragma once
#include <cuda_runtime.h>

struct SmallKernels {
  int nt, d_in, d_x, d_out, n_params, fwd_smem, bwd_smem;
  cudaError_t (*prepare)();
  cudaError_t (*occupancy)(int* fwd_blocks, int* bwd_blocks);
  cudaError_t (*fwd)(int grid, cudaStream_t st, int with_base, const float* y, const float* x, const float* params,
                     float* z, float* lp, long long B);
  cudaError_t (*bwd)(int grid, cudaStream_t st, int with_base, const float* y, const float* x, const float* params,
                     const float* g_lp, float g_scale, const float* g_z, float* partials, int npad, float* lp,
                     float* z, float* g_y, float* g_x, long long B, unsigned* tickets, int n_tickets);
};

void surj_register_small(unsigned long long key, SmallKernels k, const char* name);
const SmallKernels* surj_find_small(unsigned long long key);
int surj_registry_count();

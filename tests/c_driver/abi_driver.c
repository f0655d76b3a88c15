/* Torch-free driver of the C ABI (include/surj_b200.h): plain C + the CUDA runtime, nothing else.  It is what a caller that
 * is neither Python nor PyTorch does with the library -- the call sequence of examples/coupling_inference_surjection.py:81-90
 * (trace the chain -> log_prob -> value_and_grad -> adam step) -- and checks every result against the float64 oracle's
 * numbers stored in a fixture (tests/c_driver/make_fixture.py):
 *
 *   surj_plan_create, surj_plan_workspace_bytes, surj_chain_logprob_fwd   log-probs within 1e-5 relative
 *   surj_chain_logprob_bwd                                               loss within 1e-5, gradient within 1e-4 of its max
 *   surj_train_step (+ surj_adam_step on a copy)                         both routes give the same parameters
 *   surj_index_take                                                      bit-exact gather
 *
 * build: gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include tests/c_driver/abi_driver.c -L surjectors_b200/_lib -lsurj_b200 \
 *            -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/surjectors_b200/_lib -o build/abi_driver
 * run:   build/abi_driver tests/golden/cdrv_cfg2.bin [mode]      (mode: 0 auto, 2 per-op executor) */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "surj_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
#define SK(x) do { int s_ = (x); if (s_ != SURJ_OK) { fprintf(stderr, "surj error %d (%s) at line %d\n", s_, surj_last_error(), __LINE__); return 3; } } while (0)

static void* dev_copy(const void* h, size_t bytes) {
  void* d = NULL;
  if (cudaMalloc(&d, bytes ? bytes : 16) != cudaSuccess) return NULL;
  if (bytes && cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
  return d;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s fixture.bin [mode]\n", argv[0]); return 1; }
  const int mode = argc > 2 ? atoi(argv[2]) : SURJ_MODE_AUTO;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  char magic[8];
  int32_t hd[5];
  int64_t np_b[2], aux_bytes;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SURJFIX1", 8) || fread(hd, 4, 5, f) != 5 || fread(np_b, 8, 2, f) != 2) { fprintf(stderr, "bad fixture\n"); return 1; }
  const int n_layers = hd[0], dim_in = hd[1], dim_x = hd[2], dim_out = hd[3];
  const int64_t n_params = np_b[0], B = np_b[1];
  if (hd[4] != (int32_t)sizeof(SurjLayerDesc)) { fprintf(stderr, "sizeof(SurjLayerDesc) %d != fixture's %d\n", (int)sizeof(SurjLayerDesc), hd[4]); return 1; }
  SurjLayerDesc* layers = (SurjLayerDesc*)malloc(sizeof(SurjLayerDesc) * (size_t)n_layers);
  if (fread(layers, sizeof(SurjLayerDesc), (size_t)n_layers, f) != (size_t)n_layers || fread(&aux_bytes, 8, 1, f) != 1) return 1;
  unsigned char* aux = (unsigned char*)malloc((size_t)aux_bytes + 8);
  if (aux_bytes && fread(aux, 1, (size_t)aux_bytes, f) != (size_t)aux_bytes) return 1;
  for (int k = 0; k < n_layers; ++k) {  /* relocate the two pointer fields: stored as offset + 1 into the blob, 0 = NULL */
    const intptr_t m = (intptr_t)layers[k].mask, p = (intptr_t)layers[k].permutation;
    layers[k].mask = m ? aux + (m - 1) : NULL;
    layers[k].permutation = p ? (const int32_t*)(aux + (p - 1)) : NULL;
  }
  float* h_params = (float*)malloc(4 * (size_t)n_params);
  float* h_y = (float*)malloc(4 * (size_t)(B * dim_in));
  float* h_x = (float*)malloc(4 * (size_t)(B * (dim_x ? dim_x : 1)));
  float* h_lp = (float*)malloc(4 * (size_t)B);
  float* h_g = (float*)malloc(4 * (size_t)n_params);
  double loss_ref;
  if (fread(h_params, 4, (size_t)n_params, f) != (size_t)n_params || fread(h_y, 4, (size_t)(B * dim_in), f) != (size_t)(B * dim_in)) return 1;
  if (dim_x && fread(h_x, 4, (size_t)(B * dim_x), f) != (size_t)(B * dim_x)) return 1;
  if (fread(h_lp, 4, (size_t)B, f) != (size_t)B || fread(h_g, 4, (size_t)n_params, f) != (size_t)n_params || fread(&loss_ref, 8, 1, f) != 1) return 1;
  fclose(f);

  CK(cudaSetDevice(0));
  SurjPlan* plan = NULL;
  SK(surj_plan_create(layers, n_layers, dim_in, dim_x, n_params, mode, SURJ_BASE_STANDARD_NORMAL, &plan));
  if (surj_plan_dim_out(plan) != dim_out) { fprintf(stderr, "dim_out mismatch\n"); return 1; }
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  float* d_params = (float*)dev_copy(h_params, 4 * (size_t)n_params);
  float* d_y = (float*)dev_copy(h_y, 4 * (size_t)(B * dim_in));
  float* d_x = dim_x ? (float*)dev_copy(h_x, 4 * (size_t)(B * dim_x)) : NULL;
  float *d_lp, *d_z, *d_g, *d_loss, *d_mu, *d_nu, *d_p2, *d_mu2, *d_nu2;
  int32_t *d_count, *d_count2;
  CK(cudaMalloc((void**)&d_lp, 4 * (size_t)B)); CK(cudaMalloc((void**)&d_z, 4 * (size_t)(B * dim_out)));
  CK(cudaMalloc((void**)&d_g, 4 * (size_t)(n_params + 4))); CK(cudaMalloc((void**)&d_loss, 16));
  CK(cudaMalloc((void**)&d_mu, 4 * (size_t)n_params)); CK(cudaMalloc((void**)&d_nu, 4 * (size_t)n_params));
  CK(cudaMalloc((void**)&d_mu2, 4 * (size_t)n_params)); CK(cudaMalloc((void**)&d_nu2, 4 * (size_t)n_params));
  CK(cudaMalloc((void**)&d_p2, 4 * (size_t)n_params)); CK(cudaMalloc((void**)&d_count, 4)); CK(cudaMalloc((void**)&d_count2, 4));
  const size_t wsf = surj_plan_workspace_bytes(plan, B, 0), wsb = surj_plan_workspace_bytes(plan, B, 1);
  void* ws = NULL;
  CK(cudaMalloc(&ws, (wsf > wsb ? wsf : wsb) + 256));
  int fails = 0;

  /* 1. log_prob */
  SK(surj_chain_logprob_fwd(plan, st, d_y, d_x, d_params, d_z, d_lp, ws, wsf, B));
  float* lp = (float*)malloc(4 * (size_t)B);
  CK(cudaMemcpyAsync(lp, d_lp, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double worst = 0.0;
  for (int64_t i = 0; i < B; ++i) {
    const double e = fabs((double)lp[i] - h_lp[i]) / fmax(1.0, fabs((double)h_lp[i]));
    if (e > worst) worst = e;
  }
  printf("log_prob: executor %d, %lld samples, max relative error %.3e (bound 1e-5)\n", surj_plan_is_fused(plan), (long long)B, worst);
  fails += !(worst <= 1e-5);

  /* 2. value_and_grad of loss = -sum(lp) */
  SK(surj_chain_logprob_bwd(plan, st, d_y, d_x, d_params, NULL, -1.0f, NULL, d_g, d_loss, NULL, NULL, NULL, ws, wsb, B));
  float* g = (float*)malloc(4 * (size_t)n_params);
  float loss;
  CK(cudaMemcpyAsync(g, d_g, 4 * (size_t)n_params, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&loss, d_loss, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double gmax = 0.0, gerr = 0.0;
  for (int64_t i = 0; i < n_params; ++i) {
    if (fabs((double)h_g[i]) > gmax) gmax = fabs((double)h_g[i]);
    if (fabs((double)g[i] - h_g[i]) > gerr) gerr = fabs((double)g[i] - h_g[i]);
  }
  const double lerr = fabs((double)loss - loss_ref) / fmax(1.0, fabs(loss_ref));
  printf("value_and_grad: loss %.6f (oracle %.6f, relative error %.3e, bound 1e-5); gradient max error / max |g| = %.3e (bound 1e-4)\n",
         loss, loss_ref, lerr, gerr / gmax);
  fails += !(lerr <= 1e-5) + !(gerr <= 1e-4 * gmax);

  /* 3. one optimisation step two ways: surj_train_step, and backward + surj_adam_step on copies */
  CK(cudaMemcpy(d_p2, d_params, 4 * (size_t)n_params, cudaMemcpyDeviceToDevice));
  CK(cudaMemset(d_mu, 0, 4 * (size_t)n_params)); CK(cudaMemset(d_nu, 0, 4 * (size_t)n_params)); CK(cudaMemset(d_count, 0, 4));
  CK(cudaMemset(d_mu2, 0, 4 * (size_t)n_params)); CK(cudaMemset(d_nu2, 0, 4 * (size_t)n_params)); CK(cudaMemset(d_count2, 0, 4));
  SK(surj_adam_step(st, d_p2, d_g, d_mu2, d_nu2, d_count2, n_params, 1e-3f, 0.9f, 0.999f, 1e-8f, 1.0f));
  SK(surj_train_step(plan, st, d_y, d_x, d_params, -1.0f, d_g, d_mu, d_nu, d_count, 1e-3f, 0.9f, 0.999f, 1e-8f, d_loss, ws, wsb, B));
  float* p1 = (float*)malloc(4 * (size_t)n_params);
  float* p2 = (float*)malloc(4 * (size_t)n_params);
  int32_t count = 0;
  CK(cudaMemcpyAsync(p1, d_params, 4 * (size_t)n_params, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(p2, d_p2, 4 * (size_t)n_params, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&count, d_count, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double perr = 0.0, moved = 0.0;
  for (int64_t i = 0; i < n_params; ++i) {
    if (fabs((double)p1[i] - p2[i]) > perr) perr = fabs((double)p1[i] - p2[i]);
    if (fabs((double)p1[i] - h_params[i]) > moved) moved = fabs((double)p1[i] - h_params[i]);
  }
  printf("train step: step counter %d, parameters moved by up to %.3e, surj_train_step vs backward + surj_adam_step differ by %.3e\n", count, moved, perr);
  fails += !(count == 1) + !(moved > 1e-4 && moved < 1.1e-3) + !(perr <= 2e-6);

  /* 4. surj_index_take: reversed rows of y */
  int32_t* h_idx = (int32_t*)malloc(4 * (size_t)B);
  for (int64_t i = 0; i < B; ++i) h_idx[i] = (int32_t)(B - 1 - i);
  int32_t* d_idx = (int32_t*)dev_copy(h_idx, 4 * (size_t)B);
  float* d_t;
  int32_t* d_err;
  CK(cudaMalloc((void**)&d_t, 4 * (size_t)(B * dim_in))); CK(cudaMalloc((void**)&d_err, 4)); CK(cudaMemset(d_err, 0, 4));
  SK(surj_index_take(st, d_y, B, dim_in, d_idx, 3, B - 3, d_t, d_err));
  float* t = (float*)malloc(4 * (size_t)(B * dim_in));
  int32_t err = 0;
  CK(cudaMemcpyAsync(t, d_t, 4 * (size_t)((B - 3) * dim_in), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  int bad = err;
  for (int64_t r = 0; r < B - 3 && !bad; ++r)
    bad = memcmp(t + r * dim_in, h_y + (B - 1 - (r + 3)) * dim_in, 4 * (size_t)dim_in) != 0;
  printf("index_take: %s\n", bad ? "MISMATCH" : "bit-exact");
  fails += bad;

  surj_plan_destroy(plan);
  printf("%s\n", fails ? "FAILED" : "OK");
  return fails ? 4 : 0;
}

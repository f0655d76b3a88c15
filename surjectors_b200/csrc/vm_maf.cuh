// MaskedAutoregressive in the SAMPLING direction as one kernel (masked_autoregressive.py:63-71):
//
//     y = 0;  for _ in range(D):  theta = MADE(y, x);  y, log_det = inner(theta).forward_and_log_det(z)
//
// The reference runs D full MADE passes; pass s only changes y_s (y_<s are already final, theta_s depends on y_<s only).
// So ONE incremental pass computes the same numbers: a hidden unit is final as soon as the inputs its mask lets through
// are -- unit j of hidden layer l is "ready" at step ready_l[j] (read off the staged masks at plan time; TFP's degrees
// are non-decreasing in the unit index, so the units that become ready at one step are a contiguous range) -- and step s
//   1. finalises, layer by layer, the hidden units that become ready at s   (each weight of the MADE is used ONCE)
//   2. theta_s = b + W_out[:, s] . h_last over the ready units (masked weights are exact zeros in the staged copy)
//   3. y_s = inner(theta_s).forward(z_s), log-det accumulated -- equal to the last pass's, whose inputs are the final y.
// A warp owns a sample: y, z, the hidden vectors and the context live in its slice of shared memory, every dot product is
// lanes over the contraction index + a shuffle reduction.  The 16 warps of a CTA walk the steps together: the rows of the
// [out][in] staged weight copy that a step needs are copied into shared memory once per CTA and step.  The per-op executor
// recognises the unrolled program (OP_ZERO + D copies of the MADE block) and launches this kernel instead (vm_global.cu).
// Measured (tools/time_maf_sampling.py, D = 128, hidden [512, 512], context 16, batch 65 536): 23.4 ms per layer against
// 178 ms for the 128 unrolled passes (7.6x); 370 ms per layer at batch 2^20.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vm.cuh"
#include "vm_ops.cuh"

namespace vmaf {

constexpr int MAX_LAYERS = 4;   // masked layers of the MADE (hidden + output)
constexpr int MAX_D = 256;
constexpr int WARPS = 16;

struct Desc {
  int z_row, y_row, D, P;             // latent rows in, event rows out, width, parameters per dim (2: affine)
  int inner, skip_lp;                 // vm::INNER_AFFINE / _INV as the FORWARD map to apply; skip_lp: do not add the log-det
  int n_layers;                       // masked layers
  int wt_off[MAX_LAYERS], ldt[MAX_LAYERS], b_off[MAX_LAYERS], n_in[MAX_LAYERS], n_out[MAX_LAYERS], act[MAX_LAYERS];
  int has_ctx, cc_row, ctx_dim, ctx_wt_off, ctx_ldt, ctx_b_off;  // context Linear added to layer 0's pre-activation (made.py:80-84)
  int act0;                           // activation of layer 0 (carried by the context LINEAR when there is one)
  unsigned short cum[MAX_LAYERS - 1][MAX_D + 1];  // cum[l][s] = hidden units of layer l that are ready at step <= s
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum_i w[i] * x[i], i < n (both in shared memory, n <= padded row length): lanes over i, four independent partial sums
__device__ __forceinline__ float warp_dot(const float* w, const float* x, int n, int lane) {
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
  int i = lane;
  for (; i + 96 < n; i += 128) {
    a0 = fmaf(w[i], x[i], a0); a1 = fmaf(w[i + 32], x[i + 32], a1);
    a2 = fmaf(w[i + 64], x[i + 64], a2); a3 = fmaf(w[i + 96], x[i + 96], a3);
  }
  for (; i < n; i += 32) a0 = fmaf(w[i], x[i], a0);
  return warp_sum((a0 + a1) + (a2 + a3));
}

// floats of shared memory one warp's sample state needs
__host__ __device__ inline int warp_floats(const Desc& d) {
  int n = 2 * ((d.D + 31) & ~31) + ((d.ctx_dim + 31) & ~31);
  for (int l = 0; l + 1 < d.n_layers; ++l) n += (d.n_out[l] + 31) & ~31;
  return n;
}
// The weight rows step s needs, in the order the warps consume them: per hidden layer the rows of the units that become
// ready (n_prev valid columns, + the context row for layer 0), then the P rows of theta_s.  Every row is padded to a
// multiple of 32 floats.  f(row pointer in global memory, valid columns) is called per row; returns the floats used.
template <class F>
__host__ __device__ inline int step_rows(const Desc& d, int s, F&& f) {
  int used = 0;
  const int L = d.n_layers;
  for (int l = 0; l + 1 < L; ++l) {
    const int j0 = s > 0 ? d.cum[l][s - 1] : 0, j1 = d.cum[l][s];
    const int n_prev = l == 0 ? s : (int)d.cum[l - 1][s];
    for (int j = j0; j < j1; ++j) {
      f((long long)d.wt_off[l] + (long long)j * d.ldt[l], n_prev, used);
      used += (n_prev + 31) & ~31;
      if (l == 0 && d.has_ctx) {
        f((long long)d.ctx_wt_off + (long long)j * d.ctx_ldt, d.ctx_dim, used);
        used += (d.ctx_dim + 31) & ~31;
      }
    }
  }
  const int n_hl = L > 1 ? (int)d.cum[L - 2][s] : s;
  for (int p = 0; p < d.P; ++p) {
    f((long long)d.wt_off[L - 1] + (long long)(s * d.P + p) * d.ldt[L - 1], n_hl, used);
    used += (n_hl + 31) & ~31;
  }
  return used;
}
// the largest step's footprint (host: sizes the staging buffer)
inline int max_step_floats(const Desc& d) {
  int m = 0;
  for (int s = 0; s < d.D; ++s) m = max(m, step_rows(d, s, [](long long, int, int) {}));
  return m;
}

// The CTA's 16 warps (16 samples) walk the steps TOGETHER: the rows of a step are copied once, cooperatively and
// coalesced, into a shared-memory buffer (two buffers: step s + 1 is fetched while step s is consumed) and every warp
// reads them from there -- each weight of the MADE crosses L2 -> SM once per 16 samples.
__global__ void __launch_bounds__(32 * WARPS)
vg_maf_forward(const __grid_constant__ Desc d, float* __restrict__ rows, int C, int cols, const float* __restrict__ par, int wbuf_floats) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = d.D, Dp = (D + 31) & ~31;
  float* wbuf[2] = {smem, smem + wbuf_floats};
  float* z = smem + 2 * (size_t)wbuf_floats + (size_t)warp * warp_floats(d);
  float* y = z + Dp;
  float* ctx = y + Dp;
  float* h[MAX_LAYERS - 1];
  {
    float* p = ctx + ((d.ctx_dim + 31) & ~31);
    for (int l = 0; l + 1 < d.n_layers; ++l) { h[l] = p; p += (d.n_out[l] + 31) & ~31; }
  }
  const int L = d.n_layers;
  auto fetch = [&](int s, float* buf) {  // all threads: row by row, coalesced, zero-padded to 32
    step_rows(d, s, [&](long long off, int n, int at) {
      const int np = (n + 31) & ~31;
      for (int i = threadIdx.x; i < np; i += 32 * WARPS) buf[at + i] = i < n ? __ldg(par + off + i) : 0.0f;
    });
  };
  const int groups = (cols + WARPS - 1) / WARPS;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int col = grp * WARPS + warp;
    const bool live = col < cols;
    if (live) {
      for (int i = lane; i < D; i += 32) { z[i] = rows[(size_t)(d.z_row + i) * C + col]; y[i] = 0.0f; }
      for (int i = lane; i < d.ctx_dim; i += 32) ctx[i] = rows[(size_t)(d.cc_row + i) * C + col];
    }
    fetch(0, wbuf[0]);
    __syncthreads();
    float ld_sum = 0.0f;
    for (int s = 0; s < D; ++s) {
      const float* wb = wbuf[s & 1];
      if (s + 1 < D) fetch(s + 1, wbuf[(s + 1) & 1]);  // in flight behind this step's arithmetic
      int at = 0;
      // 1. hidden units that become ready at this step, layer by layer
      for (int l = 0; l + 1 < L; ++l) {
        const int j0 = s > 0 ? d.cum[l][s - 1] : 0, j1 = d.cum[l][s];
        const float* prev = l == 0 ? y : h[l - 1];
        const int n_prev = l == 0 ? s : (int)d.cum[l - 1][s];
        const int np = (n_prev + 31) & ~31, nc = (d.ctx_dim + 31) & ~31;
        for (int j = j0; j < j1; ++j) {
          float pre = warp_dot(wb + at, prev, n_prev, lane);
          at += np;
          if (d.b_off[l] >= 0) pre += __ldg(par + d.b_off[l] + j);
          if (l == 0 && d.has_ctx) {
            pre += warp_dot(wb + at, ctx, d.ctx_dim, lane);
            at += nc;
            if (d.ctx_b_off >= 0) pre += __ldg(par + d.ctx_b_off + j);
          }
          if (lane == 0) h[l][j] = vm::act_fwd(l == 0 ? d.act0 : d.act[l], pre);
        }
        __syncwarp();
      }
      // 2. theta of dim s from the final units of the last hidden layer
      const float* hl = L > 1 ? h[L - 2] : y;
      const int n_hl = L > 1 ? (int)d.cum[L - 2][s] : s;
      const int nh = (n_hl + 31) & ~31, o = L - 1;
      float th0 = warp_dot(wb + at, hl, n_hl, lane);
      float th1 = warp_dot(wb + at + nh, hl, n_hl, lane);
      if (d.b_off[o] >= 0) { th0 += __ldg(par + d.b_off[o] + s * d.P); th1 += __ldg(par + d.b_off[o] + s * d.P + 1); }
      // 3. the inner bijector's forward map (theta interleaved: mean / shift, log-scale)
      float out, ld;
      if (d.inner == vm::INNER_AFFINE) { out = (z[s] - th0) * expf(-th1); ld = -th1; }
      else { out = fmaf(expf(th1), z[s], th0); ld = th1; }
      ld_sum += ld;
      if (lane == 0) y[s] = out;
      __syncthreads();  // step s + 1's rows have landed; buffer s & 1 is free for step s + 2
    }
    if (live) {
      for (int i = lane; i < D; i += 32) rows[(size_t)(d.y_row + i) * C + col] = y[i];
      if (lane == 0 && !d.skip_lp) rows[(size_t)vm::ROW_LP * C + col] += ld_sum;
    }
    __syncthreads();
  }
}

}  // namespace vmaf

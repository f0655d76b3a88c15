// Whole-chain kernels built from the layer templates in smallnet.cuh, and the
// registration hook that makes one specialised chain visible to the C ABI.
#pragma once
#include "registry.h"
#include "smallnet.cuh"

namespace surj {

template <int A, int B>
struct CMax {
  static constexpr int v = A > B ? A : B;
};

// ---- left fold over the layers (chain.py:25-32) with the base log-prob at the end
template <int DX, class... Ls>
struct ChainOps;

template <int DX, class L, class... Rest>
struct ChainOps<DX, L, Rest...> {
  using Next = ChainOps<DX, Rest...>;
  static constexpr bool LAST = sizeof...(Rest) == 0;
  static constexpr int DIN = L::DIN;
  static constexpr int DFINAL = LAST ? L::DOUT : Next::DFINAL;
  static constexpr int MAXW = CMax<L::MAXW, Next::MAXW>::v;
  static constexpr int NSCR = L::NSCR + Next::NSCR;

  SURJ_DEV static void prologue(const float* __restrict__ sp, float* scr, int tid, int nt) {
    L::prologue(sp, scr, tid, nt);
    if constexpr (!LAST) Next::prologue(sp, scr + L::NSCR, tid, nt);
  }

  SURJ_DEV static void inverse(const float* __restrict__ sp, const float* scr, const float (&in)[L::DIN],
                               const float (&x)[XArr<DX>::n], float (&out)[DFINAL], float& lc) {
    if constexpr (LAST) {
      L::inverse(sp, scr, in, x, out, lc);
    } else {
      float mid[L::DOUT];
      L::inverse(sp, scr, in, x, mid, lc);
      Next::inverse(sp, scr + L::NSCR, mid, x, out, lc);
    }
  }

  // forward + backward in one recursion: layer inputs stay live on the way down, each layer's
  // internals are recomputed on the way back.  lp accumulates lc and the base log-prob;
  // g_lp is d(loss)/d(lp) for this sample.
  template <class Sink>
  SURJ_DEV static void backward(const float* __restrict__ sp, const float* scr, const float (&in)[L::DIN],
                                const float (&x)[XArr<DX>::n], float g_lp, bool with_base,
                                const float* __restrict__ g_z_row, float (&g_in)[L::DIN],
                                float (&g_x)[XArr<DX>::n], float& lp, float (&z_final)[DFINAL], Sink& sink) {
    float mid[L::DOUT], g_mid[L::DOUT];
    L::inverse(sp, scr, in, x, mid, lp);
    if constexpr (LAST) {
      // base = Independent(Normal(0, 1)): lp_z = sum(-z^2/2 - log(2pi)/2); d lp_z / dz = -z
#pragma unroll
      for (int d = 0; d < L::DOUT; ++d) {
        float g = 0.0f;
        if (with_base) {
          lp += -0.5f * mid[d] * mid[d] - SURJ_HALF_LOG_2PI;
          g = -g_lp * mid[d];
        }
        if (g_z_row != nullptr) g += g_z_row[d];
        g_mid[d] = g;
        z_final[d] = mid[d];
      }
    } else {
      Next::template backward<Sink>(sp, scr + L::NSCR, mid, x, g_lp, with_base, g_z_row, g_mid, g_x, lp, z_final,
                                    sink);
    }
    L::template backward<Sink>(sp, scr, in, x, g_mid, g_lp, g_in, g_x, sink);
  }
};

template <int DX>
struct ChainOps<DX> {
  static constexpr int DFINAL = 0, MAXW = 0, NSCR = 0;
};

template <int D0_, int DX_, int NP_, class... Ls>
struct Chain {
  static constexpr int D0 = D0_, DX = DX_, NP = NP_;
  using Ops = ChainOps<DX_, Ls...>;
  static constexpr int DZ = Ops::DFINAL;
  static constexpr int NP4 = (NP_ + 3) / 4 * 4;
  static constexpr int NSCR4 = (Ops::NSCR + 3) / 4 * 4;
};

// ---- row load / store with the widest vector the row stride allows
template <int D>
SURJ_DEV void load_row(const float* __restrict__ p, float (&v)[D]) {
  if constexpr (D % 4 == 0) {
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else if constexpr (D % 2 == 0) {
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(p) + i);
      v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = __ldg(p + i);
  }
}
template <int D>
SURJ_DEV void store_row(float* __restrict__ p, const float (&v)[D]) {
  if constexpr (D % 4 == 0) {
#pragma unroll
    for (int i = 0; i < D / 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (D % 2 == 0) {
#pragma unroll
    for (int i = 0; i < D / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) p[i] = v[i];
  }
}

template <int NP4>
SURJ_DEV void stage_params(float* sp, const float* __restrict__ params, int np, int tid, int nt) {
  const int n4 = np / 4;
  for (int i = tid; i < n4; i += nt) reinterpret_cast<float4*>(sp)[i] = __ldg(reinterpret_cast<const float4*>(params) + i);
  for (int i = 4 * n4 + tid; i < np; i += nt) sp[i] = __ldg(params + i);
}

// ------------------------------------------------------------------ forward kernel
// lp[s] = base.log_prob(z_s) + sum_layers lc_s ; optionally z.  Persistent grid-stride loop.
template <class CH, int NT>
__global__ void __launch_bounds__(NT)
chain_fwd_kernel(int with_base, const float* __restrict__ y, const float* __restrict__ x,
                 const float* __restrict__ params, float* __restrict__ z_out, float* __restrict__ lp_out, long long B) {
  extern __shared__ __align__(16) float smem[];
  float* sp = smem;
  float* scr = smem + CH::NP4;
  const int tid = threadIdx.x;
  stage_params<CH::NP4>(sp, params, CH::NP, tid, NT);
  __syncthreads();
  CH::Ops::prologue(sp, scr, tid, NT);
  __syncthreads();
  for (long long s = (long long)blockIdx.x * NT + tid; s < B; s += (long long)gridDim.x * NT) {
    float v[CH::D0], xc[XArr<CH::DX>::n], z[CH::DZ];
    load_row<CH::D0>(y + s * CH::D0, v);
    if constexpr (CH::DX > 0) load_row<CH::DX>(x + s * CH::DX, xc);
    else xc[0] = 0.0f;
    float lp = 0.0f;
    CH::Ops::inverse(sp, scr, v, xc, z, lp);
    if (with_base) {
#pragma unroll
      for (int d = 0; d < CH::DZ; ++d) lp += -0.5f * z[d] * z[d] - SURJ_HALF_LOG_2PI;
    }
    lp_out[s] = lp;
    if (z_out != nullptr) store_row<CH::DZ>(z_out + s * CH::DZ, z);
  }
}

// ----------------------------------------------------------------- backward kernel
// VJP of lp w.r.t. the parameters (and optionally y, x) with cotangent g_lp[s] (or the
// constant g_scale when g_lp == nullptr, e.g. -1 for loss = -sum lp).  Each block accumulates
// its parameter gradient in shared memory and writes one row of `partials`
// ([gridDim.x][npad], column NP holds the block's sum of lp).
template <class CH, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
chain_bwd_kernel(int with_base, const float* __restrict__ y, const float* __restrict__ x,
                 const float* __restrict__ params, const float* __restrict__ g_lp, float g_scale,
                 const float* __restrict__ g_z, float* __restrict__ partials, int npad,
                 float* __restrict__ lp_out, float* __restrict__ z_out, float* __restrict__ g_y,
                 float* __restrict__ g_x_out, long long B, unsigned* __restrict__ tickets, int n_tickets) {
  extern __shared__ __align__(16) float smem[];
  float* sp = smem;
  float* scr = sp + CH::NP4;
  float* sg = scr + CH::NSCR4;
  if (blockIdx.x == 0)  // arm the tickets of the gradient-reduction kernel that follows on the stream
    for (int i = threadIdx.x; i < n_tickets; i += NT) tickets[i] = 0u;             // [NP4 + 4]: gradient accumulator, slot NP = sum lp
  float* tile = sg + CH::NP4 + 4;
  const int tid = threadIdx.x;
  stage_params<CH::NP4>(sp, params, CH::NP, tid, NT);
  for (int i = tid; i < CH::NP4 + 4; i += NT) sg[i] = 0.0f;
  __syncthreads();
  CH::Ops::prologue(sp, scr, tid, NT);
  __syncthreads();
  BlockSink<NT> sink{sg, tile, tid};
  float lp_acc = 0.0f;
  for (long long base = (long long)blockIdx.x * NT; base < B; base += (long long)gridDim.x * NT) {
    const long long s = base + tid;
    const bool valid = s < B;
    float v[CH::D0], xc[XArr<CH::DX>::n], g_in[CH::D0], g_x[XArr<CH::DX>::n], z[CH::DZ];
    if (valid) {
      load_row<CH::D0>(y + s * CH::D0, v);
      if constexpr (CH::DX > 0) load_row<CH::DX>(x + s * CH::DX, xc);
      else xc[0] = 0.0f;
    } else {
#pragma unroll
      for (int d = 0; d < CH::D0; ++d) v[d] = 0.0f;
#pragma unroll
      for (int k = 0; k < XArr<CH::DX>::n; ++k) xc[k] = 0.0f;
    }
#pragma unroll
    for (int k = 0; k < XArr<CH::DX>::n; ++k) g_x[k] = 0.0f;
    const float g = valid ? (g_lp != nullptr ? __ldg(g_lp + s) : g_scale) : 0.0f;
    float lp = 0.0f;
    const float* g_z_row = (g_z != nullptr && valid) ? g_z + s * CH::DZ : nullptr;
    CH::Ops::template backward<BlockSink<NT>>(sp, scr, v, xc, g, with_base != 0, g_z_row, g_in, g_x, lp, z, sink);
    if (valid) {
      lp_acc += lp;
      if (lp_out != nullptr) lp_out[s] = lp;
      if (z_out != nullptr) store_row<CH::DZ>(z_out + s * CH::DZ, z);
      if (g_y != nullptr) store_row<CH::D0>(g_y + s * CH::D0, g_in);
      if constexpr (CH::DX > 0) {
        if (g_x_out != nullptr) store_row<CH::DX>(g_x_out + s * CH::DX, g_x);
      }
    }
  }
  // block sum of lp -> sg[NP]
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lp_acc += __shfl_xor_sync(0xffffffffu, lp_acc, o);
  if ((tid & 31) == 0) atomicAdd(sg + CH::NP, lp_acc);
  __syncthreads();
  float* row = partials + (size_t)blockIdx.x * npad;
  for (int i = tid; i < CH::NP + 1; i += NT) row[i] = sg[i];
}

// ------------------------------------------------------------------- registration
// MINB = CTAs per SM the backward kernel is register-capped for (more resident warps hide the
// instruction-fetch and shared-memory latency of the straight-line code)
template <class CH, int NT, int MINB = 1>
struct SmallLaunch {
  static constexpr int FWD_SMEM = (CH::NP4 + CH::NSCR4) * 4;
  static constexpr int BWD_SMEM = (CH::NP4 + CH::NSCR4 + CH::NP4 + 4 + CH::Ops::MAXW * (NT + 4)) * 4;

  static cudaError_t prepare() {
    cudaError_t e = cudaFuncSetAttribute(chain_fwd_kernel<CH, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(chain_bwd_kernel<CH, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
  }
  static cudaError_t occupancy(int* fwd_blocks, int* bwd_blocks) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(fwd_blocks, chain_fwd_kernel<CH, NT>, NT, FWD_SMEM);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bwd_blocks, chain_bwd_kernel<CH, NT, MINB>, NT, BWD_SMEM);
  }
  static cudaError_t fwd(int grid, cudaStream_t st, int with_base, const float* y, const float* x,
                         const float* params, float* z, float* lp, long long B) {
    chain_fwd_kernel<CH, NT><<<grid, NT, FWD_SMEM, st>>>(with_base, y, x, params, z, lp, B);
    return cudaGetLastError();
  }
  static cudaError_t bwd(int grid, cudaStream_t st, int with_base, const float* y, const float* x,
                         const float* params, const float* g_lp, float g_scale, const float* g_z, float* partials,
                         int npad, float* lp, float* z, float* g_y, float* g_x, long long B, unsigned* tickets,
                         int n_tickets) {
    chain_bwd_kernel<CH, NT, MINB><<<grid, NT, BWD_SMEM, st>>>(with_base, y, x, params, g_lp, g_scale, g_z, partials, npad, lp,
                                                               z, g_y, g_x, B, tickets, n_tickets);
    return cudaGetLastError();
  }
  static SmallKernels make() {
    SmallKernels k;
    k.nt = NT; k.d_in = CH::D0; k.d_x = CH::DX; k.d_out = CH::DZ; k.n_params = CH::NP;
    k.fwd_smem = FWD_SMEM; k.bwd_smem = BWD_SMEM;
    k.prepare = &prepare; k.occupancy = &occupancy; k.fwd = &fwd; k.bwd = &bwd;
    return k;
  }
};

}  // namespace surj

#define SURJ_REGISTER_CHAIN(KEY, CHAIN_T, NT, MINB)                                                          \
  namespace {                                                                                                \
  struct Reg_##KEY {                                                                                         \
    Reg_##KEY() { surj_register_small(KEY##ull, surj::SmallLaunch<CHAIN_T, NT, MINB>::make(), #CHAIN_T); }   \
  } reg_instance_##KEY;                                                                                      \
  }

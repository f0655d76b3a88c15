// Interpreter kernel of the fused small-network path (see vm.cuh for the design notes).
#pragma once
#include "vm_ops.cuh"

namespace vm {

// out[j] = act(init + b[j] + sum_i in[i] W[i][j]) for a chunk of CH outputs at a time.
template <int NT, int S, bool FAR, int CH>
VM_DEV void linear_chunks(Ctx<NT, S, FAR>& c, const float* __restrict__ par, const Op& op) {
  constexpr int T = NT * S;
  const int in_row = op.a[0], n_in = op.a[1], w_off = op.a[2], ld = op.a[3], out_row = op.a[4], n_out = op.a[5],
            b_off = op.a[6], act = op.a[7], accumulate = op.a[8], pre_row = op.a[9];
  const float* in = c.row(in_row);
  const int acc_row = pre_row >= 0 ? pre_row : out_row;
  for (int j0 = 0; j0 < n_out; j0 += CH) {
    float acc[S][CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const float b = b_off >= 0 ? par[b_off + j0 + q] : 0.0f;
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s][q] = b;
    }
    if (accumulate) {
#pragma unroll
      for (int q = 0; q < CH; ++q)
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s][q] += c.row(acc_row + j0 + q)[s * NT];
    }
    const float* w = par + w_off + j0;
#pragma unroll 4
    for (int i = 0; i < n_in; ++i) {
      float wv[CH];
#pragma unroll
      for (int q = 0; q < CH; q += 4) {
        const float4 t = *reinterpret_cast<const float4*>(w + (size_t)i * ld + q);
        wv[q] = t.x; wv[q + 1] = t.y; wv[q + 2] = t.z; wv[q + 3] = t.w;
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float a = in[(size_t)i * T + s * NT];
#pragma unroll
        for (int q = 0; q < CH; ++q) acc[s][q] = fmaf(a, wv[q], acc[s][q]);
      }
    }
    with_act(act, [&](auto tag) {
      constexpr int ACT = decltype(tag)::value;
#pragma unroll
      for (int q = 0; q < CH; ++q)
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float v = acc[s][q];
          if (pre_row >= 0) c.row(pre_row + j0 + q)[s * NT] = v;
          c.row(out_row + j0 + q)[s * NT] = act_fwd(ACT, v);
        }
    });
  }
}

template <int NT, int S, bool FAR>
VM_DEV void op_linear(Ctx<NT, S, FAR>& c, const float* __restrict__ par, const Op& op) {
  if ((op.a[5] & 15) == 0 && S <= 2) linear_chunks<NT, S, FAR, 16>(c, par, op);
  else linear_chunks<NT, S, FAR, 8>(c, par, op);
}

// Weight gradient: grad[W] += in^T g, grad[b] += sum g, contracted over the T samples of the tile by
// the whole CTA.  Work item = a 4x4 tile of (input, output) pairs x one of KS slices of the sample
// axis; the KS lanes of one tile are adjacent, combine with shuffles, and the first of them is the
// only thread of the CTA that ever updates those accumulator addresses (no atomics, deterministic).
template <int NT, int S, bool FAR>
VM_DEV void op_linear_dw(Ctx<NT, S, FAR>& c, float* __restrict__ grad, const Op& op) {
  constexpr int T = NT * S;
  const int in_row = op.a[0], n_in = op.a[1], g_row = op.a[2], n_out = op.a[3], w_off = op.a[4], ld = op.a[5],
            b_off = op.a[6];
  __syncthreads();
  const int NA = n_in + (b_off >= 0 ? 1 : 0), NG = n_out;
  const int NTA = (NA + 3) >> 2, NTG = (NG + 3) >> 2, NE = NTA * NTG;
  int KS = 1;
  while (KS < 8 && NE * KS * 2 <= NT) KS <<= 1;
  const int KLEN = T / KS, ITEMS = NE * KS;
  const float* rows = c.rows;
  for (int w0 = 0; w0 < ITEMS; w0 += NT) {
    const int w = w0 + c.col;
    const bool live = w < ITEMS;
    const int wc = live ? w : 0;
    const int e = wc / KS, ks = wc - e * KS;
    const int ta = e % NTA, tg = e / NTA;
    const float* ap[4];
    const float* gp[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int ia = min(ta + a * NTA, NA - 1);
      ap[a] = rows + (size_t)(ia < n_in ? in_row + ia : (int)ROW_ONES) * T + ks * KLEN;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) gp[g] = rows + (size_t)(g_row + min(tg + g * NTG, NG - 1)) * T + ks * KLEN;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[a][g] = 0.0f;
#pragma unroll 2
    for (int k = 0; k < KLEN; k += 4) {
      float4 av[4], gv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(ap[a] + k);
#pragma unroll
      for (int g = 0; g < 4; ++g) gv[g] = *reinterpret_cast<const float4*>(gp[g] + k);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          acc[a][g] = fmaf(av[a].x, gv[g].x, acc[a][g]);
          acc[a][g] = fmaf(av[a].y, gv[g].y, acc[a][g]);
          acc[a][g] = fmaf(av[a].z, gv[g].z, acc[a][g]);
          acc[a][g] = fmaf(av[a].w, gv[g].w, acc[a][g]);
        }
    }
    for (int o = KS >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[a][g] += __shfl_xor_sync(0xffffffffu, acc[a][g], o);
    if (live && ks == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ia = ta + a * NTA, ig = tg + g * NTG;
          if (ia < NA && ig < NG) {
            const int o = ia < n_in ? w_off + ia * ld + ig : b_off + ig;
            grad[o] += acc[a][g];
          }
        }
    }
  }
  __syncthreads();
}

template <int NT, int S>
__global__ void __launch_bounds__(NT)
vm_kernel(Program P, Io io, const float* __restrict__ params, float* __restrict__ partials, int npad, long long B,
          unsigned* __restrict__ tickets, int n_tickets) {
  constexpr int T = NT * S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int backward = io.backward;
  if (blockIdx.x == 0)  // arm the tickets of the gradient-reduction kernel that follows on the stream
    for (int i = tid; i < n_tickets; i += NT) tickets[i] = 0u;
  const int n_ops = P.n_ops;
  // carve shared memory (every region 16-byte aligned)
  auto align4 = [](int n) { return (n + 3) & ~3; };
  Op* sops = reinterpret_cast<Op*>(smem_raw);
  int32_t* sitab = reinterpret_cast<int32_t*>(sops + n_ops);
  float* par = reinterpret_cast<float*>(sitab + align4(P.n_itab));
  float* scal = par + align4(P.n_par);
  float* scal_g = scal + align4(P.n_scalars);
  float* grad = scal_g + align4(P.n_scalars);
  float* rows = grad + (backward ? align4(P.n_prim + 1) : 0);

  for (int i = tid; i < n_ops * (int)(sizeof(Op) / 4); i += NT)
    reinterpret_cast<int32_t*>(sops)[i] = reinterpret_cast<const int32_t*>(P.ops)[i];
  for (int i = tid; i < P.n_itab; i += NT) sitab[i] = P.itab[i];
  for (int i = tid; i < P.n_par; i += NT) par[i] = stage_value(P.stage[i], params);
  if (backward)
    for (int i = tid; i < P.n_prim + 1; i += NT) grad[i] = 0.0f;
  for (int i = tid; i < P.n_scalars; i += NT) scal_g[i] = 0.0f;
  // row 0 is the constant-ones row (carries the bias in the weight-gradient contraction)
  for (int i = tid; i < T; i += NT) rows[i] = 1.0f;
  __syncthreads();
  for (int j = tid; j < P.n_sjobs; j += NT) {  // LULinear log-dets (lu_linear.py:82-84)
    const ScalarJob sj = P.sjobs[j];
    float sum = 0.0f;
    for (int i = 0; i < sj.n; ++i) sum += logf(par[sj.off + i * sj.stride]);
    scal[sj.slot] = sum;
  }
  __syncthreads();

  Ctx<NT, S> c;
  c.rows = rows; c.T = T; c.col = tid; c.B = B; c.itab = sitab; c.scal = scal; c.scal_g = scal_g;
  float lp_total = 0.0f;
  const long long n_tiles = (B + T - 1) / T;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    c.base = tile * T;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      c.lp[s] = 0.0f;
      const long long g = c.sample(s);
      c.glp[s] = (backward && g < B) ? (io.g_lp != nullptr ? __ldg(io.g_lp + g) : io.g_scale) : 0.0f;
    }
    for (int pc = 0; pc < n_ops; ++pc) {
      const Op& op = sops[pc];
      if (op.code == OP_LINEAR) op_linear<NT, S, false>(c, par, op);
      else if (op.code == OP_LINEAR_DW) op_linear_dw<NT, S, false>(c, grad, op);
      else exec_simple<NT, S, false>(c, op, io, lp_total);
    }
  }
  if (!backward) return;
  // ---- write this CTA's partial gradient row: [n_params] gradients, then sum(lp)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lp_total += __shfl_xor_sync(0xffffffffu, lp_total, o);
  if ((tid & 31) == 0) atomicAdd(grad + P.n_prim, lp_total);
  float* out = partials + (size_t)blockIdx.x * npad;
  for (int i = tid; i < npad; i += NT) out[i] = 0.0f;
  __syncthreads();
  for (int i = tid; i < P.n_prim; i += NT) writeback_grad(P.grad_dst[i], P.grad_aux[i], grad[i], par[i], scal_g, params, out);
  if (tid == 0) out[P.n_params] = grad[P.n_prim];
}

}  // namespace vm

// Per-sample ("simple") ops of the chain program: everything except the matrix products.
// A thread owns S sample columns of the row workspace and touches nothing else, so these ops
// need no synchronisation.  Shared by the fused in-shared-memory interpreter (vm_kernels.cuh) and
// the per-op global-memory executor (vm_global.cu).
#pragma once
#include "vm.cuh"
#include "vm_rqs.cuh"

namespace vm {

enum : int { ROW_ONES = 0, ROW_LP = 1, ROW_GLP = 2, ROW_FIRST_FREE = 3 };

struct Io {  // caller buffers of one C-ABI call
  const float* y; const float* x; const float* g_lp; const float* g_z;
  const float* eps; int n_eps;  // sampling direction: standard-normal draws [B, n_eps] supplied by the caller
  float* z_out; float* lp_out; float* g_y; float* g_x;
  float g_scale; int with_base; int backward;
};

// FAR: the rows live in global memory (per-op executor) -- per-sample op loops then issue a block of row loads before
// the first use; in shared memory (fused interpreter) the plain one-row-at-a-time loops are cheaper in registers.
template <int NT, int S, bool FAR = false>
struct Ctx {
  static constexpr bool kFar = FAR;
  float* rows;          // [n_rows][T]
  int T;                // row stride in floats
  int col;              // first column of this thread
  long long base;       // global sample index of column 0
  long long B;          // batch size
  const int32_t* itab;
  const float* scal;    // [n_scalars] per-plan scalars (LULinear log-dets)
  float* scal_g;        // [n_scalars] accumulated cotangents of those scalars
  float lp[S];
  float glp[S];
  VM_DEV float* row(int r) const { return rows + (size_t)r * T + col; }
  VM_DEV long long sample(int s) const { return base + col + s * NT; }
};

template <int NT, int S>
struct ThRow {  // theta accessor: rows th0 + k, this thread's column
  float* p;
  int T;
  VM_DEV float& operator()(int k) const { return p[(size_t)k * T]; }
};

template <int NT, int S, bool FAR>
VM_DEV void op_load(Ctx<NT, S, FAR>& c, const Op& op, const float* __restrict__ src) {
  const int dst = op.a[0], n = op.a[1];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const long long g = c.sample(s);
    float* r = c.row(dst) + s * NT;
    if (g < c.B) {
      const float* p = src + g * n;
      if ((n & 3) == 0) {
        for (int d = 0; d < n; d += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p + d));
          r[(size_t)d * c.T] = t.x;
          r[(size_t)(d + 1) * c.T] = t.y;
          r[(size_t)(d + 2) * c.T] = t.z;
          r[(size_t)(d + 3) * c.T] = t.w;
        }
      } else {
        for (int d = 0; d < n; ++d) r[(size_t)d * c.T] = __ldg(p + d);
      }
    } else {
      for (int d = 0; d < n; ++d) r[(size_t)d * c.T] = 0.0f;
    }
  }
}

template <int NT, int S, bool FAR>
VM_DEV void op_store(Ctx<NT, S, FAR>& c, const Op& op, float* __restrict__ dst) {
  if (dst == nullptr) return;
  const int src = op.a[0], n = op.a[1];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const long long g = c.sample(s);
    if (g >= c.B) continue;
    const float* r = c.row(src) + s * NT;
    float* p = dst + g * n;
    if ((n & 3) == 0) {
      for (int d = 0; d < n; d += 4)
        *reinterpret_cast<float4*>(p + d) = make_float4(r[(size_t)d * c.T], r[(size_t)(d + 1) * c.T],
                                                        r[(size_t)(d + 2) * c.T], r[(size_t)(d + 3) * c.T]);
    } else {
      for (int d = 0; d < n; ++d) p[d] = r[(size_t)d * c.T];
    }
  }
}

template <int NT, int S, bool FAR>
VM_DEV void op_affine(Ctx<NT, S, FAR>& c, const Op& op) {
  const int in_row = op.a[0], out_row = op.a[1], n_t = op.a[2], it = op.a[3], sh = op.a[4], ss = op.a[5], ls = op.a[6],
            lss = op.a[7], inner = op.a[8];
  const float lpw = op.a[9] ? 0.0f : 1.0f;
  // U dims at a time, all loads issued before the first use: out of global memory (per-op executor) a thread's rows are
  // ~1 us away, and one load -> exp -> store chain per dim left the kernel at 1 TB/s (ncu: long_scoreboard 19-42)
  constexpr int U = FAR ? 4 : 1;
  for (int k0 = 0; k0 < n_t; k0 += U) {
    int d[U];
    float y[U][S], m[U][S], l[U][S];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = min(k0 + u, n_t - 1);
      d[u] = c.itab[it + k];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        y[u][s] = c.row(in_row + d[u])[s * NT];
        m[u][s] = c.row(sh + k * ss)[s * NT];
        l[u][s] = c.row(ls + k * lss)[s * NT];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (k0 + u >= n_t) break;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float z;
        if (inner == INNER_AFFINE) {  // ScalarAffine(m, exp(l)).inverse: z = (y - m) e^{-l}, ld = -l
          z = (y[u][s] - m[u][s]) * expf(-l[u][s]);
          c.lp[s] -= lpw * l[u][s];
        } else {  // Inverse(ScalarAffine).inverse = forward: z = e^{l} y + m, ld = +l
          z = fmaf(expf(l[u][s]), y[u][s], m[u][s]);
          c.lp[s] += lpw * l[u][s];
        }
        c.row(out_row + d[u])[s * NT] = z;
      }
    }
  }
}

// cotangents; thetas are overwritten by their cotangents
template <int NT, int S, bool FAR>
VM_DEV void op_affine_bwd(Ctx<NT, S, FAR>& c, const Op& op) {
  const int in_row = op.a[0], n_t = op.a[1], it = op.a[2], sh = op.a[3], ss = op.a[4], ls = op.a[5], lss = op.a[6],
            inner = op.a[7], gout = op.a[8], gin = op.a[9];
  constexpr int U = FAR ? 4 : 1;
  for (int k0 = 0; k0 < n_t; k0 += U) {
    int d[U];
    float y[U][S], m[U][S], l[U][S], gz[U][S];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = min(k0 + u, n_t - 1);
      d[u] = c.itab[it + k];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        y[u][s] = c.row(in_row + d[u])[s * NT];
        m[u][s] = c.row(sh + k * ss)[s * NT];
        l[u][s] = c.row(ls + k * lss)[s * NT];
        gz[u][s] = c.row(gout + d[u])[s * NT];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u;
      if (k >= n_t) break;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float glc = c.glp[s];
        float gy, gm, gs;
        if (inner == INNER_AFFINE) {
          const float e = expf(-l[u][s]);
          const float z = (y[u][s] - m[u][s]) * e;
          gy = gz[u][s] * e;
          gm = -gy;
          gs = -gz[u][s] * z - glc;
        } else {
          const float e = expf(l[u][s]);
          gy = gz[u][s] * e;
          gm = gz[u][s];
          gs = gz[u][s] * e * y[u][s] + glc;
        }
        c.row(sh + k * ss)[s * NT] = gm;
        c.row(ls + k * lss)[s * NT] = gs;
        c.row(gin + d[u])[s * NT] = gy;
      }
    }
  }
}

template <int NT, int S, bool FAR>
VM_DEV void op_rqs(Ctx<NT, S, FAR>& c, const Op& op) {
  const int in_row = op.a[0], out_row = op.a[1], n_t = op.a[2], it = op.a[3], th = op.a[4], K = op.a[5];
  const bool forward = op.a[6] == INNER_RQS_INV;
  const int P = 3 * K + 1;
  for (int k = 0; k < n_t; ++k) {
    const int d = c.itab[it + k];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float v = c.row(in_row + d)[s * NT];
      ThRow<NT, S> acc{c.row(th + k * P) + s * NT, c.T};
      RqsBin b;
      rqs_locate(acc, K, op.f0, op.f1, v, forward, b);
      float out, ld;
      rqs_eval(b, K, op.f0, op.f1, v, forward, out, ld);
      c.row(out_row + d)[s * NT] = out;
      if (!op.a[7]) c.lp[s] += ld;
    }
  }
}

template <int NT, int S, bool FAR>
VM_DEV void op_rqs_bwd(Ctx<NT, S, FAR>& c, const Op& op) {
  const int in_row = op.a[0], n_t = op.a[1], it = op.a[2], th = op.a[3], K = op.a[4];
  const bool forward = op.a[5] == INNER_RQS_INV;
  const int gout = op.a[6], gin = op.a[7];
  const int P = 3 * K + 1;
  for (int k = 0; k < n_t; ++k) {
    const int d = c.itab[it + k];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float v = c.row(in_row + d)[s * NT];
      ThRow<NT, S> acc{c.row(th + k * P) + s * NT, c.T};
      c.row(gin + d)[s * NT] = rqs_bwd(acc, K, op.f0, op.f1, v, forward, c.row(gout + d)[s * NT], c.glp[s]);
    }
  }
}

template <int NT, int S, bool FAR>
VM_DEV void op_normal_lp(Ctx<NT, S, FAR>& c, const Op& op) {
  const int v = op.a[0], n = op.a[1], mu = op.a[2], ls = op.a[3];
  for (int j = 0; j < n; ++j) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float l = c.row(ls + j)[s * NT];
      const float r = (c.row(v + j)[s * NT] - c.row(mu + j)[s * NT]) * expf(-l);
      c.lp[s] += -0.5f * r * r - l - VM_HALF_LOG_2PI;
    }
  }
}

template <int NT, int S, bool FAR>
VM_DEV void op_normal_lp_bwd(Ctx<NT, S, FAR>& c, const Op& op) {
  const int v = op.a[0], n = op.a[1], mu = op.a[2], ls = op.a[3], gv = op.a[4];
  for (int j = 0; j < n; ++j) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float* pm = c.row(mu + j) + s * NT;
      float* pl = c.row(ls + j) + s * NT;
      const float e = expf(-*pl);
      const float r = (c.row(v + j)[s * NT] - *pm) * e;
      const float g = c.glp[s];
      *pm = g * r * e;
      *pl = g * (r * r - 1.0f);
      c.row(gv + j)[s * NT] = -g * r * e;
    }
  }
}

// Executes one simple op; returns false if `op` is a matrix op the caller must handle.
// lp_total accumulates the log-probs of valid samples stored by OP_STORE_LP (for the loss).
// LITE: without the spline ops (their register footprint sets the occupancy of the whole interpreter kernel; the per-op
// executor launches runs that contain none of them -- all of them once the spline layers are fused -- on a copy built
// without).
template <int NT, int S, bool FAR, bool LITE = false>
VM_DEV bool exec_simple(Ctx<NT, S, FAR>& c, const Op& op, const Io& io, float& lp_total) {
  constexpr int UR = FAR ? 8 : 1, UA = FAR ? 4 : 1;  // rows of loads in flight in the copy-like / activation loops
  switch (op.code) {
    case OP_LOAD: op_load<NT, S, FAR>(c, op, op.a[2] == 0 ? io.y : io.x); return true;
    case OP_STORE: op_store<NT, S, FAR>(c, op, op.a[2] == 0 ? io.z_out : (op.a[2] == 1 ? io.g_y : io.g_x)); return true;
    case OP_GATHER:
      for (int k0 = 0; k0 < op.a[1]; k0 += UR) {  // UR rows in flight
        float v[UR][S];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
          const float* src = c.row(c.itab[op.a[2] + min(k0 + u, op.a[1] - 1)]);
#pragma unroll
          for (int s = 0; s < S; ++s) v[u][s] = src[s * NT];
        }
#pragma unroll
        for (int u = 0; u < UR; ++u) {
          if (k0 + u >= op.a[1]) break;
#pragma unroll
          for (int s = 0; s < S; ++s) c.row(op.a[0] + k0 + u)[s * NT] = v[u][s];
        }
      }
      return true;
    case OP_SCATTER_ADD:
      for (int k = 0; k < op.a[1]; ++k) {
        const float* src = c.row(op.a[0] + k);
        float* dst = c.row(c.itab[op.a[2] + k]);
#pragma unroll
        for (int s = 0; s < S; ++s) dst[s * NT] += src[s * NT];
      }
      return true;
    case OP_COPY:
    case OP_ADD:
      for (int k0 = 0; k0 < op.a[2]; k0 += UR) {  // UR rows in flight
        float v[UR][S];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
          const int k = min(k0 + u, op.a[2] - 1);
#pragma unroll
          for (int s = 0; s < S; ++s)
            v[u][s] = c.row(op.a[1] + k)[s * NT] + (op.code == OP_ADD ? c.row(op.a[0] + k)[s * NT] : 0.0f);
        }
#pragma unroll
        for (int u = 0; u < UR; ++u) {
          if (k0 + u >= op.a[2]) break;
#pragma unroll
          for (int s = 0; s < S; ++s) c.row(op.a[0] + k0 + u)[s * NT] = v[u][s];
        }
      }
      return true;
    case OP_ZERO:
      for (int k = 0; k < op.a[1]; ++k)
#pragma unroll
        for (int s = 0; s < S; ++s) c.row(op.a[0] + k)[s * NT] = 0.0f;
      return true;
    case OP_AFFINE: op_affine<NT, S, FAR>(c, op); return true;
    case OP_RQS: if constexpr (LITE) return false; else { op_rqs<NT, S, FAR>(c, op); return true; }
    case OP_NORMAL_LP: op_normal_lp<NT, S, FAR>(c, op); return true;
    case OP_NORMAL_SAMPLE:
      for (int j = 0; j < op.a[1]; ++j)
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const long long g = c.sample(s);
          const float e = g < c.B ? __ldg(io.eps + g * io.n_eps + op.a[4] + j) : 0.0f;
          const float l = c.row(op.a[3] + j)[s * NT];
          c.row(op.a[0] + j)[s * NT] = fmaf(expf(l), e, c.row(op.a[2] + j)[s * NT]);
          c.lp[s] += -0.5f * e * e - l - VM_HALF_LOG_2PI;
        }
      return true;
    case OP_ADD_SCALAR:
#pragma unroll
      for (int s = 0; s < S; ++s) c.lp[s] += c.scal[op.a[0]];
      return true;
    case OP_BASE:
      if (io.with_base) {
        for (int k = 0; k < op.a[1]; ++k)
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float z = c.row(op.a[0] + k)[s * NT];
            c.lp[s] += -0.5f * z * z - VM_HALF_LOG_2PI;
          }
      }
      return true;
    case OP_STORE_LP:
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const long long g = c.sample(s);
        if (g < c.B) {
          if (io.lp_out != nullptr) io.lp_out[g] = c.lp[s];
          lp_total += c.lp[s];
        }
      }
      return true;
    case OP_BASE_BWD:
      for (int k = 0; k < op.a[1]; ++k)
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const long long g = c.sample(s);
          float v = io.with_base ? -c.glp[s] * c.row(op.a[0] + k)[s * NT] : 0.0f;
          if (io.g_z != nullptr && g < c.B) v += __ldg(io.g_z + g * op.a[1] + k);
          c.row(op.a[2] + k)[s * NT] = v;
        }
      return true;
    case OP_NORMAL_LP_BWD: op_normal_lp_bwd<NT, S, FAR>(c, op); return true;
    case OP_AFFINE_BWD: op_affine_bwd<NT, S, FAR>(c, op); return true;
    case OP_RQS_BWD: if constexpr (LITE) return false; else { op_rqs_bwd<NT, S, FAR>(c, op); return true; }
    case OP_ACT_BWD:
      for (int k0 = 0; k0 < op.a[2]; k0 += UA) {  // UA rows in flight
        float g[UA][S], a[UA][S];
#pragma unroll
        for (int u = 0; u < UA; ++u) {
          const int k = min(k0 + u, op.a[2] - 1);
#pragma unroll
          for (int s = 0; s < S; ++s) { g[u][s] = c.row(op.a[0] + k)[s * NT]; a[u][s] = c.row(op.a[1] + k)[s * NT]; }
        }
#pragma unroll
        for (int u = 0; u < UA; ++u) {
          if (k0 + u >= op.a[2]) break;
#pragma unroll
          for (int s = 0; s < S; ++s) c.row(op.a[0] + k0 + u)[s * NT] = FAR ? act_bwd_fast(op.a[3], a[u][s], g[u][s]) : act_bwd(op.a[3], a[u][s], g[u][s]);
        }
      }
      return true;
    case OP_SCALAR_BWD: {
      float v = 0.0f;
#pragma unroll
      for (int s = 0; s < S; ++s) v += c.glp[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(c.scal_g + op.a[0], v);
      return true;
    }
    default: return false;
  }
}

}  // namespace vm

"""Build-time generator of fused whole-chain kernels.

For small conditioner networks the whole Chain is ONE kernel in which every width, parameter
offset, coupling mask, MADE mask and permutation is a compile-time constant
(surjectors_b200/csrc/smallnet.cuh, chain.cuh).  This module turns a `ChainSpec` into the C++
type that instantiates those templates, keyed by the canonical chain key (csrc/key.cpp) under
which the C ABI looks fused kernels up at plan time.  `__graft_entry__.build()` runs it for the
chains listed in `surjectors_b200/fused_chains.py` and compiles the result into
libsurj_b200.so; chains that are not listed run on the generic per-layer kernels.
"""

from __future__ import annotations

import numpy as np

from surjectors_b200 import _abi
from surjectors_b200._spec import ChainSpec, LayerSpec, MadeSpec, MlpSpec
from surjectors_b200.nn import made_degrees

MAX_BITS = 64


class NotFusable(ValueError):
  pass


def _u64(bits):
  v = 0
  for i, b in enumerate(bits):
    if b:
      v |= 1 << i
  return f"0x{v:x}ull"


def _lin(n_in, n_out, act, w, b, pred="surj::PAll"):
  return f"surj::Lin<{n_in}, {n_out}, {_abi.ACT[act]}, {w.offset}, {b.offset if b is not None else -1}, {pred}>"


def _mlp(m: MlpSpec, row_bits=None, col_bits=None):
  """Mlp<...> with the first layer's rows / last layer's columns pruned to the given bit sets."""
  lins = []
  n = m.n_layers
  for l in range(n):
    rows = row_bits if (l == 0 and row_bits is not None) else None
    cols = col_bits if (l == n - 1 and col_bits is not None) else None
    pred = "surj::PAll"
    if rows is not None or cols is not None:
      if m.widths[l] > MAX_BITS or m.widths[l + 1] > MAX_BITS:
        raise NotFusable("pruned layer wider than 64")
      r = _u64(rows if rows is not None else [1] * m.widths[l])
      c = _u64(cols if cols is not None else [1] * m.widths[l + 1])
      pred = f"surj::PRowCol<{r}, {c}>"
    lins.append(_lin(m.widths[l], m.widths[l + 1], m.acts[l], m.w[l], m.b[l], pred))
  return "surj::Mlp<" + ", ".join(lins) + ">"


def _made(m: MadeSpec):
  D, P = m.input_size, m.n_params
  degs = made_degrees(D, m.hidden)
  floors = []
  for l in range(len(m.hidden)):
    floors.append(min(int(np.min(degs[l])), D - 1))

  def deg_t(l):  # type of the degree function of layer l's units (0 = input)
    if l == 0:
      return f"surj::DegInput<{D}>"
    return f"surj::DegHidden<{D}, {m.hidden[l - 1]}, {floors[l - 1]}>"

  # self-check of the closed form against the numpy construction
  for l in range(1, len(m.hidden) + 1):
    U, M = m.hidden[l - 1], floors[l - 1]
    closed = [max(M, ((u + 1) * (D - 1) + U) // (U + 1)) for u in range(U)]
    if closed != [int(v) for v in degs[l]]:
      raise AssertionError("closed-form MADE degrees disagree with the TFP construction")

  nh = len(m.hidden)
  l0 = _lin(D, m.hidden[0], m.act, m.w[0], m.b[0], f"surj::PMadeHidden<{deg_t(0)}, {deg_t(1)}>")
  ctx = "surj::NoCtx"
  if m.ctx_dim > 0:
    ctx = _lin(m.ctx_dim, m.hidden[0], "identity", m.ctx_w, m.ctx_b)
  rest = []
  for l in range(1, nh):
    rest.append(_lin(m.hidden[l - 1], m.hidden[l], m.act, m.w[l], m.b[l],
                     f"surj::PMadeHidden<{deg_t(l)}, {deg_t(l + 1)}>"))
  rest.append(_lin(m.hidden[-1], D * P, "identity", m.w[nh], m.b[nh], f"surj::PMadeOut<{deg_t(nh)}, {P}>"))
  return f"surj::Made<{l0}, {ctx}, surj::Mlp<{', '.join(rest)}>>"


def _b(v):
  return "true" if v else "false"


def _theta_cols(D, layout, transformed):
  cols = [0] * (2 * D)
  for d in range(D):
    if transformed[d]:
      if layout == _abi.LAYOUT_SPLIT:
        cols[d] = cols[D + d] = 1
      else:
        cols[2 * d] = cols[2 * d + 1] = 1
  return cols


def _layer(L: LayerSpec, dx: int):
  use_x = bool(L.use_x)
  D, K = L.dim_in, L.dim_out
  if L.kind in (_abi.LAYER_COUPLING, _abi.LAYER_COUPLING_FUNNEL, _abi.LAYER_MAF, _abi.LAYER_MAF_FUNNEL):
    if L.inner not in (_abi.INNER_AFFINE, _abi.INNER_AFFINE_INV):
      raise NotFusable("spline inner bijectors run on the per-layer kernels")
  if D > MAX_BITS or dx > MAX_BITS:
    raise NotFusable("event wider than 64")
  if L.kind == _abi.LAYER_COUPLING:
    passes = [bool(v) for v in L.mask]
    rows = passes + [1] * (dx if use_x else 0)
    cols = _theta_cols(D, L.layout, [not p for p in passes])
    cond = _mlp(L.conditioner, rows, cols)
    return f"surj::Coupling<{D}, {dx}, {_b(use_x)}, {_u64(passes)}, {L.inner}, {L.layout}, {cond}>"
  if L.kind == _abi.LAYER_COUPLING_FUNNEL:
    passes = [d >= K for d in range(D)]
    rows = passes + [1] * (dx if use_x else 0)
    cols = _theta_cols(D, L.layout, [not p for p in passes])
    cond = _mlp(L.conditioner, rows, cols)
    return (f"surj::CouplingFunnel<{D}, {K}, {dx}, {_b(use_x)}, {L.inner}, {L.layout}, {cond}, "
            f"{_mlp(L.decoder)}>")
  if L.kind == _abi.LAYER_MAF:
    if L.made.n_params != 2:
      raise NotFusable("MADE with n_params != 2")
    return f"surj::Maf<{D}, {dx}, {_b(use_x)}, {L.inner}, {_made(L.made)}>"
  if L.kind == _abi.LAYER_MAF_FUNNEL:
    if L.made.n_params != 2:
      raise NotFusable("MADE with n_params != 2")
    return f"surj::MafFunnel<{D}, {K}, {dx}, {_b(use_x)}, {L.inner}, {_made(L.made)}, {_mlp(L.decoder)}>"
  if L.kind == _abi.LAYER_SLICE:
    return f"surj::SliceL<{D}, {K}, {dx}, {_b(use_x)}, {_mlp(L.decoder)}>"
  if L.kind == _abi.LAYER_LU:
    lo, up, dg = L.lu
    return f"surj::LuL<{D}, {dx}, {lo.offset}, {up.offset}, {dg.offset}>"
  if L.kind == _abi.LAYER_PERMUTATION:
    inv = np.zeros(D, dtype=np.int32)
    inv[np.asarray(L.permutation)] = np.arange(D, dtype=np.int32)  # permutation.py:31-36
    return f"surj::PermL<{dx}, {', '.join(str(int(v)) for v in inv)}>"
  if L.kind == _abi.LAYER_MLP_FUNNEL:
    raise NotFusable("MLPInferenceFunnel runs on the chain program executors")
  raise NotFusable(f"unknown layer kind {L.kind}")


def chain_type(spec: ChainSpec) -> str:
  layers = ",\n    ".join(_layer(L, spec.dim_x) for L in spec.layers)
  return f"surj::Chain<{spec.dim_in}, {spec.dim_x}, {spec.n_params},\n    {layers}>"


def emit_translation_unit(spec: ChainSpec, key: int, name: str, nt: int = 128, min_blocks: int = 1) -> str:
  tag = f"0x{key:016x}"
  return (
    f"// GENERATED by surjectors_b200/_codegen.py for chain '{name}' -- do not edit.\n"
    f'#include "../chain.cuh"\n\n'
    f"using Chain_{tag} = {chain_type(spec)};\n"
    f"SURJ_REGISTER_CHAIN({tag}, Chain_{tag}, {nt}, {min_blocks})\n"
  )

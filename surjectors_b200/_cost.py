"""Algorithmic work of a chain per sample (what bench.py divides by the measured kernel time).

FLOPs = 2 x (structurally non-zero multiply-adds) of every Linear / MaskedLinear / LU product plus
one per bias add.  "needed" removes work the reference's dense formulation spends on structural
zeros (masked MADE weights, conditioner rows of transformed dims, conditioner columns of
pass-through dims, triangular halves of L / U); "dense" is what the reference's XLA graph
multiplies.  Transcendentals and the elementwise epilogue are not counted (conservative).
A training step is counted as 3x the forward (forward + input-grad + weight-grad products).
Bytes = the float32 inputs read and the log-probs written (parameters are L2/SMEM resident).
"""

import numpy as np

from surjectors_b200 import _abi
from surjectors_b200.nn import make_dense_autoregressive_masks


def _mlp(m, rows=None, cols=None):
  need = dense = 0
  for l in range(m.n_layers):
    i, o = m.widths[l], m.widths[l + 1]
    ri = rows if (l == 0 and rows is not None) else i
    co = cols if (l == m.n_layers - 1 and cols is not None) else o
    need += 2 * ri * co + co
    dense += 2 * i * o + o
  return need, dense


def _made(m):
  masks = make_dense_autoregressive_masks(m.n_params, m.input_size, m.hidden)
  need = dense = 0
  for mk in masks:
    need += 2 * int(mk.sum()) + mk.shape[1]
    dense += 2 * mk.size + mk.shape[1]
  if m.ctx_dim > 0:
    c = 2 * m.ctx_dim * m.hidden[0] + m.hidden[0]
    need += c
    dense += c
  return need, dense


def _npar(L):
  return 2 if L.inner in (_abi.INNER_AFFINE, _abi.INNER_AFFINE_INV) else 3 * L.n_bins + 1


def forward_flops(spec):
  """(needed, dense) FLOPs per sample of one log_prob evaluation."""
  need = dense = 0
  dx = spec.dim_x
  for L in spec.layers:
    x = dx if L.use_x else 0
    D, K = L.dim_in, L.dim_out
    if L.kind in (_abi.LAYER_COUPLING, _abi.LAYER_COUPLING_FUNNEL):
      n_pass = int(np.sum(L.mask)) if L.kind == _abi.LAYER_COUPLING else D - K
      n, d = _mlp(L.conditioner, rows=n_pass + x, cols=(D - n_pass) * _npar(L))
      need, dense = need + n, dense + d
    if L.kind in (_abi.LAYER_MAF, _abi.LAYER_MAF_FUNNEL):
      n, d = _made(L.made)
      need, dense = need + n, dense + d
    if L.kind in (_abi.LAYER_SLICE, _abi.LAYER_COUPLING_FUNNEL, _abi.LAYER_MAF_FUNNEL):
      n, d = _mlp(L.decoder)
      need, dense = need + n, dense + d
    if L.kind == _abi.LAYER_LU:
      need += 2 * D * D
      dense += 4 * D * D
    if L.kind == _abi.LAYER_MLP_FUNNEL:
      n, d = _mlp(L.conditioner)
      need, dense = need + n + 2 * K * K, dense + d + 4 * K * K
      n, d = _mlp(L.decoder)
      need, dense = need + n, dense + d
  return need, dense


def forward_bytes(spec):
  return 4 * (spec.dim_in + spec.dim_x) + 4

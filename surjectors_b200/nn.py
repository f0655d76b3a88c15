"""Conditioner networks: `make_mlp`, `MADE`, `MaskedLinear` (the reference's `surjectors.nn`,
surjectors/nn/__init__.py:1-7).  Like everything in the host layer these record structure; the
arithmetic is in the CUDA kernels."""

from __future__ import annotations

import numpy as np

from surjectors_b200 import _sym, hk
from surjectors_b200._spec import MadeSpec

__all__ = ["make_mlp", "MADE", "MaskedLinear", "make_dense_autoregressive_masks"]


def _degrees(event_size, hidden_units):
  """TFP `_create_degrees(input_order='left-to-right', hidden_degrees='equal')`."""
  degrees = [np.arange(1, event_size + 1)]
  for units in hidden_units:
    min_degree = min(int(np.min(degrees[-1])), event_size - 1)
    degrees.append(
      np.maximum(min_degree, np.ceil(np.arange(1, units + 1) * (event_size - 1) / float(units + 1)).astype(np.int32))
    )
  return degrees


def made_degrees(event_size, hidden_units):
  return [np.asarray(d, dtype=np.int64) for d in _degrees(event_size, hidden_units)]


def make_dense_autoregressive_masks(params, event_size, hidden_units):
  """TFP `_make_dense_autoregressive_masks` (called at surjectors/_src/conditioners/nn/made.py:54-56):
  bool masks [in, out] per layer; output column index is d * params + p."""
  degrees = _degrees(event_size, hidden_units)
  masks = [inp[:, np.newaxis] <= out for inp, out in zip(degrees[:-1], degrees[1:])]
  last = degrees[-1][:, np.newaxis] < degrees[0]
  masks.append(np.repeat(last, params, axis=1))
  return masks


def make_mlp(dims, activation=_sym.swish, w_init=None, b_init=hk.zeros):
  """surjectors/_src/conditioners/mlp.py:7-29: hk.nets.MLP with swish, TruncatedNormal(0.01), zero bias."""
  if w_init is None:
    w_init = hk.initializers.TruncatedNormal(stddev=0.01)
  return hk.nets.MLP(output_sizes=dims, w_init=w_init, b_init=b_init, activation=activation)


class MaskedLinear(hk.Module):
  """surjectors/_src/conditioners/nn/masked_linear.py:9-58: x @ (w * mask) + b."""

  def __init__(self, mask, with_bias=True, w_init=None, b_init=None):
    super().__init__()
    self.mask = np.asarray(mask)
    self.output_size = self.mask.shape[-1]
    self.with_bias = with_bias
    self.w_init = w_init
    self.b_init = b_init or hk.zeros

  @hk.scoped
  def params(self):
    w_init = self.w_init
    if w_init is None:
      w_init = hk.initializers.TruncatedNormal(stddev=1.0 / np.sqrt(self.mask.shape[0]))
    w = hk.get_parameter("w", self.mask.shape, init=w_init)
    if not self.with_bias:
      raise _sym.Unsupported("MaskedLinear without bias is not implemented")
    b = hk.get_parameter("b", (self.mask.shape[-1],), init=self.b_init)
    return w, b


class MADE(hk.Module):
  """surjectors/_src/conditioners/nn/made.py:12-91."""

  def __init__(self, input_size, hidden_layer_sizes, n_params, w_init=None, b_init=None, activation=_sym.relu):
    super().__init__()
    self.input_size = int(input_size)
    self.output_sizes = tuple(int(h) for h in hidden_layer_sizes)
    self.n_params = int(n_params)
    self.w_init = w_init
    self.b_init = b_init
    self.activation = activation
    if len(self.output_sizes) < 1:
      raise _sym.Unsupported("MADE needs at least one hidden layer")
    masks = make_dense_autoregressive_masks(self.n_params, self.input_size, self.output_sizes)
    self.layers = tuple(MaskedLinear(mask=m, w_init=w_init, b_init=b_init) for m in masks)

  @hk.scoped
  def __call__(self, y, x=None):
    if not isinstance(y, _sym.InputSym) or (x is not None and not isinstance(x, _sym.InputSym)):
      raise _sym.Unsupported("MADE can only be used as the conditioner of a MaskedAutoregressive layer / funnel")
    if y.width != self.input_size:
      raise ValueError(f"MADE(input_size={self.input_size}) applied to an input of width {y.width}")
    ws, bs = [], []
    w, b = self.layers[0].params()
    ws.append(w)
    bs.append(b)
    ctx_w = ctx_b = None
    if x is not None:  # made.py:80-84: the context Linear is created inside __call__
      net = hk.Linear(self.output_sizes[0], w_init=self.w_init, b_init=self.b_init)(x)
      ctx_w, ctx_b = net.layers[0].w, net.layers[0].b
    for layer in self.layers[1:]:
      w, b = layer.params()
      ws.append(w)
      bs.append(b)
    spec = MadeSpec(input_size=self.input_size, n_params=self.n_params, hidden=list(self.output_sizes),
                    act=_sym.activation_name(self.activation), ctx_dim=0 if x is None else x.width, w=ws, b=bs,
                    ctx_w=ctx_w, ctx_b=ctx_b)
    out = _sym.MadeSym(self, y, x)
    out.spec = spec
    return out

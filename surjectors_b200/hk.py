"""Haiku-shaped host layer: `transform` / `init` / `apply`, module naming and parameters.

The reference is written against dm-haiku (examples/coupling_inference_surjection.py:69-76:
`hk.transform(_flow)`, `model.init(rng, method="log_prob", **batch)`, `model.apply(params, rng,
...)`; lu_linear.py:58-66 and masked_linear.py:47-52 call `hk.get_parameter`).  Under
`jax.jit` haiku re-runs the user's python to *trace* the computation.  This module keeps that
programming model but the trace produces plain-old-data descriptors (surjectors_b200/_spec.py)
for the CUDA kernels instead of XLA ops: modules called on symbolic inputs record their structure,
nothing here computes on arrays and there is no eager fallback.

Parameter trees have haiku's names and shapes ('mlp/~/linear_0': {'w': [in, out], 'b': [out]},
'made/~/masked_linear_1', 'lu_linear': {lower_entries, upper_entries, diag_entries}).  The values
are float32 torch tensors that are views into ONE flat buffer (16-byte aligned slots), which is
what the kernels read.
"""

from __future__ import annotations

import math
import re
import threading

import numpy as np
import torch

from surjectors_b200 import _sym
from surjectors_b200._spec import ParamLayout, ParamRef

_tls = threading.local()


class _Frame:
  def __init__(self, params, rng, is_init):
    self.params = params  # module -> {name: value}; numpy while initialising, tensors on apply
    self.rng = rng
    self.is_init = is_init
    self.layout = ParamLayout()
    self.counters = {}
    self.module_stack = []
    self.device = None
    self.traced = []  # ChainSpecs traced in this frame (in call order)


def current_frame() -> _Frame:
  fr = getattr(_tls, "frame", None)
  if fr is None:
    raise RuntimeError("must be called inside hk.transform (init / apply)")
  return fr


def _snake(name):
  s1 = re.sub("(.)([A-Z][a-z]+)", r"\1_\2", name)
  return re.sub("([a-z0-9])([A-Z])", r"\1_\2", s1).lower()


class _ModuleMeta(type):
  def __call__(cls, *args, **kwargs):
    obj = cls.__new__(cls)
    fr = current_frame()
    fr.module_stack.append((obj, "__init__"))
    try:
      obj.__init__(*args, **kwargs)
    finally:
      fr.module_stack.pop()
    if not hasattr(obj, "module_name"):
      raise RuntimeError("hk.Module subclasses must call super().__init__()")
    return obj


class Module(metaclass=_ModuleMeta):
  """hk.Module naming: snake_case class name, `_<n>` per scope, `parent/~/child` for children
  built in `__init__`, `parent/child` for children built in a method."""

  def __init__(self, name=None):
    fr = current_frame()
    base = name if name is not None else _snake(type(self).__name__)
    prefix = ""
    if len(fr.module_stack) >= 2:
      parent, method = fr.module_stack[-2]
      prefix = parent.module_name + ("/~/" if method == "__init__" else "/")
    key = prefix + base
    n = fr.counters.get(key, 0)
    fr.counters[key] = n + 1
    self.module_name = key if n == 0 else f"{key}_{n}"


def scoped(fn):
  """Run a module method with the module on the naming stack."""

  def wrapper(self, *a, **k):
    fr = current_frame()
    fr.module_stack.append((self, fn.__name__))
    try:
      return fn(self, *a, **k)
    finally:
      fr.module_stack.pop()

  wrapper.__name__ = fn.__name__
  wrapper.__doc__ = fn.__doc__
  return wrapper


def get_parameter(name, shape, dtype=None, init=None) -> ParamRef:
  """hk.get_parameter: creates the value on `init`, checks it on `apply`; returns the slot of
  the parameter in the flat buffer (not an array -- kernels read the flat buffer)."""
  fr = current_frame()
  if not fr.module_stack:
    raise RuntimeError("get_parameter outside of a module")
  mod = fr.module_stack[-1][0].module_name
  shape = tuple(int(s) for s in shape)
  if fr.is_init:
    bucket = fr.params.setdefault(mod, {})
    if name not in bucket:
      if init is None:
        raise ValueError(f"{mod}/{name}: initializer must be specified")
      bucket[name] = np.asarray(init(shape, np.float32), dtype=np.float32).reshape(shape)
  else:
    if mod not in fr.params or name not in fr.params[mod]:
      raise ValueError(f"parameter {mod}/{name} is missing from the params passed to apply")
    got = tuple(fr.params[mod][name].shape)
    if got != shape:
      raise ValueError(f"{mod}/{name}: expected shape {shape}, got {got}")
  return fr.layout.add(mod, name, shape)


def next_rng_key():
  return current_frame().rng


def _as_generator(rng):
  if rng is None or isinstance(rng, np.random.Generator):
    return rng
  if isinstance(rng, torch.Generator):
    return np.random.Generator(np.random.PCG64(int(rng.initial_seed())))
  return np.random.Generator(np.random.PCG64(int(rng)))


class PRNGSequence:
  def __init__(self, seed):
    self._ss = np.random.SeedSequence(int(seed))

  def __next__(self):
    (child,) = self._ss.spawn(1)
    return np.random.Generator(np.random.PCG64(child))

  next = __next__


class FlatParams(dict):
  """The params tree returned by `init`: module -> {name: tensor}, every tensor a view into
  `self.flat` (float32, layout = creation order, 16-byte aligned slots)."""

  flat: torch.Tensor
  layout: ParamLayout


def pack_params(params, layout: ParamLayout, device) -> torch.Tensor:
  """Flat float32 buffer holding `params` in `layout`.  Zero-copy when the tree already consists
  of views into one correctly laid-out buffer (what `init` and the optimizers here produce)."""
  flat = getattr(params, "flat", None)
  if (
    isinstance(flat, torch.Tensor)
    and flat.device == torch.device(device)
    and getattr(params, "layout", None) is not None
    and params.layout.signature() == layout.signature()
    and flat.numel() == layout.size
  ):
    base = flat.data_ptr()
    ok = True
    for r in layout.refs:
      t = params[r.module][r.name]
      if not (isinstance(t, torch.Tensor) and t.data_ptr() == base + 4 * r.offset and t.is_contiguous()):
        ok = False
        break
    if ok:
      return flat
  out = torch.zeros(layout.size, dtype=torch.float32, device=device)
  for r in layout.refs:
    v = params[r.module][r.name]
    v = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
    out[r.offset : r.offset + r.size].copy_(v.reshape(-1).to(device=device, dtype=torch.float32))
  return out


def tree_from_flat(flat: torch.Tensor, layout: ParamLayout) -> FlatParams:
  tree = FlatParams()
  for r in layout.refs:
    tree.setdefault(r.module, {})[r.name] = flat[r.offset : r.offset + r.size].view(r.shape)
  tree.flat = flat
  tree.layout = layout
  return tree


class Transformed:
  def __init__(self, fn, apply_rng=True):
    self._fn = fn
    self._apply_rng = apply_rng

  def _run(self, params, rng, is_init, args, kwargs):
    if getattr(_tls, "frame", None) is not None:
      raise RuntimeError("nested transforms are not supported")
    fr = _Frame(params, _as_generator(rng), is_init)
    _tls.frame = fr
    try:
      out = self._fn(*args, **kwargs)
    finally:
      _tls.frame = None
    return out, fr

  def init(self, rng, *args, **kwargs) -> FlatParams:
    raw = {}
    _, fr = self._run(raw, rng, True, args, kwargs)
    device = fr.device if fr.device is not None else torch.device("cpu")
    host = np.zeros(fr.layout.size, dtype=np.float32)
    for r in fr.layout.refs:
      host[r.offset : r.offset + r.size] = raw[r.module][r.name].reshape(-1)
    flat = torch.from_numpy(host).to(device)
    return tree_from_flat(flat, fr.layout)

  def trace(self, *args, **kwargs):
    """ChainSpecs of the flows this function evaluates (no arithmetic, no device needed) plus the
    parameter layout: what `jax.jit` would get out of tracing it."""
    _, fr = self._run({}, 0, True, args, kwargs)
    return fr.traced, fr.layout

  def apply(self, params, *args, **kwargs):
    if self._apply_rng:
      if not args:
        raise TypeError("apply(params, rng, ...) needs an rng argument (None is fine for log_prob)")
      rng, args = args[0], args[1:]
    else:
      rng = None
    out, _ = self._run(params, rng, False, args, kwargs)
    return out


def transform(fn):
  return Transformed(fn, apply_rng=True)


def without_apply_rng(t: Transformed):
  return Transformed(t._fn, apply_rng=False)


# ----------------------------------------------------------------- initializers
class initializers:  # noqa: N801  (mirrors hk.initializers)
  class TruncatedNormal:
    """N(mean, 1) truncated to +-2 sigma, times stddev."""

    def __init__(self, stddev=1.0, mean=0.0):
      self.stddev = stddev
      self.mean = mean

    def __call__(self, shape, dtype=None):
      rng = current_frame().rng
      if rng is None:
        raise ValueError("init needs an rng to draw TruncatedNormal parameters")
      n = int(np.prod(shape)) if len(shape) else 1
      out = np.empty(n, dtype=np.float64)
      filled = 0
      while filled < n:
        draw = rng.standard_normal(2 * (n - filled) + 16)
        draw = draw[np.abs(draw) <= 2.0][: n - filled]
        out[filled : filled + draw.size] = draw
        filled += draw.size
      return (out * self.stddev + self.mean).reshape(shape)

  class Constant:
    def __init__(self, constant):
      self.constant = constant

    def __call__(self, shape, dtype=None):
      return np.broadcast_to(np.asarray(self.constant, dtype=np.float64), shape).copy()

  class RandomNormal:
    def __init__(self, stddev=1.0, mean=0.0):
      self.stddev, self.mean = stddev, mean

    def __call__(self, shape, dtype=None):
      return current_frame().rng.standard_normal(shape) * self.stddev + self.mean


def zeros(shape, dtype=None):
  return np.zeros(shape)


def ones(shape, dtype=None):
  return np.ones(shape)


# ---------------------------------------------------------------------- modules
class Linear(Module):
  """hk.Linear: x @ w + b, w:[in, out]; default w_init TruncatedNormal(1/sqrt(fan_in))."""

  def __init__(self, output_size, with_bias=True, w_init=None, b_init=None, name=None):
    super().__init__(name=name)
    self.output_size = int(output_size)
    self.with_bias = with_bias
    self.w_init = w_init
    self.b_init = b_init or zeros

  @scoped
  def __call__(self, inputs):
    if isinstance(inputs, _sym.InputSym):
      source, layers, in_size = inputs, [], inputs.width
    elif isinstance(inputs, _sym.NetSym) and len(inputs.shape) == 2:
      source, layers, in_size = inputs.source, inputs.layers, inputs.width
    else:
      raise _sym.Unsupported(
        f"hk.Linear applied to {type(inputs).__name__}: conditioners must be stacks of Linear layers on the "
        "layer input (no eager array path exists)"
      )
    w_init = self.w_init
    if w_init is None:
      w_init = initializers.TruncatedNormal(stddev=1.0 / math.sqrt(in_size))
    w = get_parameter("w", [in_size, self.output_size], init=w_init)
    b = get_parameter("b", [self.output_size], init=self.b_init) if self.with_bias else None
    rec = _sym.LinearRec(w, b, in_size, self.output_size)
    return _sym.NetSym(source, layers + [rec], (None, self.output_size))


class nets:  # noqa: N801
  class MLP(Module):
    """hk.nets.MLP: Linear_0, act, ..., Linear_{n-1}; activation after the last iff activate_final."""

    def __init__(self, output_sizes, w_init=None, b_init=None, with_bias=True, activation=_sym.relu,
                 activate_final=False, name=None):
      super().__init__(name=name)
      self.activation = activation
      self.activate_final = activate_final
      self.layers = [
        Linear(int(n), with_bias=with_bias, w_init=w_init, b_init=b_init, name=f"linear_{i}")
        for i, n in enumerate(output_sizes)
      ]

    @scoped
    def __call__(self, inputs):
      out = inputs
      n = len(self.layers)
      act = _sym.activation_name(self.activation)
      for i, layer in enumerate(self.layers):
        out = layer(out)
        if i < n - 1 or self.activate_final:
          out = out.with_activation(act)
      return out


class Sequential(Module):
  def __init__(self, layers, name=None):
    super().__init__(name=name)
    self.layers = tuple(layers)

  @scoped
  def __call__(self, inputs, *args, **kwargs):
    out = inputs
    for i, layer in enumerate(self.layers):
      out = layer(out, *args, **kwargs) if i == 0 else layer(out)
    return out


class Reshape(Module):
  """hk.Reshape(output_shape, preserve_dims=1)."""

  def __init__(self, output_shape, preserve_dims=1, name=None):
    super().__init__(name=name)
    self.output_shape = tuple(int(s) for s in output_shape)
    self.preserve_dims = preserve_dims

  @scoped
  def __call__(self, inputs):
    if not isinstance(inputs, _sym.NetSym):
      raise _sym.Unsupported(f"hk.Reshape applied to {type(inputs).__name__}")
    nd = len(inputs.shape)
    keep = self.preserve_dims if self.preserve_dims > 0 else nd + self.preserve_dims
    if keep != 1:
      raise _sym.Unsupported("hk.Reshape must preserve exactly the batch dimension")
    return inputs.reshape((None,) + self.output_shape)

"""Minimal restatement of the dm-haiku 0.0.16 features the reference uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  dm-haiku is an un-vendored
dependency of the reference (pyproject.toml:21-28, uv.lock:850-852); this file
restates only what the hot path touches:

  * hk.transform / hk.without_apply_rng        (examples/coupling_inference_surjection.py:69-71)
  * hk.Module naming: snake_case class name, "_<n>" uniquification per scope,
    "parent/~/child" for children built in __init__, "parent/child" for children
    built in __call__ (pinned by masked_autoregressive_test.py:105-107:
    'made/~/masked_linear_1')
  * hk.get_parameter                            (lu_linear.py:58-66, masked_linear.py:47,52)
  * hk.Linear, hk.nets.MLP, hk.Sequential, hk.Reshape
  * hk.initializers.TruncatedNormal, hk.PRNGSequence, hk.next_rng_key

Arrays are torch CPU tensors; dtype is chosen by `transform(..., dtype=...)`.
Random streams are NumPy PCG64 (JAX PRNG streams cannot be reproduced here).
"""

from __future__ import annotations

import math
import re
import threading

import numpy as np
import torch

_tls = threading.local()


class _Frame:
  def __init__(self, params, rng, is_init, dtype):
    self.params = params
    self.rng = rng
    self.is_init = is_init
    self.dtype = dtype
    self.counters = {}
    self.module_stack = []  # (module, method_name)


def _frame() -> _Frame:
  fr = getattr(_tls, "frame", None)
  if fr is None:
    raise RuntimeError("must be called inside hk.transform")
  return fr


def _snake(name):
  s1 = re.sub("(.)([A-Z][a-z]+)", r"\1_\2", name)
  return re.sub("([a-z0-9])([A-Z])", r"\1_\2", s1).lower()


class _ModuleMeta(type):
  def __call__(cls, *args, **kwargs):
    obj = cls.__new__(cls)
    fr = _frame()
    fr.module_stack.append((obj, "__init__"))
    try:
      obj.__init__(*args, **kwargs)
    finally:
      fr.module_stack.pop()
    if not hasattr(obj, "module_name"):
      raise RuntimeError("hk.Module subclasses must call super().__init__()")
    return obj


class Module(metaclass=_ModuleMeta):
  def __init__(self, name=None):
    fr = _frame()
    base = name if name is not None else _snake(type(self).__name__)
    # the enclosing scope is the module below us on the stack (we are on top)
    prefix = ""
    if len(fr.module_stack) >= 2:
      parent, method = fr.module_stack[-2]
      sep = "/~/" if method == "__init__" else "/"
      prefix = parent.module_name + sep
    key = prefix + base
    n = fr.counters.get(key, 0)
    fr.counters[key] = n + 1
    self.module_name = key if n == 0 else f"{key}_{n}"

  def _enter(self, method="__call__"):
    _frame().module_stack.append((self, method))

  def _exit(self):
    _frame().module_stack.pop()


def _scoped(fn):
  """Run a module method with the module on the scope stack."""

  def wrapper(self, *a, **k):
    self._enter(fn.__name__)
    try:
      return fn(self, *a, **k)
    finally:
      self._exit()

  wrapper.__name__ = fn.__name__
  return wrapper


def get_parameter(name, shape, dtype=None, init=None):
  fr = _frame()
  if not fr.module_stack:
    raise RuntimeError("get_parameter outside of a module")
  mod = fr.module_stack[-1][0].module_name
  shape = tuple(int(s) for s in shape)
  if fr.is_init:
    bucket = fr.params.setdefault(mod, {})
    if name not in bucket:
      val = init(shape, fr.dtype)
      val = torch.as_tensor(np.asarray(val), dtype=fr.dtype).reshape(shape).clone()
      bucket[name] = val
    return bucket[name]
  val = fr.params[mod][name]
  if tuple(val.shape) != shape:
    raise ValueError(f"{mod}/{name}: expected shape {shape}, got {tuple(val.shape)}")
  return val


def next_rng_key():
  return _frame().rng


class PRNGSequence:
  def __init__(self, seed):
    self._ss = np.random.SeedSequence(int(seed))

  def __next__(self):
    (child,) = self._ss.spawn(1)
    return np.random.Generator(np.random.PCG64(child))

  next = __next__


def _as_generator(rng):
  if rng is None or isinstance(rng, np.random.Generator):
    return rng
  return np.random.Generator(np.random.PCG64(int(rng)))


class Transformed:
  def __init__(self, fn, dtype, apply_rng=True):
    self._fn = fn
    self._dtype = dtype
    self._apply_rng = apply_rng

  def _run(self, params, rng, is_init, args, kwargs):
    if getattr(_tls, "frame", None) is not None:
      raise RuntimeError("nested transforms are not supported")
    _tls.frame = _Frame(params, _as_generator(rng), is_init, self._dtype)
    try:
      out = self._fn(*args, **kwargs)
    finally:
      _tls.frame = None
    return out

  def init(self, rng, *args, **kwargs):
    params = {}
    self._run(params, rng, True, args, kwargs)
    return params

  def apply(self, params, *args, **kwargs):
    if self._apply_rng:
      rng, args = args[0], args[1:]
    else:
      rng = None
    return self._run(params, rng, False, args, kwargs)


def transform(fn, dtype=torch.float64):
  return Transformed(fn, dtype, apply_rng=True)


def without_apply_rng(t: Transformed):
  return Transformed(t._fn, t._dtype, apply_rng=False)


# ----------------------------------------------------------------- initializers
class initializers:  # noqa: N801  (mirrors hk.initializers)
  class TruncatedNormal:
    """N(mean, 1) truncated to +-2 sigma, times stddev (haiku semantics)."""

    def __init__(self, stddev=1.0, mean=0.0):
      self.stddev = stddev
      self.mean = mean

    def __call__(self, shape, dtype):
      rng = _frame().rng
      n = int(np.prod(shape)) if len(shape) else 1
      out = np.empty(n, dtype=np.float64)
      filled = 0
      while filled < n:
        draw = rng.standard_normal(2 * (n - filled) + 16)
        draw = draw[np.abs(draw) <= 2.0][: n - filled]
        out[filled : filled + draw.size] = draw
        filled += draw.size
      return (out * self.stddev + self.mean).reshape(shape)


def zeros(shape, dtype=None):
  return np.zeros(shape)


def ones(shape, dtype=None):
  return np.ones(shape)


# ---------------------------------------------------------------------- modules
class Linear(Module):
  """hk.Linear: x @ w + b, w:[in,out], default init TruncNormal(1/sqrt(fan_in))."""

  def __init__(self, output_size, with_bias=True, w_init=None, b_init=None, name=None):
    super().__init__(name=name)
    self.output_size = output_size
    self.with_bias = with_bias
    self.w_init = w_init
    self.b_init = b_init or zeros

  @_scoped
  def __call__(self, inputs):
    in_size = inputs.shape[-1]
    w_init = self.w_init
    if w_init is None:
      w_init = initializers.TruncatedNormal(stddev=1.0 / math.sqrt(in_size))
    w = get_parameter("w", [in_size, self.output_size], init=w_init)
    out = inputs @ w
    if self.with_bias:
      b = get_parameter("b", [self.output_size], init=self.b_init)
      out = out + b
    return out


relu_probe = None  # tests set this to a list to collect, per sample, min |pre-activation| of every ReLU call


def relu(x):
  if relu_probe is not None and x.dim() == 2:
    relu_probe.append(x.detach().abs().min(dim=-1).values)
  return torch.relu(x)


def swish(x):
  return x * torch.sigmoid(x)


class nets:  # noqa: N801
  class MLP(Module):
    """hk.nets.MLP: Linear_0, act, ..., Linear_{n-1} (act after last iff activate_final)."""

    def __init__(
      self,
      output_sizes,
      w_init=None,
      b_init=None,
      with_bias=True,
      activation=relu,
      activate_final=False,
      name=None,
    ):
      super().__init__(name=name)
      self.activation = activation
      self.activate_final = activate_final
      self.layers = [
        Linear(int(n), with_bias=with_bias, w_init=w_init, b_init=b_init, name=f"linear_{i}")
        for i, n in enumerate(output_sizes)
      ]

    @_scoped
    def __call__(self, inputs):
      out = inputs
      n = len(self.layers)
      for i, layer in enumerate(self.layers):
        out = layer(out)
        if i < n - 1 or self.activate_final:
          out = self.activation(out)
      return out


class Sequential(Module):
  def __init__(self, layers, name=None):
    super().__init__(name=name)
    self.layers = tuple(layers)

  @_scoped
  def __call__(self, inputs, *args, **kwargs):
    out = inputs
    for i, layer in enumerate(self.layers):
      out = layer(out, *args, **kwargs) if i == 0 else layer(out)
    return out


class Reshape(Module):
  """hk.Reshape(output_shape, preserve_dims=1): keep leading dims, reshape the rest."""

  def __init__(self, output_shape, preserve_dims=1, name=None):
    super().__init__(name=name)
    self.output_shape = tuple(output_shape)
    self.preserve_dims = preserve_dims

  @_scoped
  def __call__(self, inputs):
    nd = inputs.dim()
    keep = self.preserve_dims if self.preserve_dims > 0 else nd + self.preserve_dims
    return inputs.reshape(tuple(inputs.shape[:keep]) + self.output_shape)

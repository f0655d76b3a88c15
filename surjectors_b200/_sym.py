"""Symbolic stand-ins for the arrays that flow through user-supplied callables.

The reference lets `conditioner`, `bijector_fn` and `decoder` be arbitrary JAX callables
(masked_coupling.py:47-57, slice.py:54-56) and traces them under `jax.jit`.  The fused CUDA
kernels can only serve the structures they implement, so the host layer calls those same
callables with the symbols below, pattern-matches what comes back into plain-old-data
descriptors and raises `NotImplementedError` on anything else -- there is no fallback path.
"""

from __future__ import annotations

import numpy as np


class Unsupported(NotImplementedError):
  pass


def _unsupported(op):
  def fn(self, *a, **k):
    raise Unsupported(
      f"operation '{op}' on a symbolic {type(self).__name__} is not something the B200 kernels implement; "
      "supported structure: MLP / MADE conditioners, ScalarAffine(split or unstack, optionally Inverse), "
      "RationalQuadraticSpline, Independent(Normal(mu, exp(log_scale))) decoders"
    )

  return fn


class Sym:
  """Base symbolic array; shape[0] is the (unknown) batch."""

  shape: tuple

  @property
  def ndim(self):
    return len(self.shape)

  for _op in ("add", "radd", "sub", "rsub", "mul", "rmul", "truediv", "rtruediv", "neg", "pow", "matmul", "getitem"):
    locals()[f"__{_op}__"] = _unsupported(_op)
  del _op


class InputSym(Sym):
  """Input handed to a conditioner / decoder: [B, width]."""

  def __init__(self, width):
    self.width = int(width)
    self.shape = (None, self.width)


class LinearRec:
  __slots__ = ("w", "b", "n_in", "n_out", "act")

  def __init__(self, w, b, n_in, n_out, act="identity"):
    self.w, self.b, self.n_in, self.n_out, self.act = w, b, n_in, n_out, act


class NetSym(Sym):
  """Output of a stack of Linear layers (hk.nets.MLP / hk.Linear / hk.Sequential)."""

  def __init__(self, source, layers, shape):
    self.source = source  # InputSym
    self.layers = list(layers)
    self.shape = tuple(shape)

  @property
  def width(self):
    return self.layers[-1].n_out

  def with_activation(self, name):
    if name == "identity":
      return self
    if self.layers[-1].act != "identity":
      raise Unsupported("two activations in a row on a conditioner layer")
    if len(self.shape) != 2:
      raise Unsupported("activation after a reshape of the conditioner output")
    last = self.layers[-1]
    new_last = LinearRec(last.w, last.b, last.n_in, last.n_out, name)
    return NetSym(self.source, self.layers[:-1] + [new_last], self.shape)

  def reshape(self, shape):
    shape = tuple(shape)
    if int(np.prod(shape[1:])) != self.width:
      raise ValueError(f"cannot reshape conditioner output of width {self.width} to {shape[1:]}")
    return NetSym(self.source, self.layers, (None,) + shape[1:])


class MadeSym(Sym):
  """Output of a MADE: [B, input_size, n_params] (made.py:90)."""

  def __init__(self, made, y, ctx):
    self.made = made
    self.y = y
    self.ctx = ctx
    self.shape = (None, made.input_size, made.n_params)


class SplitSym(Sym):
  """jnp.split(parent, n, axis=-1)[index]"""

  def __init__(self, parent, index, n):
    if parent.shape[-1] % n:
      raise ValueError("array split does not result in an equal division")
    self.parent, self.index, self.n = parent, index, n
    self.shape = parent.shape[:-1] + (parent.shape[-1] // n,)


class UnstackSym(Sym):
  """unstack(parent, axis=-1)[index] (util.py:106-120)"""

  def __init__(self, parent, index):
    self.parent, self.index = parent, index
    self.n = parent.shape[-1]
    self.shape = parent.shape[:-1]


class ExpSym(Sym):
  def __init__(self, parent):
    self.parent = parent
    self.shape = parent.shape


class Activation:
  """Named activation: records itself on symbolic conditioner outputs, evaluates on numbers."""

  def __init__(self, name, fn):
    self.name = name
    self._fn = fn

  def __call__(self, x):
    if isinstance(x, NetSym):
      return x.with_activation(self.name)
    if isinstance(x, Sym):
      raise Unsupported(f"activation {self.name} applied to {type(x).__name__}")
    return self._fn(np.asarray(x))

  def __repr__(self):
    return f"<activation {self.name}>"


def _sigmoid(x):
  return 1.0 / (1.0 + np.exp(-x))


relu = Activation("relu", lambda x: np.maximum(x, 0))
swish = silu = Activation("swish", lambda x: x * _sigmoid(x))
tanh = Activation("tanh", np.tanh)
sigmoid = Activation("sigmoid", _sigmoid)
softplus = Activation("softplus", lambda x: np.logaddexp(x, 0))
gelu = Activation("gelu", lambda x: 0.5 * x * (1 + np.tanh(0.7978845608028654 * (x + 0.044715 * x**3))))
elu = Activation("elu", lambda x: np.where(x > 0, x, np.expm1(x)))
leaky_relu = Activation("leaky_relu", lambda x: np.where(x > 0, x, 0.01 * x))
identity = Activation("identity", lambda x: x)


def activation_name(fn):
  """Name of a user-supplied activation callable, probing plain python identities."""
  if isinstance(fn, Activation):
    return fn.name
  probe = NetSym(InputSym(1), [LinearRec(None, None, 1, 1)], (None, 1))
  try:
    out = fn(probe)
  except Unsupported:
    raise
  except Exception as e:  # noqa: BLE001
    raise Unsupported(f"activation {fn!r} is not one of the named activations in surjectors_b200.nn") from e
  if out is probe:
    return "identity"
  if isinstance(out, NetSym):
    return out.layers[-1].act
  raise Unsupported(f"activation {fn!r} is not one of the named activations in surjectors_b200.nn")


# ----------------------------------------------------------------- array helpers
def split(x, n_or_idx, axis=-1):
  """jnp.split on symbols: only `split(theta, 2, -1)` (the reference's SPLIT parameter layout,
  affine_masked_coupling.py:47) is a structure the kernels implement."""
  if isinstance(x, Sym):
    if axis not in (-1, x.ndim - 1) or not isinstance(n_or_idx, int) or n_or_idx != 2:
      raise Unsupported("only jnp.split(params, 2, axis=-1) is supported on conditioner outputs")
    return [SplitSym(x, i, 2) for i in range(2)]
  return np.split(np.asarray(x), n_or_idx, axis)


def unstack(x, axis=0):
  """surjectors.util.unstack (util.py:106-120) on symbols: INTERLEAVED layout theta[..., D, 2]."""
  if isinstance(x, Sym):
    if axis not in (-1, x.ndim - 1):
      raise Unsupported("only unstack(params, axis=-1) is supported on conditioner outputs")
    return [UnstackSym(x, i) for i in range(x.shape[-1])]
  x = np.asarray(x)
  return [np.take(x, i, axis) for i in range(x.shape[axis])]


def exp(x):
  if isinstance(x, Sym):
    return ExpSym(x)
  return np.exp(x)


# ------------------------------------------------------- inner bijectors (distrax)
class ScalarAffine:
  """distrax.ScalarAffine(shift, scale=None, log_scale=None) as a descriptor."""

  def __init__(self, shift, scale=None, log_scale=None):
    if (scale is None) == (log_scale is None):
      if scale is None:
        raise Unsupported("ScalarAffine without scale / log_scale is the identity shift; not on the hot path")
      raise ValueError("Only one of `scale` and `log_scale` can be specified, not both.")
    self.shift = shift
    self.scale = scale
    self.log_scale = log_scale


class Inverse:
  """distrax.Inverse(bijector)."""

  def __init__(self, bijector):
    self.bijector = bijector


class RationalQuadraticSpline:
  """distrax.RationalQuadraticSpline(params, range_min, range_max) with the defaults the
  reference uses (boundary_slopes='unconstrained', min_bin_size=1e-4, min_knot_slope=1e-4;
  rq_masked_coupling.py:59-61)."""

  def __init__(self, params, range_min, range_max, boundary_slopes="unconstrained", min_bin_size=1e-4,
               min_knot_slope=1e-4):
    if boundary_slopes != "unconstrained" or min_bin_size != 1e-4 or min_knot_slope != 1e-4:
      raise Unsupported("RationalQuadraticSpline: only the distrax defaults are implemented")
    if not isinstance(params, Sym):
      raise Unsupported("RationalQuadraticSpline params must be a conditioner output")
    if params.shape[-1] % 3 != 1 or params.shape[-1] < 4:
      raise ValueError("The last dimension of `params` must have size `3 * num_bins + 1` and `num_bins` must be at least 1.")
    if not range_min < range_max:
      raise ValueError("`range_min` must be less than `range_max`.")
    self.params = params
    self.range_min = float(range_min)
    self.range_max = float(range_max)
    self.num_bins = (params.shape[-1] - 1) // 3


# ------------------------------------------------------------------ distributions
class Normal:
  """distrax.Normal / tfd.Normal(loc, scale) as a descriptor."""

  def __init__(self, loc, scale):
    self.loc = loc
    self.scale = scale


class Independent:
  """distrax.Independent / tfd.Independent(distribution, reinterpreted_batch_ndims)."""

  def __init__(self, distribution, reinterpreted_batch_ndims=None):
    self.distribution = distribution
    self.reinterpreted_batch_ndims = reinterpreted_batch_ndims


class MultivariateNormalDiag:
  def __init__(self, loc, scale_diag):
    self.loc = loc
    self.scale_diag = scale_diag


# -------------------------------------------------------------- pattern matching
def match_affine(bij, theta, n_dims):
  """(inner, layout) for a bijector_fn result built on `theta`; raises Unsupported otherwise.

  inner: 0 ScalarAffine, 1 Inverse(ScalarAffine); layout: 0 SPLIT, 1 INTERLEAVED."""
  inverted = False
  if isinstance(bij, Inverse):
    inverted = True
    bij = bij.bijector
  if not isinstance(bij, ScalarAffine):
    return None
  ls = bij.log_scale
  if ls is None:
    if not isinstance(bij.scale, ExpSym):
      raise Unsupported("ScalarAffine scale must be exp(<conditioner output>) (or pass log_scale=)")
    ls = bij.scale.parent
  sh = bij.shift
  if isinstance(sh, SplitSym) and isinstance(ls, SplitSym):
    if sh.parent is not theta or ls.parent is not theta or (sh.index, ls.index) != (0, 1) or sh.n != 2:
      raise Unsupported("ScalarAffine(shift, scale) must use (first, second) half of jnp.split(params, 2, -1)")
    if tuple(theta.shape) != (None, 2 * n_dims):
      raise Unsupported(f"split layout needs conditioner output [B, {2 * n_dims}], got {theta.shape}")
    return (1 if inverted else 0), 0
  if isinstance(sh, UnstackSym) and isinstance(ls, UnstackSym):
    if sh.parent is not theta or ls.parent is not theta or (sh.index, ls.index) != (0, 1):
      raise Unsupported("ScalarAffine(shift, scale) must use unstack(params, -1)[0], [1]")
    if tuple(theta.shape) != (None, n_dims, 2):
      raise Unsupported(f"unstack layout needs conditioner output [B, {n_dims}, 2], got {theta.shape}")
    return (1 if inverted else 0), 1
  raise Unsupported("ScalarAffine arguments must come from jnp.split(params, 2, -1) or unstack(params, -1)")


def match_rqs(bij, theta, n_dims):
  """(inner, n_bins, range_min, range_max) or None."""
  inverted = False
  if isinstance(bij, Inverse):
    inverted = True
    bij = bij.bijector
  if not isinstance(bij, RationalQuadraticSpline):
    return None
  if bij.params is not theta:
    raise Unsupported("RationalQuadraticSpline params must be the conditioner output itself")
  if tuple(theta.shape) != (None, n_dims, 3 * bij.num_bins + 1):
    raise Unsupported(f"spline needs conditioner output [B, {n_dims}, 3K+1], got {theta.shape}")
  return (3 if inverted else 2), bij.num_bins, bij.range_min, bij.range_max


def match_normal_decoder(dist, source, n_out):
  """Check `dist` is Independent(Normal(mu, exp(log_sigma))) with (mu, log_sigma) =
  split(net(source), 2, -1) (examples/coupling_inference_surjection.py:24-33) and return net."""
  if isinstance(dist, Independent):
    n = dist.reinterpreted_batch_ndims
    if n not in (None, 1):
      raise Unsupported("decoder: Independent must reinterpret exactly the event axis")
    dist = dist.distribution
    loc, scale = getattr(dist, "loc", None), getattr(dist, "scale", None)
    if not isinstance(dist, Normal):
      raise Unsupported("decoder must be Independent(Normal(mu, exp(log_sigma)))")
  elif isinstance(dist, MultivariateNormalDiag):
    loc, scale = dist.loc, dist.scale_diag
  else:
    raise Unsupported("decoder must return Independent(Normal(mu, exp(log_sigma)))")
  if not isinstance(scale, ExpSym):
    raise Unsupported("decoder scale must be exp(<second half of the decoder network output>)")
  ls = scale.parent
  if not (isinstance(loc, SplitSym) and isinstance(ls, SplitSym) and loc.parent is ls.parent
          and (loc.index, ls.index) == (0, 1) and loc.n == 2):
    raise Unsupported("decoder (mu, log_sigma) must be jnp.split(net(z), 2, -1)")
  net = loc.parent
  if not isinstance(net, NetSym) or net.source is not source or tuple(net.shape) != (None, 2 * n_out):
    raise Unsupported(f"decoder network must map the layer's [z, x] to [B, {2 * n_out}]")
  return net


def check_standard_normal_base(dist, n_dims):
  """The kernels fuse base = Independent(Normal(0, 1), 1) (every example and test of the reference,
  e.g. examples/coupling_inference_surjection.py:62-65)."""
  if isinstance(dist, Independent):
    inner = dist.distribution
    if dist.reinterpreted_batch_ndims not in (None, 1) or not isinstance(inner, Normal):
      raise Unsupported("base distribution must be Independent(Normal(zeros(n), ones(n)), 1)")
    loc, scale = inner.loc, inner.scale
  elif isinstance(dist, MultivariateNormalDiag):
    loc, scale = dist.loc, dist.scale_diag
  else:
    raise Unsupported("base distribution must be Independent(Normal(zeros(n), ones(n)), 1)")
  loc = np.asarray(loc, dtype=np.float64)
  scale = np.asarray(scale, dtype=np.float64)
  shape = np.broadcast_shapes(loc.shape, scale.shape)
  if shape != (n_dims,):
    raise ValueError(f"base distribution has event shape {shape} but the flow produces {n_dims} dims")
  if np.any(loc != 0.0) or np.any(scale != 1.0):
    raise Unsupported("only the standard-normal base distribution is fused into the kernels")

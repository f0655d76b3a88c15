"""Restatement of surjectors v0.3.4's log_prob path on torch-CPU tensors.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations are relative to
/root/reference/.  The structure (generic python callables for `conditioner`,
`bijector_fn`, `decoder`; dense `w * mask` matmuls; conditioner outputs for all
dims; materialised [B, D, 3K+1] spline parameters) deliberately follows the
reference op-for-op, because this file doubles as the CPU baseline.
"""

from __future__ import annotations

import numpy as np
import torch

from oracle import distrax, hk, tfp
from oracle.distrax import Inverse, RationalQuadraticSpline, ScalarAffine  # noqa: F401  (re-exported like surjectors/__init__.py:5)


# ----------------------------------------------------------------------- util.py
def make_alternating_binary_mask(n_dim, even_idx_as_true=False):
  """surjectors/util.py:30-47: mask[i] = (i % 2 == 1), negated if even_idx_as_true."""
  mask = (np.arange(0, int(np.prod(n_dim))) % 2).reshape(n_dim).astype(bool)
  if even_idx_as_true:
    mask = np.logical_not(mask)
  return mask


def unstack(x, axis=0):
  """surjectors/util.py:106-120: list of slices along `axis` (tf.unstack)."""
  return list(torch.unbind(x, dim=axis))


def split(x, n_or_idx, axis=-1):
  """jnp.split for the two call patterns the reference uses: (x, 2, -1) and (x, [k], -1)."""
  if isinstance(n_or_idx, int):
    return list(torch.chunk(x, n_or_idx, dim=axis))
  (k,) = n_or_idx
  return [x.narrow(axis, 0, k), x.narrow(axis, k, x.shape[axis] - k)]


# --------------------------------------------------------------- conditioners/*
def make_mlp(dims, activation=hk.swish, w_init=None, b_init=hk.zeros):
  """_src/conditioners/mlp.py:7-29: hk.nets.MLP, swish, TruncNormal(0.01), zero bias."""
  if w_init is None:
    w_init = hk.initializers.TruncatedNormal(stddev=0.01)
  return hk.nets.MLP(output_sizes=dims, w_init=w_init, b_init=b_init, activation=activation)


class MaskedLinear(hk.Module):
  """_src/conditioners/nn/masked_linear.py:35-58: x @ (w * mask) + b."""

  def __init__(self, mask, with_bias=True, w_init=None, b_init=None):
    super().__init__()
    self.mask = mask
    self.with_bias = with_bias
    self.w_init = w_init
    self.b_init = b_init or hk.zeros

  @hk._scoped
  def __call__(self, inputs):
    w_init = self.w_init
    if w_init is None:
      w_init = hk.initializers.TruncatedNormal(stddev=1.0 / np.sqrt(self.mask.shape[0]))
    w = hk.get_parameter("w", self.mask.shape, init=w_init)
    mask = torch.as_tensor(self.mask, dtype=w.dtype)
    out = inputs @ (w * mask)
    if self.with_bias:
      b = hk.get_parameter("b", (self.mask.shape[-1],), init=self.b_init)
      out = out + b
    return out


class MADE(hk.Module):
  """_src/conditioners/nn/made.py:26-91."""

  def __init__(self, input_size, hidden_layer_sizes, n_params, w_init=None, b_init=None, activation=hk.relu):
    super().__init__()
    self.input_size = input_size
    self.output_sizes = tuple(hidden_layer_sizes)
    self.n_params = n_params
    self.w_init = w_init
    self.b_init = b_init
    self.activation = activation
    masks = tfp.make_dense_autoregressive_masks(n_params, input_size, self.output_sizes)  # made.py:54-56
    self.layers = tuple(MaskedLinear(mask=m, w_init=w_init, b_init=b_init) for m in masks)

  @hk._scoped
  def __call__(self, y, x=None):
    out = self.layers[0](y)
    if x is not None:  # made.py:80-84: context Linear is created inside __call__
      out = out + hk.Linear(self.output_sizes[0], w_init=self.w_init, b_init=self.b_init)(x)
    out = self.activation(out)
    rest = self.layers[1:]
    for i, layer in enumerate(rest):
      out = layer(out)
      if i < len(rest) - 1:
        out = self.activation(out)
    return out.reshape(tuple(out.shape[:-1]) + (self.input_size, self.n_params))  # hk.Reshape, made.py:90


# ------------------------------------------------------------ protocol (a2, a3)
class Surjector:
  """_src/surjectors/surjector.py:9-99."""

  def __call__(self, method, **kwargs):
    return getattr(self, method)(**kwargs)

  def inverse_and_likelihood_contribution(self, y, x=None, **kwargs):
    return self._inverse_and_likelihood_contribution(y, x=x, **kwargs)

  def forward_and_likelihood_contribution(self, z, x=None, **kwargs):
    return self._forward_and_likelihood_contribution(z, x=x, **kwargs)

  def forward(self, z, x=None, **kwargs):
    return self.forward_and_likelihood_contribution(z, x=x, **kwargs)[0]

  def inverse(self, y, x=None, **kwargs):
    return self.inverse_and_likelihood_contribution(y, x=x, **kwargs)[0]


class Bijector(Surjector):
  """_src/bijectors/bijector.py:9-40."""

  def inverse_and_log_det(self, y, x=None, **kwargs):
    return self.inverse_and_likelihood_contribution(y, x=x, **kwargs)

  def forward_and_log_det(self, z, x=None, **kwargs):
    return self.forward_and_likelihood_contribution(z, x=x, **kwargs)


class Chain(Surjector):
  """_src/surjectors/chain.py:22-67."""

  def __init__(self, transforms):
    self._transforms = transforms

  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):
    z, lcs = self._transforms[0].inverse_and_likelihood_contribution(y, x)
    for t in self._transforms[1:]:
      z, lc = t.inverse_and_likelihood_contribution(z, x)
      lcs = lcs + lc
    return z, lcs

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):
    y, ld = self._transforms[-1].forward_and_likelihood_contribution(z, x)
    for t in reversed(self._transforms[:-1]):
      y, lc = t.forward_and_likelihood_contribution(y, x)
      ld = ld + lc
    return y, ld


class TransformedDistribution:
  """_src/distributions/transformed_distribution.py:12-139 (log_prob direction)."""

  def __init__(self, base_distribution, transform):
    self.base_distribution = base_distribution
    self.transform = transform

  def __call__(self, method, **kwargs):
    return getattr(self, method)(**kwargs)

  def log_prob(self, y, x=None):
    return self.inverse_and_log_prob(y, x)[1]

  def inverse_and_log_prob(self, y, x=None):
    if x is not None:  # :83-85
      if y.dim() != x.dim() or y.shape[0] != x.shape[0]:
        raise AssertionError("rank / batch mismatch between y and x")
    if isinstance(self.transform, MaskedCoupling):
      # quirk Q1 (:87-88): distrax.Bijector instances are called WITHOUT x
      z, lc = self.transform.inverse_and_log_det(y)
    else:
      z, lc = self.transform.inverse_and_likelihood_contribution(y, x=x)
    return z, self.base_distribution.log_prob(z) + lc


# --------------------------------------------------------------- bijectors (a4-a12)
class MaskedCoupling(Bijector):
  """_src/bijectors/masked_coupling.py:12-85 over distrax.MaskedCoupling.

  mask True = dimension passes through unchanged and is visible to the conditioner.
  """

  def __init__(self, mask, conditioner, bijector_fn, event_ndims=None, inner_event_ndims=0):
    mask = np.asarray(mask)
    if mask.dtype != bool:
      raise ValueError("mask must be boolean")
    self._mask = torch.as_tensor(mask)
    self._event_mask = self._mask.reshape(tuple(mask.shape) + (1,) * inner_event_ndims)
    self._conditioner = conditioner
    self._inner_bijector = bijector_fn
    self._event_ndims = mask.ndim + inner_event_ndims if event_ndims is None else event_ndims
    self._inner_event_ndims = inner_event_ndims

  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):  # :73-85
    masked_y = torch.where(self._event_mask, y, torch.zeros((), dtype=y.dtype))
    if x is not None:
      masked_y = torch.cat([masked_y, x], dim=-1)
    params = self._conditioner(masked_y)
    z0, log_d = self._inner_bijector(params).inverse_and_log_det(y)
    z = torch.where(self._event_mask, y, z0)
    ld = distrax.sum_last(
      torch.where(self._mask, torch.zeros((), dtype=y.dtype), log_d), self._event_ndims - self._inner_event_ndims
    )
    return z, ld

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):  # :59-71
    masked_z = torch.where(self._event_mask, z, torch.zeros((), dtype=z.dtype))
    if x is not None:
      masked_z = torch.cat([masked_z, x], dim=-1)
    params = self._conditioner(masked_z)
    y0, log_d = self._inner_bijector(params).forward_and_log_det(z)
    y = torch.where(self._event_mask, z, y0)
    ld = distrax.sum_last(
      torch.where(self._mask, torch.zeros((), dtype=z.dtype), log_d), self._event_ndims - self._inner_event_ndims
    )
    return y, ld


class AffineMaskedCoupling(MaskedCoupling):
  """_src/bijectors/affine_masked_coupling.py:39-52 (SPLIT layout)."""

  def __init__(self, mask, conditioner, event_ndims=None, inner_event_ndims=0):
    def _bijector_fn(params):
      means, log_scales = split(params, 2, -1)
      return ScalarAffine(means, torch.exp(log_scales))

    super().__init__(mask, conditioner, _bijector_fn, event_ndims, inner_event_ndims)


class RationalQuadraticSplineMaskedCoupling(MaskedCoupling):
  """_src/bijectors/rq_masked_coupling.py:33-65."""

  def __init__(self, mask, conditioner, range_min, range_max, event_ndims=None, inner_event_ndims=0):
    self.range_min, self.range_max = range_min, range_max

    def _bijector_fn(params):
      return RationalQuadraticSpline(params, self.range_min, self.range_max)

    super().__init__(mask, conditioner, _bijector_fn, event_ndims, inner_event_ndims)


class MaskedAutoregressive(Bijector):
  """_src/bijectors/masked_autoregressive.py:11-79."""

  def __init__(self, conditioner, bijector_fn, event_ndims=1, inner_event_ndims=0):
    if event_ndims is not None and event_ndims < inner_event_ndims:
      raise ValueError("event_ndims should be at least as large as inner_event_ndims")
    if not isinstance(conditioner, MADE):
      raise ValueError("conditioner should be a MADE when used MaskedAutoregressive flow")
    self._event_ndims = event_ndims
    self._inner_event_ndims = inner_event_ndims
    self.conditioner = conditioner
    self._inner_bijector = bijector_fn

  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):  # :73-79
    params = self.conditioner(y, x)
    z, ld = self._inner_bijector(params).inverse_and_log_det(y)
    return z, distrax.sum_last(ld, self._event_ndims - self._inner_event_ndims)

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):  # :63-71
    y = torch.zeros_like(z)
    for _ in range(z.shape[-1]):
      params = self.conditioner(y, x)
      y, ld = self._inner_bijector(params).forward_and_log_det(z)
    return y, distrax.sum_last(ld, self._event_ndims - self._inner_event_ndims)


class AffineMaskedAutoregressive(MaskedAutoregressive):
  """_src/bijectors/affine_masked_autoregressive.py:48-58 (INTERLEAVED layout, non-inverted)."""

  def __init__(self, conditioner, event_ndims=1, inner_event_ndims=0):
    def bijector_fn(params):
      means, log_scales = unstack(params, -1)
      return ScalarAffine(means, torch.exp(log_scales))

    super().__init__(conditioner, bijector_fn, event_ndims, inner_event_ndims)


def softplus(x):
  return torch.nn.functional.softplus(x)


class LULinear(Bijector, hk.Module):
  """_src/bijectors/lu_linear.py:11-93: z = (y @ U) @ L, lc = sum(log(softplus(d)+1e-4))."""

  def __init__(self, n_dimension, with_bias=False, dtype=None):
    hk.Module.__init__(self)
    if with_bias:
      raise NotImplementedError()
    self.n_dimension = n_dimension
    n_tri = ((n_dimension - 1) * n_dimension) // 2
    self._lower_indices = np.tril_indices(n_dimension, k=-1)
    self._upper_indices = np.triu_indices(n_dimension, k=1)
    self._lower_entries = hk.get_parameter("lower_entries", [n_tri], init=hk.zeros)
    self._upper_entries = hk.get_parameter("upper_entries", [n_tri], init=hk.zeros)
    self._diag = hk.get_parameter("diag_entries", [n_dimension], init=hk.ones)

  def _matrices(self):  # :68-78
    n = self.n_dimension
    dt = self._diag.dtype
    eye = torch.eye(n, dtype=dt)
    low = torch.zeros((n, n), dtype=dt).index_put(
      (torch.as_tensor(self._lower_indices[0]), torch.as_tensor(self._lower_indices[1])), self._lower_entries
    )
    up = torch.zeros((n, n), dtype=dt).index_put(
      (torch.as_tensor(self._upper_indices[0]), torch.as_tensor(self._upper_indices[1])), self._upper_entries
    )
    upper_diag = softplus(self._diag) + 1e-4
    return low + eye, up + torch.diag(upper_diag), upper_diag

  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):  # :86-90
    L, U, upper_diag = self._matrices()
    z = (y @ U) @ L
    lc = torch.log(upper_diag).sum() * torch.ones(z.shape[0], dtype=z.dtype)
    return z, lc

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):  # :92-93
    raise NotImplementedError()


class Permutation(Bijector):
  """_src/bijectors/permutation.py:7-37."""

  def __init__(self, permutation, event_ndims_in):
    self.permutation = np.asarray(permutation)
    self.event_ndims_in = event_ndims_in

  def inverse_index(self):
    size = self.permutation.size
    inv = np.zeros(size, dtype=np.int32)  # jnp.result_type(int) is int32 by default
    inv[self.permutation] = np.arange(size, dtype=np.int32)
    return inv

  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):  # :30-37
    return y[..., torch.as_tensor(self.inverse_index(), dtype=torch.long)], torch.zeros(y.shape[:-1], dtype=y.dtype)

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):  # :27-28
    return z[..., torch.as_tensor(self.permutation, dtype=torch.long)], torch.zeros(z.shape[:-1], dtype=z.dtype)


# --------------------------------------------------------------- surjectors (a13-a15)
noise_provider = None  # tests set this to fn(shape, dtype) -> standard-normal draws for decoder samples


def _decoder_sample_and_log_prob(dist, like):
  """distrax Independent(Normal).sample_and_log_prob: loc + scale * eps, and the log-prob of that sample."""
  inner = dist.distribution
  eps = noise_provider(tuple(inner.loc.shape), inner.loc.dtype)
  sample = inner.loc + inner.scale * eps
  return sample, dist.log_prob(sample)


class Slice(Surjector):
  """_src/surjectors/slice.py:8-78."""

  def __init__(self, n_keep, decoder):
    self.n_keep = n_keep
    self.decoder = decoder

  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):  # :58-68
    z, y_minus = split(y, [self.n_keep], -1)
    z_condition = z if x is None else torch.cat([z, x], dim=-1)
    return z, self.decoder(z_condition).log_prob(y_minus)

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):  # slice.py:70-78
    z_condition = z if x is None else torch.cat([z, x], dim=-1)
    y_minus, lc = _decoder_sample_and_log_prob(self.decoder(z_condition), z)
    return torch.cat([z, y_minus], dim=-1), lc


class MLPInferenceFunnel(Surjector, hk.Module):
  """_src/surjectors/mlp.py:10-59."""

  def __init__(self, n_keep, decoder):
    hk.Module.__init__(self)
    self._r = LULinear(n_keep, False)
    self._w_prime = hk.Linear(n_keep, True)
    self.decoder = decoder
    self.n_keep = n_keep

  @hk._scoped
  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):  # :47-56
    y_plus, y_minus = split(y, [self.n_keep], -1)
    z, jac_det = self._r.inverse_and_likelihood_contribution(y_plus)
    z = z + self._w_prime(y_minus)
    lp = self.decoder(z).log_prob(y_minus)
    return z, lp + jac_det

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):  # :58-59
    raise NotImplementedError()


class MaskedCouplingInferenceFunnel(Surjector):
  """_src/surjectors/masked_coupling_inference_funnel.py:9-106."""

  def __init__(self, n_keep, decoder, conditioner, bijector_fn):
    self.n_keep = n_keep
    self.decoder = decoder
    self.conditioner = conditioner
    self.bijector_fn = bijector_fn

  def _mask(self, array):  # :72-75
    return np.arange(array.shape[-1]) >= self.n_keep

  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):  # :80-91
    inner = MaskedCoupling(self._mask(y), self.conditioner, self.bijector_fn)
    faux, jac_det = inner.inverse_and_log_det(y, x=x)
    z = faux[:, : self.n_keep]
    z_condition = z if x is None else torch.cat([z, x], dim=-1)
    lc = self.decoder(z_condition).log_prob(y[:, self.n_keep :])
    return z, lc + jac_det

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):  # :93-106
    z_condition = z if x is None else torch.cat([z, x], dim=-1)
    y_minus, jac_det = _decoder_sample_and_log_prob(self.decoder(z_condition), z)
    z_tilde = torch.cat([z, y_minus], dim=-1)
    inner = MaskedCoupling(self._mask(z_tilde), self.conditioner, self.bijector_fn)
    y, lc = inner.forward_and_log_det(z_tilde, x)
    return y, lc + jac_det


class AffineMaskedCouplingInferenceFunnel(MaskedCouplingInferenceFunnel):
  """_src/surjectors/affine_masked_coupling_inference_funnel.py:11-55."""

  def __init__(self, n_keep, decoder, conditioner):
    def bijector_fn(params):
      shift, log_scale = split(params, 2, -1)
      return ScalarAffine(shift, torch.exp(log_scale))

    super().__init__(n_keep, decoder, conditioner, bijector_fn)


class RationalQuadraticSplineMaskedCouplingInferenceFunnel(MaskedCouplingInferenceFunnel):
  """_src/surjectors/rq_masked_coupling_inference_funnel.py:11-86."""

  def __init__(self, n_keep, decoder, conditioner, range_min, range_max):
    self.range_min, self.range_max = range_min, range_max

    def _bijector_fn(params):
      return RationalQuadraticSpline(params, self.range_min, self.range_max)

    super().__init__(n_keep, decoder, conditioner, _bijector_fn)


class MaskedAutoregressiveInferenceFunnel(Surjector):
  """_src/surjectors/masked_autoregressive_inference_funnel.py:10-108."""

  def __init__(self, n_keep, decoder, conditioner, bijector_fn):
    self.n_keep = n_keep
    self.decoder = decoder
    self.conditioner = conditioner
    self.bijector_fn = bijector_fn

  def _inverse_and_likelihood_contribution(self, y, x=None, **kwargs):  # :77-90
    y_plus, y_minus = y[..., : self.n_keep], y[..., self.n_keep :]
    y_cond = y_minus if x is None else torch.cat([y_minus, x], dim=-1)
    z, jac_det = MaskedAutoregressive(self.conditioner, self.bijector_fn).inverse_and_log_det(y_plus, y_cond)
    z_condition = z if x is None else torch.cat([z, x], dim=-1)
    lc = self.decoder(z_condition).log_prob(y_minus)
    return z, lc + jac_det

  def _forward_and_likelihood_contribution(self, z, x=None, **kwargs):
    raise NotImplementedError("sampling direction is SURVEY section 8f (next); broken upstream (quirk Q9)")


class AffineMaskedAutoregressiveInferenceFunnel(MaskedAutoregressiveInferenceFunnel):
  """_src/surjectors/affine_masked_autoregressive_inference_funnel.py:13-65."""

  def __init__(self, n_keep, decoder, conditioner):
    def bijector_fn(params):
      shift, log_scale = unstack(params, axis=-1)
      return ScalarAffine(shift, torch.exp(log_scale))

    super().__init__(n_keep, decoder, conditioner, bijector_fn)


# ------------------------------------------------------------- common user idioms
def normal_decoder(n_dim, hidden=(4, 4), conditioner_factory=None):
  """The `_decoder_fn` idiom of examples/coupling_inference_surjection.py:24-33."""
  net = make_mlp(tuple(hidden) + (n_dim * 2,)) if conditioner_factory is None else conditioner_factory()

  def _fn(z):
    params = net(z)
    mu, log_scale = split(params, 2, -1)
    return distrax.Independent(distrax.Normal(mu, torch.exp(log_scale)))

  return _fn


def standard_normal_base(n_dim, dtype=torch.float64):
  """tfd.Independent(tfd.Normal(zeros(n), ones(n)), 1) (examples/...:62-65)."""
  return distrax.Independent(
    distrax.Normal(torch.zeros(n_dim, dtype=dtype), torch.ones(n_dim, dtype=dtype)), reinterpreted_batch_ndims=1
  )

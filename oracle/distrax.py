"""Restatement of the distrax 0.1.7 pieces on the reference's hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  distrax is un-vendored
(uv.lock:506-508); reference call sites: masked_coupling.py:3-4,12,65,79,
affine_masked_coupling.py:48, rq_masked_coupling.py:59-61,
affine_masked_autoregressive.py:56, funnels :53 / :82-84.  Semantics follow the
published distrax sources for that version.
PARITY UNPINNED by upstream fixtures except ScalarAffine.forward (MAF KATs).
"""

import math

import torch


def sum_last(x, ndims):
  """distrax.math.sum_last: sum over the last `ndims` axes."""
  if ndims == 0:
    return x
  return x.sum(dim=tuple(range(-ndims, 0)))


class ScalarAffine:
  """y = scale * x + shift; fwd log|scale|; inverse uses 1/scale."""

  def __init__(self, shift, scale=None, log_scale=None):
    self._shift = shift
    if scale is not None:
      self._scale = scale
      self._inv_scale = 1.0 / scale
      self._log_scale = torch.log(torch.abs(scale))
    else:
      self._scale = torch.exp(log_scale)
      self._inv_scale = torch.exp(-log_scale)
      self._log_scale = log_scale

  def forward_and_log_det(self, x):
    return self._scale * x + self._shift, self._log_scale.expand(torch.broadcast_shapes(self._log_scale.shape, x.shape))

  def inverse_and_log_det(self, y):
    x = self._inv_scale * (y - self._shift)
    ld = -self._log_scale
    return x, ld.expand(torch.broadcast_shapes(ld.shape, y.shape))


class Inverse:
  def __init__(self, bijector):
    self._b = bijector

  def forward_and_log_det(self, x):
    return self._b.inverse_and_log_det(x)

  def inverse_and_log_det(self, y):
    return self._b.forward_and_log_det(y)


def _normalize_bin_sizes(u, total_size, min_bin_size):
  k = u.shape[-1]
  if k * min_bin_size > total_size:
    raise ValueError("too many bins for the range")
  return torch.softmax(u, dim=-1) * (total_size - k * min_bin_size) + min_bin_size


def _normalize_knot_slopes(u, min_knot_slope):
  offset = math.log(math.exp(1.0 - min_knot_slope) - 1.0)
  return torch.nn.functional.softplus(u + offset) + min_knot_slope


class RationalQuadraticSpline:
  """distrax.RationalQuadraticSpline(params[...,3K+1], range_min, range_max,
  boundary_slopes='unconstrained', min_bin_size=1e-4, min_knot_slope=1e-4)."""

  def __init__(self, params, range_min, range_max, min_bin_size=1e-4, min_knot_slope=1e-4):
    if params.shape[-1] % 3 != 1 or params.shape[-1] < 4:
      raise ValueError("last dim of params must be 3*num_bins+1")
    k = (params.shape[-1] - 1) // 3
    self.num_bins = k
    uw, uh, us = params[..., :k], params[..., k : 2 * k], params[..., 2 * k :]
    size = range_max - range_min
    bw = _normalize_bin_sizes(uw, size, min_bin_size)
    bh = _normalize_bin_sizes(uh, size, min_bin_size)
    lo = torch.full(params.shape[:-1] + (1,), float(range_min), dtype=params.dtype)
    hi = torch.full(params.shape[:-1] + (1,), float(range_max), dtype=params.dtype)
    self.x_pos = torch.cat([lo, range_min + torch.cumsum(bw[..., :-1], -1), hi], -1)
    self.y_pos = torch.cat([lo, range_min + torch.cumsum(bh[..., :-1], -1), hi], -1)
    self.knot_slopes = _normalize_knot_slopes(us, min_knot_slope)

  def _gather(self, v, pos):
    """One-hot bin selection exactly as distrax: pos[k] <= v < pos[k+1]; if none, first bin."""
    vv = v.unsqueeze(-1)
    correct = (vv >= pos[..., :-1]) & (vv < pos[..., 1:])
    none = ~correct.any(dim=-1, keepdim=True)
    first = torch.zeros_like(correct)
    first[..., 0] = True
    correct = torch.where(none, first, correct).to(v.dtype)

    def pick(arr):
      return (correct * arr[..., :-1]).sum(-1), (correct * arr[..., 1:]).sum(-1)

    return pick(self.x_pos), pick(self.y_pos), pick(self.knot_slopes)

  def forward_and_log_det(self, x):
    (x0, x1), (y0, y1), (d0, d1) = self._gather(x, self.x_pos)
    bw, bh = x1 - x0, y1 - y0
    s = bh / bw
    z = torch.clamp((x - x0) / bw, 0.0, 1.0)
    sq_z = z * z
    z1mz = z - sq_z
    sq_1mz = (1.0 - z) ** 2
    t = d1 + d0 - 2.0 * s
    num = bh * (s * sq_z + d0 * z1mz)
    den = s + t * z1mz
    y = y0 + num / den
    ld = 2.0 * torch.log(s) + torch.log(d1 * sq_z + 2.0 * s * z1mz + d0 * sq_1mz) - 2.0 * torch.log(den)
    below = x <= self.x_pos[..., 0]
    above = x >= self.x_pos[..., -1]
    k0, k1 = self.knot_slopes[..., 0], self.knot_slopes[..., -1]
    y = torch.where(below, (x - self.x_pos[..., 0]) * k0 + self.y_pos[..., 0], y)
    y = torch.where(above, (x - self.x_pos[..., -1]) * k1 + self.y_pos[..., -1], y)
    ld = torch.where(below, torch.log(k0), ld)
    ld = torch.where(above, torch.log(k1), ld)
    return y, ld

  def inverse_and_log_det(self, y):
    (x0, x1), (y0, y1), (d0, d1) = self._gather(y, self.y_pos)
    bw, bh = x1 - x0, y1 - y0
    s = bh / bw
    w = torch.clamp((y - y0) / bh, 0.0, 1.0)
    t = d1 + d0 - 2.0 * s
    c = -s * w
    b = d0 - t * w
    a = s - b
    z = -2.0 * c / (b + torch.sqrt(b * b - 4.0 * a * c))
    z = torch.clamp(z, 0.0, 1.0)
    x = bw * z + x0
    sq_z = z * z
    z1mz = z - sq_z
    sq_1mz = (1.0 - z) ** 2
    den = s + t * z1mz
    ld = -2.0 * torch.log(s) - torch.log(d1 * sq_z + 2.0 * s * z1mz + d0 * sq_1mz) + 2.0 * torch.log(den)
    below = y <= self.y_pos[..., 0]
    above = y >= self.y_pos[..., -1]
    k0, k1 = self.knot_slopes[..., 0], self.knot_slopes[..., -1]
    x = torch.where(below, (y - self.y_pos[..., 0]) / k0 + self.x_pos[..., 0], x)
    x = torch.where(above, (y - self.y_pos[..., -1]) / k1 + self.x_pos[..., -1], x)
    ld = torch.where(below, -torch.log(k0), ld)
    ld = torch.where(above, -torch.log(k1), ld)
    return x, ld


class Normal:
  """distrax.Normal: log_prob = -0.5*((v-m)/s)^2 - (0.5*log(2pi) + log s)."""

  def __init__(self, loc, scale):
    self.loc = loc
    self.scale = scale

  def log_prob(self, value):
    z = (value - self.loc) / self.scale
    return -0.5 * z * z - (0.5 * math.log(2.0 * math.pi) + torch.log(self.scale))

  @property
  def batch_shape(self):
    return torch.broadcast_shapes(tuple(self.loc.shape), tuple(self.scale.shape))


class Independent:
  """distrax.Independent(d, n=None): n defaults to all batch axes but the first
  (tfd.Independent has the same default)."""

  def __init__(self, distribution, reinterpreted_batch_ndims=None):
    self.distribution = distribution
    if reinterpreted_batch_ndims is None:
      reinterpreted_batch_ndims = max(0, len(distribution.batch_shape) - 1)
    self.n = reinterpreted_batch_ndims

  def log_prob(self, value):
    return sum_last(self.distribution.log_prob(value), self.n)

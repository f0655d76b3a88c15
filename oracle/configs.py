"""The five BASELINE.json configs, written against the oracle exactly as the
reference's examples write them against surjectors.  TEST INFRASTRUCTURE.

cfg1  README.md:30-45                       Chain([Slice(5, dec), LULinear(5)]), dim 10
cfg2  examples/coupling_inference_surjection.py:40-67     20 -> 10 -> 10 -> 5
cfg3  examples/autoregressive_inference_surjection.py:35-67
cfg4  docs/notebooks/dimension_reduction.ipynb:600-640 scaled
cfg5  examples/conditional_density_estimation.py:28-76 scaled

Each `make_cfgN(**scale)` returns `flow(y=..., x=...) -> lp` to be wrapped in
`hk.transform`.  Scale arguments default to BASELINE sizes; tests pass smaller ones.
"""

import numpy as np
import torch

from oracle import hk
from oracle import surjectors as S


def _dec(n_dim, hidden):
  return S.normal_decoder(n_dim, hidden)


def make_cfg1(dim=10, n_keep=5, dec_hidden=(32, 32)):
  def flow(y, x=None):
    transform = S.Chain([S.Slice(n_keep, _dec(dim - n_keep, dec_hidden)), S.LULinear(n_keep)])
    td = S.TransformedDistribution(S.standard_normal_base(n_keep, y.dtype), transform)
    return td.log_prob(y, x)

  return flow


def make_cfg2(dim=20, cond_hidden=(8, 8), dec_hidden=(4, 4)):
  def bijector_fn(params):
    shift, log_scale = S.split(params, 2, -1)
    return S.ScalarAffine(shift, torch.exp(log_scale))

  def flow(y, x=None):
    n_dim = dim
    layers = []
    for i in range(3):
      if i != 1:
        layer = S.AffineMaskedCouplingInferenceFunnel(
          n_keep=int(n_dim / 2),
          decoder=_dec(int(n_dim / 2), dec_hidden),
          conditioner=S.make_mlp(tuple(cond_hidden) + (n_dim * 2,)),
        )
        n_dim = int(n_dim / 2)
      else:
        layer = S.MaskedCoupling(
          mask=S.make_alternating_binary_mask(n_dim, i % 2 == 0),
          bijector_fn=bijector_fn,
          conditioner=S.make_mlp(list(cond_hidden) + [n_dim * 2]),
        )
      layers.append(layer)
    td = S.TransformedDistribution(S.standard_normal_base(n_dim, y.dtype), S.Chain(layers))
    return td.log_prob(y, x)

  return flow


def make_cfg3(dim=20, made_hidden=(8, 8), dec_hidden=(4, 4)):
  def made_bijector_fn(params):
    means, log_scales = S.unstack(params, -1)
    return S.Inverse(S.ScalarAffine(means, torch.exp(log_scales)))

  def flow(y, x=None):
    n_dim = dim
    layers = []
    for i in range(3):
      if i != 1:
        layer = S.AffineMaskedAutoregressiveInferenceFunnel(
          n_keep=int(n_dim / 2),
          decoder=_dec(int(n_dim / 2), dec_hidden),
          conditioner=S.MADE(int(n_dim / 2), list(made_hidden), 2),
        )
        n_dim = int(n_dim / 2)
      else:
        layer = S.MaskedAutoregressive(conditioner=S.MADE(n_dim, list(made_hidden), 2), bijector_fn=made_bijector_fn)
      layers.append(layer)
    td = S.TransformedDistribution(S.standard_normal_base(n_dim, y.dtype), S.Chain(layers))
    return td.log_prob(y, x)

  return flow


def make_cfg4(dim=64, n_keep=16, hidden=(256, 256), n_bins=64, n_pre=4, n_post=4, rng=(-3.0, 3.0)):
  n_par = 3 * n_bins + 1

  def cond(d):
    return hk.Sequential([S.make_mlp(tuple(hidden) + (d * n_par,)), hk.Reshape((d, n_par))])

  def flow(y, x=None):
    layers = []
    for i in range(n_pre):
      layers.append(
        S.RationalQuadraticSplineMaskedCoupling(
          S.make_alternating_binary_mask(dim, i % 2 == 0), cond(dim), rng[0], rng[1]
        )
      )
    layers.append(
      S.RationalQuadraticSplineMaskedCouplingInferenceFunnel(
        n_keep, _dec(dim - n_keep, hidden), cond(dim), rng[0], rng[1]
      )
    )
    for i in range(n_post):
      layers.append(
        S.RationalQuadraticSplineMaskedCoupling(
          S.make_alternating_binary_mask(n_keep, i % 2 == 0), cond(n_keep), rng[0], rng[1]
        )
      )
    td = S.TransformedDistribution(S.standard_normal_base(n_keep, y.dtype), S.Chain(layers))
    return td.log_prob(y, x)

  return flow


def make_cfg5(dim=128, hidden=(512, 512), n_layers=16):
  def bijector_fn(params):
    means, log_scales = S.unstack(params, -1)
    return S.ScalarAffine(means, torch.exp(log_scales))

  def flow(y, x=None):
    layers = []
    order = np.arange(dim)
    for _ in range(n_layers):
      layers.append(
        S.MaskedAutoregressive(
          bijector_fn=bijector_fn,
          conditioner=S.MADE(
            dim, list(hidden), 2, w_init=hk.initializers.TruncatedNormal(0.01), b_init=hk.zeros
          ),
        )
      )
      layers.append(S.LULinear(dim))
      order = order[::-1]
      layers.append(S.Permutation(order, 1))
    layers = layers[:-1]
    td = S.TransformedDistribution(S.standard_normal_base(dim, y.dtype), S.Chain(layers))
    return td.log_prob(y, x)

  return flow


def make_mlpf(dim=20, n_keep=10, dec_hidden=(8, 8), n_lu=2):
  def flow(y, x=None):
    layers = [S.LULinear(dim) for _ in range(n_lu)]
    layers.append(S.MLPInferenceFunnel(n_keep, _dec(dim - n_keep, dec_hidden)))
    layers += [S.LULinear(n_keep) for _ in range(n_lu)]
    td = S.TransformedDistribution(S.standard_normal_base(n_keep, y.dtype), S.Chain(layers))
    return td.log_prob(y, x)

  return flow


def make_cond(dim=12, dx=3, hidden=(8, 8), n_bins=4):
  def _unstack_fn(params):
    means, log_scales = S.unstack(params, -1)
    return S.ScalarAffine(means, torch.exp(log_scales))

  def _split_inv(params):
    a, b = S.split(params, 2, -1)
    return S.Inverse(S.ScalarAffine(a, torch.exp(b)))

  def flow(y, x=None):
    d = dim
    layers = [S.AffineMaskedAutoregressiveInferenceFunnel(d - 2, _dec(2, hidden), S.MADE(d - 2, list(hidden), 2))]
    d -= 2
    layers += [
      S.MaskedCoupling(S.make_alternating_binary_mask(d, True), S.make_mlp(list(hidden) + [2 * d]), _split_inv),
      S.RationalQuadraticSplineMaskedCoupling(
        S.make_alternating_binary_mask(d),
        hk.Sequential([S.make_mlp(list(hidden) + [d * (3 * n_bins + 1)]), hk.Reshape((d, 3 * n_bins + 1))]), -3.0, 3.0),
      S.AffineMaskedCouplingInferenceFunnel(d - 3, _dec(3, hidden), S.make_mlp(list(hidden) + [2 * d])),
    ]
    d -= 3
    layers += [
      S.MaskedAutoregressive(S.MADE(d, list(hidden), 2), _unstack_fn),
      S.LULinear(d),
      S.Permutation(np.arange(d)[::-1], 1),
      S.Slice(d - 2, _dec(2, hidden)),
    ]
    d -= 2
    td = S.TransformedDistribution(S.standard_normal_base(d, y.dtype), S.Chain(layers))
    return td.log_prob(y, x)

  return flow


def make_cpl(dim=12, hidden=(16, 16), dec_hidden=(8, 8), activation="tanh"):
  act = {"tanh": torch.tanh, "relu": hk.relu, "swish": hk.swish}[activation]

  def _split_inv(params):
    a, b = S.split(params, 2, -1)
    return S.Inverse(S.ScalarAffine(a, torch.exp(b)))

  def _unstack_fn(params):
    means, log_scales = S.unstack(params, -1)
    return S.ScalarAffine(means, torch.exp(log_scales))

  def _mlp(dims):
    return S.make_mlp(dims, activation=act)

  def flow(y, x=None):
    d = dim
    layers = [
      S.MaskedCoupling(S.make_alternating_binary_mask(d, True), _mlp(list(hidden) + [2 * d]), _split_inv),
      S.AffineMaskedCouplingInferenceFunnel(
        d - 4, S.normal_decoder(4, conditioner_factory=lambda: _mlp(list(dec_hidden) + [8])), _mlp(list(hidden) + [2 * d])),
    ]
    d -= 4
    layers += [
      S.MaskedCoupling(S.make_alternating_binary_mask(d), hk.Sequential([_mlp(list(hidden) + [2 * d]), hk.Reshape((d, 2))]),
                       _unstack_fn),
      S.AffineMaskedCouplingInferenceFunnel(
        d - 3, S.normal_decoder(3, conditioner_factory=lambda: _mlp(list(dec_hidden) + [6])), _mlp(list(hidden) + [2 * d])),
    ]
    d -= 3
    td = S.TransformedDistribution(S.standard_normal_base(d, y.dtype), S.Chain(layers))
    return td.log_prob(y, x)

  return flow


def randomize(params, seed, w_std=0.3, b_val=None, b_std=0.1, literal=False):
  """Replace haiku-default inits by non-trivial values so every transform is exercised
  (mirrors masked_coupling_test.py:78-82: TruncatedNormal(1.0) weights, ones bias)."""
  rng = np.random.Generator(np.random.PCG64(seed))
  out = {}
  for mod, bucket in params.items():
    out[mod] = {}
    for name, val in bucket.items():
      if name == "w":
        fan_in = val.shape[0]
        if literal:  # per-weight sigma = w_std, truncated at 2 sigma: hk.initializers.TruncatedNormal(stddev=w_std)
          new = np.clip(rng.standard_normal(val.shape), -2.0, 2.0) * w_std
        else:
          new = rng.standard_normal(val.shape) * (w_std / np.sqrt(max(fan_in, 1)) * np.sqrt(8.0))
      elif name == "b":
        new = rng.standard_normal(val.shape) * b_std if b_val is None else np.full(val.shape, b_val)
      elif name in ("lower_entries", "upper_entries"):
        n_dim = max(1.0, (1.0 + np.sqrt(1.0 + 8.0 * val.shape[0])) / 2.0)  # entries = n (n - 1) / 2
        new = rng.standard_normal(val.shape) * (0.5 / n_dim)
      elif name == "diag_entries":
        new = 0.5413 + 0.2 * rng.standard_normal(val.shape)  # softplus(0.5413) = 1: keeps deep chains conditioned
      else:
        new = val.numpy()
      out[mod][name] = torch.as_tensor(np.asarray(new), dtype=val.dtype).reshape(val.shape).clone()
  return out

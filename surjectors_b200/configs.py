"""The five BASELINE.json configs written against the public API exactly as the reference's
examples write them against surjectors (model definitions only; citations relative to
/root/reference/):

cfg1  README.md:30-45                                     Chain([Slice(5, dec), LULinear(5)]), dim 10
cfg2  examples/coupling_inference_surjection.py:40-67     20 -> 10 -> 10 -> 5
cfg3  examples/autoregressive_inference_surjection.py:35-67
cfg4  docs/notebooks/dimension_reduction.ipynb:600-640 scaled to 8 spline layers, 64 bins, hidden 256, 64 -> 16
cfg5  examples/conditional_density_estimation.py:28-76 scaled to a conditional 16-layer MAF, dim 128

Each `make_cfgN(**scale)` returns `flow(method, **kwargs)` to be wrapped in `hk.transform`.
"""

import numpy as np

from surjectors_b200 import (
  AffineMaskedAutoregressiveInferenceFunnel,
  AffineMaskedCouplingInferenceFunnel,
  Chain,
  LULinear,
  MLPInferenceFunnel,
  MaskedAutoregressive,
  MaskedCoupling,
  Permutation,
  RationalQuadraticSplineMaskedCoupling,
  RationalQuadraticSplineMaskedCouplingInferenceFunnel,
  Slice,
  TransformedDistribution,
  hk,
)
from surjectors_b200 import _sym
from surjectors_b200.compat import distrax, jnp, tfd
from surjectors_b200.nn import MADE, make_mlp
from surjectors_b200.util import make_alternating_binary_mask, unstack


def _decoder_fn(n_dim, hidden=(4, 4)):
  decoder_net = make_mlp(list(hidden) + [n_dim * 2])

  def _fn(z):
    params = decoder_net(z)
    mu, log_scale = jnp.split(params, 2, -1)
    return tfd.Independent(tfd.Normal(mu, jnp.exp(log_scale)))

  return _fn


def _base(n_dim):
  return tfd.Independent(tfd.Normal(jnp.zeros(n_dim), jnp.ones(n_dim)), 1)


def make_cfg1(dim=10, n_keep=5, dec_hidden=(32, 32)):
  def flow(method, **kwargs):
    transform = Chain([Slice(n_keep, _decoder_fn(dim - n_keep, dec_hidden)), LULinear(n_keep)])
    return TransformedDistribution(_base(n_keep), transform)(method=method, **kwargs)

  return flow


def make_cfg2(dim=20, cond_hidden=(8, 8), dec_hidden=(4, 4)):
  def _bijector_fn(params):
    shift, log_scale = jnp.split(params, 2, -1)
    return distrax.ScalarAffine(shift, jnp.exp(log_scale))

  def flow(method, **kwargs):
    n_dim = dim
    layers = []
    for i in range(3):
      if i != 1:
        layer = AffineMaskedCouplingInferenceFunnel(
          n_keep=int(n_dim / 2),
          decoder=_decoder_fn(int(n_dim / 2), dec_hidden),
          conditioner=make_mlp(list(cond_hidden) + [n_dim * 2]),
        )
        n_dim = int(n_dim / 2)
      else:
        layer = MaskedCoupling(
          mask=make_alternating_binary_mask(n_dim, i % 2 == 0),
          bijector_fn=_bijector_fn,
          conditioner=make_mlp(list(cond_hidden) + [n_dim * 2]),
        )
      layers.append(layer)
    return TransformedDistribution(_base(n_dim), Chain(layers))(method=method, **kwargs)

  return flow


def make_cfg3(dim=20, made_hidden=(8, 8), dec_hidden=(4, 4)):
  def _made_bijector_fn(params):
    means, log_scales = unstack(params, -1)
    return distrax.Inverse(distrax.ScalarAffine(means, jnp.exp(log_scales)))

  def flow(method, **kwargs):
    n_dim = dim
    layers = []
    for i in range(3):
      if i != 1:
        layer = AffineMaskedAutoregressiveInferenceFunnel(
          n_keep=int(n_dim / 2),
          decoder=_decoder_fn(int(n_dim / 2), dec_hidden),
          conditioner=MADE(int(n_dim / 2), list(made_hidden), 2),
        )
        n_dim = int(n_dim / 2)
      else:
        layer = MaskedAutoregressive(conditioner=MADE(n_dim, list(made_hidden), 2), bijector_fn=_made_bijector_fn)
      layers.append(layer)
    return TransformedDistribution(_base(n_dim), Chain(layers))(method=method, **kwargs)

  return flow


def make_cfg4(dim=64, n_keep=16, hidden=(256, 256), n_bins=64, n_pre=4, n_post=4, rng=(-3.0, 3.0)):
  n_par = 3 * n_bins + 1

  def cond(d):
    return hk.Sequential([make_mlp(list(hidden) + [d * n_par]), hk.Reshape((d, n_par))])

  def flow(method, **kwargs):
    layers = []
    for i in range(n_pre):
      layers.append(
        RationalQuadraticSplineMaskedCoupling(make_alternating_binary_mask(dim, i % 2 == 0), cond(dim), rng[0], rng[1])
      )
    layers.append(
      RationalQuadraticSplineMaskedCouplingInferenceFunnel(
        n_keep, _decoder_fn(dim - n_keep, hidden), cond(dim), rng[0], rng[1]
      )
    )
    for i in range(n_post):
      layers.append(
        RationalQuadraticSplineMaskedCoupling(
          make_alternating_binary_mask(n_keep, i % 2 == 0), cond(n_keep), rng[0], rng[1]
        )
      )
    return TransformedDistribution(_base(n_keep), Chain(layers))(method=method, **kwargs)

  return flow


def make_cfg5(dim=128, hidden=(512, 512), n_layers=16):
  def _bijector_fn(params):
    means, log_scales = unstack(params, -1)
    return distrax.ScalarAffine(means, jnp.exp(log_scales))

  def flow(method, **kwargs):
    layers = []
    order = np.arange(dim)
    for _ in range(n_layers):
      layers.append(
        MaskedAutoregressive(
          bijector_fn=_bijector_fn,
          conditioner=MADE(dim, list(hidden), 2, w_init=hk.initializers.TruncatedNormal(0.01), b_init=hk.zeros),
        )
      )
      layers.append(LULinear(dim))
      order = order[::-1]
      layers.append(Permutation(order, 1))
    layers = layers[:-1]
    return TransformedDistribution(_base(dim), Chain(layers))(method=method, **kwargs)

  return flow


def make_mlpf(dim=20, n_keep=10, dec_hidden=(8, 8), n_lu=2):
  """docs/notebooks/dimension_reduction.ipynb:397-420: LULinear layers around an MLPInferenceFunnel (section 8f.3)."""

  def flow(method, **kwargs):
    layers = [LULinear(dim) for _ in range(n_lu)]
    layers.append(MLPInferenceFunnel(n_keep, _decoder_fn(dim - n_keep, dec_hidden)))
    layers += [LULinear(n_keep) for _ in range(n_lu)]
    return TransformedDistribution(_base(n_keep), Chain(layers))(method=method, **kwargs)

  return flow


def make_cond(dim=12, dx=3, hidden=(8, 8), n_bins=4):
  """Every layer kind on the path with a conditioning input x (the reference's conditional density estimation
  pattern, examples/conditional_density_estimation.py:28-76): AR funnel, coupling, spline coupling, coupling funnel,
  MAF, LULinear, Permutation, Slice."""

  def _unstack_fn(params):
    means, log_scales = unstack(params, -1)
    return distrax.ScalarAffine(means, jnp.exp(log_scales))

  def flow(method, **kwargs):
    d = dim
    layers = [
      AffineMaskedAutoregressiveInferenceFunnel(d - 2, _decoder_fn(2, hidden), MADE(d - 2, list(hidden), 2)),
    ]
    d -= 2
    layers += [
      MaskedCoupling(make_alternating_binary_mask(d, True), make_mlp(list(hidden) + [2 * d]),
                     lambda p: distrax.Inverse(distrax.ScalarAffine(*(lambda a, b: (a, jnp.exp(b)))(*jnp.split(p, 2, -1))))),
      RationalQuadraticSplineMaskedCoupling(
        make_alternating_binary_mask(d), hk.Sequential([make_mlp(list(hidden) + [d * (3 * n_bins + 1)]),
                                                        hk.Reshape((d, 3 * n_bins + 1))]), -3.0, 3.0),
      AffineMaskedCouplingInferenceFunnel(d - 3, _decoder_fn(3, hidden), make_mlp(list(hidden) + [2 * d])),
    ]
    d -= 3
    layers += [
      MaskedAutoregressive(MADE(d, list(hidden), 2), _unstack_fn),
      LULinear(d),
      Permutation(np.arange(d)[::-1], 1),
      Slice(d - 2, _decoder_fn(2, hidden)),
    ]
    d -= 2
    return TransformedDistribution(_base(d), Chain(layers))(method=method, **kwargs)

  return flow


def make_cpl(dim=12, hidden=(16, 16), dec_hidden=(8, 8), activation="tanh"):
  """A user-style affine coupling / funnel chain off the BASELINE shapes: another activation and hidden widths, an
  Inverse(ScalarAffine) inner bijector and an unstack-ed (interleaved) conditioner output -- the family of the compact
  fused kernel (csrc/compact.cuh)."""
  act = {"tanh": _sym.tanh, "relu": _sym.relu, "swish": _sym.swish}[activation]

  def _split_inv(params):
    a, b = jnp.split(params, 2, -1)
    return distrax.Inverse(distrax.ScalarAffine(a, jnp.exp(b)))

  def _unstack_fn(params):
    means, log_scales = unstack(params, -1)
    return distrax.ScalarAffine(means, jnp.exp(log_scales))

  def _dec(n_dim):
    net = make_mlp(list(dec_hidden) + [n_dim * 2], activation=act)

    def _fn(z):
      mu, log_scale = jnp.split(net(z), 2, -1)
      return tfd.Independent(tfd.Normal(mu, jnp.exp(log_scale)))

    return _fn

  def flow(method, **kwargs):
    d = dim
    layers = [
      MaskedCoupling(make_alternating_binary_mask(d, True), make_mlp(list(hidden) + [2 * d], activation=act), _split_inv),
      AffineMaskedCouplingInferenceFunnel(d - 4, _dec(4), make_mlp(list(hidden) + [2 * d], activation=act)),
    ]
    d -= 4
    layers += [
      MaskedCoupling(make_alternating_binary_mask(d), hk.Sequential([make_mlp(list(hidden) + [2 * d], activation=act),
                                                                      hk.Reshape((d, 2))]), _unstack_fn),
      AffineMaskedCouplingInferenceFunnel(d - 3, _dec(3), make_mlp(list(hidden) + [2 * d], activation=act)),
    ]
    d -= 3
    return TransformedDistribution(_base(d), Chain(layers))(method=method, **kwargs)

  return flow


CONFIGS = {
  "cpl": dict(make=make_cpl, dim=12, dim_x=0, batch=4096),
  "cfg1": dict(make=make_cfg1, dim=10, dim_x=0, batch=4096),
  "cfg2": dict(make=make_cfg2, dim=20, dim_x=0, batch=65536),
  "cfg3": dict(make=make_cfg3, dim=20, dim_x=0, batch=65536),
  "cfg4": dict(make=make_cfg4, dim=64, dim_x=0, batch=1 << 20),
  "cfg5": dict(make=make_cfg5, dim=128, dim_x=16, batch=1 << 23),
  "mlpf": dict(make=make_mlpf, dim=20, dim_x=0, batch=65536),
  "cond": dict(make=make_cond, dim=12, dim_x=3, batch=65536),
}

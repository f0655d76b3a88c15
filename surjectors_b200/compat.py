"""Stand-ins for the names user code imports next to surjectors (`jax.numpy as jnp`, `jax.nn`,
`distrax`, `tfp.distributions as tfd`, `haiku as hk`) when it writes conditioners, bijector_fns,
decoders and base distributions -- e.g. examples/coupling_inference_surjection.py:24-37:

    def _decoder_fn(n_dim):
      decoder_net = make_mlp([4, 4, n_dim * 2])
      def _fn(z):
        params = decoder_net(z)
        mu, log_scale = jnp.split(params, 2, -1)
        return tfd.Independent(tfd.Normal(mu, jnp.exp(log_scale)))
      return _fn

works unchanged with `from surjectors_b200.compat import jnp, tfd`.  On conditioner outputs the
functions build symbolic descriptors; on numbers they are NumPy."""

import types

import numpy as np

from surjectors_b200 import _sym
from surjectors_b200 import hk  # noqa: F401

jnp = types.SimpleNamespace(
  split=_sym.split, exp=_sym.exp, zeros=np.zeros, ones=np.ones, arange=np.arange, array=np.asarray,
  asarray=np.asarray, float32=np.float32, int32=np.int32, logical_not=np.logical_not, concatenate=np.concatenate,
)

_nn = types.SimpleNamespace(
  relu=_sym.relu, swish=_sym.swish, silu=_sym.silu, tanh=_sym.tanh, sigmoid=_sym.sigmoid, softplus=_sym.softplus,
  gelu=_sym.gelu, elu=_sym.elu, leaky_relu=_sym.leaky_relu,
)
jax = types.SimpleNamespace(nn=_nn, numpy=jnp)

distrax = types.SimpleNamespace(
  ScalarAffine=_sym.ScalarAffine, Inverse=_sym.Inverse, RationalQuadraticSpline=_sym.RationalQuadraticSpline,
  Normal=_sym.Normal, Independent=_sym.Independent, MultivariateNormalDiag=_sym.MultivariateNormalDiag,
)
tfd = types.SimpleNamespace(
  Normal=_sym.Normal, Independent=_sym.Independent, MultivariateNormalDiag=_sym.MultivariateNormalDiag,
)

"""Sampling direction on the CUDA path: the reference's four MAF *forward*
known-answer values (masked_autoregressive_test.py:113-114,149-150) reproduced by the kernels, and
forward/inverse round trips of bijective chains through both executors."""

import json
import os

import numpy as np
import pytest
import torch

import surjectors_b200 as sj
from surjectors_b200 import execution_mode, hk
from surjectors_b200.compat import distrax, jnp
from surjectors_b200.nn import MADE, make_mlp
from surjectors_b200.util import make_alternating_binary_mask, unstack

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat.json")))


def _maf(hidden):
  def bij(params):
    means, log_scales = unstack(params, -1)
    return distrax.ScalarAffine(means, jnp.exp(log_scales))

  def fn(**kw):
    layer = sj.MaskedAutoregressive(
      bijector_fn=bij, conditioner=MADE(3, hidden, 2, w_init=hk.ones, b_init=hk.zeros, activation=lambda v: v))
    return layer.forward(**kw)

  return hk.transform(fn)


@pytest.mark.parametrize("mode", ["fused", "layered"])
@pytest.mark.parametrize("hidden", [[2, 4, 2], [2]])
def test_reference_maf_forward_kats(hidden, mode):
  key = "[" + ",".join(str(h) for h in hidden) + "]"
  m = _maf(hidden)
  z = torch.ones((1, 3), device=DEV)
  with execution_mode(mode):
    params = m.init(0, z=z)
    y = m.apply(params, None, z=z)
    assert np.allclose(y.cpu().numpy(), np.array(KAT["maf_forward"][key]), rtol=1e-5)
    x = torch.tensor([[0.1, 0.2]], device=DEV)
    params = m.init(0, z=z, x=x)
    y = m.apply(params, None, z=z, x=x)
    # the last entry is e^42.5: ONE float32 ulp of that exponent is already 3.8e-6 relative and the
    # all-ones weights multiply upstream rounding by the path count, so float32 implementations
    # legitimately differ by a few 1e-5 here (the float64 oracle matches the reference to 1e-8)
    assert np.allclose(y.cpu().numpy(), np.array(KAT["conditional_maf_forward"][key]), rtol=5e-5)


def _bijective_chain(dim=6):
  def bij_unstack(params):
    means, log_scales = unstack(params, -1)
    return distrax.Inverse(distrax.ScalarAffine(means, jnp.exp(log_scales)))

  def rq_cond(d, k):
    return hk.Sequential([make_mlp([16, 16, d * (3 * k + 1)]), hk.Reshape((d, 3 * k + 1))])

  def fn(method, **kw):
    layers = [
      sj.AffineMaskedCoupling(make_alternating_binary_mask(dim, True), make_mlp([16, 16, 2 * dim])),
      sj.Permutation(np.arange(dim)[::-1], 1),
      sj.RationalQuadraticSplineMaskedCoupling(make_alternating_binary_mask(dim), rq_cond(dim, 6), -3.0, 3.0),
      sj.MaskedAutoregressive(MADE(dim, [16, 16], 2), bij_unstack),
      sj.AffineMaskedAutoregressive(MADE(dim, [12], 2)),
    ]
    base = distrax.Independent(distrax.Normal(jnp.zeros(dim), jnp.ones(dim)), 1)
    return sj.TransformedDistribution(base, sj.Chain(layers))(method=method, **kw)

  return hk.transform(fn)


def _randomize(params, seed, scale=0.4):
  g = torch.Generator().manual_seed(seed)
  for mod in params.values():
    for k, t in mod.items():
      t.copy_((torch.randn(t.shape, generator=g) * (scale / max(1.0, float(t.shape[0]) ** 0.5 if t.dim() > 1 else 1.0))).to(t.device))


@pytest.mark.parametrize("mode", ["fused", "layered"])
@pytest.mark.parametrize("conditional", [False, True])
def test_forward_inverts_inverse(mode, conditional):
  dim, B = 6, 333
  m = _bijective_chain(dim)
  g = torch.Generator().manual_seed(3)
  y = torch.randn((B, dim), generator=g).to(DEV)
  kw = dict(x=torch.randn((B, 3), generator=g).to(DEV)) if conditional else {}
  with execution_mode(mode):
    params = m.init(1, method="log_prob", y=y, **kw)
    _randomize(params, 5)
    z, lp = m.apply(params, None, method="inverse_and_log_prob", y=y, **kw)

    y2, lp2 = m.apply(params, 7, method="sample_and_log_prob", sample_shape=(B,), **kw)
    # log_prob of the drawn samples through the inverse direction agrees with what sampling reported
    lp_back = m.apply(params, None, method="log_prob", y=y2, **kw)
    assert torch.isfinite(y2).all() and torch.allclose(lp_back, lp2, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("mode", ["fused", "layered"])
def test_layer_forward_round_trip(mode):
  dim, B = 5, 200

  def build():
    return sj.Chain([
      sj.AffineMaskedCoupling(make_alternating_binary_mask(dim), make_mlp([8, 8, 2 * dim])),
      sj.AffineMaskedAutoregressive(MADE(dim, [8, 8], 2)),
      sj.Permutation(np.array([3, 0, 4, 1, 2]), 1),
    ])

  inv = hk.transform(lambda **kw: build().inverse_and_likelihood_contribution(**kw))
  fwd = hk.transform(lambda **kw: build().forward_and_likelihood_contribution(**kw))
  y = torch.randn((B, dim), generator=torch.Generator().manual_seed(0)).to(DEV)
  with execution_mode(mode):
    params = inv.init(0, y=y)
    _randomize(params, 11)
    z, lc = inv.apply(params, None, y=y)
    y2, fl = fwd.apply(params, None, z=z)
  assert torch.allclose(y2, y, rtol=1e-4, atol=1e-4), float((y2 - y).abs().max())
  assert torch.allclose(fl, -lc, rtol=1e-4, atol=1e-4), float((fl + lc).abs().max())


def test_lulinear_and_surjections_have_no_sampling_direction():
  def fn(**kw):
    return sj.Chain([sj.LULinear(4)]).forward_and_likelihood_contribution(**kw)

  m = hk.transform(fn)
  z = torch.zeros((2, 4), device=DEV)
  params = m.init(0, z=z)
  with pytest.raises(NotImplementedError):  # lu_linear.py:92-93 raises upstream as well
    m.apply(params, None, z=z)


@pytest.mark.parametrize("mode", ["fused", "layered"])
@pytest.mark.parametrize("name", ["cfg2", "cfg4"])
def test_sampling_through_surjections_matches_oracle(name, mode):
  """Slice / coupling-funnel sampling (slice.py:70-78, masked_coupling_inference_funnel.py:93-106) with the
  decoder noise passed in explicitly: same z, same eps -> same y and same contribution as the oracle."""
  from oracle import surjectors as OS
  from oracle import hk as ohk
  from oracle import configs as ocfg
  from surjectors_b200 import _runtime, configs as pcfg, fused_chains
  from tests import parity

  scale = parity.SMALL.get(name, {})
  pm, pparams, om, oparams, y64, x64 = parity.build_pair(name, 4, 0, 0.3, **scale)
  c = pcfg.CONFIGS[name]
  dim = scale.get("dim", c["dim"])
  spec = fused_chains.spec_of(c["make"](**scale), dim, c["dim_x"])
  B = 257
  g = torch.Generator().manual_seed(2)
  with execution_mode(mode):
    plan = _runtime.get_plan(spec, _runtime.BASE_NONE)
    n_eps, dz = plan.noise_dim(), plan.dim_out
    assert n_eps > 0
    z = torch.randn((B, dz), generator=g, dtype=torch.float64)
    eps = torch.randn((B, n_eps), generator=g, dtype=torch.float64)
    flat = parity.to_device(pparams, torch.device(DEV)).flat
    y, fl = plan.forward(z.float().to(DEV).contiguous(), None, flat, eps.float().to(DEV).contiguous())
  # oracle: same chain object, forward direction, noise handed out in the order the layers ask for it
  cursor = {"k": 0}

  def provider(shape, dtype):
    n = shape[-1]
    out = eps[:, cursor["k"] : cursor["k"] + n].to(dtype)
    cursor["k"] += n
    return out

  chain_holder = {}
  orig_td = OS.TransformedDistribution

  class Capture(orig_td):
    def log_prob(self, yy, xx=None):
      chain_holder["y"], chain_holder["ld"] = self.transform.forward_and_likelihood_contribution(z, xx)
      return torch.zeros(yy.shape[0], dtype=yy.dtype)

  OS.TransformedDistribution = Capture
  OS.noise_provider = provider
  try:
    om.apply(oparams, y=y64)
  finally:
    OS.TransformedDistribution = orig_td
    OS.noise_provider = None
  assert cursor["k"] == n_eps
  yo, lo = chain_holder["y"], chain_holder["ld"]
  assert torch.allclose(y.cpu().double(), yo, rtol=2e-5, atol=2e-5), float((y.cpu().double() - yo).abs().max())
  assert parity.lp_error(fl.cpu().numpy(), lo.numpy()) <= 2e-5


def test_sample_api_on_cfg2():
  from surjectors_b200 import configs as pcfg

  m = hk.transform(pcfg.CONFIGS["cfg2"]["make"]())
  y0 = torch.zeros((4, 20), device=DEV)
  params = m.init(3, method="log_prob", y=y0)
  y, lp = m.apply(params, 11, method="sample_and_log_prob", sample_shape=(1000,))
  assert tuple(y.shape) == (1000, 20) and tuple(lp.shape) == (1000,) and bool(torch.isfinite(y).all())
  y2 = m.apply(params, 11, method="sample", sample_shape=(1000,))
  assert torch.equal(y, y2)  # same rng -> same draws


@pytest.mark.parametrize("conditional", [False, True])
@pytest.mark.parametrize("dim,hidden", [(24, [48, 48]), (64, [96]), (7, [16, 16, 16])])
def test_maf_sampling_recursion_kernel_matches_the_unrolled_passes(dim, hidden, conditional):
  """masked_autoregressive.py:63-71 runs D MADE passes; the per-op executor runs ONE incremental pass (csrc/vm_maf.cuh).
  Same parameters, same latents: the recursion kernel against the D unrolled passes (fusion off), and a round trip
  through the inverse direction."""
  from surjectors_b200 import _abi

  def _inv_affine(params):  # the examples' Inverse(ScalarAffine) variant (examples/autoregressive_inference_surjection.py:35-37)
    means, log_scales = unstack(params, -1)
    return distrax.Inverse(distrax.ScalarAffine(means, jnp.exp(log_scales)))

  def build():
    return sj.Chain([sj.AffineMaskedAutoregressive(MADE(dim, hidden, 2)), sj.Permutation(np.arange(dim)[::-1].copy(), 1),
                     sj.MaskedAutoregressive(MADE(dim, hidden, 2), _inv_affine)])

  inv = hk.transform(lambda **kw: build().inverse_and_likelihood_contribution(**kw))
  fwd = hk.transform(lambda **kw: build().forward_and_likelihood_contribution(**kw))
  B = 300
  g = torch.Generator().manual_seed(dim)
  z = torch.randn((B, dim), generator=g).to(DEV)
  kw = dict(x=torch.randn((B, 5), generator=g).to(DEV)) if conditional else {}
  lib = _abi.lib()
  res = {}
  with execution_mode("layered"):
    params = fwd.init(0, z=z, **kw)
    _randomize(params, 13, scale=0.7 if dim < 32 else 0.25)  # (a 64-step recursion amplifies: keep the scales of order one)
    for on in (0, 1):
      prev = lib.surj_set_wide_fusion(on)
      try:
        res[on] = fwd.apply(params, None, z=z, **kw)
      finally:
        lib.surj_set_wide_fusion(prev)
    y1, fl1 = res[1]
    z_back, lc = inv.apply(params, None, y=y1, **kw)
  y0, fl0 = res[0]
  assert torch.isfinite(y1).all()
  tol = 2e-5 if dim < 32 else 2e-4  # two float32 evaluation orders of a D-step recursion
  assert torch.allclose(y1, y0, rtol=tol, atol=tol), float((y1 - y0).abs().max())
  assert torch.allclose(fl1, fl0, rtol=tol, atol=tol), float((fl1 - fl0).abs().max())
  assert torch.allclose(z_back, z, rtol=10 * tol, atol=10 * tol), float((z_back - z).abs().max())
  assert torch.allclose(fl1, -lc, rtol=10 * tol, atol=10 * tol)

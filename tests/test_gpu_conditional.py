"""Conditional flows (x given) through every layer kind, the latent z output, and the cotangents of the inputs
(g_y, g_x) against the oracle's autograd -- both program executors."""

import numpy as np
import pytest
import torch

from surjectors_b200 import execution_mode
from tests import parity

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("mode", ["fused", "layered"])
def test_conditional_chain_all_layer_kinds(mode):
  out = parity.check_config("cond", batch=333, check_grad=True, mode=mode)
  print("cond", mode, out)


@pytest.mark.parametrize("mode", ["auto", "fused", "layered"])
@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cond", "cfg5"])
def test_latent_and_input_cotangents(name, mode):
  scale = parity.SMALL.get(name, {})
  pm, pparams, om, oparams, y64, x64 = parity.build_pair(name, 128, 0, 0.3, **scale)
  keep = parity.differentiable_samples(om, oparams, dict(y=y64) if x64 is None else dict(y=y64, x=x64))
  if keep is not None:
    y64 = y64[keep]
    x64 = None if x64 is None else x64[keep]
  params = parity.to_device(pparams, DEV)
  y = y64.float().to(DEV).requires_grad_(True)
  x = None if x64 is None else x64.float().to(DEV).requires_grad_(True)
  kw = dict(y=y) if x is None else dict(y=y, x=x)
  g = torch.Generator().manual_seed(4)
  with execution_mode(mode):
    z, lp = pm.apply(params, None, method="inverse_and_log_prob", **kw)
  wz = torch.randn(z.shape, generator=g, dtype=torch.float64)
  wl = torch.randn(lp.shape, generator=g, dtype=torch.float64)
  obj = (z * wz.float().to(DEV)).sum() + (lp * wl.float().to(DEV)).sum()
  grads = torch.autograd.grad(obj, [y] + ([x] if x is not None else []))
  # oracle: same objective in float64
  yo = y64.clone().requires_grad_(True)
  xo = None if x64 is None else x64.clone().requires_grad_(True)
  import oracle.surjectors as OS

  holder = {}
  orig = OS.TransformedDistribution

  class Capture(orig):
    def log_prob(self, yy, xx=None):
      holder["z"], lp_ = self.inverse_and_log_prob(yy, xx)
      return lp_

  OS.TransformedDistribution = Capture
  try:
    lpo = om.apply(oparams, **(dict(y=yo) if xo is None else dict(y=yo, x=xo)))
  finally:
    OS.TransformedDistribution = orig
  zo = holder["z"]
  assert torch.allclose(z.detach().cpu().double(), zo.detach(), rtol=2e-5, atol=2e-5), float((z.detach().cpu().double() - zo).abs().max())
  objo = (zo * wz).sum() + (lpo * wl).sum()
  gos = torch.autograd.grad(objo, [yo] + ([xo] if xo is not None else []))
  for got, ref in zip(grads, gos):
    scale_ = max(1.0, float(ref.abs().max()))
    assert float((got.cpu().double() - ref).abs().max()) <= 1e-4 * scale_

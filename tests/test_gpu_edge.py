"""Edge cases the seeded-normal parity runs only touch by chance (sigma in {0.3, 1.0}, bias 1;
spline inputs exactly on knots, on the range boundaries and in the linear tails; large |log_scale|)."""

import numpy as np
import pytest
import torch

from oracle import distrax as odx
from tests import parity, recipes

pytestmark = pytest.mark.gpu


_SIGMA1_LAYERS = {
  "coupling": dict(kind="coupling", d=6, hidden=(8,), dec_hidden=(4, 4), even=True, inv=False, n_bins=4, unstack=False),
  "coupling_rqs": dict(kind="coupling_rqs", d=6, hidden=(8,), dec_hidden=(4, 4), even=False, inv=False, n_bins=4, unstack=False),
  "maf": dict(kind="maf", d=6, hidden=(8,), dec_hidden=(4, 4), inv=True),
}


@pytest.mark.parametrize("mode", ["fused", "layered"])
@pytest.mark.parametrize("kind", sorted(_SIGMA1_LAYERS))
def test_reference_style_sigma_one_bias_one(kind, mode):
  """The parameters of the reference's own coupling test, literally (masked_coupling_test.py:78-82): ONE layer whose
  conditioner is make_mlp([8, 2 n]) with w_init = TruncatedNormal(stddev=1.0) and b_init = ones -- per-weight sigma 1,
  biases 1, log-scales of order 10.  (Chains of such layers overflow even float64: the reference tests one layer.)"""
  L = _SIGMA1_LAYERS[kind]
  name = "sigma1_" + kind
  recipes.register(name, dict(layers=[L], act="swish", d_out=L.get("n_keep", 6)), 6, 0)
  out = parity.check_config(name, batch=257, check_grad=True, mode=mode, w_std=1.0, b_val=1.0, literal=True)
  print(name, mode, out)


@pytest.mark.parametrize("mode", ["auto", "layered"])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "cond"])
def test_bias_one(name, mode):
  """Every config with all biases 1 on top of the usual parity weights (fan-in-scaled, per-layer gain 0.85): transforms far
  from the identity, |lp| up to ~400.  Larger gains (0.4 / 0.5 x sqrt(8)) push the chained log-scales past exp overflow
  in float64 already (checked with the oracle), so this is the strongest setting whose reference value exists."""
  out = parity.check_config(name, batch=257, check_grad=True, mode=mode, w_std=0.3, b_val=1.0, fp32_slack=True,
                            **parity.SMALL.get(name, {}))
  print(name, mode, out)


def _rqs_recipe(n_bins, funnel):
  layers = [dict(kind="coupling_rqs", d=6, hidden=(8, 8), dec_hidden=(4, 4), even=True, inv=False, n_bins=n_bins, unstack=False)]
  d_out = 6
  if funnel:
    layers.append(dict(kind="funnel_rqs", d=6, hidden=(8, 8), dec_hidden=(4, 4), n_keep=3, n_bins=n_bins, inv=False))
    d_out = 3
  return dict(layers=layers, act="tanh", d_out=d_out), 6, 0


@pytest.mark.parametrize("mode", ["fused", "layered"])
@pytest.mark.parametrize("n_bins,funnel", [(4, False), (64, False), (5, True)])
def test_rqs_knots_boundaries_and_tails(n_bins, funnel, mode):
  """Spline inputs ON the y-knots (log_prob runs the spline inverse, which bins by the y-knots), one float32 ulp either
  side of them, exactly on range_min / range_max, and in both linear tails (|y| up to 10).  The spline is C1, so value
  and log-det are continuous across knots and boundaries: forward parity must hold everywhere even when float32 and
  float64 pick neighbouring bins; parameter gradients are only piecewise smooth, so the gradient comparison leaves out
  the inputs that sit on (or one ulp from) a knot or a boundary."""
  name = f"rqs_edge_{n_bins}_{int(funnel)}"
  recipe, dim, dx = _rqs_recipe(n_bins, funnel)
  recipes.register(name, recipe, dim, dx)
  # knots of the first layer for a pass-through input of zeros: build the oracle, read theta, place inputs there
  pm, pparams, om, oparams, _, _ = parity.build_pair(name, 4, 0, 0.3)
  rng = np.random.Generator(np.random.PCG64(5))
  base = rng.standard_normal((1, dim))
  # theta of layer 0 depends on the pass-through dims only (mask True = odd dims for even=True ... either way: fix them)
  from oracle import surjectors as S

  captured = {}
  orig = S.RationalQuadraticSpline if hasattr(S, "RationalQuadraticSpline") else None
  real_cls = odx.RationalQuadraticSpline

  class Spy(real_cls):
    def __init__(self, params, lo, hi, *a, **k):
      super().__init__(params, lo, hi, *a, **k)
      captured.setdefault("y_pos", self.y_pos.detach().clone())

  odx.RationalQuadraticSpline = Spy
  if orig is not None:
    S.RationalQuadraticSpline = Spy
  try:
    om.apply(oparams, y=torch.as_tensor(base, dtype=torch.float64))
  finally:
    odx.RationalQuadraticSpline = real_cls
    if orig is not None:
      S.RationalQuadraticSpline = orig
  assert "y_pos" in captured, "oracle spline was not reached"
  mask = np.array([(i % 2 == 0) for i in range(dim)])  # make_alternating_binary_mask(d, even_idx_as_true=True): True = pass-through
  t_dims = np.where(~mask)[0]
  # the oracle, like the reference, builds the spline for every dim and discards the masked ones (masked_coupling.py:78-84)
  y_pos = captured["y_pos"].reshape(dim, n_bins + 1)[t_dims]  # [transformed dims, K + 1] for the pass-through values of `base`
  n_t = len(t_dims)
  rows, smooth = [], []
  def add(vals, ok):
    r = base[0].copy()
    r[t_dims] = vals
    rows.append(r)
    smooth.append(ok)
  for k in range(n_bins + 1):
    knot = y_pos[:, k].numpy().astype(np.float32)
    add(knot.astype(np.float64), False)
    add(np.nextafter(knot, np.float32(10)).astype(np.float64), False)
    add(np.nextafter(knot, np.float32(-10)).astype(np.float64), False)
  for v, ok in ((-3.0, False), (3.0, False), (-3.5, True), (3.5, True), (-10.0, True), (10.0, True), (0.0, True),
                (2.999, True), (-2.999, True)):
    add(np.full(n_t, v), ok)
  for _ in range(64):
    add(rng.uniform(-4.0, 4.0, n_t), True)
  ys = np.stack(rows)
  out = parity.check_config(name, batch=len(ys), check_grad=True, mode=mode, y_override=ys, grad_keep=np.array(smooth),
                            fp32_slack=True)
  # the tails were really exercised
  assert (np.abs(ys[:, t_dims]) > 3.0).any()
  print(name, mode, out)

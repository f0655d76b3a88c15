"""Self-consistency of the CPU oracle where the reference has no numeric pins:
closed-form identities, inverse/forward round trips, log-det against finite differences, and the
committed golden fixtures (tests/golden, regenerate with tests/golden/make_golden.py)."""

import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import configs as ocfg
from oracle import distrax as od
from oracle import hk as ohk
from oracle import surjectors as S
from tests import parity

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f64 = torch.float64


def _apply(fn, seed=0, **batch):
  m = ohk.without_apply_rng(ohk.transform(fn, dtype=f64))
  p = m.init(seed, **batch)
  return m, p


def test_lulinear_at_init_closed_form():  # z = 1.3133617 y, lc = n * 0.27259
  y = torch.randn((7, 5), dtype=f64)
  m, p = _apply(lambda y: S.LULinear(5).inverse_and_likelihood_contribution(y), y=y)
  z, lc = m.apply(p, y=y)
  u = math.log1p(math.e) + 1e-4
  assert torch.allclose(z, u * y, atol=1e-12) and torch.allclose(lc, torch.full((7,), 5 * math.log(u), dtype=f64))


def test_zero_conditioner_makes_affine_layers_identity():
  y = torch.randn((6, 4), dtype=f64)

  def fn(y):
    layer = S.AffineMaskedCoupling(S.make_alternating_binary_mask(4), S.make_mlp([8, 8], w_init=ohk.zeros))
    return layer.inverse_and_likelihood_contribution(y)

  m, p = _apply(fn, y=y)
  z, lc = m.apply(p, y=y)
  assert torch.equal(z, y) and torch.all(lc == 0)


def test_permutation_reverse_twice_is_identity_and_int32():
  y = torch.randn((3, 6), dtype=f64)
  perm = S.Permutation(np.arange(6)[::-1], 1)
  z, lc = perm._inverse_and_likelihood_contribution(y)
  z2, _ = perm._inverse_and_likelihood_contribution(z)
  assert torch.equal(z2, y) and torch.all(lc == 0) and perm.inverse_index().dtype == np.int32


@pytest.mark.parametrize("k", [1, 4, 64])
def test_rqs_round_trip_logdet_and_identity(k):
  g = torch.Generator().manual_seed(k)
  params = torch.randn((50, 3 * k + 1), generator=g, dtype=f64)
  spline = od.RationalQuadraticSpline(params, -3.0, 3.0)
  x = torch.randn(50, generator=g, dtype=f64) * 2.0  # some samples land in the linear tails
  yv, ld_f = spline.forward_and_log_det(x)
  xr, ld_i = spline.inverse_and_log_det(yv)
  assert torch.allclose(xr, x, atol=1e-9) and torch.allclose(ld_f + ld_i, torch.zeros_like(x), atol=1e-9)
  eps = 1e-6
  fd = (spline.forward_and_log_det(x + eps)[0] - spline.forward_and_log_det(x - eps)[0]) / (2 * eps)
  assert torch.allclose(torch.log(fd), ld_f, atol=1e-5)
  assert bool(torch.all(fd > 0))  # monotone
  ident = od.RationalQuadraticSpline(torch.zeros((5, 3 * k + 1), dtype=f64), -3.0, 3.0)  # zero params: identity
  xi = torch.linspace(-2.9, 2.9, 5, dtype=f64)
  yi, ldi = ident.forward_and_log_det(xi)
  assert torch.allclose(yi, xi, atol=1e-9) and torch.allclose(ldi, torch.zeros_like(xi), atol=1e-9)


def test_normal_independent_log_prob():
  v, mu, ls = torch.randn((4, 3), dtype=f64), torch.randn((4, 3), dtype=f64), 0.3 * torch.randn((4, 3), dtype=f64)
  lp = od.Independent(od.Normal(mu, torch.exp(ls))).log_prob(v)
  ref = (-0.5 * ((v - mu) / torch.exp(ls)) ** 2 - ls - 0.5 * math.log(2 * math.pi)).sum(-1)
  assert torch.allclose(lp, ref, atol=1e-12)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_oracle_reproduces_golden(name):
  g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
  pm, pparams, om, oparams, y64, x64 = parity.build_pair(name, g["y"].shape[0], 0, 0.3, **parity.SMALL.get(name, {}))
  assert np.array_equal(y64.numpy(), g["y"])
  for m in oparams:
    for k in oparams[m]:
      assert np.array_equal(oparams[m][k].numpy(), g[f"param:{m}:{k}"]), (m, k)
  ob = dict(y=y64) if x64 is None else dict(y=y64, x=x64)
  lp = om.apply(oparams, **ob).numpy()
  assert np.allclose(lp, g["lp"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_oracle_gradients_against_central_differences(name):
  from oracle import autodiff as oad

  pm, pparams, om, oparams, y64, x64 = parity.build_pair(name, 6, 0, 0.3, **parity.SMALL.get(name, {}))
  ob = dict(y=y64) if x64 is None else dict(y=y64, x=x64)
  _, grads, _ = oad.value_and_grad(om, oparams, **ob)
  rng = np.random.default_rng(0)
  for m in list(oparams)[:4]:
    w = oparams[m]["w"]
    idx = tuple(int(rng.integers(0, s)) for s in w.shape)
    eps = 1e-6
    orig = float(w[idx])
    w[idx] = orig + eps
    lp1 = -om.apply(oparams, **ob).sum()
    w[idx] = orig - eps
    lp0 = -om.apply(oparams, **ob).sum()
    w[idx] = orig
    fd = float(lp1 - lp0) / (2 * eps)
    assert abs(fd - float(grads[m]["w"][idx])) <= 1e-5 * max(1.0, abs(fd)), (m, idx)


def test_kat_file_matches_reference_tests():
  kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
  assert kat["made"]["[2,4,2]"] == [[0.0, 2.0, 10.0]] and kat["conditional_made"]["[2]"] == [[0.0, 3.0, 7.0]]

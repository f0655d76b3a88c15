"""Random chains for executor fuzzing: a chain is a *recipe* (plain data), built once through the product API and once
through the oracle, so any layer sequence the reference accepts can be checked for parity.  TEST INFRASTRUCTURE."""

import numpy as np
import torch


def random_recipe(seed, max_layers=5):
  """(recipe, dim, dim_x): a random chain over every layer kind on the path; widths stay small so the float64 oracle is fast."""
  rng = np.random.Generator(np.random.PCG64(seed))
  dim = int(rng.integers(6, 13))
  dx = int(rng.choice([0, 0, 2]))
  act = str(rng.choice(["swish", "relu", "tanh"]))
  d = dim
  layers = []
  for _ in range(int(rng.integers(2, max_layers + 1))):
    kinds = ["coupling", "coupling_rqs", "maf", "lu", "perm"]
    if d >= 5:
      kinds += ["funnel", "funnel_rqs", "maf_funnel", "slice"]
    k = str(rng.choice(kinds))
    h = int(rng.choice([4, 8, 12]))
    L = dict(kind=k, d=d, hidden=(h, h), dec_hidden=(int(rng.choice([4, 8])),) * 2)
    if k in ("coupling", "coupling_rqs"):
      L["even"] = bool(rng.integers(0, 2))
      L["inv"] = bool(rng.integers(0, 2))
      L["n_bins"] = int(rng.choice([3, 5]))
      L["unstack"] = bool(rng.integers(0, 2)) and k == "coupling"
    elif k in ("funnel", "funnel_rqs", "maf_funnel", "slice"):
      L["n_keep"] = int(rng.integers(2, d - 1))
      L["n_bins"] = int(rng.choice([3, 5]))
      L["inv"] = bool(rng.integers(0, 2))
    elif k == "maf":
      L["inv"] = bool(rng.integers(0, 2))
    elif k == "perm":
      L["perm"] = [int(v) for v in rng.permutation(d)]
    layers.append(L)
    if "n_keep" in L:
      d = L["n_keep"]
  return dict(layers=layers, act=act, d_out=d), dim, dx


def product_flow(recipe):
  from surjectors_b200 import (AffineMaskedAutoregressiveInferenceFunnel, AffineMaskedCouplingInferenceFunnel, Chain, LULinear,
                               MaskedAutoregressive, MaskedCoupling, Permutation, RationalQuadraticSplineMaskedCoupling,
                               RationalQuadraticSplineMaskedCouplingInferenceFunnel, Slice, TransformedDistribution, _sym, hk)
  from surjectors_b200.compat import distrax, jnp, tfd
  from surjectors_b200.nn import MADE, make_mlp
  from surjectors_b200.util import make_alternating_binary_mask, unstack

  act = {"swish": _sym.swish, "relu": _sym.relu, "tanh": _sym.tanh}[recipe["act"]]

  def affine(inv, unstacked):
    def fn(p):
      a, b = unstack(p, -1) if unstacked else jnp.split(p, 2, -1)
      bij = distrax.ScalarAffine(a, jnp.exp(b))
      return distrax.Inverse(bij) if inv else bij
    return fn

  def dec(n, hidden):
    net = make_mlp(list(hidden) + [2 * n], activation=act)

    def fn(z):
      mu, ls = jnp.split(net(z), 2, -1)
      return tfd.Independent(tfd.Normal(mu, jnp.exp(ls)))
    return fn

  def rqs_cond(d, hidden, K):
    return hk.Sequential([make_mlp(list(hidden) + [d * (3 * K + 1)], activation=act), hk.Reshape((d, 3 * K + 1))])

  def flow(method, **kw):
    layers = []
    for L in recipe["layers"]:
      d, k, hid = L["d"], L["kind"], list(L["hidden"])
      if k == "coupling":
        cond = make_mlp(hid + [2 * d], activation=act)
        if L["unstack"]:
          cond = hk.Sequential([cond, hk.Reshape((d, 2))])
        layers.append(MaskedCoupling(make_alternating_binary_mask(d, L["even"]), cond, affine(L["inv"], L["unstack"])))
      elif k == "coupling_rqs":
        layers.append(RationalQuadraticSplineMaskedCoupling(make_alternating_binary_mask(d, L["even"]),
                                                            rqs_cond(d, hid, L["n_bins"]), -3.0, 3.0))
      elif k == "funnel":
        layers.append(AffineMaskedCouplingInferenceFunnel(L["n_keep"], dec(d - L["n_keep"], L["dec_hidden"]),
                                                          make_mlp(hid + [2 * d], activation=act)))
      elif k == "funnel_rqs":
        layers.append(RationalQuadraticSplineMaskedCouplingInferenceFunnel(
          L["n_keep"], dec(d - L["n_keep"], L["dec_hidden"]), rqs_cond(d, hid, L["n_bins"]), -3.0, 3.0))
      elif k == "maf":
        layers.append(MaskedAutoregressive(MADE(d, hid, 2, activation=act), affine(L["inv"], True)))
      elif k == "maf_funnel":
        layers.append(AffineMaskedAutoregressiveInferenceFunnel(L["n_keep"], dec(d - L["n_keep"], L["dec_hidden"]),
                                                                MADE(L["n_keep"], hid, 2, activation=act)))
      elif k == "lu":
        layers.append(LULinear(d))
      elif k == "perm":
        layers.append(Permutation(np.asarray(L["perm"]), 1))
      elif k == "slice":
        layers.append(Slice(L["n_keep"], dec(d - L["n_keep"], L["dec_hidden"])))
    n = recipe["d_out"]
    base = tfd.Independent(tfd.Normal(jnp.zeros(n), jnp.ones(n)), 1)
    return TransformedDistribution(base, Chain(layers))(method=method, **kw)

  return flow


def oracle_flow(recipe):
  from oracle import hk
  from oracle import surjectors as S

  act = {"swish": hk.swish, "relu": hk.relu, "tanh": torch.tanh}[recipe["act"]]

  def affine(inv, unstacked):
    def fn(p):
      a, b = S.unstack(p, -1) if unstacked else S.split(p, 2, -1)
      bij = S.ScalarAffine(a, torch.exp(b))
      return S.Inverse(bij) if inv else bij
    return fn

  def mlp(dims):
    return S.make_mlp(dims, activation=act)

  def dec(n, hidden):
    return S.normal_decoder(n, conditioner_factory=lambda: mlp(list(hidden) + [2 * n]))

  def rqs_cond(d, hidden, K):
    return hk.Sequential([mlp(list(hidden) + [d * (3 * K + 1)]), hk.Reshape((d, 3 * K + 1))])

  def flow(y, x=None):
    layers = []
    for L in recipe["layers"]:
      d, k, hid = L["d"], L["kind"], list(L["hidden"])
      if k == "coupling":
        cond = mlp(hid + [2 * d])
        if L["unstack"]:
          cond = hk.Sequential([cond, hk.Reshape((d, 2))])
        layers.append(S.MaskedCoupling(S.make_alternating_binary_mask(d, L["even"]), cond, affine(L["inv"], L["unstack"])))
      elif k == "coupling_rqs":
        layers.append(S.RationalQuadraticSplineMaskedCoupling(S.make_alternating_binary_mask(d, L["even"]),
                                                              rqs_cond(d, hid, L["n_bins"]), -3.0, 3.0))
      elif k == "funnel":
        layers.append(S.AffineMaskedCouplingInferenceFunnel(L["n_keep"], dec(d - L["n_keep"], L["dec_hidden"]),
                                                            mlp(hid + [2 * d])))
      elif k == "funnel_rqs":
        layers.append(S.RationalQuadraticSplineMaskedCouplingInferenceFunnel(
          L["n_keep"], dec(d - L["n_keep"], L["dec_hidden"]), rqs_cond(d, hid, L["n_bins"]), -3.0, 3.0))
      elif k == "maf":
        layers.append(S.MaskedAutoregressive(S.MADE(d, hid, 2, activation=act), affine(L["inv"], True)))
      elif k == "maf_funnel":
        layers.append(S.AffineMaskedAutoregressiveInferenceFunnel(L["n_keep"], dec(d - L["n_keep"], L["dec_hidden"]),
                                                                  S.MADE(L["n_keep"], hid, 2, activation=act)))
      elif k == "lu":
        layers.append(S.LULinear(d))
      elif k == "perm":
        layers.append(S.Permutation(np.asarray(L["perm"]), 1))
      elif k == "slice":
        layers.append(S.Slice(L["n_keep"], dec(d - L["n_keep"], L["dec_hidden"])))
    td = S.TransformedDistribution(S.standard_normal_base(recipe["d_out"], y.dtype), S.Chain(layers))
    return td.log_prob(y, x)

  return flow


def register(name, recipe, dim, dx):
  """Makes the recipe visible to tests/parity.py as config `name` on both sides."""
  from oracle import configs as ocfg
  from surjectors_b200 import configs as pcfg

  pcfg.CONFIGS[name] = dict(make=lambda **kw: product_flow(recipe), dim=dim, dim_x=dx, batch=256)
  setattr(ocfg, "make_" + name, lambda **kw: oracle_flow(recipe))

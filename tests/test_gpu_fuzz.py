"""Executor fuzzing: random chains over every layer kind on the path (tests/recipes.py), built through the product API
and through the oracle with equal parameters; log-probs (1e-5) and gradients (1e-4) must agree on both program
executors.  Seeds are fixed, so the set of chains is the same on every run."""

import pytest

from tests import parity, recipes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fused", "layered"])
@pytest.mark.parametrize("seed", range(16))
def test_random_chain_parity(seed, mode):
  recipe, dim, dx = recipes.random_recipe(seed)
  recipes.register("fuzz", recipe, dim, dx)
  out = parity.check_config("fuzz", batch=150, check_grad=True, seed=seed, mode=mode)
  print(seed, mode, [L["kind"] for L in recipe["layers"]], out)


@pytest.mark.parametrize("seed", range(100, 108))
def test_random_coupling_chain_on_the_compact_kernel(seed):
  """Chains restricted to the compact kernel's family (affine couplings and coupling funnels, one activation and one
  hidden width per role), with every mask / layout / inner-bijector combination the recipe generator produces."""
  import numpy as np

  from surjectors_b200 import _runtime, configs, fused_chains

  rng = np.random.Generator(np.random.PCG64(seed))
  dim = int(rng.integers(6, 17))
  hc, hd = [(4, 4), (8, 4), (8, 8), (16, 8), (16, 16)][int(rng.integers(0, 5))]
  d, layers = dim, []
  for _ in range(int(rng.integers(2, 6))):
    if d >= 5 and rng.integers(0, 2):
      k = int(rng.integers(2, d - 1))
      layers.append(dict(kind="funnel", d=d, hidden=(hc, hc), dec_hidden=(hd, hd), n_keep=k))
      d = k
    else:
      layers.append(dict(kind="coupling", d=d, hidden=(hc, hc), dec_hidden=(hd, hd), even=bool(rng.integers(0, 2)),
                         inv=bool(rng.integers(0, 2)), unstack=bool(rng.integers(0, 2)), n_bins=3))
  recipe = dict(layers=layers, act=str(rng.choice(["swish", "relu", "tanh"])), d_out=d)
  recipes.register("fuzz", recipe, dim, 0)
  spec = fused_chains.spec_of(configs.CONFIGS["fuzz"]["make"](), dim, 0)
  assert _runtime.get_plan(spec).executor == "compact"
  out = parity.check_config("fuzz", batch=int(rng.integers(1, 600)), check_grad=True, seed=seed)
  print(seed, [L["kind"] for L in layers], recipe["act"], hc, hd, out)

"""Parity of the CUDA path (through the C ABI) against the CPU oracle on seeded inputs."""

import numpy as np
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
@pytest.mark.parametrize("batch", [1, 37, 1000])
def test_fused_small_configs(name, batch):
  out = parity.check_config(name, batch=batch, check_grad=True)
  print(name, batch, out)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_fused_plan_is_used(name):
  from surjectors_b200 import _runtime, fused_chains, configs

  c = configs.CONFIGS[name]
  spec = fused_chains.spec_of(c["make"](), c["dim"], c["dim_x"])
  assert _runtime.get_plan(spec).executor == "gen1"


@pytest.mark.parametrize("mode", ["fused", "layered"])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_small_configs_both_executors(name, mode):
  out = parity.check_config(name, batch=300, check_grad=True, mode=mode)
  print(name, mode, out)


@pytest.mark.parametrize("mode", ["fused", "layered"])
@pytest.mark.parametrize("name", ["cfg4", "cfg5"])
def test_scaled_wide_configs(name, mode):
  """cfg4 (spline couplings + spline funnel) and cfg5 (conditional MAF + LULinear + Permutation)
  at oracle-sized shapes, through both executors."""
  out = parity.check_config(name, batch=200, check_grad=True, mode=mode, **parity.SMALL[name])
  print(name, mode, out)


@pytest.mark.parametrize("mode", ["fused", "layered"])
def test_mlp_inference_funnel(mode):
  """MLPInferenceFunnel (mlp.py:40-56) between LULinear layers, as in docs/notebooks/dimension_reduction.ipynb:397-420."""
  out = parity.check_config("mlpf", batch=300, check_grad=True, mode=mode)
  print("mlpf", mode, out)


def test_user_built_generated_kernel_parity(tmp_path):
  """fused_chains.build_fused_kernel: a non-BASELINE chain gets its own generated straight-line kernel (nvcc at the
  user's request), which must agree with the oracle like the built-in ones."""
  import torch

  from surjectors_b200 import configs, fused_chains, hk

  scale = dict(dim=12, cond_hidden=(6, 6), dec_hidden=(4, 4))
  m = hk.transform(configs.make_cfg2(**scale))
  fused_chains.build_fused_kernel(m, dict(y=torch.zeros((2, 12))), cache_dir=str(tmp_path))
  out = parity.check_config("cfg2", batch=500, check_grad=True, mode="gen1", **scale)
  print("user-built", out)

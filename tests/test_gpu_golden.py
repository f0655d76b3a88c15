"""The CUDA path against the committed golden fixtures (tests/golden/*.npz): log-probs within
1e-5 relative, loss gradients within 1e-4, through every executor that accepts the chain."""

import os

import numpy as np
import pytest
import torch

from surjectors_b200 import execution_mode, value_and_grad
from tests import parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("mode", ["auto", "fused", "layered"])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_golden(name, mode):
  g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
  pm, pparams, _, _, _, _ = parity.build_pair(name, 2, 0, 0.3, **parity.SMALL.get(name, {}))
  for m in pparams:
    for k in pparams[m]:
      pparams[m][k].copy_(torch.as_tensor(g[f"param:{m}:{k}"], dtype=torch.float32))
  dev = torch.device("cuda:0")
  params = parity.to_device(pparams, dev)
  batch = dict(y=torch.as_tensor(g["y"], dtype=torch.float32, device=dev))
  if "x" in g.files:
    batch["x"] = torch.as_tensor(g["x"], dtype=torch.float32, device=dev)
  with execution_mode(mode):
    lp = pm.apply(params, None, method="log_prob", **batch)
    assert parity.lp_error(lp.cpu().numpy(), g["lp"]) <= parity.LP_RTOL
    loss, grads = value_and_grad(lambda p: -torch.sum(pm.apply(p, None, method="log_prob", **batch)))(params)
  assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
  top = max(float(np.abs(g[f"grad:{m}:{k}"]).max()) for m in grads for k in grads[m])
  for m in grads:
    for k in grads[m]:
      r = g[f"grad:{m}:{k}"]
      scale = max(float(np.abs(r).max()), parity.GRAD_FLOOR * top)  # true relative bound, see tests/parity.py
      assert float(np.abs(grads[m][k].cpu().double().numpy() - r).max()) <= parity.GRAD_RTOL * scale, (m, k)
      if "masked_linear" in m and k == "w":  # masked MADE weights get exactly zero gradient -- and ONLY those do
        got = grads[m][k].cpu().numpy()
        assert np.all(got[r == 0] == 0), (m, k, "a structurally masked weight has a non-zero gradient")

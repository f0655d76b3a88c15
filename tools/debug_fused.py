"""Fused conditioner + spline kernel against the unfused per-op path (same float32 inputs) and the float64 oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surjectors_b200 import _abi, execution_mode, value_and_grad
from tests import parity

lib = _abi.lib()
dev = torch.device("cuda:0")
grad = "--grad" in sys.argv
cases = [("cfg4", 300, parity.SMALL["cfg4"], 0.3), ("cfg4", 192, {}, 0.15)]
if "--small" in sys.argv:
  cases = cases[:1]
for name, batch, scale, w_std in cases:
  pm, pparams, om, oparams, y64, x64 = parity.build_pair(name, batch, 0, w_std, **scale)
  params = parity.to_device(pparams, dev)
  y = y64.float().to(dev)
  ref = om.apply(oparams, y=y64).numpy()
  res = {}
  with execution_mode("layered"):
    for fus in (0, 1):
      lib.surj_set_wide_fusion(fus)
      t0 = time.time()
      lp = pm.apply(params, None, method="log_prob", y=y)
      torch.cuda.synchronize()
      res[fus] = lp.cpu().double().numpy()
      print(name, scale and "small" or "full", "fusion", fus, "lp err vs oracle", parity.lp_error(res[fus], ref), "finite", np.isfinite(res[fus]).all(), f"{time.time() - t0:.2f}s", flush=True)
      if grad:
        loss, g = value_and_grad(lambda p: -torch.sum(pm.apply(p, None, method="log_prob", y=y)))(params)
        torch.cuda.synchronize()
        res[("g", fus)] = g.flat.cpu().double().numpy()
        res[("l", fus)] = float(loss)
    print("  fused vs unfused: max |dlp| / max(1,|lp|)", float(np.max(np.abs(res[1] - res[0]) / np.maximum(1, np.abs(res[0])))))
    if grad:
      g0, g1 = res[("g", 0)], res[("g", 1)]
      print("  grad fused vs unfused: max |dg| / max |g|", float(np.abs(g1 - g0).max() / np.abs(g0).max()), "loss", res[("l", 0)], res[("l", 1)])
lib.surj_set_wide_fusion(1)

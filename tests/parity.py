"""Shared parity harness: the same model built twice -- through the product API
(surjectors_b200, CUDA) and through the CPU oracle (oracle/, float64) -- with identical
parameters and inputs.  TEST INFRASTRUCTURE: the only place product and oracle meet.

Tolerances (BASELINE.json north_star): log-probs within 1e-5 relative, gradients within 1e-4
relative, FP32.  "Relative" is measured against max(1, |reference|) for log-probs (per sample) and
against max(1e-3.., max|reference|) per parameter tensor for gradients, i.e. relative to the scale of
the tensor, so exact zeros (masked MADE weights) must come out as exact zeros.
"""

import numpy as np
import torch

from oracle import autodiff as oad
from oracle import configs as ocfg
from oracle import hk as ohk
from surjectors_b200 import configs as pcfg
from surjectors_b200 import hk as phk
from surjectors_b200 import execution_mode, value_and_grad

LP_RTOL = 1e-5
GRAD_RTOL = 1e-4

SMALL = {  # scaled-down shapes of the two wide configs for oracle-speed parity
  "cfg4": dict(dim=12, n_keep=4, hidden=(16, 16), n_bins=5, n_pre=2, n_post=2),
  "cfg5": dict(dim=8, hidden=(16, 16), n_layers=3),
}


def synthetic(dim, dim_x, batch, seed, scale=1.0):
  rng = np.random.Generator(np.random.PCG64(seed))
  y = rng.standard_normal((batch, dim)) * scale
  x = rng.standard_normal((batch, dim_x)) if dim_x else None
  return y, x


def build_pair(name, batch, seed=0, w_std=0.3, **scale):
  """(product model, product params on CPU, oracle model, oracle params f64, y, x) with equal params."""
  c = pcfg.CONFIGS[name]
  kw = dict(scale)
  dim = kw.get("dim", c["dim"])
  dim_x = c["dim_x"]
  y, x = synthetic(dim, dim_x, batch, 1000 + seed)
  y64 = torch.as_tensor(y, dtype=torch.float64)
  x64 = None if x is None else torch.as_tensor(x, dtype=torch.float64)
  om = ohk.without_apply_rng(ohk.transform(getattr(ocfg, "make_" + name)(**kw), dtype=torch.float64))
  ob = dict(y=y64) if x64 is None else dict(y=y64, x=x64)
  oparams = ocfg.randomize(om.init(2000 + seed, **ob), 2000 + seed, w_std=w_std)
  pm = phk.transform(c["make"](**kw))
  pb = dict(y=y64[:2].float()) if x64 is None else dict(y=y64[:2].float(), x=x64[:2].float())
  pparams = pm.init(2000 + seed, method="log_prob", **pb)
  copy_params(oparams, pparams)
  return pm, pparams, om, oparams, y64, x64


def copy_params(oparams, pparams):
  assert set(oparams) == set(pparams), (sorted(oparams), sorted(pparams))
  for m in oparams:
    assert set(oparams[m]) == set(pparams[m]), (m, sorted(oparams[m]), sorted(pparams[m]))
    for k in oparams[m]:
      assert tuple(oparams[m][k].shape) == tuple(pparams[m][k].shape), (m, k)
      pparams[m][k].copy_(oparams[m][k].to(torch.float32))


def to_device(pparams, device):
  return phk.tree_from_flat(pparams.flat.to(device), pparams.layout)


def lp_error(lp, ref):
  lp = np.asarray(lp, dtype=np.float64)
  ref = np.asarray(ref, dtype=np.float64)
  return float(np.max(np.abs(lp - ref) / np.maximum(1.0, np.abs(ref)))) if ref.size else 0.0


def grad_errors(grads, ref):
  out = {}
  for m in ref:
    for k in ref[m]:
      g = grads[m][k].detach().cpu().double().numpy()
      r = ref[m][k].detach().cpu().double().numpy()
      scale = max(1.0, float(np.max(np.abs(r)))) if r.size else 1.0
      out[f"{m}/{k}"] = float(np.max(np.abs(g - r))) / scale if r.size else 0.0
  return out


def differentiable_samples(om, oparams, ob, eps=2e-5):
  """Mask of samples none of whose ReLU pre-activations lies within `eps` of the kink.  At a kink
  the gradient is discontinuous and float32 rounding decides the side (the float64 oracle holds
  pre-activations as small as 5e-10 in the 16-layer MADE config), so gradient parity is only defined
  away from it; samples are independent, so the others are compared as usual."""
  ohk.relu_probe = []
  try:
    om.apply(oparams, **ob)
    probes = ohk.relu_probe
  finally:
    ohk.relu_probe = None
  if not probes:
    return None
  return torch.stack(probes).min(dim=0).values >= eps


def check_config(name, batch=256, device="cuda:0", check_grad=True, seed=0, w_std=0.3, mode="auto", **scale):
  with execution_mode(mode):
    return _check_config(name, batch, device, check_grad, seed, w_std, **scale)


def _check_config(name, batch, device, check_grad, seed, w_std, **scale):
  pm, pparams, om, oparams, y64, x64 = build_pair(name, batch, seed, w_std, **scale)
  dev = torch.device(device)
  params = to_device(pparams, dev)
  y = y64.float().to(dev)
  x = None if x64 is None else x64.float().to(dev)
  pb = dict(y=y) if x is None else dict(y=y, x=x)
  ob = dict(y=y64) if x64 is None else dict(y=y64, x=x64)
  lp = pm.apply(params, None, method="log_prob", **pb)
  ref = om.apply(oparams, **ob)
  e = lp_error(lp.cpu().numpy(), ref.numpy())
  assert np.isfinite(ref.numpy()).all()
  assert e <= LP_RTOL, f"{name}: log_prob relative error {e:.3e} > {LP_RTOL}"
  out = {"lp_err": e}
  if check_grad:
    keep = differentiable_samples(om, oparams, ob)
    if keep is not None and not bool(keep.all()):
      out["kink_samples_dropped"] = int((~keep).sum())
      ob = {k: v[keep] for k, v in ob.items()}
      pb = {k: v[keep.to(v.device)] for k, v in pb.items()}

    def loss_fn(p):
      return -torch.sum(pm.apply(p, None, method="log_prob", **pb))

    loss, grads = value_and_grad(loss_fn)(params)
    oloss, ograds, _ = oad.value_and_grad(om, oparams, **ob)
    le = abs(float(loss) - float(oloss)) / max(1.0, abs(float(oloss)))
    assert le <= 1e-5, f"{name}: loss relative error {le:.3e}"
    ge = grad_errors(grads, ograds)
    worst = max(ge, key=ge.get)
    assert ge[worst] <= GRAD_RTOL, f"{name}: gradient error {ge[worst]:.3e} at {worst}"
    out["grad_err"] = ge[worst]
    out["grad_worst"] = worst
  return out

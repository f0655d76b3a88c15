"""Parity of the compact looped fused kernel (csrc/compact.cuh) against the oracle, through the C ABI: affine
MaskedCoupling / MaskedCouplingInferenceFunnel chains with two-hidden-layer MLPs.  The kernel owns two samples per
thread and tiles the batch per CTA, so batch sizes around those boundaries are covered explicitly."""

import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch", [1, 2, 63, 449, 1000, 4097])
def test_cfg2_compact_batches(batch):
  out = parity.check_config("cfg2", batch=batch, check_grad=batch > 1, mode="compact")
  print("cfg2 compact", batch, out)


@pytest.mark.parametrize(
  "scale",
  [
    dict(),                                                             # tanh, hidden 16 / 8, dim 12
    dict(activation="relu", hidden=(8, 8), dec_hidden=(8, 8)),
    dict(activation="swish", hidden=(16, 16), dec_hidden=(16, 16), dim=16),
    dict(activation="tanh", hidden=(4, 4), dec_hidden=(4, 4), dim=9),
  ],
)
def test_user_chain_runs_on_the_compact_kernel(scale):
  """cpl: Inverse(ScalarAffine) inner, interleaved (unstack) conditioner output, plain couplings between funnels.
  No generated kernel exists for these chains, so AUTO must pick the compact kernel."""
  from surjectors_b200 import _runtime, configs, fused_chains

  c = configs.CONFIGS["cpl"]
  spec = fused_chains.spec_of(c["make"](**scale), scale.get("dim", c["dim"]), 0)
  assert _runtime.get_plan(spec).executor == "compact"
  out = parity.check_config("cpl", batch=700, check_grad=True, **scale)
  print("cpl", scale, out)


def test_outside_the_family_is_refused():
  """Forcing the compact kernel on a chain it does not cover is an error, not a fallback."""
  from surjectors_b200 import _runtime, configs, execution_mode, fused_chains

  c = configs.CONFIGS["cfg3"]  # MADE conditioners
  spec = fused_chains.spec_of(c["make"](), c["dim"], 0)
  with execution_mode("compact"):
    with pytest.raises(Exception, match="compact"):
      _runtime.get_plan(spec)


def test_compact_z_gy_and_cotangents():
  """z, lp, the input cotangent g_y and the parameter gradient under non-trivial cotangents (g_lp, g_z) agree with the
  interpreter executor on the same inputs (both are parity-tested against the oracle separately)."""
  from surjectors_b200 import _runtime, configs, execution_mode, fused_chains

  c = configs.CONFIGS["cfg2"]
  spec = fused_chains.spec_of(c["make"](), c["dim"], 0)
  dev = torch.device("cuda:0")
  g = torch.Generator().manual_seed(11)
  B = 333
  y = torch.randn((B, 20), generator=g).to(dev)
  res = {}
  for mode in ("compact", "fused"):
    with execution_mode(mode):
      plan = _runtime.get_plan(spec)
    assert plan.executor == mode
    flat = (0.1 * torch.randn(plan.n_params, generator=torch.Generator().manual_seed(5))).to(dev)
    z, lp = plan.fwd(y, None, flat)
    g_lp = torch.linspace(-1.0, 2.0, B, device=dev)
    g_z = torch.cos(z)
    gp, gy, _, lp2 = plan.bwd(y, None, flat, g_lp=g_lp, g_z=g_z, want_gy=True, want_lp=True)
    torch.cuda.synchronize()
    assert torch.equal(lp, lp2)
    res[mode] = [t.cpu() for t in (z, lp, gy, gp)]
  for a, b, what in zip(res["compact"], res["fused"], ("z", "lp", "g_y", "g_params")):
    scale = max(1.0, float(b.abs().max()))
    assert float((a - b).abs().max()) <= 2e-5 * scale, what


def test_compact_matches_generated_kernel_at_full_size():
  """BASELINE batch and beyond (several tiles per CTA, ragged tail): log-probs and the summed gradient of the compact
  kernel agree with the generated straight-line kernel, which is oracle-checked at oracle-sized batches."""
  from surjectors_b200 import _runtime, configs, execution_mode, fused_chains

  c = configs.CONFIGS["cfg2"]
  spec = fused_chains.spec_of(c["make"](), c["dim"], 0)
  dev = torch.device("cuda:0")
  for B in (65536, 200003):
    y = torch.randn((B, 20), generator=torch.Generator().manual_seed(B)).to(dev)
    out = {}
    for mode in ("compact", "gen1"):
      with execution_mode(mode):
        plan = _runtime.get_plan(spec)
      assert plan.executor == mode
      flat = (0.1 * torch.randn(plan.n_params, generator=torch.Generator().manual_seed(5))).to(dev)
      _, lp = plan.fwd(y, None, flat, want_z=False)
      gp, _, _, lp2 = plan.bwd(y, None, flat, g_scale=-1.0 / B, want_lp=True)
      torch.cuda.synchronize()
      assert torch.equal(lp, lp2)
      out[mode] = (lp.cpu().double(), gp.cpu().double())
    lp_a, gp_a = out["compact"]
    lp_b, gp_b = out["gen1"]
    assert float(((lp_a - lp_b).abs() / lp_b.abs().clamp(min=1.0)).max()) <= 1e-5
    assert float((gp_a - gp_b).abs().max()) <= 1e-4 * float(gp_b.abs().max())

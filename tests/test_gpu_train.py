"""The training step (fused log_prob+VJP kernel, gradient reduction, fused Adam) against the oracle's
value_and_grad + optax-style adam over several steps, for a generated-kernel config, the interpreter
kernel and the per-op executor; plus the host-fed path and CUDA-graph replay."""

import numpy as np
import pytest
import torch

from oracle import autodiff as oad
from surjectors_b200 import execution_mode, hk
from surjectors_b200.train import TrainStep
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,mode", [("cfg2", "auto"), ("cfg2", "fused"), ("cfg3", "layered"), ("cfg5", "fused")])
def test_train_steps_match_oracle_adam(name, mode):
  B, lr = 96, 1e-2  # a large lr makes the updates visible in float32
  pm, pparams, om, oparams, y64, x64 = parity.build_pair(name, B, 0, 0.3, **parity.SMALL.get(name, {}))
  dev = torch.device("cuda:0")
  params = parity.to_device(pparams, dev)
  batch = dict(y=y64.float().to(dev))
  ob = dict(y=y64)
  if x64 is not None:
    batch["x"] = x64.float().to(dev)
    ob["x"] = x64
  with execution_mode(mode):
    ts = TrainStep(pm, params, batch, lr=lr)
  state = oad.adam_init(oparams)
  for step in range(4):
    loss = ts.step(batch["y"], batch.get("x"))
    oloss, grads, _ = oad.value_and_grad(om, oparams, **ob)
    oparams, state = oad.adam_update(oparams, grads, state, lr=lr)
    assert abs(float(loss) - float(oloss)) <= 2e-5 * max(1.0, abs(float(oloss))), (step, float(loss), float(oloss))
  assert int(ts.opt.count.item()) == 4
  for r in params.layout.refs:
    got = params[r.module][r.name].cpu().double().numpy()
    ref = oparams[r.module][r.name].numpy()
    assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), (r.module, r.name)


def test_host_fed_steps_and_graph_replay_agree_with_resident_steps():
  B = 4096
  pm, pparams, _, _, y64, _ = parity.build_pair("cfg2", B, 0, 0.3)
  dev = torch.device("cuda:0")
  y = y64.float()
  runs = []
  for kind in ("resident", "host", "graph"):
    params = hk.tree_from_flat(pparams.flat.clone().to(dev), pparams.layout)
    ts = TrainStep(pm, params, dict(y=y[:2].to(dev)), lr=1e-3)
    yd = y.to(dev)
    if kind == "resident":
      for _ in range(3):
        ts.step(yd)
    elif kind == "host":
      yh = y.pin_memory()
      out = torch.zeros(3).pin_memory()
      for i in range(3):
        ts.step_host(yh, loss_out=out[i : i + 1])
      ts.drain()
      torch.cuda.synchronize()
      assert bool(torch.isfinite(out).all())
    else:
      ts.step(yd)  # warm-up outside capture (allocations), then rewind the state
      ts.flat.copy_(pparams.flat.to(dev)); ts.opt.mu.zero_(); ts.opt.nu.zero_(); ts.opt.count.zero_()
      s = torch.cuda.Stream()
      with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
          ts.step(yd)
        ts.flat.copy_(pparams.flat.to(dev)); ts.opt.mu.zero_(); ts.opt.nu.zero_(); ts.opt.count.zero_()
        for _ in range(3):
          g.replay()
      s.synchronize()
    torch.cuda.synchronize()
    runs.append((ts.flat.clone(), float(ts.loss.item()), int(ts.opt.count.item())))
  for flat, loss, count in runs[1:]:
    assert count == 3 and torch.equal(flat, runs[0][0]) and loss == runs[0][1]

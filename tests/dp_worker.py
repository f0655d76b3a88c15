"""Worker of tests/test_gpu_dp.py: one process per GPU (spawned), NCCL rendezvous on 127.0.0.1.  Every case runs the
data-parallel train step on contiguous batch shards and the SAME global batch through a single-rank TrainStep on this
rank's GPU, and compares.  TEST INFRASTRUCTURE."""

import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def _case(name, exchange, mode, reduction, batch, steps, dev, group, rank, world):
  import numpy as np
  import torch
  import torch.distributed as dist

  from surjectors_b200 import configs, execution_mode, hk, parallel
  from surjectors_b200.train import TrainStep
  from tests import parity

  os.environ["SURJ_B200_EXCHANGE"] = exchange
  c = configs.CONFIGS[name]
  scale = parity.SMALL.get(name, {})
  dim = scale.get("dim", c["dim"])
  rng = np.random.Generator(np.random.PCG64(77))  # same data on every rank; each takes its shard
  ys = [torch.from_numpy(rng.standard_normal((batch, dim)).astype(np.float32)).to(dev) for _ in range(steps)]
  xs = [torch.from_numpy(rng.standard_normal((batch, c["dim_x"])).astype(np.float32)).to(dev) if c["dim_x"] else None
        for _ in range(steps)]
  model = hk.transform(c["make"](**scale))
  ex = dict(y=ys[0]) if xs[0] is None else dict(y=ys[0], x=xs[0])
  lo, hi = parallel.shard_bounds(batch, rank, world)
  with execution_mode(mode):
    p_dp = model.init(11, method="log_prob", **ex)
    parity.copy_params(parity.ocfg.randomize({m: {k: v.cpu().double() for k, v in b.items()} for m, b in p_dp.items()}, 3, w_std=0.2), p_dp)
    p_one = hk.tree_from_flat(p_dp.flat.clone(), p_dp.layout)
    dp = TrainStep(model, p_dp, ex, lr=1e-3, reduction=reduction, process_group=group)
    one = TrainStep(model, p_one, ex, lr=1e-3, reduction=reduction)
    assert (dp.peer is not None) == (exchange == "peer" and dp.plan.is_fused), (exchange, dp.plan.executor)
    l_dp, l_one = [], []
    for y, x in zip(ys, xs):
      l_dp.append(float(dp.step(y[lo:hi], None if x is None else x[lo:hi]).item()))
      l_one.append(float(one.step(y, x).item()))
  torch.cuda.synchronize(dev)
  dp.check()
  gathered = [torch.zeros_like(dp.flat) for _ in range(world)]
  dist.all_gather(gathered, dp.flat, group=group)
  identical = all(torch.equal(gathered[0], g) for g in gathered)
  pscale = float(one.flat.abs().max())
  perr = float((dp.flat - one.flat).abs().max()) / max(pscale, 1e-30)
  lerr = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(l_dp, l_one))
  res = dict(case=f"{name}/{exchange}/{mode}/{reduction}/B{batch}", executor=dp.plan.executor, replicas_identical=identical,
             param_rel_err=perr, loss_rel_err=lerr, peer=dp.peer is not None)
  assert identical, res
  assert lerr <= 1e-5, res
  assert perr <= 1e-5, res  # steps * lr = 4e-3 of accumulated updates; float32 summation order differs between 1 and N ranks
  return res


def _delayed_rank(dev, group, rank, world):
  """A rank that falls behind by more than the exchange timeout: the waiting rank must NOT apply a stale update, must
  leave its step counter alone, and must raise on its next step."""
  import numpy as np
  import torch

  from surjectors_b200 import _abi, configs, hk
  from surjectors_b200.train import TrainStep

  os.environ["SURJ_B200_EXCHANGE"] = "peer"
  c = configs.CONFIGS["cfg2"]
  rng = np.random.Generator(np.random.PCG64(5 + rank))
  y = torch.from_numpy(rng.standard_normal((1024, c["dim"])).astype(np.float32)).to(dev)
  model = hk.transform(c["make"]())
  params = model.init(3, method="log_prob", y=y)
  ts = TrainStep(model, params, dict(y=y), lr=1e-3, process_group=group)
  assert ts.peer is not None
  ts.peer.set_timeout_ms(300)
  for _ in range(3):
    ts.step(y)
  torch.cuda.synchronize(dev)
  ts.check()
  before = ts.flat.clone()
  count_before = int(ts.opt.count.item())
  raised = False
  if rank == 1:
    time.sleep(2.0)  # rank 0's reduction kernel waits 300 ms for this rank's tagged gradient, then gives up
  try:
    for _ in range(4):
      ts.step(y)
      torch.cuda.synchronize(dev)
      time.sleep(0.05)
    ts.check()
  except _abi.SurjError:
    raised = True
  torch.cuda.synchronize(dev)
  res = dict(case="delayed_rank", rank=rank, raised=raised, err_flag=ts.peer.error(), count_before=count_before,
             count_after=int(ts.opt.count.item()), params_unchanged=bool(torch.equal(before, ts.flat)))
  if rank == 0:
    assert raised and res["err_flag"] == 1, res
    assert res["params_unchanged"] and res["count_after"] == count_before, res
  else:
    assert raised, res  # rank 0 stopped sending: this rank times out one step later
  return res


def main(rank, world, port, out_path):
  import json

  import torch
  import torch.distributed as dist

  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  results, ok = [], True
  try:
    g = dist.group.WORLD
    for args in (("cfg2", "peer", "auto", "sum", 4096, 4), ("cfg2", "nccl", "auto", "sum", 4096, 4),
                 ("cfg2", "peer", "auto", "mean", 1001, 4), ("cfg3", "peer", "fused", "sum", 999, 3),
                 ("cfg2", "nccl", "layered", "mean", 1001, 3), ("cfg4", "nccl", "layered", "sum", 600, 3),
                 ("cfg5", "nccl", "layered", "sum", 515, 3)):
      results.append(_case(*args, dev, g, rank, world))
    results.append(_delayed_rank(dev, g, rank, world))
  except Exception:  # noqa: BLE001
    ok = False
    results.append(dict(error=traceback.format_exc()))
  with open(f"{out_path}.{rank}", "w") as f:
    json.dump(dict(ok=ok, results=results), f)
  sys.stdout.flush()
  os._exit(0 if ok else 1)  # the delayed-rank case leaves the peer exchange dead: no collective teardown

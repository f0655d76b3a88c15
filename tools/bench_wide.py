"""Throughput of the two wide BASELINE configs (cfg4, cfg5) through the per-op executor:
log_prob samples/s and train-step samples/s, with the needed-FLOP rate beside them."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from surjectors_b200 import _cost, configs, hk
from surjectors_b200.train import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg4")
ap.add_argument("--batch", type=int, default=1 << 18)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
c = configs.CONFIGS[a.config]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
y = torch.randn((a.batch, c["dim"]), generator=g).to(dev)
x = torch.randn((a.batch, c["dim_x"]), generator=g).to(dev) if c["dim_x"] else None
b = dict(y=y) if x is None else dict(y=y, x=x)
model = hk.transform(c["make"]())
params = model.init(2000, method="log_prob", **{k: v[:2] for k, v in b.items()})
ts = TrainStep(model, params, b, lr=1e-4)
need, dense = _cost.forward_flops(ts.spec)


def timed(fn):
  fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(a.reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / a.reps


ms_f = timed(lambda: model.apply(params, None, method="log_prob", **b))
ms_t = timed(lambda: ts.step(y, x))
out = {
  "config": a.config, "batch": a.batch, "executor": ts.plan.executor, "tc": os.environ.get("SURJ_B200_TC", "1"),
  "params": ts.flat.numel(), "needed_flops_per_sample_fwd": need, "dense_flops_per_sample_fwd": dense,
  "log_prob_ms": ms_f, "log_prob_samples_per_s": a.batch / (ms_f * 1e-3), "log_prob_needed_tflops": need * a.batch / (ms_f * 1e-3) / 1e12,
  "train_ms": ms_t, "train_samples_per_s": a.batch / (ms_t * 1e-3), "train_needed_tflops": 3 * need * a.batch / (ms_t * 1e-3) / 1e12,
  "loss": float(ts.loss.item()),
}
print(json.dumps(out))

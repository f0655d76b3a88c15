"""Small driver for ncu: a few train steps of one BASELINE config through the public TrainStep API."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from surjectors_b200 import configs, hk
from surjectors_b200.train import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--fwd", action="store_true")
a = ap.parse_args()
c = configs.CONFIGS[a.config]
dev = torch.device("cuda:0")
rng = np.random.Generator(np.random.PCG64(1))
ys = [torch.from_numpy(rng.standard_normal((a.batch, c["dim"]), dtype=np.float32)).to(dev) for _ in range(a.steps)]
xs = [torch.from_numpy(rng.standard_normal((a.batch, c["dim_x"]), dtype=np.float32)).to(dev) if c["dim_x"] else None
      for _ in range(a.steps)]
model = hk.transform(c["make"]())
ex = dict(y=ys[0]) if xs[0] is None else dict(y=ys[0], x=xs[0])
params = model.init(2002, method="log_prob", **ex)
ts = TrainStep(model, params, ex)
for i in range(a.steps):
  if a.fwd:
    lp = model.apply(params, None, method="log_prob", **(dict(y=ys[i]) if xs[i] is None else dict(y=ys[i], x=xs[i])))
  else:
    ts.step(ys[i], xs[i])
torch.cuda.synchronize()
print("ok", float(ts.loss.item()))

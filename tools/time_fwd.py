"""log_prob-only timing of the small BASELINE configs at one batch size (CUDA graph of N calls, CUDA events)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from surjectors_b200 import _runtime, configs, fused_chains

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--calls", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda:0")
for name in ("cfg1", "cfg2", "cfg3", "cpl"):
  c = configs.CONFIGS[name]
  spec = fused_chains.spec_of(c["make"](), c["dim"], c["dim_x"])
  plan = _runtime.get_plan(spec)
  rng = np.random.Generator(np.random.PCG64(1))
  ys = [torch.from_numpy(rng.standard_normal((a.batch, c["dim"]), dtype=np.float32)).to(dev) for _ in range(8)]
  flat = torch.from_numpy(0.1 * rng.standard_normal(plan.n_params, dtype=np.float32)).to(dev)
  lp = torch.empty(a.batch, device=dev)
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    for i in range(3):
      plan.fwd(ys[i % 8], None, flat, want_z=False, lp_out=lp)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
      for i in range(a.calls):
        plan.fwd(ys[i % 8], None, flat, want_z=False, lp_out=lp)
    g.replay()
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10):
      g.replay()
    e1.record(s)
    s.synchronize()
  us = e0.elapsed_time(e1) * 1e3 / (10 * a.calls)
  print(f"{name} executor={plan.executor} batch={a.batch} log_prob {us:.2f} us/call {a.batch / us * 1e-3:.2f} G samples/s")

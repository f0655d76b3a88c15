"""Is the host-fed train step (TrainStep.step_host) bound by the CPU enqueue, or by the H2D copy?  Prints the enqueue time
per step, the end-to-end time per step and the raw pinned H2D time of one batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surjectors_b200 import configs, hk
from surjectors_b200.train import TrainStep
c = configs.CONFIGS["cfg2"]; B = 65536
dev = torch.device("cuda:0")
host = torch.randn((8, B, 20)).pin_memory()
model = hk.transform(c["make"]())
params = model.init(1, method="log_prob", y=host[0, :2].to(dev))
ts = TrainStep(model, params, dict(y=host[0].to(dev)))
for i in range(20): ts.step_host(host[i % 8])
ts.drain(); torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for i in range(n): ts.step_host(host[i % 8])
t1 = time.perf_counter()
ts.drain(); torch.cuda.synchronize()
t2 = time.perf_counter()
print("enqueue us/step", (t1 - t0) / n * 1e6, "total us/step", (t2 - t0) / n * 1e6)
# raw H2D of the same buffers
d = torch.empty((B, 20), device=dev)
s = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
  for i in range(10): d.copy_(host[i % 8], non_blocking=True)
  e0.record(s)
  for i in range(100): d.copy_(host[i % 8], non_blocking=True)
  e1.record(s)
s.synchronize()
print("raw H2D us/copy", e0.elapsed_time(e1) * 10, "GB/s", B * 20 * 4 / (e0.elapsed_time(e1) * 10) / 1e3)

"""Per-kernel-kind time of one log_prob call and one train step of a wide config (surj_profile_begin / _end)."""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surjectors_b200 import _abi, configs, hk
from surjectors_b200.train import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg4")
ap.add_argument("--batch", type=int, default=1 << 20)
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
lib = _abi.lib()
lp = torch.empty(a.batch, device=dev)


def prof(fn, what):
  fn(); torch.cuda.synchronize()
  lib.surj_profile_begin()
  fn()
  buf = C.create_string_buffer(1 << 16)
  lib.surj_profile_end(buf, len(buf))
  k = json.loads(buf.value.decode())
  tot = sum(v["ms"] for v in k.values())
  print(f"== {what}: {tot:.1f} ms")
  for name, v in sorted(k.items(), key=lambda kv: -kv[1]["ms"]):
    if v["launches"]:
      tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
      print(f"  {name:36s} {v['launches']:5d} launches {v['ms']:9.2f} ms  {100 * v['ms'] / tot:5.1f} %  {tf:7.1f} needed TFLOP/s")


prof(lambda: ts.plan.fwd(y, x, ts.flat, want_z=False, lp_out=lp), f"{a.config} log_prob B={a.batch}")
prof(lambda: ts.step(y, x), f"{a.config} train step B={a.batch}")

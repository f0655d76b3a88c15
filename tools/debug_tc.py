import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity
from oracle import autodiff as oad
from surjectors_b200 import value_and_grad
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
pm, pparams, om, oparams, y64, x64 = parity.build_pair(name, 192, 0, 0.15)
dev = torch.device("cuda:0")
params = parity.to_device(pparams, dev)
pb = dict(y=y64.float().to(dev)); ob = dict(y=y64)
if x64 is not None: pb["x"] = x64.float().to(dev); ob["x"] = x64
lp = pm.apply(params, None, method="log_prob", **pb)
ref = om.apply(oparams, **ob)
print("TC mode", os.environ.get("SURJ_B200_TC", "1"), "lp err", parity.lp_error(lp.cpu().numpy(), ref.numpy()))
loss, grads = value_and_grad(lambda p: -torch.sum(pm.apply(p, None, method="log_prob", **pb)))(params)
oloss, ograds, _ = oad.value_and_grad(om, oparams, **ob)
ge = parity.grad_errors(grads, ograds)
for k in sorted(ge, key=ge.get, reverse=True)[:5]:
  print("  ", k, "%.3e" % ge[k])

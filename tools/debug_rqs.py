import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity
from surjectors_b200 import execution_mode
kw = dict(parity.SMALL["cfg4"])
for variant in [dict(n_pre=1, n_post=0), dict(n_pre=0, n_post=1), dict(n_pre=2, n_post=2)]:
  kw2 = dict(kw); kw2.update(variant)
  pm, pparams, om, oparams, y64, x64 = parity.build_pair("cfg4", 2000, 0, 0.3, **kw2)
  dev = torch.device("cuda:0")
  params = parity.to_device(pparams, dev)
  y = y64.float().to(dev)
  with execution_mode("fused"):
    z, lp = pm.apply(params, None, method="inverse_and_log_prob", y=y)
  ref = om.apply(oparams, y=y64)
  # float32 oracle for comparison
  from oracle import hk as ohk, configs as ocfg
  om32 = ohk.without_apply_rng(ohk.transform(ocfg.make_cfg4(**kw2), dtype=torch.float32))
  op32 = {m: {k: v.float() for k, v in b.items()} for m, b in oparams.items()}
  ref32 = om32.apply(op32, y=y64.float())
  err = (lp.cpu().double() - ref).abs() / ref.abs().clamp(min=1)
  err32 = (ref32.double() - ref).abs() / ref.abs().clamp(min=1)
  i = int(err.argmax())
  print(variant, "gpu max rel err", float(err.max()), "at", i, "lp", float(lp[i]), "ref", float(ref[i]), "| torch-f32 oracle max rel err", float(err32.max()), "median gpu", float(err.median()), "median f32", float(err32.median()))
  print("   n(err>1e-5) gpu:", int((err > 1e-5).sum()), " f32 oracle:", int((err32 > 1e-5).sum()), " |y| max at worst:", float(y64[i].abs().max()))

"""Sampling direction of BASELINE cfg5 (16 x [MAF(128, hidden 512), LULinear, Permutation]) has no forward for LULinear
upstream, so the timing uses its MAF + Permutation layers: D = 128, hidden [512, 512], conditional on x[B, 16]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import surjectors_b200 as sj
from surjectors_b200 import _abi, execution_mode, hk
from surjectors_b200.nn import MADE

dim, hidden, n_layers, B = 128, [512, 512], int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
def build():
  layers = []
  for _ in range(n_layers):
    layers += [sj.AffineMaskedAutoregressive(MADE(dim, hidden, 2)), sj.Permutation(np.arange(dim)[::-1].copy(), 1)]
  return sj.Chain(layers[:-1])
fwd = hk.transform(lambda **kw: build().forward_and_likelihood_contribution(**kw))
g = torch.Generator().manual_seed(0)
z = torch.randn((B, dim), generator=g).cuda(); x = torch.randn((B, 16), generator=g).cuda()
lib = _abi.lib()
with execution_mode("layered"):
  params = fwd.init(0, z=z[:8], x=x[:8])
  for on in (1, 0):
    if on == 0 and "--unrolled" not in sys.argv: continue
    lib.surj_set_wide_fusion(on)
    bb = B if on else min(B, 1 << 13)
    fwd.apply(params, None, z=z[:bb], x=x[:bb]); torch.cuda.synchronize()
    t0 = time.perf_counter(); y, fl = fwd.apply(params, None, z=z[:bb], x=x[:bb]); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{'recursion kernel' if on else 'D unrolled passes'}: {n_layers} MAF layers, D={dim}, hidden {hidden}, batch {bb}: {1e3 * dt:.1f} ms = {1e3 * dt / n_layers:.2f} ms per layer, {bb / dt / 1e6:.3f} M samples/s")
lib.surj_set_wide_fusion(1)
import ctypes as C, json
with execution_mode("layered"):
  lib.surj_profile_begin()
  fwd.apply(params, None, z=z, x=x)
  buf = C.create_string_buffer(1 << 16); lib.surj_profile_end(buf, len(buf))
  for k, v in json.loads(buf.value.decode()).items():
    if v["launches"]: print("  ", k, v["launches"], "launches", round(v["ms"], 2), "ms")

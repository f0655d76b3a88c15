"""Writes the fixtures of the torch-free C driver (tests/c_driver/abi_driver.c): a chain's SurjLayerDesc[] exactly as the
host layer hands it to surj_plan_create (pointer fields replaced by offsets into an attached blob), its parameters, inputs,
and the float64 oracle's log-probs, loss and gradient.  TEST INFRASTRUCTURE.  Run on the CPU:

  python tests/c_driver/make_fixture.py            # writes tests/golden/cdrv_<name>.bin

File (little-endian): "SURJFIX1", i32 n_layers, dim_in, dim_x, dim_out, sizeof(SurjLayerDesc), i64 n_params, batch,
desc bytes, i64 aux_bytes, aux, f32 params[n_params], y[B, dim_in], x[B, dim_x], lp_ref[B], g_ref[n_params], f64 loss_ref."""

import ctypes as C
import os
import struct
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import autodiff as oad  # noqa: E402
from surjectors_b200 import _abi  # noqa: E402
from tests import parity  # noqa: E402

CASES = {  # name -> (config, batch, scale kwargs, w_std)
  "cfg2": ("cfg2", 300, {}, 0.3),
  "cfg3": ("cfg3", 300, {}, 0.3),
  "cfg4_small": ("cfg4", 300, parity.SMALL["cfg4"], 0.3),
  "cfg5_small": ("cfg5", 200, parity.SMALL["cfg5"], 0.3),
}


def write(name, cfg, batch, scale, w_std, out_dir):
  pm, pparams, om, oparams, y64, x64 = parity.build_pair(cfg, batch, 0, w_std, **scale)
  ex = dict(y=y64[:1].float()) if x64 is None else dict(y=y64[:1].float(), x=x64[:1].float())
  traced, _ = pm.trace(method="log_prob", **ex)
  spec = traced[0]
  arr, keep = spec.to_c()
  n = len(spec.layers)
  raw = bytearray(C.string_at(C.addressof(arr), C.sizeof(arr)))
  aux = bytearray()
  sz = C.sizeof(_abi.LayerDesc)
  for k, L in enumerate(spec.layers):
    for field, data in (("mask", None if L.mask is None else np.asarray(L.mask).astype(np.uint8)),
                        ("permutation", None if L.permutation is None else np.asarray(L.permutation).astype(np.int32))):
      off = getattr(_abi.LayerDesc, field).offset
      val = 0
      if data is not None:
        while len(aux) % 8:
          aux.append(0)
        val = len(aux) + 1  # offset + 1; 0 = NULL
        aux += data.tobytes()
      raw[k * sz + off:k * sz + off + 8] = struct.pack("<q", val)
  ob = dict(y=y64) if x64 is None else dict(y=y64, x=x64)
  lp = om.apply(oparams, **ob).numpy()
  loss, grads, _ = oad.value_and_grad(om, oparams, **ob)
  g = np.zeros(spec.n_params, dtype=np.float32)
  flat = pparams.flat.numpy()
  gflat = torch.zeros_like(pparams.flat)
  gtree = parity.phk.tree_from_flat(gflat, pparams.layout)
  for m in grads:
    for kk in grads[m]:
      gtree[m][kk].copy_(grads[m][kk].to(torch.float32))
  g = gflat.numpy()
  dim_x = 0 if x64 is None else x64.shape[1]
  path = os.path.join(out_dir, f"cdrv_{name}.bin")
  with open(path, "wb") as f:
    f.write(b"SURJFIX1")
    f.write(struct.pack("<5i", n, spec.dim_in, dim_x, spec.dim_out, sz))
    f.write(struct.pack("<2q", spec.n_params, batch))
    f.write(bytes(raw))
    f.write(struct.pack("<q", len(aux)))
    f.write(bytes(aux))
    f.write(flat.astype(np.float32).tobytes())
    f.write(y64.numpy().astype(np.float32).tobytes())
    if dim_x:
      f.write(x64.numpy().astype(np.float32).tobytes())
    f.write(lp.astype(np.float32).tobytes())
    f.write(g.astype(np.float32).tobytes())
    f.write(struct.pack("<d", float(loss)))
  del keep
  return path, os.path.getsize(path)


if __name__ == "__main__":
  out = os.path.join(ROOT, "tests", "golden")
  for name, (cfg, batch, scale, w_std) in CASES.items():
    print(*write(name, cfg, batch, scale, w_std, out))

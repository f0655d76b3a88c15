"""One parity case through the per-op executor with the workspace filled with 0xFF bytes (float NaN) before every ABI call:
whatever a kernel reads without having written it in the same call shows up as NaN."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surjectors_b200 import _abi, _runtime
from tests import parity
name = sys.argv[1]; batch = int(sys.argv[2]); small = sys.argv[3] == "small"; fusion = int(sys.argv[4])
_abi.lib().surj_set_wide_fusion(fusion)
orig = _runtime.Plan.workspace
def poisoned(self, b, backward, device):
  buf, n = orig(self, b, backward, device)
  buf.fill_(0xFF)
  return buf, n
if "--poison" in sys.argv:
  _runtime.Plan.workspace = poisoned
try:
  print(parity.check_config(name, batch=batch, check_grad=True, mode="layered", w_std=0.3 if small else 0.15, kink_max_frac=0.9,
                            **(parity.SMALL.get(name, {}) if small else {})))
except AssertionError as e:
  print("FAIL", str(e)[:300])

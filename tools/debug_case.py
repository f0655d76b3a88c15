import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surjectors_b200 import _abi
from tests import parity
name = sys.argv[1]; batch = int(sys.argv[2]); small = sys.argv[3] == "small"; fusion = int(sys.argv[4])
_abi.lib().surj_set_wide_fusion(fusion)
try:
  print(parity.check_config(name, batch=batch, check_grad=True, mode="layered", w_std=0.3 if small else 0.15, **(parity.SMALL.get(name, {}) if small else {})))
except AssertionError as e:
  print("FAIL", str(e)[:300])

"""Chains that get a fused whole-chain kernel compiled into libsurj_b200.so.

Each entry builds a model through the public API, traces it on CPU (no arithmetic) and hands the
resulting ChainSpec to the code generator.  The BASELINE.json small-network configs are listed at
their stated shapes; every other chain runs on the generic per-layer kernels.
"""

import os

import torch

from surjectors_b200 import configs, hk


def spec_of(flow, dim, dim_x=0):
  """ChainSpec of `flow` (a function for hk.transform taking method=, y=, x=) without running it."""
  y = torch.zeros((1, dim))
  batch = dict(y=y) if dim_x == 0 else dict(y=y, x=torch.zeros((1, dim_x)))
  traced, _ = hk.transform(flow).trace(method="log_prob", **batch)
  return traced[-1]


def fused_chain_specs():
  out = []
  for name in ("cfg1", "cfg2", "cfg3"):
    c = configs.CONFIGS[name]
    nt = int(os.environ.get("SURJ_GEN1_NT", "64"))
    minb = int(os.environ.get("SURJ_GEN1_MINB", "8"))
    out.append((name, spec_of(c["make"](), c["dim"], c["dim_x"]), nt, minb))
  return out

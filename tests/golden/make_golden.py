"""Generates the golden fixtures in this directory from the CPU oracle (float64).

Run from the repo root:  python tests/golden/make_golden.py
Each fixture holds the seeded inputs, the parameters (by haiku name) and the oracle's log-probs,
latents and loss gradients for one BASELINE config at a small batch.  The real reference
(JAX/haiku/distrax) cannot be imported in this image, so these vectors come
from the restatement, which is itself pinned by the reference's 8 known-answer tests
(tests/test_oracle_kat.py); the KAT values are stored verbatim in kat.json.
"""

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import autodiff as oad  # noqa: E402
from tests import parity  # noqa: E402

BATCH = 48


def main():
  for name in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
    scale = parity.SMALL.get(name, {})
    pm, pparams, om, oparams, y64, x64 = parity.build_pair(name, BATCH, seed=0, w_std=0.3, **scale)
    ob = dict(y=y64) if x64 is None else dict(y=y64, x=x64)
    lp = om.apply(oparams, **ob)
    loss, grads, _ = oad.value_and_grad(om, oparams, **ob)
    out = {"y": y64.numpy(), "lp": lp.numpy(), "loss": np.float64(loss)}
    if x64 is not None:
      out["x"] = x64.numpy()
    for m in oparams:
      for k in oparams[m]:
        out[f"param:{m}:{k}"] = oparams[m][k].numpy()
        out[f"grad:{m}:{k}"] = grads[m][k].numpy()
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "lp[:3] =", lp[:3].tolist(), "loss =", float(loss))
  kat = {
    "source": "surjectors v0.3.4 made_test.py:43-44,76-77; masked_autoregressive_test.py:113-114,149-150",
    "made": {"[2,4,2]": [[0.0, 2.0, 10.0]], "[2]": [[0.0, 1.0, 3.0]]},
    "conditional_made": {"[2,4,2]": [[0.0, 6.0, 26.0]], "[2]": [[0.0, 3.0, 7.0]]},
    "maf_forward": {"[2,4,2]": [[1.0, 9.3890562, 4.2617353e11]], "[2]": [[1.0, 3.7182817, 310.09976]]},
    "conditional_maf_forward": {"[2,4,2]": [[1.0, 16.063736, 2.9474317e18]], "[2]": [[1.0, 4.9692965, 1945.3466]]},
  }
  with open(os.path.join(HERE, "kat.json"), "w") as f:
    json.dump(kat, f, indent=1)


if __name__ == "__main__":
  main()

"""Every loader variant of the gradient-phase kernel (csrc/vm_gemm3.cuh) against the float64 oracle.

The default run only sends contractions of K >= 1024 to `vg_gemm3<0>` and uses the TMA-fed one-CTA variant; the cp.async
loaders (the fallback when a tensor map cannot describe an operand), the CTA-pair kernel, the variant that splits the
weight operand in the SM and the fused activation derivative are reached through environment switches that the library
reads once per process -- so each variant runs the oracle-parity cases of `test_gpu_wide_fusion.py` in a child process,
with `SURJ_B200_GEMM3_MINK=1` so that the small test chains (K = 16 ... 772, ragged in k and in the sample tile) go
through the kernel as well."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
  "tma": {},
  "tma_split_weights_in_sm": {"SURJ_B200_GEMM3_BLO": "0"},
  "cp_async": {"SURJ_B200_GEMM3_TMA": "0"},
  "cta_pairs": {"SURJ_B200_GEMM3_TMA": "15"},
  "act_bwd_in_epilogue": {"SURJ_B200_FUSE_ACT_BWD": "1"},
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_variant_against_oracle(variant):
  env = dict(os.environ)
  env.update(VARIANTS[variant])
  env["SURJ_B200_GEMM3_MINK"] = "1"
  r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_wide_fusion.py"), "-q", "-x", "-m", "gpu",
                      "-k", "against_oracle or against_unfused or three_launches", "-p", "no:cacheprovider"],
                     cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
  tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
  assert r.returncode == 0, f"variant {variant} ({VARIANTS[variant]}):\n{tail}"
  assert " passed" in r.stdout, tail

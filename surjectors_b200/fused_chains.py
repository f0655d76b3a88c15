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


def build_fused_kernel(model, example_batch, method="log_prob", name="user", nt=64, min_blocks=8, cache_dir=None):
  """Ahead-of-time build of the generated straight-line kernel for ONE more chain (the same code generator and
  templates `__graft_entry__.build()` uses for the BASELINE configs): traces `model` (an hk.transform), emits the
  chain's template instantiation, compiles it with nvcc for sm_100a into a small shared object next to the main
  library and loads it, which registers the kernel under the chain's key -- plans created afterwards pick it up
  (executor "gen1").  Affine chains with small conditioners only (see _codegen.NotFusable).  Returns the .so path."""
  import ctypes
  import subprocess

  from surjectors_b200 import _abi, _codegen

  traced, _ = model.trace(method=method, **{k: v[:1].cpu() for k, v in example_batch.items() if v is not None})
  spec = traced[-1]
  key = spec.key(_abi.keylib() if not os.path.exists(_abi.LIB_PATH) else _abi.lib())
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out_dir = cache_dir or os.path.join(_abi.LIB_DIR, "user_chains")
  os.makedirs(out_dir, exist_ok=True)
  so = os.path.join(out_dir, f"chain_{key:016x}.so")
  if not os.path.exists(so):
    src = os.path.join(out_dir, f"chain_{key:016x}.cu")
    text = _codegen.emit_translation_unit(spec, key, name, nt, min_blocks).replace(
      '#include "../chain.cuh"', f'#include "{os.path.join(root, "surjectors_b200", "csrc", "chain.cuh")}"')
    with open(src, "w") as f:
      f.write(text)
    cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
           "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared", "-I", os.path.join(root, "include"),
           src, "-o", so, "-L", _abi.LIB_DIR, "-lsurj_b200", f"-Xlinker=-rpath={_abi.LIB_DIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError(f"nvcc failed for chain {key:016x}:\n{r.stdout}\n{r.stderr}")
  _abi.lib()  # the main library must be loaded (globally) first: the chain object calls surj_register_small in it
  ctypes.CDLL(so, mode=ctypes.RTLD_GLOBAL)
  return so

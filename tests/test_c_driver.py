"""The C ABI without Python or PyTorch on the calling side: tests/c_driver/abi_driver.c (plain C99 + the CUDA runtime) is
compiled with gcc, linked against libsurj_b200.so and run on fixtures holding the float64 oracle's numbers
(tests/c_driver/make_fixture.py).  CPU part: it compiles and links against every symbol it uses; GPU part: it runs."""

import os
import shutil
import subprocess

import pytest

from surjectors_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_driver", "abi_driver.c")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
FIXTURES = ["cfg2", "cfg3", "cfg4_small", "cfg5_small"]


def _build(tmp_path):
  if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
    pytest.skip("gcc or the CUDA headers are missing")
  exe = str(tmp_path / "abi_driver")
  libdir = os.path.dirname(_abi.LIB_PATH)
  cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), SRC,
         "-L", libdir, "-lsurj_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{libdir}",
         f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}", "-o", exe]
  r = subprocess.run(cmd, capture_output=True, text=True)
  assert r.returncode == 0, r.stderr
  return exe


def test_driver_compiles_and_links(tmp_path, built_library):
  exe = _build(tmp_path)
  out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
  assert "libsurj_b200.so" in out and "libtorch" not in out and "libpython" not in out and "libc10" not in out, out
  for name in FIXTURES:
    assert os.path.exists(os.path.join(ROOT, "tests", "golden", f"cdrv_{name}.bin")), name


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("name", FIXTURES)
def test_driver_runs_against_oracle_fixture(tmp_path, name, mode):
  exe = _build(tmp_path)
  r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", f"cdrv_{name}.bin"), str(mode)], capture_output=True, text=True, timeout=300)
  print(r.stdout, r.stderr)
  assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stdout, r.stderr)

"""The C-ABI library loads on a box without a GPU and exports every symbol include/surj_b200.h
declares; the ctypes mirror of the descriptor structs matches what a C compiler lays out; plans
validate descriptors the way the reference's constructors do.  No compute calls here."""

import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from surjectors_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "surj_b200.h")


@pytest.fixture(scope="module")
def lib():
  if not os.path.exists(_abi.LIB_PATH):
    import __graft_entry__ as entry

    entry.build()
  return _abi.lib()


def declared_functions():
  text = open(HEADER).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(surj_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
  names = declared_functions()
  assert len(names) >= 15
  raw = C.CDLL(_abi.LIB_PATH)
  for n in names:
    assert hasattr(raw, n), f"{n} is declared in include/surj_b200.h but not exported"
  assert set(names) == set(_abi.EXPORTED_SYMBOLS), set(names) ^ set(_abi.EXPORTED_SYMBOLS)


def test_header_is_plain_c_and_struct_layout_matches(tmp_path):
  src = tmp_path / "sizes.c"
  src.write_text(
    '#include <stdio.h>\n#include <stddef.h>\n#include "surj_b200.h"\n'
    "int main(void){printf(\"%zu %zu %zu %zu %zu %zu\\n\", sizeof(SurjMlpDesc), sizeof(SurjMadeDesc), "
    "sizeof(SurjLayerDesc), offsetof(SurjLayerDesc, conditioner), offsetof(SurjLayerDesc, made), "
    "offsetof(SurjLayerDesc, lu_lower_off)); return 0;}\n"
  )
  exe = tmp_path / "sizes"
  subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                 check=True)
  got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
  want = [C.sizeof(_abi.MlpDesc), C.sizeof(_abi.MadeDesc), C.sizeof(_abi.LayerDesc), _abi.LayerDesc.conditioner.offset,
          _abi.LayerDesc.made.offset, _abi.LayerDesc.lu_lower_off.offset]
  assert got == want


def test_abi_version_and_error_string(lib):
  assert lib.surj_abi_version() == 1
  out = C.c_void_p()
  st = lib.surj_plan_create(None, 0, 0, 0, 0, 0, 0, C.byref(out))
  assert st == _abi.E_INVALID
  assert b"bad arguments" in lib.surj_last_error()


def test_made_masks_match_tfp_construction(lib):
  from oracle import tfp
  from surjectors_b200.nn import make_dense_autoregressive_masks

  for n_params, d, hidden in [(2, 3, [2, 4, 2]), (2, 3, [2]), (2, 10, [8, 8]), (2, 5, [8, 8]), (3, 7, [16]), (2, 128, [64, 32]),
                              (193, 4, [9, 5]), (2, 1, [4])]:
    h = (C.c_int32 * len(hidden))(*hidden)
    n = lib.surj_made_masks(n_params, d, h, len(hidden), None)
    buf = (C.c_uint8 * n)()
    assert lib.surj_made_masks(n_params, d, h, len(hidden), buf) == n
    flat = np.frombuffer(buf, dtype=np.uint8)
    ref = tfp.make_dense_autoregressive_masks(n_params, d, hidden)
    ours = make_dense_autoregressive_masks(n_params, d, hidden)
    off = 0
    for r, o in zip(ref, ours):
      assert np.array_equal(np.asarray(r, dtype=bool), np.asarray(o, dtype=bool))
      got = flat[off : off + r.size].reshape(r.shape).astype(bool)
      assert np.array_equal(got, np.asarray(r, dtype=bool))  # bit-exact
      off += r.size
    assert off == n


def test_permutation_inverse_is_bit_exact(lib):
  rng = np.random.default_rng(0)
  for n in (1, 2, 5, 128):
    perm = rng.permutation(n).astype(np.int32)
    inv = np.zeros(n, dtype=np.int32)
    st = lib.surj_permutation_inverse(perm.ctypes.data_as(C.POINTER(C.c_int32)), n, inv.ctypes.data_as(C.POINTER(C.c_int32)))
    assert st == 0
    ref = np.zeros(n, dtype=np.int32)
    ref[perm] = np.arange(n, dtype=np.int32)  # permutation.py:31-36
    assert np.array_equal(inv, ref) and inv.dtype == np.int32
  bad = np.array([0, 0, 1], dtype=np.int32)
  inv = np.zeros(3, dtype=np.int32)
  assert lib.surj_permutation_inverse(bad.ctypes.data_as(C.POINTER(C.c_int32)), 3, inv.ctypes.data_as(C.POINTER(C.c_int32))) == 0

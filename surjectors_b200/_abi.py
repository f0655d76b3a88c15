"""ctypes mirror of include/surj_b200.h and the loader of the in-tree shared library.

There is no CPU fallback: if `libsurj_b200.so` is missing, importing a compute entry point
raises.  Structures must stay field-for-field identical to the header.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.environ.get("SURJ_B200_LIB", os.path.join(LIB_DIR, "libsurj_b200.so"))  # override: A/B builds
KEYLIB_PATH = os.path.join(LIB_DIR, "_surjkey.so")

MAX_NET_LAYERS = 8

# enums (include/surj_b200.h)
LAYER_COUPLING, LAYER_COUPLING_FUNNEL, LAYER_MAF, LAYER_MAF_FUNNEL, LAYER_SLICE, LAYER_LU, LAYER_PERMUTATION, \
  LAYER_MLP_FUNNEL = range(8)
INNER_AFFINE, INNER_AFFINE_INV, INNER_RQS, INNER_RQS_INV = range(4)
LAYOUT_SPLIT, LAYOUT_INTERLEAVED = range(2)
ACT = {
  "identity": 0, "relu": 1, "swish": 2, "tanh": 3, "sigmoid": 4, "softplus": 5, "gelu": 6, "elu": 7,
  "leaky_relu": 8,
}
MODE_AUTO, MODE_FUSED, MODE_LAYERED, MODE_FUSED_GEN1, MODE_FUSED_COMPACT = range(5)
OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_WORKSPACE = 0, -1, -2, -3, -4


class MlpDesc(C.Structure):
  _fields_ = [
    ("n_layers", C.c_int32),
    ("widths", C.c_int32 * (MAX_NET_LAYERS + 1)),
    ("act", C.c_int32 * MAX_NET_LAYERS),
    ("w_off", C.c_int64 * MAX_NET_LAYERS),
    ("b_off", C.c_int64 * MAX_NET_LAYERS),
  ]


class MadeDesc(C.Structure):
  _fields_ = [
    ("input_size", C.c_int32),
    ("n_params", C.c_int32),
    ("n_hidden", C.c_int32),
    ("hidden", C.c_int32 * MAX_NET_LAYERS),
    ("act", C.c_int32),
    ("ctx_dim", C.c_int32),
    ("w_off", C.c_int64 * (MAX_NET_LAYERS + 1)),
    ("b_off", C.c_int64 * (MAX_NET_LAYERS + 1)),
    ("ctx_w_off", C.c_int64),
    ("ctx_b_off", C.c_int64),
  ]


class LayerDesc(C.Structure):
  _fields_ = [
    ("kind", C.c_int32),
    ("dim_in", C.c_int32),
    ("dim_out", C.c_int32),
    ("use_x", C.c_int32),
    ("inner", C.c_int32),
    ("layout", C.c_int32),
    ("n_bins", C.c_int32),
    ("range_min", C.c_float),
    ("range_max", C.c_float),
    ("mask", C.POINTER(C.c_uint8)),
    ("permutation", C.POINTER(C.c_int32)),
    ("conditioner", MlpDesc),
    ("made", MadeDesc),
    ("decoder", MlpDesc),
    ("lu_lower_off", C.c_int64),
    ("lu_upper_off", C.c_int64),
    ("lu_diag_off", C.c_int64),
  ]


class SurjError(RuntimeError):
  pass


_lib = None
_keylib = None

_SIGNATURES = {
  "surj_plan_create": (C.c_int, [C.POINTER(LayerDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                 C.c_int32, C.POINTER(C.c_void_p)]),
  "surj_plan_destroy": (None, [C.c_void_p]),
  "surj_plan_key": (C.c_uint64, [C.c_void_p]),
  "surj_chain_key": (C.c_int, [C.POINTER(LayerDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                               C.POINTER(C.c_uint64)]),
  "surj_plan_is_fused": (C.c_int32, [C.c_void_p]),
  "surj_plan_dim_out": (C.c_int32, [C.c_void_p]),
  "surj_plan_launches": (C.c_int32, [C.c_void_p, C.c_int32]),
  "surj_plan_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int32]),
  "surj_chain_logprob_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64]),
  "surj_chain_logprob_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64]),
  "surj_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                               C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
  "surj_chain_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64]),
  "surj_plan_noise_dim": (C.c_int32, [C.c_void_p]),
  "surj_plan_forward_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
  "surj_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64]),
  "surj_dp_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
  "surj_dp_connect": (C.c_int, [C.c_void_p, C.c_char_p]),
  "surj_dp_destroy": (None, [C.c_void_p]),
  "surj_dp_error": (C.c_int32, [C.c_void_p]),
  "surj_dp_error_nosync": (C.c_int32, [C.c_void_p]),
  "surj_dp_set_timeout_ms": (C.c_int, [C.c_void_p, C.c_int64]),
  "surj_train_step_dp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                   C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64]),
  "surj_made_masks": (C.c_int64, [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_uint8)]),
  "surj_permutation_inverse": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
  "surj_index_take": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
  "surj_bench_fp32_peak": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_double)]),
  "surj_bench_tf32_peak": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_double)]),
  "surj_profile_begin": (C.c_int, []),
  "surj_profile_end": (C.c_int64, [C.c_char_p, C.c_int64]),
  "surj_plan_sampling_recursions": (C.c_int32, [C.c_void_p]),
  "surj_set_chunk_samples": (C.c_int64, [C.c_int64]),
  "surj_set_wide_fusion": (C.c_int32, [C.c_int32]),
  "surj_last_error": (C.c_char_p, []),
  "surj_abi_version": (C.c_int32, []),
  "surj_registry_size": (C.c_int32, []),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def _bind(lib, names):
  for name in names:
    fn = getattr(lib, name)
    fn.restype, fn.argtypes = _SIGNATURES[name]
  return lib


def lib():
  """The product library.  Raises (loudly) when it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise SurjError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(there is no CPU fallback)"
      )
    _lib = _bind(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL), _SIGNATURES)
  return _lib


def keylib():
  """Host-only helper library (chain key) used by the build-time code generator."""
  global _keylib
  if _keylib is None:
    if os.path.exists(LIB_PATH) and _lib is not None:
      return _lib
    path = KEYLIB_PATH if os.path.exists(KEYLIB_PATH) else LIB_PATH
    if not os.path.exists(path):
      raise SurjError(f"neither {KEYLIB_PATH} nor {LIB_PATH} exists; run the build first")
    _keylib = _bind(C.CDLL(path), ["surj_chain_key"])
  return _keylib


def check(status, what=""):
  if status == OK:
    return
  msg = lib().surj_last_error().decode("utf-8", "replace")
  text = f"{what}: {msg}" if what else msg
  if status == E_UNSUPPORTED:
    raise NotImplementedError(text)
  if status == E_INVALID:
    raise ValueError(text)
  raise SurjError(f"{text} (status {status})")

"""Plain-old-data description of a Chain: what the host layer hands to the C ABI.

A `ChainSpec` is the structure the reference would trace under `jax.jit`
(examples/coupling_inference_surjection.py:81): the list of layers, their conditioner /
decoder networks and the float offset of every haiku parameter inside one flat float32
buffer.  `to_c()` produces the `SurjLayerDesc[]` of include/surj_b200.h.
"""

from __future__ import annotations

import ctypes as C
import dataclasses as dc
from typing import List, Optional, Tuple

import numpy as np

from surjectors_b200 import _abi

ALIGN = 4  # every parameter starts on a 16-byte boundary of the flat buffer


@dc.dataclass
class ParamRef:
  module: str
  name: str
  shape: Tuple[int, ...]
  offset: int = -1

  @property
  def size(self):
    return int(np.prod(self.shape)) if len(self.shape) else 1


class ParamLayout:
  """Flat-buffer layout: haiku (module, name) -> float offset, in creation order."""

  def __init__(self):
    self.refs: List[ParamRef] = []
    self._index = {}
    self.size = 0

  def add(self, module, name, shape) -> ParamRef:
    key = (module, name)
    if key in self._index:
      ref = self._index[key]
      if tuple(ref.shape) != tuple(shape):
        raise ValueError(f"{module}/{name}: shape {tuple(shape)} does not match earlier use {ref.shape}")
      return ref
    ref = ParamRef(module, name, tuple(int(s) for s in shape), self.size)
    self.size += (ref.size + ALIGN - 1) // ALIGN * ALIGN
    self.refs.append(ref)
    self._index[key] = ref
    return ref

  def signature(self):
    return tuple((r.module, r.name, r.shape, r.offset) for r in self.refs)


@dc.dataclass
class MlpSpec:
  widths: List[int]  # widths[0] = input width
  acts: List[str]  # activation after layer l
  w: List[ParamRef]
  b: List[Optional[ParamRef]]

  @property
  def n_layers(self):
    return len(self.w)

  def to_c(self):
    d = _abi.MlpDesc()
    if self.n_layers > _abi.MAX_NET_LAYERS:
      raise NotImplementedError(f"networks deeper than {_abi.MAX_NET_LAYERS} Linear layers are not supported")
    d.n_layers = self.n_layers
    for i, wd in enumerate(self.widths):
      d.widths[i] = wd
    for l in range(self.n_layers):
      d.act[l] = _abi.ACT[self.acts[l]]
      d.w_off[l] = self.w[l].offset
      d.b_off[l] = self.b[l].offset if self.b[l] is not None else -1
    return d


@dc.dataclass
class MadeSpec:
  input_size: int
  n_params: int
  hidden: List[int]
  act: str
  ctx_dim: int
  w: List[ParamRef]
  b: List[ParamRef]
  ctx_w: Optional[ParamRef] = None
  ctx_b: Optional[ParamRef] = None

  def to_c(self):
    d = _abi.MadeDesc()
    if len(self.hidden) > _abi.MAX_NET_LAYERS:
      raise NotImplementedError(f"MADE with more than {_abi.MAX_NET_LAYERS} hidden layers is not supported")
    d.input_size = self.input_size
    d.n_params = self.n_params
    d.n_hidden = len(self.hidden)
    for i, h in enumerate(self.hidden):
      d.hidden[i] = h
    d.act = _abi.ACT[self.act]
    d.ctx_dim = self.ctx_dim
    for l in range(len(self.hidden) + 1):
      d.w_off[l] = self.w[l].offset
      d.b_off[l] = self.b[l].offset
    d.ctx_w_off = self.ctx_w.offset if self.ctx_w is not None else -1
    d.ctx_b_off = self.ctx_b.offset if self.ctx_b is not None else -1
    return d


@dc.dataclass
class LayerSpec:
  kind: int
  dim_in: int
  dim_out: int
  use_x: bool = False
  inner: int = _abi.INNER_AFFINE
  layout: int = _abi.LAYOUT_SPLIT
  n_bins: int = 0
  range_min: float = 0.0
  range_max: float = 0.0
  mask: Optional[np.ndarray] = None  # bool [dim_in]
  permutation: Optional[np.ndarray] = None  # int32 [dim_in]
  conditioner: Optional[MlpSpec] = None
  made: Optional[MadeSpec] = None
  decoder: Optional[MlpSpec] = None
  lu: Optional[Tuple[ParamRef, ParamRef, ParamRef]] = None  # lower, upper, diag

  def structure(self):
    """Hashable structural identity (used to cache plans)."""

    def mlp(m):
      if m is None:
        return None
      return (tuple(m.widths), tuple(m.acts), tuple(w.offset for w in m.w),
              tuple(-1 if b is None else b.offset for b in m.b))

    def made(m):
      if m is None:
        return None
      return (m.input_size, m.n_params, tuple(m.hidden), m.act, m.ctx_dim, tuple(w.offset for w in m.w),
              tuple(b.offset for b in m.b), -1 if m.ctx_w is None else m.ctx_w.offset,
              -1 if m.ctx_b is None else m.ctx_b.offset)

    return (
      self.kind, self.dim_in, self.dim_out, bool(self.use_x), self.inner, self.layout, self.n_bins,
      float(np.float32(self.range_min)), float(np.float32(self.range_max)),
      None if self.mask is None else tuple(bool(v) for v in self.mask),
      None if self.permutation is None else tuple(int(v) for v in self.permutation),
      mlp(self.conditioner), made(self.made), mlp(self.decoder),
      None if self.lu is None else tuple(r.offset for r in self.lu),
    )


class ChainSpec:
  def __init__(self, layers: List[LayerSpec], dim_in: int, dim_x: int, layout: ParamLayout):
    self.layers = layers
    self.dim_in = int(dim_in)
    self.dim_x = int(dim_x)
    self.layout = layout
    self.dim_out = layers[-1].dim_out if layers else dim_in

  @property
  def n_params(self):
    return self.layout.size

  def structure(self):
    return (self.dim_in, self.dim_x, self.n_params, tuple(l.structure() for l in self.layers))

  def to_c(self):
    """(SurjLayerDesc array, keep-alive list).  The arrays behind mask / permutation pointers
    must outlive the call that consumes the descriptors."""
    n = len(self.layers)
    arr = (_abi.LayerDesc * n)()
    keep = []
    for k, L in enumerate(self.layers):
      d = arr[k]
      d.kind, d.dim_in, d.dim_out, d.use_x = L.kind, L.dim_in, L.dim_out, int(bool(L.use_x))
      d.inner, d.layout, d.n_bins = L.inner, L.layout, L.n_bins
      d.range_min, d.range_max = float(L.range_min), float(L.range_max)
      if L.mask is not None:
        m = np.ascontiguousarray(np.asarray(L.mask).astype(np.uint8))
        keep.append(m)
        d.mask = m.ctypes.data_as(C.POINTER(C.c_uint8))
      if L.permutation is not None:
        p = np.ascontiguousarray(np.asarray(L.permutation).astype(np.int32))
        keep.append(p)
        d.permutation = p.ctypes.data_as(C.POINTER(C.c_int32))
      if L.conditioner is not None:
        d.conditioner = L.conditioner.to_c()
      if L.made is not None:
        d.made = L.made.to_c()
      if L.decoder is not None:
        d.decoder = L.decoder.to_c()
      if L.lu is not None:
        d.lu_lower_off, d.lu_upper_off, d.lu_diag_off = (r.offset for r in L.lu)
      else:
        d.lu_lower_off = d.lu_upper_off = d.lu_diag_off = -1
    return arr, keep

  def key(self, keylib=None):
    lib = keylib if keylib is not None else _abi.keylib()
    arr, keep = self.to_c()
    out = C.c_uint64(0)
    st = lib.surj_chain_key(arr, len(self.layers), self.dim_in, self.dim_x, self.n_params, C.byref(out))
    del keep
    if st != 0:
      raise ValueError("surj_chain_key failed")
    return int(out.value)

"""surjectors/util.py: the index helpers on the hot path (bit-exact integer / bool semantics) and the batch iterator of
the reference's training loops as an on-device gather."""

from collections import namedtuple

import numpy as np

from surjectors_b200._sym import unstack  # noqa: F401  (util.py:106-120)

__all__ = ["make_alternating_binary_mask", "as_batch_iterator", "unstack", "named_dataset"]

named_dataset = namedtuple("named_dataset", "y x")  # util.py:12


def make_alternating_binary_mask(n_dim, even_idx_as_true=False):
  """util.py:30-47: mask[i] = (i % 2 == 1); negated if even_idx_as_true."""
  mask = np.arange(0, int(np.prod(n_dim))) % 2
  mask = np.reshape(mask, n_dim).astype(bool)
  if even_idx_as_true:
    mask = np.logical_not(mask)
  return mask


class _DataLoader:
  """util.py:16-27: `loader(idx)` is batch `idx`; `num_batches`, `idxs` as upstream."""

  def __init__(self, num_batches, idxs, get_batch):
    self.num_batches = num_batches
    self.idxs = idxs
    self.get_batch = get_batch

  def __call__(self, idx, idxs=None):
    if idxs is None:
      idxs = self.idxs
    return self.get_batch(idx, idxs)


def batch_layout(n, batch_size):
  """(num_batches, batch_size) exactly as util.py:80-87 computes them."""
  if n < batch_size:
    return 1, n
  if n % batch_size == 0:
    return int(n // batch_size), batch_size
  return int(n // batch_size) + 1, batch_size


def as_batch_iterator(rng_key, data, batch_size, shuffle=True):
  """util.py:50-103 with the data set resident in HBM: the (shuffled) index vector lives on the device and a batch is
  cut out by ONE row-gather launch per field (`surj_index_take`, csrc/gather.cu), so the training loop's data step
  stays on the GPU -- no host round trip per batch.

  `data` is a NamedTuple of CUDA float32 tensors (`y`, optionally `x`, ...) with equal leading dimension.  `rng_key`: an
  int seed or a `numpy.random.Generator` (JAX's `jr.permutation` stream is not reproducible outside JAX: the shuffle is
  *a* uniform permutation, drawn on the host once per iterator).  Returns the reference's `_DataLoader`."""
  import torch

  from surjectors_b200 import _abi

  y = data.y
  if not (isinstance(y, torch.Tensor) and y.is_cuda):
    raise _abi.SurjError("as_batch_iterator gathers on the device: the data set must be CUDA tensors (no CPU fallback)")
  n = int(y.shape[0])
  fields = list(zip(data._fields, data))
  for name, a in fields:
    if a is None:
      continue
    if not (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float32 and a.device == y.device):
      raise ValueError(f"field {name!r} must be a float32 CUDA tensor on the device of `y`")
    if int(a.shape[0]) != n:
      raise ValueError(f"field {name!r} has {a.shape[0]} rows, `y` has {n}")
  num_batches, batch_size = batch_layout(n, int(batch_size))
  idx_host = np.arange(n, dtype=np.int32)
  if shuffle:
    rng = rng_key if isinstance(rng_key, np.random.Generator) else np.random.Generator(np.random.PCG64(int(rng_key)))
    idx_host = rng.permutation(n).astype(np.int32)
  idxs = torch.from_numpy(idx_host).to(y.device)
  err = torch.zeros(1, dtype=torch.int32, device=y.device)
  lib = _abi.lib()
  flat = {name: a.contiguous().reshape(n, -1) for name, a in fields if a is not None}

  def get_batch(idx, idxs=idxs):
    idx = int(idx)
    if not 0 <= idx < num_batches:
      raise IndexError(f"batch {idx} of {num_batches}")
    start = idx * batch_size
    step = min(n - start, batch_size)
    out = {}
    with torch.cuda.device(y.device):
      stream = torch.cuda.current_stream(y.device).cuda_stream
      for name, a in flat.items():
        dst = torch.empty((step,) + tuple(dict(fields)[name].shape[1:]), dtype=torch.float32, device=y.device)
        _abi.check(lib.surj_index_take(stream, a.data_ptr(), n, int(a.shape[1]), idxs.data_ptr(), start, step, dst.data_ptr(),
                                       err.data_ptr()), "surj_index_take")
        out[name] = dst
    return out

  loader = _DataLoader(num_batches, idxs, get_batch)
  loader.error_flag = err  # device int32: 1 if an index was out of range (user-supplied idxs)
  return loader

"""TEST INFRASTRUCTURE -- numpy restatement of surjectors/util.py:50-103 (as_batch_iterator) for a GIVEN index vector:
the batch arithmetic (:80-87) and the per-field `lax.index_take(array, (ret_idx,), axes=(0,))` (:88-95).  The shuffle itself
(`jr.permutation`, :90-91) is JAX's PRNG stream and cannot be reproduced here: parity is on (idxs -> batches)."""

import numpy as np


def batch_layout(n, batch_size):
  if n < batch_size:  # util.py:80-82
    return 1, n
  if n % batch_size == 0:  # :83-84
    return int(n // batch_size), batch_size
  return int(n // batch_size) + 1, batch_size  # :85-86


def get_batch(fields, idxs, idx, batch_size):
  """fields: dict name -> numpy array [n, ...]; returns dict name -> array[ret_idx] (util.py:93-101)."""
  n = len(idxs)
  start = idx * batch_size
  step = min(n - start, batch_size)
  ret_idx = np.asarray(idxs)[start:start + step]
  return {name: np.take(a, ret_idx, axis=0) for name, a in fields.items()}

"""surjectors/util.py: the index helpers on the hot path (bit-exact integer / bool semantics)."""

import numpy as np

from surjectors_b200._sym import unstack  # noqa: F401  (util.py:106-120)

__all__ = ["make_alternating_binary_mask", "unstack"]


def make_alternating_binary_mask(n_dim, even_idx_as_true=False):
  """util.py:30-47: mask[i] = (i % 2 == 1); negated if even_idx_as_true."""
  mask = np.arange(0, int(np.prod(n_dim))) % 2
  mask = np.reshape(mask, n_dim).astype(bool)
  if even_idx_as_true:
    mask = np.logical_not(mask)
  return mask

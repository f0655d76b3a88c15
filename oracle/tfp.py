"""Restatement of the TFP pieces the reference imports (tfp-nightly 0.26.0.dev20250828).

TEST INFRASTRUCTURE (see oracle/__init__.py).

`_make_dense_autoregressive_masks` is imported by the reference at
surjectors/_src/conditioners/nn/made.py:7-9 and called at made.py:54-56 with
(params=n_params, event_size=input_size, hidden_units=hidden_layer_sizes) and
TFP's defaults input_order='left-to-right', hidden_degrees='equal'.
KAT-pinned by made_test.py:43-44,76-77 (tests/test_oracle_kat.py).
"""

import numpy as np


def _create_degrees(input_size, hidden_units):
  """TFP `_create_degrees` with input_order='left-to-right', hidden_degrees='equal'."""
  degrees = [np.arange(1, input_size + 1)]
  for units in hidden_units:
    min_degree = min(int(np.min(degrees[-1])), input_size - 1)
    degrees.append(
      np.maximum(
        min_degree,
        np.ceil(np.arange(1, units + 1) * (input_size - 1) / float(units + 1)).astype(np.int32),
      )
    )
  return degrees


def make_dense_autoregressive_masks(params, event_size, hidden_units):
  """List of bool masks, shape [in, out] per layer; last one has event_size*params columns.

  Output column index is d*params + p (each event column tiled `params` times).
  """
  degrees = _create_degrees(event_size, hidden_units)
  masks = [inp[:, np.newaxis] <= out for inp, out in zip(degrees[:-1], degrees[1:])]
  masks.append(degrees[-1][:, np.newaxis] < degrees[0])
  last = masks[-1]
  last = np.reshape(np.tile(last[..., np.newaxis], [1, 1, params]), [last.shape[0], event_size * params])
  masks[-1] = last
  return masks


def normal_log_prob(value, loc, scale):
  """tfd.Normal.log_prob: -0.5*(v/s - m/s)^2 - (0.5*log(2pi) + log s)."""
  import math

  import torch

  z = value / scale - loc / scale
  return -0.5 * z * z - (0.5 * math.log(2.0 * math.pi) + torch.log(scale))

"""Pin the oracle against every numeric known-answer value in the reference's tests.

The 8 KATs: made_test.py:43-44,76-77 and
masked_autoregressive_test.py:113-114,149-150.  Expected values are quoted from
those files (they are data, not code).
"""

import numpy as np
import pytest
import torch

from oracle import hk
from oracle import surjectors as S


def _made(shape, **kwargs):
  inp, hidden, out = shape

  def fn(**kw):
    made = S.MADE(inp, hidden, out, w_init=hk.ones, activation=lambda v: v)
    return made(**kw)

  return hk.without_apply_rng(hk.transform(fn))


@pytest.mark.parametrize("shape,expected", [((3, [2, 4, 2], 2), [[0.0, 2.0, 10.0]]), ((3, [2], 2), [[0.0, 1.0, 3.0]])])
def test_made_output(shape, expected):  # made_test.py:41-70
  m = _made(shape)
  y = torch.ones((1, 3), dtype=torch.float64)
  params = m.init(0, y=y)
  out = m.apply(params, y=y)
  assert tuple(out.shape) == (1, 3, 2)  # made_test.py:37
  a, b = S.unstack(out, -1)
  assert np.array_equal(a.numpy(), np.array(expected))
  assert np.array_equal(b.numpy(), np.array(expected))


@pytest.mark.parametrize("shape,expected", [((3, [2, 4, 2], 2), [[0.0, 6.0, 26.0]]), ((3, [2], 2), [[0.0, 3.0, 7.0]])])
def test_conditional_made_output(shape, expected):  # made_test.py:73-107
  m = _made(shape)
  y = torch.ones((1, 3), dtype=torch.float64)
  x = torch.ones((1, 2), dtype=torch.float64)
  params = m.init(0, y=y, x=x)
  a, b = S.unstack(m.apply(params, y=y, x=x), -1)
  assert np.array_equal(a.numpy(), np.array(expected))
  assert np.array_equal(b.numpy(), np.array(expected))


def _maf_forward(shape):
  inp, hidden, out = shape

  def bij(params):
    means, log_scales = S.unstack(params, -1)
    return S.ScalarAffine(means, torch.exp(log_scales))

  def fn(**kw):
    layer = S.MaskedAutoregressive(
      bijector_fn=bij, conditioner=S.MADE(inp, hidden, out, w_init=hk.ones, b_init=hk.zeros, activation=lambda v: v)
    )
    return layer.forward(**kw)

  return hk.without_apply_rng(hk.transform(fn, dtype=torch.float32))


@pytest.mark.parametrize(
  "shape,expected",
  [((3, [2, 4, 2], 2), [[1.0, 9.3890562, 4.2617353e11]]), ((3, [2], 2), [[1.0, 3.7182817, 310.09976]])],
)
def test_unconditional_maf_forward(shape, expected):  # masked_autoregressive_test.py:108-143
  f = _maf_forward(shape)
  z = torch.ones((1, 3), dtype=torch.float32)
  params = f.init(0, z=z)
  y = f.apply(params, z=z)
  assert np.allclose(y.numpy(), np.array(expected, dtype=np.float32), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize(
  "shape,expected",
  [((3, [2, 4, 2], 2), [[1.0, 1.6063736e01, 2.9474317e18]]), ((3, [2], 2), [[1.0, 4.9692965, 1.9453466e03]])],
)
def test_conditional_maf_forward(shape, expected):  # masked_autoregressive_test.py:146-182
  f = _maf_forward(shape)
  z = torch.ones((1, 3), dtype=torch.float32)
  x = torch.tensor([[0.1, 0.2]], dtype=torch.float32)
  params = f.init(0, z=z, x=x)
  y = f.apply(params, z=z, x=x)
  assert np.allclose(y.numpy(), np.array(expected, dtype=np.float32), rtol=1e-5, atol=1e-8)


def test_param_names_and_shapes():  # masked_autoregressive_test.py:105-107,199-201
  def bij(params):
    means, log_scales = S.unstack(params, -1)
    return S.Inverse(S.ScalarAffine(means, torch.exp(log_scales)))

  def fn(**kw):
    layer = S.MaskedAutoregressive(bijector_fn=bij, conditioner=S.MADE(2, [8], 2))
    return S.TransformedDistribution(S.standard_normal_base(2), layer).log_prob(**kw)

  f = hk.without_apply_rng(hk.transform(fn))
  y = torch.zeros((100, 2), dtype=torch.float64)
  x = torch.zeros((100, 1), dtype=torch.float64)
  params = f.init(0, y=y, x=x)
  assert tuple(params["made/~/masked_linear_1"]["w"].shape) == (8, 4)
  assert tuple(params["made/~/masked_linear"]["w"].shape) == (2, 8)
  assert tuple(params["made/linear"]["w"].shape) == (1, 8)

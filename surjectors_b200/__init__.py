"""surjectors_b200: the batched flow log-likelihood path of dirmeier/surjectors (and its VJP) as
fused sm_100a CUDA kernels behind the reference's own class names.

Public names mirror surjectors/__init__.py:45-68 for the layers on the hot path (SURVEY.md
section 8a).  `hk`, `jnp`, `jax`, `distrax` and `tfd` in `surjectors_b200.compat` stand in for the
reference's dependencies where user code builds conditioners, bijector_fns and decoders.
"""

__version__ = "0.1.0"

from surjectors_b200 import hk, nn, util  # noqa: F401
from surjectors_b200._abi import SurjError  # noqa: F401
from surjectors_b200._layers import (  # noqa: F401
  AffineMaskedAutoregressive,
  AffineMaskedAutoregressiveInferenceFunnel,
  AffineMaskedCoupling,
  AffineMaskedCouplingInferenceFunnel,
  Bijector,
  Chain,
  LULinear,
  MLPInferenceFunnel,
  MaskedAutoregressive,
  MaskedAutoregressiveInferenceFunnel,
  MaskedCoupling,
  MaskedCouplingInferenceFunnel,
  Permutation,
  RationalQuadraticSplineMaskedCoupling,
  RationalQuadraticSplineMaskedCouplingInferenceFunnel,
  Slice,
  Surjector,
  TransformedDistribution,
)
from surjectors_b200._runtime import execution_mode, value_and_grad  # noqa: F401
from surjectors_b200._sym import Inverse, RationalQuadraticSpline, ScalarAffine  # noqa: F401

__all__ = [
  "Chain",
  "Permutation",
  "TransformedDistribution",
  "AffineMaskedAutoregressive",
  "MaskedAutoregressive",
  "MaskedAutoregressiveInferenceFunnel",
  "AffineMaskedAutoregressiveInferenceFunnel",
  "MaskedCoupling",
  "AffineMaskedCoupling",
  "RationalQuadraticSplineMaskedCoupling",
  "MaskedCouplingInferenceFunnel",
  "AffineMaskedCouplingInferenceFunnel",
  "RationalQuadraticSplineMaskedCouplingInferenceFunnel",
  "LULinear",
  "MLPInferenceFunnel",
  "Slice",
  "ScalarAffine",
  "RationalQuadraticSpline",
  "Inverse",
]

"""CPU oracle for the surjectors log_prob hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch-CPU tensors, float64 by default,
float32 on request) of the algorithm that dirmeier/surjectors v0.3.4 runs for
`TransformedDistribution.log_prob` through a `Chain`, *including* the parts of
the arithmetic that live in un-vendored third-party packages:

  * distrax 0.1.7      -> oracle/distrax.py  (ScalarAffine, Inverse,
                          RationalQuadraticSpline, Normal, Independent)
  * tfp-nightly 0.26   -> oracle/tfp.py      (_make_dense_autoregressive_masks)
  * dm-haiku 0.0.16    -> oracle/hk.py       (transform/init/apply, param
                          naming, Linear, nets.MLP, Sequential, Reshape)
  * surjectors itself  -> oracle/surjectors.py (every layer on the path; each
                          function cites the reference file:line it follows)

Gradients come from torch autograd over the float64 restatement (the reference
uses `jax.value_and_grad`, examples/coupling_inference_surjection.py:87 -- no
custom VJP exists upstream).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and
`bench.py` (`cpu_baseline` leg / `--impl reference`).  The product package
`surjectors_b200` never imports it and has no CPU fallback.

PARITY PINNING.  The JAX stack is not installable in this image, so the
restatement cannot be run against the real library.  It is pinned against the
only numeric known-answer values the reference's own test-suite holds -- the
8 KATs of `made_test.py:43-44,76-77` and `masked_autoregressive_test.py:113-114,
149-150` (tests/test_oracle_kat.py) -- which cover the TFP mask construction,
MADE wiring (context path, [D,P] reshape order), `unstack` layout, ScalarAffine
forward and the MAF forward recursion.  Everything else (RationalQuadraticSpline,
ScalarAffine inverse, MaskedCoupling masking, Normal/Independent log_prob,
LULinear, Permutation, Slice, the funnels, Chain, all gradients) is **parity
unpinned** by upstream fixtures; those parts are checked only by closed-form
identities and self-consistency (tests/test_oracle_properties.py).
"""

"""Host-side mirror of the reference interface: constructor signatures, haiku parameter names and
shapes, masks, error behaviour and the reference's quirks -- written after the reference's own
tests (citations: /root/reference/surjectors/_src/...).  Runs without a GPU: `init` and plan
compilation need no device; `apply` on CPU tensors must fail loudly."""

import numpy as np
import pytest
import torch

import surjectors_b200 as sj
from surjectors_b200 import _abi, _runtime, fused_chains, configs, hk
from surjectors_b200.compat import distrax, jnp
from surjectors_b200.nn import MADE, make_mlp
from surjectors_b200.util import make_alternating_binary_mask, unstack


def _base(n):
  return distrax.Independent(distrax.Normal(jnp.zeros(n), jnp.ones(n)), 1)


def test_public_names_match_reference_all():
  # surjectors/__init__.py:45-68 minus what SURVEY.md section 2 marks out of scope
  for name in ["Chain", "Permutation", "TransformedDistribution", "AffineMaskedAutoregressive", "MaskedAutoregressive",
               "MaskedAutoregressiveInferenceFunnel", "AffineMaskedAutoregressiveInferenceFunnel", "MaskedCoupling",
               "AffineMaskedCoupling", "RationalQuadraticSplineMaskedCoupling", "MaskedCouplingInferenceFunnel",
               "AffineMaskedCouplingInferenceFunnel", "RationalQuadraticSplineMaskedCouplingInferenceFunnel", "LULinear",
               "Slice", "ScalarAffine", "RationalQuadraticSpline", "Inverse"]:
    assert hasattr(sj, name), name
    assert name in sj.__all__


def test_alternating_mask_bit_exact():  # util.py:30-47
  m = make_alternating_binary_mask(5)
  assert m.dtype == bool and m.tolist() == [False, True, False, True, False]
  assert make_alternating_binary_mask(4, True).tolist() == [True, False, True, False]


def test_made_param_names_and_shapes():
  # masked_autoregressive_test.py:105-107: params["made/~/masked_linear_1"]["w"].shape == (8, n_dim * n_params)
  n_dim, n_params = 2, 2

  def fn(method, **kw):
    layer = sj.MaskedAutoregressive(
      conditioner=MADE(n_dim, [8, 8], n_params),
      bijector_fn=lambda p: distrax.ScalarAffine(*(lambda a, b: (a, jnp.exp(b)))(*unstack(p, -1))),
    )
    return sj.TransformedDistribution(_base(n_dim), layer)(method=method, **kw)

  m = hk.transform(fn)
  y = torch.zeros((10, n_dim))
  params = m.init(0, method="log_prob", y=y)
  assert sorted(params) == ["made/~/masked_linear", "made/~/masked_linear_1", "made/~/masked_linear_2"]
  assert tuple(params["made/~/masked_linear"]["w"].shape) == (n_dim, 8)
  assert tuple(params["made/~/masked_linear_1"]["w"].shape) == (8, 8)
  assert tuple(params["made/~/masked_linear_2"]["w"].shape) == (8, n_dim * n_params)
  # conditional: the context Linear is created inside __call__ (made.py:80-84) -> "made/linear"
  params = m.init(0, method="log_prob", y=y, x=torch.zeros((10, 3)))
  assert tuple(params["made/linear"]["w"].shape) == (3, 8)
  assert list(params)[:2] == ["made/~/masked_linear", "made/linear"]


def test_mlp_and_lu_param_tree():
  def fn(method, **kw):
    t = sj.Chain([sj.Slice(5, configs._decoder_fn(5, (32, 32))), sj.LULinear(5)])
    return sj.TransformedDistribution(_base(5), t)(method=method, **kw)

  params = hk.transform(fn).init(1, method="log_prob", y=torch.zeros((4, 10)))
  assert {k: tuple(v.shape) for k, v in params["lu_linear"].items()} == {
    "lower_entries": (10,), "upper_entries": (10,), "diag_entries": (5,)}  # lu_linear.py:58-66
  assert torch.all(params["lu_linear"]["diag_entries"] == 1) and torch.all(params["lu_linear"]["lower_entries"] == 0)
  assert tuple(params["mlp/~/linear_0"]["w"].shape) == (5, 32) and tuple(params["mlp/~/linear_2"]["b"].shape) == (10,)
  assert torch.all(params["mlp/~/linear_0"]["b"] == 0)  # make_mlp: zero bias, TruncatedNormal(0.01) weights
  w = params["mlp/~/linear_1"]["w"]
  assert float(w.abs().max()) <= 0.02 + 1e-6 and float(w.std()) > 0.005
  # every tensor is a view into one flat, 16-byte aligned buffer
  base = params.flat.data_ptr()
  for mod in params.values():
    for t in mod.values():
      assert (t.data_ptr() - base) % 16 == 0 and 0 <= t.data_ptr() - base < params.flat.numel() * 4


def test_constructor_errors_match_reference():
  def run(build):
    def fn(method, **kw):
      return sj.TransformedDistribution(_base(4), build())(method=method, **kw)

    return hk.transform(fn).init(0, method="log_prob", y=torch.zeros((2, 4)))

  with pytest.raises(ValueError, match="conditioner should be a MADE"):  # masked_autoregressive.py:54-57
    run(lambda: sj.MaskedAutoregressive(conditioner=make_mlp([4, 8]), bijector_fn=lambda p: None))
  with pytest.raises(ValueError, match="event_ndims"):  # masked_autoregressive.py:49-53
    run(lambda: sj.MaskedAutoregressive(MADE(4, [8], 2), lambda p: None, event_ndims=0, inner_event_ndims=1))
  with pytest.raises(NotImplementedError):  # lu_linear.py:44-47
    run(lambda: sj.LULinear(4, with_bias=True))
  with pytest.raises(ValueError, match="bool"):  # distrax.MaskedCoupling: mask must be boolean
    run(lambda: sj.AffineMaskedCoupling(np.array([0, 1, 0, 1]), make_mlp([8, 8])))
  with pytest.raises(NotImplementedError):  # unknown inner bijector: no fallback path
    run(lambda: sj.MaskedCoupling(make_alternating_binary_mask(4), make_mlp([8, 8]), lambda p: object()))
  with pytest.raises(ValueError):  # conditioner output width does not fit the bijector_fn
    run(lambda: sj.AffineMaskedCoupling(make_alternating_binary_mask(4), make_mlp([8, 7])))
  with pytest.raises(NotImplementedError):  # non-standard base distribution
    def fn(method, **kw):
      base = distrax.Independent(distrax.Normal(jnp.zeros(4), 2 * jnp.ones(4)), 1)
      return sj.TransformedDistribution(base, sj.LULinear(4))(method=method, **kw)
    hk.transform(fn).init(0, method="log_prob", y=torch.zeros((2, 4)))


def test_rank_and_batch_checks():  # transformed_distribution.py:83-85
  def fn(method, **kw):
    return sj.TransformedDistribution(_base(4), sj.LULinear(4))(method=method, **kw)

  m = hk.transform(fn)
  with pytest.raises(AssertionError):
    m.init(0, method="log_prob", y=torch.zeros((2, 4)), x=torch.zeros((3, 2)))
  with pytest.raises(AssertionError):
    m.init(0, method="log_prob", y=torch.zeros((2, 4)), x=torch.zeros((2,)))


def test_quirk_q1_bare_coupling_ignores_x():
  # transformed_distribution.py:87-88: a distrax.Bijector transform is called WITHOUT x, so the
  # conditional smoke test (affine_masked_coupling_test.py:48-55) initialises an unconditional conditioner
  def fn(method, **kw):
    layer = sj.AffineMaskedCoupling(make_alternating_binary_mask(2), make_mlp([8, 8, 4]))
    return sj.TransformedDistribution(_base(2), layer)(method=method, **kw)

  p = hk.transform(fn).init(0, method="log_prob", y=torch.zeros((10, 2)), x=torch.zeros((10, 2)))
  assert tuple(p["mlp/~/linear_0"]["w"].shape) == (2, 8)

  def fn_chain(method, **kw):
    layer = sj.AffineMaskedCoupling(make_alternating_binary_mask(2), make_mlp([8, 8, 4]))
    return sj.TransformedDistribution(_base(2), sj.Chain([layer]))(method=method, **kw)

  p = hk.transform(fn_chain).init(0, method="log_prob", y=torch.zeros((10, 2)), x=torch.zeros((10, 2)))
  assert tuple(p["mlp/~/linear_0"]["w"].shape) == (4, 8)  # inside a Chain x is honoured (chain.py:36-41)


def test_quirk_q1_is_one_sided_sampling_conditions_on_x():
  # transformed_distribution.py:137: sample_and_log_prob always calls transform.forward_and_likelihood_contribution(z, x=x),
  # so the SAME bare coupling that ignores x in log_prob (:87-88) does condition on x when sampling
  def fn(method, **kw):
    layer = sj.AffineMaskedCoupling(make_alternating_binary_mask(2), make_mlp([8, 8, 4]))
    return sj.TransformedDistribution(_base(2), layer)(method=method, **kw)

  p = hk.transform(fn).init(0, method="sample_and_log_prob", x=torch.zeros((10, 3)))
  assert tuple(p["mlp/~/linear_0"]["w"].shape) == (2 + 3, 8)
  traced, _ = hk.transform(fn).trace(method="sample_and_log_prob", x=torch.zeros((10, 3)))
  assert traced[0].dim_x == 3


def test_smoke_shapes_of_reference_tests_compile():
  """The constructor call patterns of the reference's smoke tests all trace and compile to a plan."""
  def rq_cond(d, k):
    return hk.Sequential([hk.nets.MLP([8, 8], activate_final=True), hk.Linear(d * (3 * k + 1)),
                          hk.Reshape((d, 3 * k + 1))])  # rq_masked_coupling_inference_funnel_test.py:34-47

  builders = {
    "rq_coupling": (2, lambda: sj.RationalQuadraticSplineMaskedCoupling(make_alternating_binary_mask(2), rq_cond(2, 4), -2.0, 2.0)),
    "slice": (10, lambda: sj.Slice(3, configs._decoder_fn(7))),
    "affine_funnel": (4, lambda: sj.AffineMaskedCouplingInferenceFunnel(2, configs._decoder_fn(2), make_mlp([4, 4, 8]))),
    "rq_funnel": (4, lambda: sj.RationalQuadraticSplineMaskedCouplingInferenceFunnel(2, configs._decoder_fn(2), rq_cond(4, 1), -1.0, 1.0)),
    "ar_funnel": (4, lambda: sj.AffineMaskedAutoregressiveInferenceFunnel(2, configs._decoder_fn(2), MADE(2, [8, 8], 2))),
    "maf": (3, lambda: sj.AffineMaskedAutoregressive(MADE(3, [8, 8], 2))),
    "perm": (3, lambda: sj.Permutation(np.array([2, 0, 1]), 1)),
  }
  for name, (dim, build) in builders.items():
    def fn(method, **kw):
      t = sj.Chain([build()])
      return sj.TransformedDistribution(_base(t._transforms[0]._trace(dim, kw["x"].shape[-1] if "x" in kw else 0)[1]), t)(
        method=method, **kw)

    def fn2(method, **kw):
      layer = build()
      d_out = {"slice": 3, "affine_funnel": 2, "rq_funnel": 2, "ar_funnel": 2}.get(name, dim)
      return sj.TransformedDistribution(_base(d_out), sj.Chain([layer]))(method=method, **kw)

    for batch in (dict(y=torch.zeros((10, dim))), dict(y=torch.zeros((10, dim)), x=torch.zeros((10, 2)))):
      traced, layout = hk.transform(fn2).trace(method="log_prob", **batch)
      plan = _runtime.get_plan(traced[0])
      assert plan.executor in ("fused", "layered", "gen1", "compact"), name
      assert plan.dim_in == dim


def test_every_baseline_config_compiles_to_a_plan():
  for name, c in configs.CONFIGS.items():
    spec = fused_chains.spec_of(c["make"](), c["dim"], c["dim_x"])
    plan = _runtime.get_plan(spec)
    want = {"cpl": "compact", "cfg1": "gen1", "cfg2": "gen1", "cfg3": "gen1", "cfg4": "layered", "cfg5": "layered", "mlpf": "fused", "cond": "fused"}[name]
    assert plan.executor == want, (name, plan.executor)
    assert plan.launches(False) >= 1 and plan.launches(True) >= 2


def test_mlp_inference_funnel_param_tree():
  m = hk.transform(configs.CONFIGS["mlpf"]["make"](n_lu=1))
  p = m.init(0, method="log_prob", y=torch.zeros((2, 20)))
  assert tuple(p["mlp_inference_funnel/~/lu_linear"]["diag_entries"].shape) == (10,)
  assert tuple(p["mlp_inference_funnel/~/linear"]["w"].shape) == (10, 10)
  assert tuple(p["mlp_inference_funnel/~/linear"]["b"].shape) == (10,)
  assert "lu_linear" in p and "lu_linear_1" in p


def test_no_cpu_fallback():
  c = configs.CONFIGS["cfg1"]
  m = hk.transform(c["make"]())
  y = torch.zeros((4, 10))
  params = m.init(0, method="log_prob", y=y)
  with pytest.raises(_abi.SurjError, match="no CPU fallback"):
    m.apply(params, None, method="log_prob", y=y)
  with pytest.raises(_abi.SurjError, match="no CPU fallback"):  # the sampling direction has no CPU path either
    m.apply(params, 0, method="sample", sample_shape=(2,))


def test_bad_descriptor_is_rejected_by_the_c_abi():
  c = configs.CONFIGS["cfg2"]
  spec = fused_chains.spec_of(c["make"](), c["dim"], c["dim_x"])
  spec.layers[1].dim_in = 11  # the chain delivers 10
  with pytest.raises(ValueError, match="delivers"):
    _runtime.Plan(spec)


def test_build_fused_kernel_for_a_user_chain(tmp_path):
  """A chain that is not a BASELINE config gets its own generated kernel at the user's request (nvcc, no GPU needed)."""
  def flow(method, **kw):
    t = sj.Chain([sj.Slice(3, configs._decoder_fn(4, (6, 6))), sj.LULinear(3)])
    return sj.TransformedDistribution(_base(3), t)(method=method, **kw)

  m = hk.transform(flow)
  batch = dict(y=torch.zeros((2, 7)))
  traced, _ = m.trace(method="log_prob", **batch)
  assert _runtime.Plan(traced[0]).executor == "fused"
  fused_chains.build_fused_kernel(m, batch, cache_dir=str(tmp_path))
  assert _runtime.Plan(traced[0]).executor == "gen1"


def test_forced_executors_refuse_chains_they_do_not_cover():
  """Forcing an executor never falls back: the compact kernel's family is affine coupling / coupling-funnel chains with
  two-equal-hidden-layer MLPs, the generated kernels exist only for the compiled-in chains (plan creation is host-only)."""
  for name, mode, ok in (("cfg2", "compact", True), ("cpl", "compact", True), ("cfg3", "compact", False),
                         ("cfg4", "compact", False), ("cond", "compact", False), ("cpl", "gen1", False),
                         ("cfg2", "gen1", True)):
    c = configs.CONFIGS[name]
    spec = fused_chains.spec_of(c["make"](), c["dim"], c["dim_x"])
    if ok:
      assert _runtime.Plan(spec, mode=_runtime.MODES[mode]).executor == mode, (name, mode)
    else:
      with pytest.raises(NotImplementedError):  # SURJ_E_UNSUPPORTED, as the reference raises for unsupported structure
        _runtime.Plan(spec, mode=_runtime.MODES[mode])


def test_sampling_recursion_is_recognised():
  """masked_autoregressive.py:63-71 (D MADE passes upstream): the per-op executor must run every affine MaskedAutoregressive
  of a chain as ONE incremental kernel (csrc/vm_maf.cuh) -- a host-side property of the compiled plan."""
  import numpy as np

  import surjectors_b200 as sj
  from surjectors_b200 import _abi, _runtime
  from surjectors_b200.nn import MADE

  for dim, hidden, dx in ((128, [512, 512], 16), (24, [48], 0), (6, [16, 16, 16], 3)):
    def build():
      return sj.Chain([sj.AffineMaskedAutoregressive(MADE(dim, hidden, 2)), sj.Permutation(np.arange(dim)[::-1].copy(), 1),
                       sj.AffineMaskedAutoregressive(MADE(dim, hidden, 2))])

    fwd = hk.transform(lambda **kw: build().forward_and_likelihood_contribution(**kw))
    kw = dict(z=torch.zeros(2, dim))
    if dx:
      kw["x"] = torch.zeros(2, dx)
    traced, _ = fwd.trace(**kw)
    plan = _runtime.Plan(traced[0], _runtime.BASE_NONE, _abi.MODE_LAYERED)
    assert _abi.lib().surj_plan_sampling_recursions(plan._h) == 2, (dim, hidden, dx)

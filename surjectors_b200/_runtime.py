"""Plans, launches and autograd glue above the C ABI (include/surj_b200.h).

torch is used for what the prompt calls plumbing: device memory, the current CUDA stream and
autograd bookkeeping (`torch.autograd.Function` plays the role `jax.custom_vjp` around a
`jax.ffi.ffi_call` would play in the reference).  All arithmetic is in libsurj_b200.so; there is
no CPU or eager path -- a call on CPU tensors, or with the library missing, raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from surjectors_b200 import _abi, _sym, hk
from surjectors_b200._spec import ChainSpec

BASE_STANDARD_NORMAL, BASE_NONE = 0, 1

_plans = {}
_plans_lock = threading.Lock()
_default_mode = _abi.MODE_AUTO

MODES = {"auto": _abi.MODE_AUTO, "fused": _abi.MODE_FUSED, "layered": _abi.MODE_LAYERED, "gen1": _abi.MODE_FUSED_GEN1,
         "compact": _abi.MODE_FUSED_COMPACT}
if os.environ.get("SURJ_B200_MODE"):  # A/B measurements of the executors without touching the caller's code
  _default_mode = MODES[os.environ["SURJ_B200_MODE"]]


class execution_mode:
  """Context manager selecting which executor new plans use: "auto" (default), "fused" (the
  in-shared-memory interpreter kernel), "layered" (one kernel per op out of global memory) or "gen1"
  (the build-time generated kernel of a BASELINE config).  For tests and A/B measurements."""

  def __init__(self, mode):
    self.mode = MODES[mode] if isinstance(mode, str) else int(mode)

  def __enter__(self):
    global _default_mode
    self._prev, _default_mode = _default_mode, self.mode
    return self

  def __exit__(self, *a):
    global _default_mode
    _default_mode = self._prev


def _on(t):
  """Make the device of `t` current for the duration of an ABI call: the C side launches on, allocates plan tables
  on and checks against the CURRENT device (a plan is bound to the device of its first launch)."""
  return torch.cuda.device(t.device)


class chunk_samples:
  """Context manager: samples per chunk of the per-op executor (tests run multi-chunk cases at small batches)."""

  def __init__(self, samples):
    self.samples = int(samples)

  def __enter__(self):
    self._prev = int(_abi.lib().surj_set_chunk_samples(self.samples))
    return self

  def __exit__(self, *a):
    _abi.lib().surj_set_chunk_samples(self._prev)


class Plan:
  """Owns a `SurjPlan*`: the compiled description of one chain (what jax.jit's trace cache holds).  One Plan per
  (chain, device): the C object keeps device-resident tables."""

  def __init__(self, spec: ChainSpec, base=BASE_STANDARD_NORMAL, mode=_abi.MODE_AUTO):
    lib = _abi.lib()
    arr, keep = spec.to_c()
    handle = C.c_void_p()
    st = lib.surj_plan_create(arr, len(spec.layers), spec.dim_in, spec.dim_x, spec.n_params, mode, base,
                              C.byref(handle))
    del keep
    _abi.check(st, "surj_plan_create")
    self._lib = lib
    self._h = handle
    self.spec = spec
    self.base = base
    self.dim_in, self.dim_x, self.dim_out, self.n_params = spec.dim_in, spec.dim_x, spec.dim_out, spec.n_params
    self.executor = {0: "layered", 1: "gen1", 2: "fused", 3: "compact"}[int(lib.surj_plan_is_fused(handle))]
    self.is_fused = self.executor != "layered"
    self.key = int(lib.surj_plan_key(handle))
    self._ws = {}

  def __del__(self):
    h, self._h = getattr(self, "_h", None), None
    if h:
      try:
        self._lib.surj_plan_destroy(h)
      except Exception:  # noqa: BLE001  (interpreter shutdown)
        pass

  def launches(self, backward=False):
    return int(self._lib.surj_plan_launches(self._h, int(backward)))

  def workspace(self, batch, backward, device):
    """Cached scratch buffer (the caller of the C ABI owns all device memory)."""
    n = int(self._lib.surj_plan_workspace_bytes(self._h, batch, int(backward)))
    key = (batch, bool(backward), str(device))
    buf = self._ws.get(key)
    if buf is None or buf.numel() < n:
      buf = torch.empty(max(n, 16), dtype=torch.uint8, device=device)
      self._ws[key] = buf
    return buf, n

  # -- raw launches on the current stream; all arguments are contiguous float32 CUDA tensors
  def fwd(self, y, x, flat, want_z=True, lp_out=None, z_out=None):
    B = y.shape[0]
    _check_inputs(self, y, x, flat)
    lp = lp_out if lp_out is not None else torch.empty(B, dtype=torch.float32, device=y.device)
    z = z_out
    if z is None and want_z:
      z = torch.empty((B, self.dim_out), dtype=torch.float32, device=y.device)
    if B == 0:
      return z, lp
    with _on(y):
      ws, n = self.workspace(B, False, y.device)
      st = self._lib.surj_chain_logprob_fwd(
        self._h, _stream(y), y.data_ptr(), _ptr(x), flat.data_ptr(), _ptr(z), lp.data_ptr(), ws.data_ptr(), n, B)
    _abi.check(st, "surj_chain_logprob_fwd")
    return z, lp

  def bwd(self, y, x, flat, g_lp=None, g_scale=1.0, g_z=None, want_gy=False, want_gx=False, want_lp=False,
          g_params=None, loss_out=None):
    B = y.shape[0]
    _check_inputs(self, y, x, flat)
    dev = y.device
    gp = g_params if g_params is not None else torch.empty(self.n_params, dtype=torch.float32, device=dev)
    gy = torch.empty_like(y) if want_gy else None
    gx = torch.empty_like(x) if (want_gx and x is not None) else None
    lp = torch.empty(B, dtype=torch.float32, device=dev) if want_lp else None
    if B == 0:
      gp.zero_()
      return gp, gy, gx, lp
    with _on(y):
      ws, n = self.workspace(B, True, dev)
      st = self._lib.surj_chain_logprob_bwd(
        self._h, _stream(y), y.data_ptr(), _ptr(x), flat.data_ptr(), _ptr(g_lp), float(g_scale), _ptr(g_z),
        gp.data_ptr(), _ptr(loss_out), _ptr(lp), _ptr(gy), _ptr(gx), ws.data_ptr(), n, B)
    _abi.check(st, "surj_chain_logprob_bwd")
    return gp, gy, gx, lp


def _plan_forward(self, z, x, flat, eps=None):
  """surj_chain_forward: (y, summed forward contribution) for latents z (sampling direction); eps[B, noise_dim]
  are the standard-normal draws the decoders of Slice / funnel layers consume."""
  B = z.shape[0]
  if not z.is_cuda:
    raise _abi.SurjError("surjectors_b200 computes on CUDA tensors only (there is no CPU fallback)")
  if z.dim() != 2 or z.shape[1] != self.dim_out or z.dtype != torch.float32 or not z.is_contiguous():
    raise ValueError(f"z must be a contiguous float32 [B, {self.dim_out}] tensor, got {tuple(z.shape)} {z.dtype}")
  y = torch.empty((B, self.dim_in), dtype=torch.float32, device=z.device)
  ld = torch.empty(B, dtype=torch.float32, device=z.device)
  if B == 0:
    return y, ld
  with _on(z):
    n = int(self._lib.surj_plan_forward_workspace_bytes(self._h, B))
  key = ("gen", B, str(z.device))
  ws = self._ws.get(key)
  if ws is None or ws.numel() < n:
    ws = torch.empty(max(n, 16), dtype=torch.uint8, device=z.device)
    self._ws[key] = ws
  n_eps = self.noise_dim()
  if n_eps > 0 and (eps is None or tuple(eps.shape) != (B, n_eps) or eps.dtype != torch.float32 or not eps.is_contiguous()
                    or eps.device != z.device):
    raise ValueError(f"eps must be a contiguous float32 [B, {n_eps}] tensor on the device of z")
  with _on(z):
    st = self._lib.surj_chain_forward(self._h, _stream(z), z.data_ptr(), _ptr(x), _ptr(eps) if n_eps else None,
                                      flat.data_ptr(), y.data_ptr(), ld.data_ptr(), ws.data_ptr(), n, B)
  _abi.check(st, "surj_chain_forward")
  return y, ld


Plan.forward = _plan_forward
Plan.noise_dim = lambda self: int(self._lib.surj_plan_noise_dim(self._h))


def _plan_train_step(self, y, x, flat, g_scale, grads, opt, loss_out):
  """surj_train_step: fused log_prob+VJP kernel, gradient reduction with the Adam update in its last stage."""
  B = y.shape[0]
  _check_inputs(self, y, x, flat)
  with _on(y):
    ws, n = self.workspace(B, True, y.device)
    st = self._lib.surj_train_step(
      self._h, _stream(y), y.data_ptr(), _ptr(x), flat.data_ptr(), float(g_scale), grads.data_ptr(), opt.mu.data_ptr(),
      opt.nu.data_ptr(), opt.count.data_ptr(), opt.lr, opt.b1, opt.b2, opt.eps, _ptr(loss_out), ws.data_ptr(), n, B)
  _abi.check(st, "surj_train_step")


Plan.train_step = _plan_train_step


def _plan_train_step_dp(self, dp, y, x, flat, g_scale, grads, opt, loss_out):
  """surj_train_step_dp: the same two launches with the sum over ranks fused in (NVLink peer mailboxes)."""
  B = y.shape[0]
  _check_inputs(self, y, x, flat)
  with _on(y):
    ws, n = self.workspace(B, True, y.device)
    st = self._lib.surj_train_step_dp(
      self._h, dp.handle, _stream(y), y.data_ptr(), _ptr(x), flat.data_ptr(), float(g_scale), grads.data_ptr(),
      opt.mu.data_ptr(), opt.nu.data_ptr(), opt.count.data_ptr(), opt.lr, opt.b1, opt.b2, opt.eps, _ptr(loss_out),
      ws.data_ptr(), n, B)
  _abi.check(st, "surj_train_step_dp")


Plan.train_step_dp = _plan_train_step_dp


def _ptr(t):
  return None if t is None else t.data_ptr()


def _stream(t):
  return torch.cuda.current_stream(t.device).cuda_stream


def _check_inputs(plan, y, x, flat):
  if not y.is_cuda:
    raise _abi.SurjError("surjectors_b200 computes on CUDA tensors only (there is no CPU fallback)")
  if y.dim() != 2 or y.shape[1] != plan.dim_in or y.dtype != torch.float32 or not y.is_contiguous():
    raise ValueError(f"y must be a contiguous float32 [B, {plan.dim_in}] tensor, got {tuple(y.shape)} {y.dtype}")
  if plan.dim_x > 0:
    if x is None or x.dim() != 2 or x.shape != (y.shape[0], plan.dim_x) or x.dtype != torch.float32 \
        or not x.is_contiguous() or x.device != y.device:
      raise ValueError(f"x must be a contiguous float32 [B, {plan.dim_x}] tensor on the device of y")
  if flat.numel() != plan.n_params or flat.dtype != torch.float32 or flat.device != y.device \
      or not flat.is_contiguous():
    raise ValueError("flat parameter buffer does not match the plan")


def get_plan(spec: ChainSpec, base=BASE_STANDARD_NORMAL, mode=None, device=None) -> Plan:
  """The plan of (chain, base, executor) ON `device`: the C plan keeps device-resident tables and per-device launch
  facts, so the cache is keyed by the device index as well (the same chain on cuda:1 gets its own plan)."""
  mode = _default_mode if mode is None else mode
  dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
  idx = dev.index if dev.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
  key = (spec.structure(), base, mode, idx)
  with _plans_lock:
    plan = _plans.get(key)
    if plan is None:
      plan = Plan(spec, base, mode)
      _plans[key] = plan
    return plan


class _LogProbFn(torch.autograd.Function):
  """(y, x, flat) -> (z, lp) with the hand-written VJP kernel as backward."""

  @staticmethod
  def forward(ctx, plan, y, x, flat):
    z, lp = plan.fwd(y, x, flat, want_z=True)
    ctx.plan = plan
    ctx.save_for_backward(y, x, flat)
    return z, lp

  @staticmethod
  def backward(ctx, g_z, g_lp):
    plan = ctx.plan
    y, x, flat = ctx.saved_tensors
    need_y, need_x = ctx.needs_input_grad[1], ctx.needs_input_grad[2] and x is not None
    if g_lp is None:
      g_lp = torch.zeros(y.shape[0], dtype=torch.float32, device=y.device)
    g_lp = g_lp.contiguous()
    if g_z is not None:
      g_z = g_z.contiguous()
    gp, gy, gx, _ = plan.bwd(y, x, flat, g_lp=g_lp, g_z=g_z, want_gy=need_y, want_gx=need_x)
    return None, gy, gx, (gp if ctx.needs_input_grad[3] else None)


def _prep(t, what):
  if t is None:
    return None
  if not isinstance(t, torch.Tensor):
    raise TypeError(f"{what} must be a torch tensor (device memory is managed through torch), got {type(t).__name__}")
  if t.dtype != torch.float32:
    t = t.to(torch.float32)
  return t.contiguous()


def run_transform(transform, y, x, honour_x=True, base=None):
  """Trace `transform` in the current hk frame and run it: returns (z, lp-or-lc)."""
  from surjectors_b200 import _layers

  fr = hk.current_frame()
  y = _prep(y, "y")
  x = _prep(x, "x")
  if y.dim() != 2:  # lu_linear.py:89 / funnels :87,90 index [:, :k]: the path is defined for [B, D]
    raise ValueError(f"y must have shape [batch, event], got {tuple(y.shape)}")
  dx = x.shape[-1] if (x is not None and honour_x) else 0
  spec = _layers.trace_chain(transform, y.shape[-1], dx, honour_x=True)
  if base is not None:
    _sym.check_standard_normal_base(base, spec.dim_out)
  if fr.is_init:
    fr.device = y.device
    B = y.shape[0]
    return (torch.zeros((B, spec.dim_out), dtype=torch.float32, device=y.device),
            torch.zeros(B, dtype=torch.float32, device=y.device))
  if not y.is_cuda:
    raise _abi.SurjError("surjectors_b200 computes on CUDA tensors only (there is no CPU fallback)")
  plan = get_plan(spec, BASE_STANDARD_NORMAL if base is not None else BASE_NONE, device=y.device)
  flat = hk.pack_params(fr.params, fr.layout, y.device)
  xx = x if dx > 0 else None
  if torch.is_grad_enabled() and (flat.requires_grad or y.requires_grad or (xx is not None and xx.requires_grad)):
    return _LogProbFn.apply(plan, y, xx, flat)
  return plan.fwd(y, xx, flat, want_z=True)


def run_forward(transform, z, x, honour_x=True, eps=None):
  """Sampling direction: trace `transform` (its structure is direction-independent) and run
  forward_and_likelihood_contribution: returns (y, forward contribution).  Decoder noise `eps` defaults to
  standard-normal draws from the haiku rng of this apply (the reference draws with hk.next_rng_key())."""
  from surjectors_b200 import _layers

  fr = hk.current_frame()
  z = _prep(z, "z")
  x = _prep(x, "x")
  if z.dim() != 2:
    raise ValueError(f"z must have shape [batch, event], got {tuple(z.shape)}")
  dx = x.shape[-1] if (x is not None and honour_x) else 0
  d_in = _layers.input_dim_for_latent(transform, z.shape[-1], dx)
  spec = _layers.trace_chain(transform, d_in, dx)
  if spec.dim_out != z.shape[-1]:
    raise ValueError(f"latent has {z.shape[-1]} dims but the flow's latent space has {spec.dim_out}")
  if fr.is_init:
    fr.device = z.device
    B = z.shape[0]
    return (torch.zeros((B, d_in), dtype=torch.float32, device=z.device),
            torch.zeros(B, dtype=torch.float32, device=z.device))
  if not z.is_cuda:
    raise _abi.SurjError("surjectors_b200 computes on CUDA tensors only (there is no CPU fallback)")
  plan = get_plan(spec, BASE_NONE, device=z.device)
  flat = hk.pack_params(fr.params, fr.layout, z.device)
  n_eps = plan.noise_dim()
  if n_eps > 0 and eps is None:
    rng = hk.next_rng_key()
    if rng is None:
      raise ValueError("sampling through a surjection draws decoder noise: pass an rng to apply")
    g = torch.Generator(device="cpu").manual_seed(int(rng.integers(0, 2**62)))
    eps = torch.randn((z.shape[0], n_eps), generator=g, dtype=torch.float32).to(z.device)
  return plan.forward(z, x if dx > 0 else None, flat, _prep(eps, "eps") if n_eps else None)


def value_and_grad(fun, has_aux=False):
  """`jax.value_and_grad(fun)` for loss functions of a params tree built from `model.apply`
  (examples/coupling_inference_surjection.py:83-87).  Gradients come from the VJP kernels."""

  def wrapped(params, *args, **kwargs):
    layout = getattr(params, "layout", None)
    if layout is not None and isinstance(getattr(params, "flat", None), torch.Tensor):
      leaf = params.flat.detach().requires_grad_(True)
      tree = hk.tree_from_flat(leaf, layout)
      with torch.enable_grad():
        out = fun(tree, *args, **kwargs)
        loss = out[0] if has_aux else out
        (g,) = torch.autograd.grad(loss, [leaf], allow_unused=True)
      if g is None:
        g = torch.zeros_like(leaf)
      grads = hk.tree_from_flat(g, layout)
    else:
      leaves = {m: {k: v.detach().requires_grad_(True) for k, v in b.items()} for m, b in params.items()}
      with torch.enable_grad():
        out = fun(leaves, *args, **kwargs)
        loss = out[0] if has_aux else out
        flat = [v for b in leaves.values() for v in b.values()]
        gs = torch.autograd.grad(loss, flat, allow_unused=True)
      it = iter(gs)
      grads = {m: {k: (lambda g, v: torch.zeros_like(v) if g is None else g)(next(it), v) for k, v in b.items()}
               for m, b in leaves.items()}
    if has_aux:
      return (loss.detach(), out[1]), grads
    return loss.detach(), grads

  return wrapped

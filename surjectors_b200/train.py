"""The training step of the reference's examples as device-side work only:
`loss, grads = jax.value_and_grad(loss_fn)(params)` with `loss = -sum(lp)` followed by
`optax.adam(lr)` (examples/coupling_inference_surjection.py:78-90), i.e. fused log_prob+VJP kernel,
gradient reduction, (data-parallel: one NCCL all-reduce of the flat gradient) and a fused Adam
update on the flat parameter buffer.  Enqueue-only, CUDA-graph capturable."""

from __future__ import annotations

import ctypes as C

import torch

from surjectors_b200 import _abi, _runtime, hk, parallel


class Adam:
  """optax.adam(lr, b1=0.9, b2=0.999, eps=1e-8) state over one flat buffer (optax 0.2.4 semantics:
  bias-corrected moments, eps outside the sqrt, eps_root = 0)."""

  def __init__(self, flat: torch.Tensor, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    self.lr, self.b1, self.b2, self.eps = float(lr), float(b1), float(b2), float(eps)
    self.mu = torch.zeros_like(flat)
    self.nu = torch.zeros_like(flat)
    self.count = torch.zeros(1, dtype=torch.int32, device=flat.device)

  def update_(self, flat, grads, grad_scale=1.0):
    st = _abi.lib().surj_adam_step(
      torch.cuda.current_stream(flat.device).cuda_stream, flat.data_ptr(), grads.data_ptr(), self.mu.data_ptr(),
      self.nu.data_ptr(), self.count.data_ptr(), flat.numel(), self.lr, self.b1, self.b2, self.eps, float(grad_scale))
    _abi.check(st, "surj_adam_step")


class TrainStep:
  """One optimisation step on a batch shard.  With a process group the flat gradient (and the
  loss) are summed over ranks before the update, which is the whole multi-GPU story of this
  path: samples are independent, parameters are replicated."""

  def __init__(self, model: hk.Transformed, params, example_batch, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8,
               reduction="sum", process_group=None, method="log_prob"):
    if not isinstance(getattr(params, "flat", None), torch.Tensor):
      raise ValueError("params must be the tree returned by model.init (views into one flat buffer)")
    traced, layout = model.trace(method=method, **{k: v[:1].cpu() for k, v in example_batch.items() if v is not None})
    if len(traced) != 1:
      raise ValueError("the model function must evaluate exactly one flow log_prob")
    if layout.signature() != params.layout.signature():
      raise ValueError("params do not belong to this model")
    self.spec = traced[0]
    self.params = params
    self.flat = params.flat
    if not self.flat.is_cuda:
      raise _abi.SurjError("TrainStep needs parameters on a CUDA device (there is no CPU fallback)")
    self.plan = _runtime.get_plan(self.spec, device=self.flat.device)
    self.opt = Adam(self.flat, lr, b1, b2, eps)
    self.grads = torch.zeros(self.flat.numel() + 4, dtype=torch.float32, device=self.flat.device)
    self.reduction = reduction
    self.group = process_group
    self.world = 1
    self.peer = None
    if process_group is not None:
      import os

      import torch.distributed as dist

      self.world = dist.get_world_size(process_group)
      # KB-sized gradients: exchange fused into the reduction kernel over NVLink peer memory; large ones
      # (wide networks, per-op executor): one NCCL all-reduce.  SURJ_B200_EXCHANGE=nccl forces the latter.
      # The peer exchange needs every rank on one node with CUDA peer access to every other rank's GPU; the probe is
      # collective, so all ranks take the same branch (anything else falls back to the all-reduce).
      if (self.world > 1 and self.plan.is_fused and self.flat.numel() <= (1 << 16)
          and os.environ.get("SURJ_B200_EXCHANGE", "peer") != "nccl"
          and parallel.peer_access_everywhere(self.flat.device, process_group)):
        with torch.cuda.device(self.flat.device):
          self.peer = parallel.PeerExchange(self.flat.numel(), process_group)
    self._global_batch = {}

  @property
  def loss(self):
    """Device scalar: the (global) loss of the last step."""
    return self.grads[self.flat.numel():self.flat.numel() + 1]

  def launches_per_step(self):
    if self.plan.is_fused and (self.world == 1 or self.peer is not None):
      return 2
    return self.plan.launches(backward=True) + 2 + (1 if self.world > 1 else 0)

  def check(self):
    """Raise if the data-parallel exchange reported a peer that never arrived (non-blocking poll of a host-mapped
    flag).  A failed step leaves parameters, moments and the step counter untouched on this rank."""
    if self.peer is not None and self.peer.error_nosync():
      raise _abi.SurjError("data-parallel exchange timed out: a peer rank did not deliver its gradient "
                           "(no update was applied on this rank since); stop and restart from a checkpoint")

  def _mean_scale(self, B):
    """-1 / (global sample count): shards made by shard_bounds differ by up to one sample, so the local batch times
    the world size is not the global batch.  One all-reduce per distinct local batch size, cached."""
    if self.world == 1:
      return -1.0 / B
    if B not in self._global_batch:
      import torch.distributed as dist

      t = torch.tensor([B], dtype=torch.int64, device=self.flat.device)
      dist.all_reduce(t, group=self.group)
      self._global_batch[B] = int(t.item())
    return -1.0 / self._global_batch[B]

  def step(self, y, x=None):
    n = self.flat.numel()
    B = y.shape[0]
    self.check()
    g_scale = -1.0 if self.reduction == "sum" else self._mean_scale(B)
    xx = x if self.plan.dim_x > 0 else None
    if self.world == 1:  # the whole step is two launches: backward kernel, reduction + Adam
      self.plan.train_step(y, xx, self.flat, g_scale, self.grads, self.opt, self.grads[n:n + 1])
      return self.loss
    if self.peer is not None:  # same two launches; the sum over ranks happens inside the reduction kernel
      self.plan.train_step_dp(self.peer, y, xx, self.flat, g_scale, self.grads, self.opt, self.grads[n:n + 1])
      return self.loss
    self.plan.bwd(y, xx, self.flat, g_lp=None, g_scale=g_scale, g_params=self.grads, loss_out=self.grads[n:n + 1])
    if self.world > 1:
      parallel.allreduce_flat_(self.grads, self.group)  # the one collective of the path: flat gradient + loss
    self.opt.update_(self.flat, self.grads)
    return self.loss

  # ---- host-fed steps: the batch lives in (pinned) host memory, the loss is read back
  def _staging(self, y_host, x_host):
    shapes = (tuple(y_host.shape), None if x_host is None else tuple(x_host.shape))
    if getattr(self, "_stage", None) is None or self._stage_shapes != shapes:
      self._stage_shapes = shapes
      dev = self.flat.device
      self._copy_stream = torch.cuda.Stream(dev)
      self._stage = []
      for _ in range(2):
        ys = torch.empty(y_host.shape, dtype=torch.float32, device=dev)
        xs = None if x_host is None else torch.empty(x_host.shape, dtype=torch.float32, device=dev)
        self._stage.append((ys, xs, torch.cuda.Event(), torch.cuda.Event()))
      self._n_host = 0
    return self._stage[self._n_host % 2]

  def step_host(self, y_host, x_host=None, loss_out=None):
    """One step on a host batch: H2D copy on a side stream into one of two staging buffers (so the
    copy of batch i+1 overlaps the kernels of batch i), the step, and an async D2H copy of the loss
    into `loss_out` (a pinned host tensor) if given.  Enqueue-only; `drain()` blocks the host until every
    enqueued copy (the D2H loss reads included) has completed -- call it before reading `loss_out`."""
    ys, xs, ready, free = self._staging(y_host, x_host)
    cur = torch.cuda.current_stream(self.flat.device)
    if self._n_host >= 2:
      self._copy_stream.wait_event(free)
    with torch.cuda.stream(self._copy_stream):
      ys.copy_(y_host, non_blocking=True)
      if xs is not None:
        xs.copy_(x_host, non_blocking=True)
      ready.record(self._copy_stream)
    cur.wait_event(ready)
    loss = self.step(ys, xs)
    free.record(cur)
    if loss_out is not None:
      loss_out.copy_(loss, non_blocking=True)
    if getattr(self, "_done", None) is None:
      self._done = torch.cuda.Event()
    self._done.record(cur)  # after the step's kernels and its D2H loss copy
    self._n_host += 1
    return loss

  def drain(self):
    """Host-blocking: returns when every host-fed step enqueued so far has finished, its loss copy included."""
    if getattr(self, "_copy_stream", None) is not None:
      torch.cuda.current_stream(self.flat.device).wait_stream(self._copy_stream)
    if getattr(self, "_done", None) is not None:
      self._done.synchronize()
    self.check()

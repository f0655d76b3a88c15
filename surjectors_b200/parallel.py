"""Data parallelism of the path: samples are independent, so the batch is
cut into contiguous shards (one per GPU / process), parameters and optimizer state are replicated,
and the only collective is ONE all-reduce(sum) of the flat parameter-gradient buffer (with the loss
riding in its last slot) per training step.  log_prob needs no collective at all."""

import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
  """[lo, hi) of rank's contiguous shard; shard sizes differ by at most one sample."""
  base, rem = divmod(int(batch), int(world))
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat_(flat: torch.Tensor, group=None):
  """Sum the flat gradient (+ loss) buffer over the ranks of `group`, in place."""
  dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  return flat


def adam_reference_(p, g, mu, nu, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
  """The arithmetic of surj_adam_kernel (optax.adam semantics) on host tensors, in place; used to
  check that replicated updates stay bit-identical across ranks."""
  mu.mul_(b1).add_(g, alpha=1 - b1)
  nu.mul_(b2).addcmul_(g, g, value=1 - b2)
  c1, c2 = 1.0 / (1.0 - b1**t), 1.0 / (1.0 - b2**t)
  p.sub_(lr * (mu * c1) / (torch.sqrt(nu * c2) + eps))
  return p


def peer_access_everywhere(device, group=None):
  """True on every rank iff all ranks sit on one host and every rank's GPU can map every other rank's GPU
  (cudaDeviceCanAccessPeer).  Collective: all ranks must call it; all get the same answer."""
  import socket

  me = (socket.gethostname(), torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device())
  world = dist.get_world_size(group)
  everyone = [None] * world
  dist.all_gather_object(everyone, me, group=group)
  ok = len({h for h, _ in everyone}) == 1
  if ok:
    for _, idx in everyone:
      if idx != me[1] and not torch.cuda.can_device_access_peer(me[1], idx):
        ok = False
  flags = [None] * world
  dist.all_gather_object(flags, bool(ok), group=group)
  return all(flags)


class PeerExchange:
  """The fused alternative to the all-reduce for KB-sized gradients: every rank's reduced gradient is stored
  straight into its peers' mailboxes over NVLink peer memory by the gradient-reduction kernel itself, which
  then sums the mailbox and applies Adam (include/surj_b200.h: surj_dp_create / surj_dp_connect /
  surj_train_step_dp).  torch.distributed is only used once, to exchange the CUDA IPC handles."""

  def __init__(self, n_params, group=None):
    import ctypes as C

    from surjectors_b200 import _abi

    self._lib = _abi.lib()
    self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
    handle = C.create_string_buffer(64)
    self._h = C.c_void_p()
    _abi.check(self._lib.surj_dp_create(self.rank, self.world, int(n_params), C.byref(self._h), handle), "surj_dp_create")
    handles = [None] * self.world
    dist.all_gather_object(handles, bytes(handle.raw), group=group)
    _abi.check(self._lib.surj_dp_connect(self._h, b"".join(handles)), "surj_dp_connect")
    dist.barrier(group=group)

  @property
  def handle(self):
    return self._h

  def error(self):
    """Synchronising read of the device-side error flag."""
    return int(self._lib.surj_dp_error(self._h))

  def error_nosync(self):
    """Non-blocking poll (host-mapped flag): 1 once a peer missed the exchange timeout."""
    return int(self._lib.surj_dp_error_nosync(self._h))

  def set_timeout_ms(self, ms):
    """0 = wait for ever (what a NCCL collective does); default 30 s (SURJ_B200_DP_TIMEOUT_MS)."""
    from surjectors_b200 import _abi

    _abi.check(self._lib.surj_dp_set_timeout_ms(self._h, int(ms)), "surj_dp_set_timeout_ms")

  def close(self):
    if self._h:
      self._lib.surj_dp_destroy(self._h)
      self._h = None

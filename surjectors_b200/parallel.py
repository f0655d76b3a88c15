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

"""The multi-GPU story of the path on CPU (gloo, world_size 2): contiguous batch shards, replicated
parameters, ONE all-reduce(sum) of the flat gradient (+ loss) per step, identical Adam updates on
every rank.  The per-shard gradients come from the oracle here (no GPU); the host logic under test
is surjectors_b200.parallel (sharding + flat all-reduce) and the Adam arithmetic the CUDA kernel
implements."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from surjectors_b200 import parallel


def test_shard_bounds_cover_batch_exactly():
  for B in (1, 7, 64, 65536, 1000003):
    for W in (1, 2, 3, 8):
      b = [parallel.shard_bounds(B, r, W) for r in range(W)]
      assert b[0][0] == 0 and b[-1][1] == B and all(b[i][1] == b[i + 1][0] for i in range(W - 1))
      sizes = [hi - lo for lo, hi in b]
      assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from oracle import autodiff as oad
    from tests import parity

    B = 64
    pm, pparams, om, oparams, y64, x64 = parity.build_pair("cfg2", B, 0, 0.3)
    lo, hi = parallel.shard_bounds(B, rank, world)
    loss, grads, _ = oad.value_and_grad(om, oparams, y=y64[lo:hi])
    layout = pparams.layout
    flat = torch.zeros(layout.size + 1, dtype=torch.float64)
    for r in layout.refs:
      flat[r.offset : r.offset + r.size] = grads[r.module][r.name].reshape(-1)
    flat[layout.size] = loss
    parallel.allreduce_flat_(flat, dist.group.WORLD)
    # reference: the full batch on one process
    loss_all, grads_all, _ = oad.value_and_grad(om, oparams, y=y64)
    ref = torch.zeros_like(flat)
    for r in layout.refs:
      ref[r.offset : r.offset + r.size] = grads_all[r.module][r.name].reshape(-1)
    ref[layout.size] = loss_all
    err = float((flat - ref).abs().max())
    # replicated Adam: identical parameters on every rank after the update
    p = pparams.flat.double().clone()
    mu, nu = torch.zeros_like(p), torch.zeros_like(p)
    parallel.adam_reference_(p, flat[: layout.size], mu, nu, 1, lr=1e-4)
    gathered = [torch.zeros_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
      out.put((err, same))
  finally:
    dist.destroy_process_group()


def test_gradient_allreduce_equals_full_batch_gradient():
  ctx = mp.get_context("spawn")
  out = ctx.SimpleQueue()
  port = 29500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(180)
    assert p.exitcode == 0
  err, same = out.get()
  assert err <= 1e-9 and same

"""The data-parallel train step on 2 real GPUs: batch shards, replicated parameters, ONE sum of
the flat gradient per step -- through the fused NVLink peer-memory exchange (fused executors) and through the NCCL
all-reduce (per-op executor, or forced) -- against the single-rank step on the whole batch: replicas bit-identical,
losses and parameters equal to the 1-rank run within float32 summation order.  Also: a rank delayed beyond the exchange
timeout makes its peer raise instead of applying a stale gradient.  Needs 2 GPUs (gpurun --gpus 2); skipped on one."""

import json
import os
import socket
import time

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_train_step_matches_single_rank(tmp_path):
  import torch.multiprocessing as mp

  from tests import dp_worker

  out = str(tmp_path / "dp")
  ctx = mp.spawn(dp_worker.main, args=(2, _free_port(), out), nprocs=2, join=False)
  deadline = time.time() + 600
  while not ctx.join(timeout=5):  # join() returns as soon as ONE process has exited: keep joining until all have
    assert time.time() < deadline, "data-parallel workers did not finish"
  res = [json.load(open(f"{out}.{r}")) for r in range(2)]
  for r in res:
    for item in r["results"]:
      print(item)
  assert all(r["ok"] for r in res), res

"""2+ GPU check of the fused peer-memory gradient exchange against the NCCL all-reduce path:
both must produce identical losses and (near-)identical parameters after a few steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from surjectors_b200 import configs, hk
from surjectors_b200.train import TrainStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
c = configs.CONFIGS["cfg2"]
B = 4096
rng = np.random.Generator(np.random.PCG64(100 + rank))
ys = [torch.from_numpy(rng.standard_normal((B, c["dim"]), dtype=np.float32)).to(dev) for _ in range(6)]
model = hk.transform(c["make"]())
results = {}
for mode in ("nccl", "peer"):
  os.environ["SURJ_B200_EXCHANGE"] = mode
  params = model.init(7, method="log_prob", y=ys[0])
  ts = TrainStep(model, params, dict(y=ys[0]), lr=1e-3, process_group=dist.group.WORLD)
  assert (ts.peer is not None) == (mode == "peer")
  losses = []
  for y in ys:
    losses.append(float(ts.step(y).item()))
  torch.cuda.synchronize()
  if ts.peer is not None:
    assert ts.peer.error() == 0
  results[mode] = (losses, ts.flat.clone())
l0, p0 = results["nccl"]
l1, p1 = results["peer"]
gathered = [torch.zeros_like(p1) for _ in range(world)]
dist.all_gather(gathered, p1)
same_across_ranks = all(torch.equal(gathered[0], g) for g in gathered)
err = float((p0 - p1).abs().max())
lerr = max(abs(a - b) / max(1.0, abs(a)) for a, b in zip(l0, l1))
if rank == 0:
  print(f"dp_check world={world}: loss rel diff {lerr:.2e}, param max diff {err:.2e}, replicas identical: {same_across_ranks}")
assert lerr < 1e-5 and err < 1e-5 and same_across_ranks
dist.barrier()
os._exit(0)

#!/usr/bin/env python
"""Benchmark of the one hot path: the train step (log_prob + VJP + Adam) of BASELINE.json
configs[1] -- examples/coupling_inference_surjection.py: AffineMaskedCouplingInferenceFunnel +
MaskedCoupling chain, dim 20 -> 10 -> 10 -> 5, batch 65536 per GPU -- on N B200s (weak scaling: the
batch is sharded, parameters replicated, one NCCL all-reduce of the flat gradient per step).

  python bench.py --gpus N --steps K --warmup W              our arm   (CUDA kernels via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...    reference arm (CPU restatement of the
                                                             reference on the host cores; rank 0 only)

Prints ONE JSON line.  `value` = samples/s with the batches resident in HBM (a pool of batches
larger than L2 is cycled, so no step re-reads a cached batch); `e2e` = the same step driven from
pinned HOST batches through `TrainStep.step_host` (H2D copy of every batch and D2H read of every
loss inside the timed region); `roofline` = the fused log_prob+VJP kernel timed alone with CUDA
events; `cpu_baseline` = the oracle (a torch-CPU restatement of the reference; JAX is not
installable here) on the box's cores.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
_RESULT_OUT = sys.stdout
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

CONFIG = "cfg2"
BATCH = 65536
LR = 1e-4
DATA_SEED, PARAM_SEED = 1002, 2002


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=256)
  ap.add_argument("--warmup", type=int, default=16)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--config", default=CONFIG)
  ap.add_argument("--batch", type=int, default=BATCH)
  ap.add_argument("--pool", type=int, default=32, help="device-resident batches cycled through (pool bytes > L2)")
  ap.add_argument("--no-graph", action="store_true", help="launch every step from python instead of CUDA graphs")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--cpu-seconds", type=float, default=12.0)
  return ap.parse_args()


def workload_name(cfg, batch):
  return (f"{cfg}: examples/coupling_inference_surjection.py train step (log_prob + VJP + adam(1e-4)), "
          f"dim 20->10->10->5, batch {batch} per GPU")


# --------------------------------------------------------------------- reference arm
def cpu_train_baseline(cfg, batch, seconds=None, steps=None, warmup=2):
  """The reference's train step restated on the host (oracle/, float32, all cores)."""
  import numpy as np
  import torch

  from oracle import autodiff as oad
  from oracle import configs as ocfg
  from oracle import hk as ohk
  from surjectors_b200 import configs as pcfg

  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  c = pcfg.CONFIGS[cfg]
  rng = np.random.Generator(np.random.PCG64(DATA_SEED))
  y = torch.as_tensor(rng.standard_normal((batch, c["dim"])), dtype=torch.float32)
  b = dict(y=y)
  if c["dim_x"]:
    b["x"] = torch.as_tensor(rng.standard_normal((batch, c["dim_x"])), dtype=torch.float32)
  model = ohk.without_apply_rng(ohk.transform(getattr(ocfg, "make_" + cfg)(), dtype=torch.float32))
  params = model.init(PARAM_SEED, **b)
  state = oad.adam_init(params)

  def step():
    nonlocal params, state
    loss, grads, _ = oad.value_and_grad(model, params, **b)
    params, state = oad.adam_update(params, grads, state, lr=LR)
    return float(loss)

  for _ in range(warmup):
    step()
  t0 = time.perf_counter()
  n = 0
  while True:
    step()
    n += 1
    el = time.perf_counter() - t0
    if (steps is not None and n >= steps) or (steps is None and el >= seconds and n >= 3):
      break
  return dict(samples_per_s=n * batch / el, ms_per_step=1e3 * el / n, steps=n, cores=cores, seconds=el)


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  r = cpu_train_baseline(args.config, args.batch, steps=args.steps, warmup=args.warmup)
  sample = f"{r['steps']} train steps of batch {args.batch} (float32, torch-CPU restatement of the reference, not JAX)"
  line = {
    "impl": "reference", "metric": "train_samples_per_sec", "value": r["samples_per_s"], "unit": "samples/s",
    "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": {"workload": workload_name(args.config, args.batch), "batch_per_gpu": args.batch},
    "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                     "sample": sample},
    "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }
  print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ----------------------------------------------------------------------- clock sampler
class ClockSampler:
  Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index = index
    self.rows = []
    self._stop = threading.Event()
    self._t = None

  def _loop(self):
    while not self._stop.is_set():
      try:
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [v.strip() for v in out.strip().split(",")]
        if len(f) >= 7:
          self.rows.append(f)
      except Exception:  # noqa: BLE001
        pass
      self._stop.wait(0.05)

  def __enter__(self):
    self._t = threading.Thread(target=self._loop, daemon=True)
    self._t.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    self._t.join(timeout=10)

  def summary(self):
    import statistics

    if not self.rows:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
    mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------- our arm
def _dbg(msg):
  if os.environ.get("SURJ_BENCH_DEBUG"):
    print(f"[bench r{os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def run_b200(args):
  import numpy as np
  import torch

  import __graft_entry__ as entry
  from surjectors_b200 import _abi, _cost, configs, hk
  from surjectors_b200.train import TrainStep

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not os.path.exists(_abi.LIB_PATH):
    entry.build()
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  group = None
  if world > 1:
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
  if world != args.gpus and rank == 0:
    print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

  cfg = configs.CONFIGS[args.config]
  B, P, D, DX = args.batch, args.pool, cfg["dim"], cfg["dim_x"]
  rng = np.random.Generator(np.random.PCG64(DATA_SEED + rank))
  host_y = torch.empty((P, B, D), dtype=torch.float32).pin_memory()
  host_y.copy_(torch.from_numpy(rng.standard_normal((P, B, D), dtype=np.float32)))
  host_x = None
  if DX:
    host_x = torch.empty((P, B, DX), dtype=torch.float32).pin_memory()
    host_x.copy_(torch.from_numpy(rng.standard_normal((P, B, DX), dtype=np.float32)))
  dev_y = host_y.to(dev)
  dev_x = host_x.to(dev) if DX else None

  model = hk.transform(cfg["make"]())
  ex = dict(y=dev_y[0]) if not DX else dict(y=dev_y[0], x=dev_x[0])
  params = model.init(PARAM_SEED, method="log_prob", **ex)  # same seed on every rank: replicated parameters
  ts = TrainStep(model, params, ex, lr=LR, process_group=group)
  assert ts.plan.is_fused, "bench config must run on the fused whole-chain kernel"

  def step(i):
    k = i % P
    return ts.step(dev_y[k], dev_x[k] if DX else None)

  _dbg("model + data ready")
  # --- CUDA graphs: one graph = one pass over the pool (P consecutive steps); singles for the remainder
  use_graph = not args.no_graph
  full = singles = None
  stream = torch.cuda.Stream(dev)
  torch.cuda.synchronize(dev)
  with torch.cuda.stream(stream):
    for i in range(3):  # allocate workspaces, load kernels, NCCL warm-up outside capture
      step(i)
  torch.cuda.synchronize(dev)
  _dbg("eager warm-up done")
  if use_graph:
    try:
      with torch.cuda.stream(stream):
        full = torch.cuda.CUDAGraph()
        with torch.cuda.graph(full, stream=stream):
          for i in range(P):
            step(i)
        singles = []
        for i in range(min(P, max(args.steps % P, args.warmup % P))):
          g = torch.cuda.CUDAGraph()
          with torch.cuda.graph(g, stream=stream):
            step(i)
          singles.append(g)
    except Exception as e:  # noqa: BLE001
      if rank == 0:
        print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); launching steps from python",
              file=sys.stderr)
      use_graph, full, singles = False, None, None
      torch.cuda.synchronize(dev)

  def run_steps(n):
    if use_graph:
      for _ in range(n // P):
        full.replay()
      for i in range(n % P):
        singles[i].replay()
    else:
      for i in range(n):
        step(i)

  def barrier():
    if world > 1:
      import torch.distributed as dist

      dist.barrier(device_ids=[local])

  def max_over_ranks(ms):
    if world == 1:
      return ms
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  _dbg(f"graphs captured: {use_graph}")
  # --- timed region: resident inputs
  with torch.cuda.stream(stream):
    run_steps(args.warmup)
    torch.cuda.synchronize(dev)
    barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
      e0.record(stream)
      run_steps(args.steps)
      e1.record(stream)
      torch.cuda.synchronize(dev)
      # keep the sampler alive for at least a few samples when the region is very short
      ms = e0.elapsed_time(e1)
      if ms < 1000.0:
        # same load, outside the timed region, until nvidia-smi has a few samples; the number of extra
        # passes is agreed between ranks (every step contains a collective)
        extra = int(max_over_ranks(1000.0 / max(ms / args.steps, 1e-3) / P)) + 1
        for _ in range(min(extra, 4096)):
          run_steps(P)
        torch.cuda.synchronize(dev)
    barrier()
    torch.cuda.synchronize(dev)
  _dbg("timed region done")
  ms = max_over_ranks(ms)
  value = world * B * args.steps / (ms * 1e-3)
  loss_resident = float(ts.loss.item())

  # --- roofline: the fused log_prob+VJP kernel alone, CUDA events on its launch stream
  n_par = ts.flat.numel()
  reps = 10 * P
  with torch.cuda.stream(stream):
    for i in range(P):
      ts.plan.bwd(dev_y[i], dev_x[i] if DX else None, ts.flat, g_scale=-1.0, g_params=ts.grads,
                  loss_out=ts.grads[n_par:n_par + 1])
    torch.cuda.synchronize(dev)
    if use_graph:
      kg = torch.cuda.CUDAGraph()
      with torch.cuda.graph(kg, stream=stream):
        for i in range(P):
          ts.plan.bwd(dev_y[i], dev_x[i] if DX else None, ts.flat, g_scale=-1.0, g_params=ts.grads,
                      loss_out=ts.grads[n_par:n_par + 1])
      kg.replay()
      torch.cuda.synchronize(dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for r in range(reps // P):
      if use_graph:
        kg.replay()
      else:
        for i in range(P):
          ts.plan.bwd(dev_y[i], dev_x[i] if DX else None, ts.flat, g_scale=-1.0, g_params=ts.grads,
                      loss_out=ts.grads[n_par:n_par + 1])
    k1.record(stream)
    torch.cuda.synchronize(dev)
    bwd_pair_ms = k0.elapsed_time(k1) / reps  # fused kernel + its partial-reduction kernel
    # FP32 pipe peak, measured with the probe kernel of the library
    lib = _abi.lib()
    import ctypes as C

    sink = torch.zeros(4, device=dev)
    fl = C.c_double(0.0)
    for _ in range(2):
      lib.surj_bench_fp32_peak(stream.cuda_stream, 4096, sink.data_ptr(), C.byref(fl))
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(5):
      lib.surj_bench_fp32_peak(stream.cuda_stream, 4096, sink.data_ptr(), C.byref(fl))
    p1.record(stream)
    torch.cuda.synchronize(dev)
    fp32_peak = 5 * fl.value / (p0.elapsed_time(p1) * 1e-3) / 1e12

  # --- the other half of BASELINE.json's metric: log_prob (inference) samples/s, same pool of resident batches
  lp_buf = torch.empty(B, dtype=torch.float32, device=dev)
  with torch.cuda.stream(stream):
    def lp_pass():
      for i in range(P):
        ts.plan.fwd(dev_y[i], dev_x[i] if DX else None, ts.flat, want_z=False, lp_out=lp_buf)
    lp_pass()
    torch.cuda.synchronize(dev)
    lg = None
    if use_graph:
      lg = torch.cuda.CUDAGraph()
      with torch.cuda.graph(lg, stream=stream):
        lp_pass()
      lg.replay()
      torch.cuda.synchronize(dev)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(stream)
    for _ in range(10):
      lg.replay() if lg is not None else lp_pass()
    l1.record(stream)
    torch.cuda.synchronize(dev)
    lp_ms = max_over_ranks(l0.elapsed_time(l1) / (10 * P))
  logprob = {"samples_per_sec": world * B / (lp_ms * 1e-3), "us_per_call": 1e3 * lp_ms,
             "api": "surj_chain_logprob_fwd (fused forward kernel), no collective"}
  _dbg("roofline done")
  need, dense = _cost.forward_flops(ts.spec)
  flops_launch = 3.0 * need * B
  bytes_launch = float(_cost.forward_bytes(ts.spec) * B + 4 * n_par)
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except Exception:  # noqa: BLE001
    pass
  hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
  traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of the kernel, from the committed ncu --set full capture
  try:
    t = json.load(open(os.path.join(ROOT, "profiles", "r1_cfg2_bwd_traffic.json")))
    if args.config == "cfg2" and B == BATCH:
      traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
  except Exception:  # noqa: BLE001
    pass
  ach_tf = flops_launch / (bwd_pair_ms * 1e-3) / 1e12
  ach_gb = bytes_launch / (bwd_pair_ms * 1e-3) / 1e9
  roofline = {
    "kernel": "surj::chain_bwd_kernel<cfg2> (+ surj_reduce_partials_kernel)",
    "bound": "fp32", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak,
    "peak_source": "measured here with surj_bench_fp32_peak (FFMA probe); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
    "flops_per_launch": flops_launch, "flops_dense_as_reference_per_launch": 3.0 * dense * B,
    "us_per_launch": 1e3 * bwd_pair_ms, "traffic": traffic,
    "hbm": {"bound": "hbm", "achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gb / hbm_peak,
            "bytes_per_launch": bytes_launch,
            "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"},
  }

  # --- e2e: pinned host batches through TrainStep.step_host (H2D of every batch, D2H of every loss)
  losses_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
  with torch.cuda.stream(stream):
    for i in range(max(3, min(args.warmup, 8))):
      ts.step_host(host_y[i % P], host_x[i % P] if DX else None)
    ts.drain()
    torch.cuda.synchronize(dev)
    barrier()
    torch.cuda.synchronize(dev)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record(stream)
    for i in range(args.steps):
      ts.step_host(host_y[i % P], host_x[i % P] if DX else None, loss_out=losses_host[i:i + 1])
    ts.drain()
    h1.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
  _dbg("e2e done")
  e2e_ms = max_over_ranks(h0.elapsed_time(h1))
  e2e = {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": "samples/s",
         "h2d_bytes_per_step": 4 * B * (D + DX), "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps,
         "api": "surjectors_b200.train.TrainStep.step_host (pinned host batch in, loss out)"}
  assert np.isfinite(losses_host.numpy()).all()

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    r = cpu_train_baseline(args.config, B, seconds=args.cpu_seconds)
    cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
           "sample": f"{r['steps']} train steps of batch {B} in {r['seconds']:.1f} s (float32 torch-CPU restatement "
                     "of the reference, all cores; JAX is not installable in this image)"}

  if rank == 0:
    line = {
      "metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
      "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": workload_name(args.config, B), "batch_per_gpu": B, "global_batch": B * world,
                 "parallelism": f"dp{world}", "params": n_par,
                 "l2": f"pool of {P} distinct batches ({P * B * (D + DX) * 4 / 1e6:.0f} MB > 126 MB L2) cycled",
                 "cuda_graph": bool(use_graph), "train_steps_per_sec": args.steps / (ms * 1e-3)},
      "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": args.steps * ts.launches_per_step(),
      "roofline": roofline, "cpu_baseline": cpu, "log_prob": logprob, "final_loss": loss_resident,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)
  if world > 1:
    # CUDA graphs that captured NCCL collectives keep the communicator busy at teardown
    # (destroy_process_group was seen to hang); everything is measured and printed: leave hard.
    torch.cuda.synchronize(dev)
    barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
  # stdout carries exactly ONE line, the JSON result: everything libraries print there (NCCL's version banner, ...)
  # is sent to stderr by pointing file descriptor 1 at stderr for the run and keeping the real stdout aside.
  global _RESULT_OUT
  sys.stdout.flush()
  _RESULT_OUT = os.fdopen(os.dup(1), "w")
  os.dup2(2, 1)
  args = parse()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_b200(args)


if __name__ == "__main__":
  main()

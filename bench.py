#!/usr/bin/env python
"""Benchmark of the one hot path -- the flow log_prob and its train step (log_prob + VJP + Adam) -- on N B200s.

  python bench.py --gpus N --steps K --warmup W              our arm   (CUDA kernels through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...    reference arm (CPU restatement of the reference on the
                                                             host cores, micro-batched; rank 0 only)

The headline workload is BASELINE.json configs[3], the largest configuration that fits one GPU:
RationalQuadraticSplineMaskedCouplingInferenceFunnel + 8 spline MaskedCoupling layers, 64 bins, hidden 256, dim
64 -> 16, batch 2^20 per GPU -- a dense contraction (19.8 M parameters, 18.9 MFLOP needed per sample and direction),
tensor-pipe bound.  Weak scaling: the batch is sharded, parameters replicated, one NCCL all-reduce of the 79 MB flat
gradient per step.  `--config cfg2|cfg5|...` selects another BASELINE config; at N = 1 the default run also reports the
other four configs under "other_configs" (each with its own ms_per_step and roofline fraction; cfg1-3 additionally at
B = 2^24, the steady state where the kernel and not the launch is measured).

Prints ONE JSON line.  `value` = samples/s with the batches resident in HBM (a pool of distinct batches larger than L2
is cycled, so no step re-reads a cached batch); `e2e` = the same step driven from pinned HOST batches through
`TrainStep.step_host` (H2D copy of every batch and D2H read of every loss inside the timed region); `roofline` = the
dominant kernel timed with CUDA events on its launch stream (surj_profile_begin / surj_profile_end bracket every launch
of the per-op executor) against the tensor-pipe peak measured in the same run (surj_bench_tf32_peak, divided by 3:
FP32-grade products are three tensor-core MMAs); `cpu_baseline` = the oracle (a torch-CPU float32 restatement of the
reference; JAX is not installable here) on the box's cores, micro-batched as SURVEY.md 8d prescribes.
"""

import argparse
import ctypes as C
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

CONFIG = "cfg4"
LR = 1e-4
SEEDS = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}  # data seed 1000 + cfg, params 2000 + cfg
PER_GPU_BATCH = {"cfg5": 1 << 20}  # BASELINE: 8 M samples over 8 GPUs
CPU_MICRO = {"cfg4": 1 << 14, "cfg5": 1 << 14}  # the host cannot hold [B, D, 3K+1] at 2^20 -- micro-batches
WORKLOADS = {
  "cfg1": "cfg1: README flow Chain([Slice(5, dec 5->[32,32,10]), LULinear(5)]), dim 10",
  "cfg2": "cfg2: examples/coupling_inference_surjection.py, affine coupling funnels, dim 20->10->10->5",
  "cfg3": "cfg3: examples/autoregressive_inference_surjection.py, affine MADE funnels, dim 20->10->10->5",
  "cfg4": ("cfg4: RationalQuadraticSplineMaskedCouplingInferenceFunnel + 8 spline MaskedCoupling layers, 64 bins, "
           "hidden 256, dim 64->16"),
  "cfg5": ("cfg5: examples/conditional_density_estimation.py scaled: conditional 16-layer MAF (MADE hidden 512) + "
           "LULinear + Permutation, dim 128, dim_x 16"),
}


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=8)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--config", default=CONFIG)
  ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: BASELINE's)")
  ap.add_argument("--pool", type=int, default=0, help="distinct device-resident batches cycled (default: enough to exceed L2)")
  ap.add_argument("--no-graph", action="store_true", help="launch every step from python instead of CUDA graphs")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs (N = 1 default run measures them)")
  ap.add_argument("--cpu-seconds", type=float, default=15.0)
  a = ap.parse_args()
  if a.batch <= 0:
    a.batch = default_batch(a.config)
  return a


def default_batch(cfg):
  from surjectors_b200 import configs

  return PER_GPU_BATCH.get(cfg, configs.CONFIGS[cfg]["batch"])


def workload_name(cfg, batch):
  return f"{WORKLOADS.get(cfg, cfg)}; train step (log_prob + VJP + adam(1e-4)), batch {batch} per GPU"


def seeds(cfg):
  k = SEEDS.get(cfg, 9)
  return 1000 + k, 2000 + k


# --------------------------------------------------------------------- reference arm
def cpu_train_baseline(cfg, batch, seconds=None, steps=None, warmup=1):
  """The reference's train step restated on the host (oracle/, float32, all cores).  Wide configs run in micro-batches
  of CPU_MICRO samples (a 64-dim spline layer materialises 49 KB of spline parameters per sample: 51.8 GB at 2^20);
  each micro-batch is a complete value_and_grad + Adam update, throughput is samples/s over the timed micro-batches."""
  import numpy as np
  import torch

  from oracle import autodiff as oad
  from oracle import configs as ocfg
  from oracle import hk as ohk
  from surjectors_b200 import configs as pcfg

  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  c = pcfg.CONFIGS[cfg]
  micro = min(batch, CPU_MICRO.get(cfg, batch))
  dseed, pseed = seeds(cfg)
  rng = np.random.Generator(np.random.PCG64(dseed))
  y = torch.as_tensor(rng.standard_normal((micro, c["dim"])), dtype=torch.float32)
  b = dict(y=y)
  if c["dim_x"]:
    b["x"] = torch.as_tensor(rng.standard_normal((micro, c["dim_x"])), dtype=torch.float32)
  model = ohk.without_apply_rng(ohk.transform(getattr(ocfg, "make_" + cfg)(), dtype=torch.float32))
  params = model.init(pseed, **{k: v[:8] for k, v in b.items()})
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
  return dict(samples_per_s=n * micro / el, ms_per_step=1e3 * el / n, steps=n, cores=cores, seconds=el, micro=micro)


def cpu_sample_text(r, batch):
  mb = "" if r["micro"] == batch else f" (micro-batches of the {batch}-sample workload, throughput extrapolates linearly)"
  return (f"{r['steps']} train steps of {r['micro']} samples{mb} in {r['seconds']:.1f} s; float32 torch-CPU restatement of "
          f"the reference, {r['cores']} threads; not JAX (not installable in this image)")


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  r = cpu_train_baseline(args.config, args.batch, steps=args.steps, warmup=min(args.warmup, 2))
  line = {
    "impl": "reference", "metric": "train_samples_per_sec", "value": r["samples_per_s"], "unit": "samples/s",
    "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": config_dict(args.config, args.batch, args.gpus, None),
    "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                     "sample": cpu_sample_text(r, args.batch)},
    "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }
  print(json.dumps(line), file=_RESULT_OUT, flush=True)


def config_dict(cfg, batch, world, n_par):
  """The SAME dictionary in both arms (the driver compares them): what the workload is, nothing about how it ran."""
  d = {"workload": workload_name(cfg, batch), "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"dp{world}"}
  if n_par is not None:
    d["params"] = n_par
  return d


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
    pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "power_w": statistics.median(pw) if pw else None, "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------- our arm
def _dbg(msg):
  if os.environ.get("SURJ_BENCH_DEBUG"):
    print(f"[bench r{os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


class Ctx:
  """Process-wide facts of the run: device, ranks, measured peaks."""

  def __init__(self):
    import torch

    self.world = int(os.environ.get("WORLD_SIZE", "1"))
    self.rank = int(os.environ.get("RANK", "0"))
    self.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(self.local)
    self.dev = torch.device("cuda", self.local)
    self.group = None
    if self.world > 1:
      import torch.distributed as dist

      os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
      dist.init_process_group("nccl", device_id=self.dev)
      self.group = dist.group.WORLD
    self.stream = torch.cuda.Stream(self.dev)
    self._peaks = None

  def barrier(self):
    if self.world > 1:
      import torch.distributed as dist

      dist.barrier(device_ids=[self.local])

  def max_over_ranks(self, v):
    if self.world == 1:
      return v
    import torch
    import torch.distributed as dist

    t = torch.tensor([v], dtype=torch.float64, device=self.dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  def peaks(self):
    """FP32-FMA and TF32 tensor-pipe peaks measured here with the library's probe kernels (TFLOP/s)."""
    if self._peaks is not None:
      return self._peaks
    import torch

    from surjectors_b200 import _abi

    lib = _abi.lib()
    sink = torch.zeros(4, device=self.dev)
    out = {}
    with torch.cuda.stream(self.stream):
      for name, fn, iters in (("fp32", lib.surj_bench_fp32_peak, 4096), ("tf32", lib.surj_bench_tf32_peak, 20000)):
        fl = C.c_double(0.0)
        for _ in range(2):
          _abi.check(fn(self.stream.cuda_stream, iters, sink.data_ptr(), C.byref(fl)), "peak probe")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(5):
          fn(self.stream.cuda_stream, iters, sink.data_ptr(), C.byref(fl))
        e1.record(self.stream)
        torch.cuda.synchronize(self.dev)
        out[name] = 5 * fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
    try:
      out["file"] = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
      out["file"] = {}
    self._peaks = out
    return out


def profile_kinds(fn, dev):
  """Run fn() once between surj_profile_begin / surj_profile_end: {kernel kind: {launches, ms, flops}} of the per-op
  executor's launches, each bracketed by CUDA events on its launch stream."""
  import torch

  from surjectors_b200 import _abi

  lib = _abi.lib()
  torch.cuda.synchronize(dev)
  lib.surj_profile_begin()
  fn()
  buf = C.create_string_buffer(1 << 16)
  n = lib.surj_profile_end(buf, len(buf))
  return json.loads(buf.value.decode()) if n > 0 else {}


class Workload:
  """One BASELINE config on this rank: model, parameters, a pool of distinct batches (device + pinned host)."""

  def __init__(self, ctx, cfg_name, batch, pool=0, host=True):
    import numpy as np
    import torch

    from surjectors_b200 import _cost, configs, hk
    from surjectors_b200.train import TrainStep

    self.ctx, self.name, self.B = ctx, cfg_name, batch
    cfg = configs.CONFIGS[cfg_name]
    self.D, self.DX = cfg["dim"], cfg["dim_x"]
    bytes_batch = 4 * batch * (self.D + self.DX)
    l2 = 126e6
    self.P = pool if pool > 0 else max(2, min(32, int(l2 // bytes_batch) + 2))
    dseed, pseed = seeds(cfg_name)
    g = torch.Generator().manual_seed(dseed + 7919 * ctx.rank)
    mk = (lambda shape: torch.empty(shape, dtype=torch.float32).pin_memory()) if host else (lambda shape: torch.empty(shape, dtype=torch.float32))
    self.host_y = mk((self.P, batch, self.D))
    self.host_y.normal_(generator=g)
    self.host_x = None
    if self.DX:
      self.host_x = mk((self.P, batch, self.DX))
      self.host_x.normal_(generator=g)
    self.dev_y = self.host_y.to(ctx.dev)
    self.dev_x = self.host_x.to(ctx.dev) if self.DX else None
    self.model = hk.transform(cfg["make"]())
    ex = dict(y=self.dev_y[0]) if not self.DX else dict(y=self.dev_y[0], x=self.dev_x[0])
    self.params = self.model.init(pseed, method="log_prob", **ex)  # same seed on every rank: replicated parameters
    self.ts = TrainStep(self.model, self.params, ex, lr=LR, process_group=ctx.group)
    self.need, self.dense = _cost.forward_flops(self.ts.spec)
    self.bytes_fwd = _cost.forward_bytes(self.ts.spec)
    self.n_par = self.ts.flat.numel()
    self.l2_note = (f"pool of {self.P} distinct batches ({self.P * bytes_batch / 1e6:.0f} MB > 126 MB L2) cycled" if bytes_batch < l2 else
                    f"every batch is {bytes_batch / 1e6:.0f} MB > 126 MB L2; pool of {self.P} distinct batches cycled")

  def x(self, k):
    return self.dev_x[k] if self.DX else None

  def step(self, i):
    k = i % self.P
    return self.ts.step(self.dev_y[k], self.x(k))

  def free(self):
    import torch

    self.ts = self.params = self.model = self.dev_y = self.dev_x = self.host_y = self.host_x = None
    torch.cuda.empty_cache()


def make_runner(w, fn, n_steps, use_graph):
  """run(n): n consecutive calls fn(i).  Fused plans (a step is 2 launches of tens of microseconds) replay CUDA graphs:
  one graph of exactly `n_steps` steps for the timed region plus one-pool graphs for warm-up; the per-op executor's
  steps last hundreds of milliseconds and are launched directly."""
  import torch

  ctx = w.ctx
  if not use_graph:
    def run(n):
      for i in range(n):
        fn(i)
    return run, False
  try:
    graphs = {}
    with torch.cuda.stream(ctx.stream):
      for n in sorted({n_steps, w.P}):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=ctx.stream):
          for i in range(n):
            fn(i)
        graphs[n] = g
  except Exception as e:  # noqa: BLE001
    if ctx.rank == 0:
      print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); launching steps from python", file=sys.stderr)
    torch.cuda.synchronize(ctx.dev)
    return make_runner(w, fn, n_steps, False)

  def run(n):
    while n >= n_steps and n_steps >= w.P:
      graphs[n_steps].replay()
      n -= n_steps
    while n >= w.P:
      graphs[w.P].replay()
      n -= w.P
    if n == n_steps:
      graphs[n_steps].replay()
      n = 0
    for i in range(n):
      fn(i)
  return run, True


def timed(ctx, run, n, sampler=None, align=True):
  """ms of run(n) on ctx.stream, max over ranks.  The ranks are aligned on the DEVICE: barrier, then one untimed pass
  enqueued right before the start event with no host sync in between (every train step of a data-parallel run contains
  the gradient exchange, which is itself a device-side barrier), so host skew at barrier exit stays out of the region."""
  import torch

  with torch.cuda.stream(ctx.stream):
    torch.cuda.synchronize(ctx.dev)
    ctx.barrier()
    torch.cuda.synchronize(ctx.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if align:
      run(1)
    e0.record(ctx.stream)
    run(n)
    e1.record(ctx.stream)
    torch.cuda.synchronize(ctx.dev)
  return ctx.max_over_ranks(e0.elapsed_time(e1))


def measure(ctx, cfg_name, batch, steps, warmup, pool=0, no_graph=False, e2e=True, with_clocks=False, quick=False):
  """Train step (resident + host-fed), log_prob and the roofline leg of one config.  Returns a dict."""
  import numpy as np
  import torch

  w = Workload(ctx, cfg_name, batch, pool, host=e2e)
  ts = w.ts
  fused = ts.plan.is_fused
  out = {"executor": ts.plan.executor, "params": w.n_par, "batch_per_gpu": batch, "l2": w.l2_note,
         "needed_flops_per_sample_fwd": w.need, "dense_flops_per_sample_fwd": w.dense}
  with torch.cuda.stream(ctx.stream):
    for i in range(2 if not fused else 3):  # allocate workspaces, load kernels, NCCL warm-up outside capture
      w.step(i)
  torch.cuda.synchronize(ctx.dev)
  _dbg(f"{cfg_name}: eager warm-up done")
  run, graphed = make_runner(w, w.step, steps, fused and not no_graph)
  out["cuda_graph"] = graphed

  # --- timed region: resident inputs
  with torch.cuda.stream(ctx.stream):
    run(max(0, warmup - 2))
  if with_clocks:
    with ClockSampler(ctx.local) as clocks:
      ms = timed(ctx, run, steps)
      if ms < 1000.0:  # keep the same load up until nvidia-smi has a few samples (outside the timed region)
        extra = int(ctx.max_over_ranks(1000.0 / max(ms / steps, 1e-3) / w.P)) + 1
        with torch.cuda.stream(ctx.stream):
          for _ in range(min(extra, 4096)):
            run(w.P)
          torch.cuda.synchronize(ctx.dev)
    out["clocks"] = clocks.summary()
  else:
    ms = timed(ctx, run, steps)
  out["ms_per_step"] = ms / steps
  out["train_samples_per_sec"] = ctx.world * batch * steps / (ms * 1e-3)
  out["final_loss"] = float(ts.loss.item())
  out["gpu_launches"] = steps * ts.launches_per_step()
  _dbg(f"{cfg_name}: resident train {out['ms_per_step']:.3f} ms/step")

  # --- log_prob (inference half of BASELINE's metric), same pool of resident batches, no collective
  lp_buf = torch.empty(batch, dtype=torch.float32, device=ctx.dev)

  def lp_call(i):
    k = i % w.P
    ts.plan.fwd(w.dev_y[k], w.x(k), ts.flat, want_z=False, lp_out=lp_buf)

  with torch.cuda.stream(ctx.stream):
    lp_call(0)
  n_lp = w.P if not quick else min(w.P, 2)
  lp_run, _ = make_runner(w, lp_call, n_lp, fused and not no_graph)
  lp_ms = timed(ctx, lp_run, n_lp) / n_lp
  out["log_prob"] = {"samples_per_sec": ctx.world * batch / (lp_ms * 1e-3), "ms_per_call": lp_ms,
                     "needed_tflops": w.need * batch / (lp_ms * 1e-3) / 1e12,
                     "api": "surj_chain_logprob_fwd, no collective"}
  _dbg(f"{cfg_name}: log_prob {lp_ms:.3f} ms")

  # --- roofline of the dominant kernel
  peaks = ctx.peaks()
  step_flops = 3.0 * w.need * batch
  if fused:
    n_par = w.n_par

    def bwd_call(i):
      k = i % w.P
      ts.plan.bwd(w.dev_y[k], w.x(k), ts.flat, g_scale=-1.0, g_params=ts.grads, loss_out=ts.grads[n_par:n_par + 1])

    with torch.cuda.stream(ctx.stream):
      bwd_call(0)
    k_run, _ = make_runner(w, bwd_call, w.P, not no_graph)
    reps = 4 if quick else 10
    pair_ms = timed(ctx, lambda n: [k_run(w.P) for _ in range(n)], reps) / (reps * w.P)
    ach = step_flops / (pair_ms * 1e-3) / 1e12
    bytes_launch = float(w.bytes_fwd * batch + 4 * w.n_par)
    hbm_peak = float(peaks["file"].get("hbm_gbs", 6650.0))
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of the kernel, from the committed ncu --set full capture
      t = json.load(open(os.path.join(ROOT, "profiles", "r1_cfg2_bwd_traffic.json")))
      if cfg_name == "cfg2" and batch == 65536:
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:  # noqa: BLE001
      pass
    out["roofline"] = {
      "kernel": f"surj::chain_bwd_kernel<{cfg_name}> (+ surj_reduce_partials_kernel): log_prob + VJP of one batch",
      "bound": "fp32", "achieved": ach, "peak": peaks["fp32"], "unit": "TFLOP/s", "frac": ach / peaks["fp32"],
      "peak_source": "FP32 FMA pipe measured in this run (surj_bench_fp32_peak); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
      "flops_per_launch": step_flops, "flops_dense_as_reference_per_launch": 3.0 * w.dense * batch,
      "us_per_launch": 1e3 * pair_ms, "traffic": traffic,
      "hbm": {"bound": "hbm", "achieved": bytes_launch / (pair_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
              "frac": bytes_launch / (pair_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": bytes_launch,
              "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks["file"] else "fallback (B200_PROFILING.md)"},
    }
  else:
    with torch.cuda.stream(ctx.stream):
      kinds = profile_kinds(lambda: w.step(0), ctx.dev)
    total_ms = sum(v["ms"] for v in kinds.values())
    # tensor-pipe peak per kernel kind: FP32-grade products are THREE tensor-core MMAs, so peak = dense rate / 3.
    #   kind::tf32 kernels (3xTF32: vm_gemm3.cuh, vm_tc.cuh)       dense TF32 tcgen05.mma rate measured in this run
    #   kind::f16 kernel (fp16x3: vm_fused.cuh)                    MEASURED_PEAKS.json dense bf16/fp16 rate (sustained: the
    #                                                              kernel runs inside a long step), else the fallback
    f16_dense = float(peaks["file"].get("bf16_tflops_sustained", peaks["file"].get("bf16_tflops", 1670.0)))
    f16_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks["file"] else "fallback 1670 (B200_PROFILING.md)"
    tensor_peak = peaks["tf32"] / 3.0

    def peak_of(kind):
      return (f16_dense / 3.0, f"{f16_src} ({f16_dense:.0f} TFLOP/s, kind::f16) / 3") if kind.startswith("fused_") else \
             (tensor_peak, f"dense TF32 tcgen05.mma rate measured in this run (surj_bench_tf32_peak: {peaks['tf32']:.0f} TFLOP/s) / 3")

    top = max(kinds, key=lambda k: kinds[k]["ms"]) if kinds else None
    for k, v in kinds.items():
      v["share"] = v["ms"] / total_ms if total_ms > 0 else 0.0
      v["needed_tflops"] = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 and v["flops"] > 0 else None
      if v["needed_tflops"] is not None and "tcgen05" in k:
        v["peak_tflops"] = peak_of(k)[0]
        v["frac"] = v["needed_tflops"] / v["peak_tflops"]
    if top is not None:
      t = kinds[top]
      pk, pk_src = peak_of(top)
      per_launch_flops = t["flops"] / max(t["launches"], 1)
      per_launch_ms = t["ms"] / max(t["launches"], 1)
      ach = (t["flops"] / (t["ms"] * 1e-3) / 1e12) if t["flops"] > 0 else 0.0
      traffic = None
      try:
        tj = json.load(open(os.path.join(ROOT, "profiles", f"r2_{cfg_name}_top_traffic.json")))
        traffic = tj.get(top, {}).get("dram_bytes_per_launch")
      except Exception:  # noqa: BLE001
        pass
      out["roofline"] = {
        "kernel": f"{top} (per-op executor; {t['launches']} launches, {100 * t['share']:.1f} % of the train step)",
        "bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk if pk else None,
        "peak_source": pk_src + " -- FP32-grade products are 3 tensor-core MMAs; numerator = NEEDED FLOPs (structural zeros excluded)",
        "flops_per_launch": per_launch_flops, "us_per_launch": 1e3 * per_launch_ms, "traffic": traffic,
        "whole_step": {"achieved": step_flops / (out["ms_per_step"] * 1e-3) / 1e12, "peak": tensor_peak,
                       "frac": step_flops / (out["ms_per_step"] * 1e-3) / 1e12 / tensor_peak if tensor_peak else None,
                       "flops": step_flops, "peak_note": "TF32 / 3 for the whole step (most of its products are 3xTF32)"},
        "kernels": kinds,
      }
  _dbg(f"{cfg_name}: roofline done")

  # --- e2e: pinned host batches through TrainStep.step_host (H2D of every batch, D2H of every loss)
  if e2e:
    losses_host = torch.zeros(steps, dtype=torch.float32).pin_memory()

    def host_steps(n, record=True):
      for i in range(n):
        k = i % w.P
        ts.step_host(w.host_y[k], w.host_x[k] if w.DX else None, loss_out=losses_host[i:i + 1] if record else None)
      ts.drain()

    with torch.cuda.stream(ctx.stream):
      host_steps(2 if not fused else max(3, min(warmup, 8)), record=False)
    e2e_ms = timed(ctx, host_steps, steps, align=False)
    out["e2e"] = {"value": ctx.world * batch * steps / (e2e_ms * 1e-3), "unit": "samples/s",
                  "h2d_bytes_per_step": 4 * batch * (w.D + w.DX), "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / steps,
                  "api": "surjectors_b200.train.TrainStep.step_host (pinned host batch in, loss out)"}
    assert np.isfinite(losses_host.numpy()).all()
    _dbg(f"{cfg_name}: e2e {e2e_ms / steps:.3f} ms/step")

  # --- data-parallel correctness riding along with every multi-GPU run (SCALE carries it)
  if ctx.world > 1:
    out["dp_check"] = dp_check(ctx, w)
  w.free()
  return out


def dp_check(ctx, w):
  """(i) replicas bit-identical after the timed steps; (ii) on a small batch, the summed shard gradients / losses of the
  N ranks against ONE rank evaluating the whole batch (the property the data-parallel step relies on)."""
  import torch
  import torch.distributed as dist

  from surjectors_b200 import parallel

  ts = w.ts
  flats = [torch.empty_like(ts.flat) for _ in range(ctx.world)]
  dist.all_gather(flats, ts.flat)
  identical = all(torch.equal(flats[0], f) for f in flats)
  nb = 512 * ctx.world + 3
  g = torch.Generator().manual_seed(4242)
  y = torch.randn((nb, w.D), generator=g).to(ctx.dev)
  x = torch.randn((nb, w.DX), generator=g).to(ctx.dev) if w.DX else None
  lo, hi = parallel.shard_bounds(nb, ctx.rank, ctx.world)
  n = w.n_par
  g_shard = torch.zeros(n + 4, dtype=torch.float32, device=ctx.dev)
  g_full = torch.zeros(n + 4, dtype=torch.float32, device=ctx.dev)
  ts.plan.bwd(y[lo:hi], None if x is None else x[lo:hi], ts.flat, g_lp=None, g_scale=-1.0, g_params=g_shard, loss_out=g_shard[n:n + 1])
  dist.all_reduce(g_shard)
  ts.plan.bwd(y, x, ts.flat, g_lp=None, g_scale=-1.0, g_params=g_full, loss_out=g_full[n:n + 1])
  torch.cuda.synchronize(ctx.dev)
  gerr = float((g_shard[:n] - g_full[:n]).abs().max() / g_full[:n].abs().max().clamp_min(1e-30))
  lerr = float((g_shard[n] - g_full[n]).abs() / g_full[n].abs().clamp_min(1.0))
  exch = "peer-memory mailbox (fused into the reduction kernel)" if ts.peer is not None else "NCCL all-reduce of the flat gradient"
  return {"replicas_identical": bool(identical), "grad_vs_single_rank_rel": gerr, "loss_vs_single_rank_rel": lerr,
          "exchange": exch, "check_batch": nb}


def run_b200(args):
  import torch

  import __graft_entry__ as entry
  from surjectors_b200 import _abi

  if not os.path.exists(_abi.LIB_PATH):
    entry.build()
  ctx = Ctx()
  if ctx.world != args.gpus and ctx.rank == 0:
    print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={ctx.world}; using WORLD_SIZE", file=sys.stderr)
  main_cfg, B = args.config, args.batch
  r = measure(ctx, main_cfg, B, args.steps, args.warmup, args.pool, args.no_graph, e2e=True, with_clocks=True)

  others = None
  if ctx.world == 1 and not args.no_others and main_cfg == CONFIG:
    others = {}
    for name in ("cfg5", "cfg2", "cfg3", "cfg1"):
      b = default_batch(name)
      wide = name in ("cfg4", "cfg5")
      try:
        o = measure(ctx, name, b, 3 if wide else 64, 3, 0, args.no_graph, e2e=False, quick=True)
        entry_ = {"workload": workload_name(name, b), "ms_per_step": o["ms_per_step"], "train_samples_per_sec": o["train_samples_per_sec"],
                  "log_prob_samples_per_sec": o["log_prob"]["samples_per_sec"], "log_prob_ms": o["log_prob"]["ms_per_call"],
                  "executor": o["executor"], "roofline_bound": o["roofline"]["bound"], "roofline_frac": o["roofline"]["frac"],
                  "roofline_achieved_tflops": o["roofline"]["achieved"], "roofline_peak_tflops": o["roofline"]["peak"]}
        if not wide:  # steady state: the kernel, not the launch
          big = 1 << 24
          s = measure(ctx, name, big, 4, 3, 2, True, e2e=False, quick=True)
          entry_["steady_state_b2p24"] = {"batch": big, "ms_per_step": s["ms_per_step"], "train_samples_per_sec": s["train_samples_per_sec"],
                                          "log_prob_samples_per_sec": s["log_prob"]["samples_per_sec"],
                                          "roofline_frac": s["roofline"]["frac"], "roofline_achieved_tflops": s["roofline"]["achieved"]}
        others[name] = entry_
      except Exception as e:  # noqa: BLE001  (an extra must never take the headline line down)
        others[name] = {"error": f"{type(e).__name__}: {e}"}
      _dbg(f"other config {name} done")

  cpu = None
  if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
    c = cpu_train_baseline(main_cfg, B, seconds=args.cpu_seconds)
    cpu = {"value": c["samples_per_s"], "unit": "samples/s", "cores": c["cores"], "kind": "port", "sample": cpu_sample_text(c, B)}

  if ctx.rank == 0:
    cfgd = config_dict(main_cfg, B, ctx.world, None)  # identical in both arms (the driver compares them)
    line = {
      "metric": "train_samples_per_sec", "value": r["train_samples_per_sec"], "unit": "samples/s", "n_gpus": ctx.world,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgd,
      "run": {"params": r["params"], "l2": r["l2"], "cuda_graph": r["cuda_graph"], "executor": r["executor"], "train_steps_per_sec": 1e3 / r["ms_per_step"],
              "needed_flops_per_sample_fwd": r["needed_flops_per_sample_fwd"], "dense_flops_per_sample_fwd": r["dense_flops_per_sample_fwd"]},
      "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "roofline": r["roofline"], "cpu_baseline": cpu,
      "log_prob": r["log_prob"], "final_loss": r["final_loss"],
    }
    if "dp_check" in r:
      line["dp_check"] = r["dp_check"]
    if others is not None:
      line["other_configs"] = others
    print(json.dumps(line), file=_RESULT_OUT, flush=True)
  if ctx.world > 1:
    # CUDA graphs that captured NCCL collectives keep the communicator busy at teardown
    # (destroy_process_group was seen to hang); everything is measured and printed: leave hard.
    torch.cuda.synchronize(ctx.dev)
    ctx.barrier()
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

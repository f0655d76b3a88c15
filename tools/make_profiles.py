"""Turns the ncu artefacts a gpurun call left in gpurun_out/ into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launch_summary(path, top=8):
  rows = list(csv.reader(open(path)))
  hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
  h = rows[hi]
  kn, mv = h.index("Kernel Name"), h.index("Metric Value")
  agg = collections.OrderedDict()
  for r in rows[hi + 1:]:
    if len(r) <= mv:
      continue
    try:
      v = float(r[mv].replace(",", ""))
    except ValueError:
      continue
    a = agg.setdefault(r[kn].split("(")[0][:60], [0, 0.0])
    a[0] += 1
    a[1] += v
  tot = sum(a[1] for a in agg.values())
  out = ["| launches | total us | avg us | share | kernel |", "|---|---|---|---|---|"]
  for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    out.append(f"| {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | {100 * t / tot:.1f}% | `{k}` |")
  return "\n".join(out)


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
  txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(txt.splitlines()))
  hdr, units, r = rows[0], rows[1], rows[-1]
  d = {h: (r[i], units[i]) for i, h in enumerate(hdr) if h in WANT}
  st = []
  for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
      try:
        st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
      except ValueError:
        pass
  return d, sorted(st, reverse=True)[:6]


def to_bytes(v):
  x, u = float(v[0]), v[1].lower()
  return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
  shutil.copy(os.path.join(G, "launches_r1_bench.csv"), os.path.join(P, "r1_bench_launches_ncu.csv"))
  shutil.copy(os.path.join(G, "bench_r1_n1.json"), os.path.join(P, "r1_bench_n1.json"))
  d, st = raw(os.path.join(G, "prof_r1_cfg2_bwd_final.ncu-rep"))
  rd, wr = to_bytes(d["dram__bytes_read.sum"]), to_bytes(d["dram__bytes_write.sum"])
  json.dump({"kernel": "chain_bwd_kernel<cfg2>", "dram_bytes_read": rd, "dram_bytes_write": wr,
             "source": "profiles/r1_cfg2_bwd_ncu.md"}, open(os.path.join(P, "r1_cfg2_bwd_traffic.json"), "w"))
  table = "\n".join(f"| `{k}` | {v[0]} {v[1]} |" for k, v in d.items())
  open(os.path.join(P, "r1_cfg2_bwd_ncu.md"), "w").write(f"""# ncu --set full, `surj::chain_bwd_kernel<cfg2>` (generated straight-line kernel, NT=64, 8 CTAs/SM)

Command (B200, `gpurun`): `ncu --set full --clock-control none --import-source on -k regex:chain_bwd -s 3 -c 1 python tools/prof_step.py`
(one train step of BASELINE cfg2, batch 65536; the same command exited 0 without ncu first).  Regenerate this file with
`python tools/make_profiles.py` after such a call.

| metric | value |
|---|---|
{table}

Top warp stall reasons (warps stalled per issue-active cycle): {', '.join(f'{n} {v:.2f}' for v, n in st)}.

Reading: DRAM traffic per launch = {rd / 1e6:.2f} MB read + {wr / 1e6:.2f} MB written vs 5.51 MB algorithmic
(84 B/sample x 65536 + parameters): no re-reads, the kernel is not memory bound.  It is latency bound on two fronts:
`short_scoreboard` (the warp-uniform LDS.128 weight loads and the shared-memory contraction feeding dependent FFMA2 /
FFMA) and `no_instruction` (the straight-line stream is ~140 KB per warp, SM instruction-cache hit rate
{d['sm__icc_request_hit_rate.pct'][0][:5]} %), at 13.8 resident warps per SM (batch 65 536 = 443 samples per SM; register-limited
at 128); FMA pipe about a fifth active.
History of this kernel in round 1 (CUDA-event time of kernel + reduction): 248 us (shared-memory float atomics in the gradient
contraction) -> 92 us (shuffle-combined, atomics-free contraction) -> 62 us (register cap 128, 64-thread CTAs, 8 CTAs/SM;
128/256/512-thread CTAs measured 79/118/145 us, 5-6 CTAs/SM 87 us) -> 59 us (one conditioner forward per layer in the backward)
-> 57 us (ticketed reduction with fused Adam) -> 53 us (approximate sigmoid for hidden activations only) -> 52 us
(MUFU-only sigmoid, 6-instruction expf-grade exp; packing the matvecs into FFMA2 over output pairs then removed another
8 % of the instructions -- 60.4 vs 64.8 us under ncu -- but under 1 % of the warm, graph-replayed step).  See DESIGN.md section 2 for the instruction-delivery microbenchmarks and
the two alternative kernels measured against this one.
""")
  open(os.path.join(P, "r1_launch_lists.md"), "w").write(f"""# ncu launch lists (`--metrics gpu__time_duration.sum --clock-control none`), round 1

Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.

## `python bench.py --steps 64 --warmup 8 --no-cpu-baseline` (cfg2 train step; raw list: r1_bench_launches_ncu.csv)

{launch_summary(os.path.join(P, 'r1_bench_launches_ncu.csv'))}

`chain_bwd_kernel` + `surj_reduce_partials_kernel` are the train step (2 launches; the Adam update is fused into the last
stage of the reduction); `chain_fwd_kernel` is the log_prob-only measurement of the same run.  `bench.py` measures the
step pair live with CUDA events (see r1_bench_n1.json: `roofline.us_per_launch`).

## `python tools/bench_wide.py --config cfg4 --batch 131072 --reps 1` (per-op executor, defaults; r1_cfg4_launches_ncu.csv)

{launch_summary(os.path.join(P, 'r1_cfg4_launches_ncu.csv'))}

## `python tools/bench_wide.py --config cfg5 --batch 131072 --reps 1` (r1_cfg5_launches_ncu.csv)

{launch_summary(os.path.join(P, 'r1_cfg5_launches_ncu.csv'))}

`vg_linear_pipe` = TMA-fed pipelined tcgen05 LINEAR on operand planes tiled by `vg_tile`; `vg_dw_tc` = register-staged
tcgen05 weight-gradient contraction; `vg_elem` = fused runs of per-sample ops (splines, affine, gathers, activations'
derivatives); `vg_linear` / `vg_dw` = FP32 SIMT tiles for operands narrower than 32.  Earlier in the round, with the
register-staged LINEAR and 16 384-sample chunks, the shares for cfg4 were `vg_linear_tc` 54 %, `vg_dw_tc` 31 %, `vg_elem` 13 %.
""")
  print(open(os.path.join(P, "r1_cfg2_bwd_ncu.md")).read())


if __name__ == "__main__":
  main()

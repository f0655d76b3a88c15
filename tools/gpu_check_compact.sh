set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/cc_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/cc_tests.log
tail -8 gpurun_out/cc_tests.log
timeout 300 python bench.py --steps 512 --warmup 32 --no-cpu-baseline > gpurun_out/cc_bench_gen1.json 2> gpurun_out/cc_bench.err; tail -3 gpurun_out/cc_bench.err
SURJ_B200_MODE=compact timeout 300 python bench.py --steps 512 --warmup 32 --no-cpu-baseline > gpurun_out/cc_bench_compact.json 2> gpurun_out/cc_bench.err; tail -3 gpurun_out/cc_bench.err
python - <<'PY'
import json
for n in ("gen1","compact"):
  d=json.load(open(f'gpurun_out/cc_bench_{n}.json'))
  print(n, d['ms_per_step'], d['value'], d['roofline']['us_per_launch'], d['log_prob']['us_per_call'])
PY

# Round-end evidence pass on one B200: bench line (with CPU baseline), ncu launch list of a short bench run, and one
# full ncu capture of the train kernel.  Outputs land in gpurun_out/; tools/make_profiles.py turns them into profiles/.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
timeout 600 python bench.py --steps 512 --warmup 32 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err || exit 1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_bench.csv \
    python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/prof_step.py > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:chain_bwd -s 3 -c 1 -o gpurun_out/prof_r1_cfg2_bwd_final -f \
    python tools/prof_step.py > gpurun_out/ncu_full.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['cpu_baseline']['value'], d['log_prob']['samples_per_sec'], d['clocks'])"

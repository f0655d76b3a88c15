# Round-end evidence pass on one B200: GPU tests, bench line (with CPU baseline), the reference arm, the ncu launch list of
# two train steps and one full ncu capture of the gradient-phase kernels.  Outputs land in gpurun_out/; copy what is to be
# judged into profiles/ (tools/make_profiles.py: launch_summary; tools/sass_mnemonics.py).
set -x
R=${R:-r2}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gpu_tests.log 2>&1; tail -3 gpurun_out/${R}_gpu_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err || exit 1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err
python tools/prof_step.py --config cfg4 --batch 262144 --steps 2 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_cfg4_launches_ncu.csv \
    python tools/prof_step.py --config cfg4 --batch 262144 --steps 2 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:vg_gemm3 -s 20 -c 4 -o gpurun_out/${R}_gemm3 -f \
    python tools/prof_step.py --config cfg4 --batch 262144 --steps 1 > gpurun_out/ncu_full.log 2>&1 && \
  ncu -i gpurun_out/${R}_gemm3.ncu-rep --page raw --csv > gpurun_out/${R}_gemm3_raw.csv && rm -f gpurun_out/${R}_gemm3.ncu-rep
python -c "
import json; d=json.load(open('gpurun_out/${R}_bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['cpu_baseline']['value'], d['log_prob']['samples_per_sec'], d['clocks'])"

# A/B of library variants under surjectors_b200/_lib/variants/ against the default build (bench.py, cfg2 train step)
for v in default $(ls surjectors_b200/_lib/variants/ | sed 's/\.so$//'); do
  if [ "$v" = default ]; then unset SURJ_B200_LIB; else export SURJ_B200_LIB=$PWD/surjectors_b200/_lib/variants/$v.so; fi
  timeout 300 python bench.py --steps 512 --warmup 32 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', d['ms_per_step'], d['roofline']['us_per_launch'], d['log_prob']['us_per_call'])"
done

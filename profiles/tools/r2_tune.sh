#!/bin/bash
# quick A/B on the GPU box: value and per-kernel ms (serialised pass) for each "VARIANT[:ENV=V[,ENV=V]]" argument
mkdir -p gpurun_out
: > gpurun_out/r2_tune.txt
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  [ "$v" == "default" ] && v=""
  line=$(env GBN_LIB_VARIANT=$v $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 --roofline-steps 1 2>gpurun_out/r2_tune_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_sweep'].items()}, 'rhat', d['max_gelman_rubin'])")
  echo "$spec  $line" | tee -a gpurun_out/r2_tune.txt
done

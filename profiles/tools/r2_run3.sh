#!/bin/bash
# parity on both fast paths (small-network persistent kernel on / off), bounds-checked build, then benches
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
tail -15 gpurun_out/r2_tests.log
GBN_SMALL=2 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests_nosmall.log 2>&1; echo "nosmall rc=$?" >> gpurun_out/r2_tests_nosmall.log
tail -5 gpurun_out/r2_tests_nosmall.log
GBN_LIB_VARIANT=chk python -m pytest tests -q -m gpu > gpurun_out/r2_tests_chk.log 2>&1; echo "chk rc=$?" >> gpurun_out/r2_tests_chk.log
tail -5 gpurun_out/r2_tests_chk.log
bash profiles/tools/r2_tune.sh default
for sh in "C1 3" "C2 64" "C2 512"; do
  set -- $sh
  for sm in 1 2; do
    echo "== $1 x $2 chains GBN_SMALL=$sm"
    GBN_SMALL=$sm python bench.py --shape $1 --chains $2 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --roofline-steps 1 2>gpurun_out/r2_small_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'launches', d['gpu_launches'], 'rhat', d['max_gelman_rubin'])"
  done
done 2>&1 | tee gpurun_out/r2_small.txt

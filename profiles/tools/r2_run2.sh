#!/bin/bash
# parity (default + bounds-checked build), then bench of the default and the 80-register S kernel
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
tail -5 gpurun_out/r2_tests.log
GBN_LIB_VARIANT=chk python -m pytest tests -q -m gpu > gpurun_out/r2_tests_chk.log 2>&1; echo "chk rc=$?" >> gpurun_out/r2_tests_chk.log
tail -5 gpurun_out/r2_tests_chk.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
for v in "$@"; do
  GBN_LIB_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err
done
python - "$@" <<'PY'
import json, sys
for v in ["default"] + sys.argv[1:]:
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, "%.3e" % d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_sweep"])
    except Exception as e:
        print(v, "failed", e)
PY

#!/bin/bash
# A/B of the two forms of the S kernel (k_fast_s: GBN_S_V=1, k_fast_s2: default) and of build variants of the second:
# the GPU suite, then one short bench line per arm (value, kernel ms per sweep)
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then shift
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s2_tests.log 2>&1; echo "tests rc=$?"
  tail -3 gpurun_out/r2_s2_tests.log
fi
SHORT="python bench.py --steps 6 --warmup 3 --e2e-steps 2 --roofline-steps 2 --no-cpu-baseline --no-secondary"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["roofline"]["kernel_ms_per_sweep"])'
GBN_S_V=1 $SHORT 2>/dev/null | python -c "$pick" v1
$SHORT 2>/dev/null | python -c "$pick" v2
for v in "$@"; do GBN_LIB_VARIANT=$v $SHORT 2>/dev/null | python -c "$pick" v2-$v; done

#!/bin/bash
# quick A/B on the GPU box: value and per-kernel ms (serialised pass) for each environment setting
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 --roofline-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_sweep'].items()})"; }
run GBN_X=1 "$@"

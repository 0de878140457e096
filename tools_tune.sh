#!/bin/bash
# tuning sweep on the GPU box: prints value and per-kernel ms for each setting
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3e' % d['value'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_sweep'].items()})"; }
run GBN_S_NT=256
run GBN_S_NT=128 GBN_TILE_TARGETS=256 GBN_TILE_SLOTS=4096
run GBN_S_NT=128 GBN_TILE_TARGETS=256 GBN_TILE_SLOTS=4096 GBN_STAGE_FX=2
run GBN_S_NT=256 GBN_STAGE_FX=2
run GBN_S_NT=128 GBN_TILE_TARGETS=128 GBN_TILE_SLOTS=2560 GBN_STAGE_FX=2
run GBN_S_NT=512 GBN_TILE_TARGETS=1024 GBN_TILE_SLOTS=16384
run GBN_LIB_VARIANT=fma GBN_S_NT=256

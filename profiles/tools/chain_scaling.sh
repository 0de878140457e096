#!/bin/bash
# edge-updates/s against the number of chains on one GPU (C3 network); writes gpurun_out/chain_scaling.txt
mkdir -p gpurun_out
: > gpurun_out/chain_scaling.txt
for c in 64 128 256 512 1024 4096; do
  st=3; [ $c -ge 1024 ] && st=2
  python bench.py --chains $c --steps $st --warmup 2 --no-cpu-baseline --e2e-steps 1 --roofline-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($c, '%.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_sweep'].items()})" | tee -a gpurun_out/chain_scaling.txt
done

#!/bin/bash
# Run on the GPU box (gpurun): the plain command first, then the same command under ncu.
#   full capture (source-level stalls) of one X, T, FX, S launch with the kernels serialised on one stream
#   (stream groups = 1), so that one launch covers all 256 chains exactly as in bench.py's per-kernel pass
# usage: profile_r2.sh <tag> [skip]
set -u
TAG=${1:-r2}
SKIP=${2:-18}
mkdir -p gpurun_out
SHORT="python bench.py --steps 1 --warmup 1 --sweeps-per-step 5 --e2e-steps 1 --roofline-steps 1 --no-cpu-baseline --no-secondary"
GBN_STREAM_GROUPS=1 $SHORT > gpurun_out/plain_full_$TAG.log 2>&1 &&
GBN_STREAM_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:k_fast_ -s $SKIP -c 4 -o gpurun_out/${TAG}_full $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_full_$TAG.log

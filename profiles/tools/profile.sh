#!/bin/bash
# Run on the GPU box (gpurun): the plain command first, then the same command under ncu.
#   1. launch list with device time per launch (cold-cache, serialised: compare SHARES)
#   2. full capture of the sweep kernels with the kernels serialised on one stream (stream groups = 1),
#      so that one launch covers all 256 chains exactly as in bench.py's per-kernel pass
#      (2 refresh kernels at model creation + 3 sweeps of 4 kernels are skipped: X, T, FX, S of sweep 4)
set -u
mkdir -p gpurun_out
SHORT="python bench.py --steps 1 --warmup 1 --sweeps-per-step 5 --e2e-steps 1 --roofline-steps 1 --no-cpu-baseline"
$SHORT > gpurun_out/plain_launchlist.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv $SHORT > gpurun_out/ncu_launchlist.log 2>&1
GBN_STREAM_GROUPS=1 $SHORT > gpurun_out/plain_full.log 2>&1 &&
GBN_STREAM_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:k_fast_ -s 14 -c 4 -o gpurun_out/r1_full $SHORT > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_launchlist.log gpurun_out/ncu_full.log

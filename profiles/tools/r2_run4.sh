#!/bin/bash
# round 2, single-chain X kernel: GPU suite, then the bench with the X phase on k_fast_x1 (default) and on the pair kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/r2_tests_x1.log; cat gpurun_out/r2_tests_x1.log
rm -f gpurun_out/r2_tune.txt
bash profiles/tools/r2_tune.sh default "default:GBN_X_LPT=4" "default:GBN_X_LPT=16"
python profiles/tools/flip_regimes.py > gpurun_out/r2_flip_regimes_x1.txt 2>&1; cat gpurun_out/r2_flip_regimes_x1.txt
python profiles/tools/x_track_stats.py > gpurun_out/r2_x_track_x1.txt 2>&1; cat gpurun_out/r2_x_track_x1.txt

#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_tune.txt
bash profiles/tools/r2_tune.sh default
python profiles/tools/flip_regimes.py > gpurun_out/r2_flip_regimes_x1.txt 2>&1; cat gpurun_out/r2_flip_regimes_x1.txt
GBN_X_LPT=4 python profiles/tools/flip_regimes.py > gpurun_out/r2_flip_regimes_x1_lpt4.txt 2>&1; cat gpurun_out/r2_flip_regimes_x1_lpt4.txt

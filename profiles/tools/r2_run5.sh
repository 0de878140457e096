#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
rm -f gpurun_out/r2_tune.txt
bash profiles/tools/r2_tune.sh default "default:GBN_X_LPT=4" "default:GBN_X1_NT=1024" "default:GBN_X1=2"
python profiles/tools/flip_regimes.py 2>&1 | cut -c1-420
python profiles/tools/c4_phases.py

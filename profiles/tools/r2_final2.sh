#!/bin/bash
# one-GPU evidence of round 2 (final kernels): the bench line, the reference arm, the flip-regime table, the window /
# resolve statistics, the C4 phase split, the launch list and the full ncu capture of the sweep kernels (each ncu run
# directly after the same command has exited 0 without ncu)
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python profiles/tools/flip_regimes.py > gpurun_out/r2_flip_regimes.txt 2>&1; echo "flip rc=$?"
GBN_X1=2 python profiles/tools/flip_regimes.py > gpurun_out/r2_flip_regimes_pair.txt 2>&1
python profiles/tools/x_track_stats.py > gpurun_out/r2_x_track.txt 2>&1
python profiles/tools/small_phases.py > gpurun_out/r2_small_phases.txt 2>&1
python profiles/tools/c4_phases.py > gpurun_out/r2_c4_phases.txt 2>&1
C4_CONTRASTS=64 python profiles/tools/c4_phases.py >> gpurun_out/r2_c4_phases.txt 2>&1
SHORT="python bench.py --steps 1 --warmup 1 --sweeps-per-step 5 --e2e-steps 1 --roofline-steps 1 --no-cpu-baseline --no-secondary"
$SHORT > gpurun_out/plain_launchlist_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $SHORT > gpurun_out/ncu_launchlist_r2.log 2>&1
bash profiles/tools/profile_r2.sh r2e 24

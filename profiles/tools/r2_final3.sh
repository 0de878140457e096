#!/bin/bash
# one-GPU evidence of round 2 with the second form of the S kernel (k_fast_s2): the whole GPU suite, the bench line, the
# reference arm, the flip-regime table, the window / resolve statistics, the small-shape and C4 phase splits, the
# launch list and the full ncu capture of the sweep kernels (each ncu run directly after the same command has exited 0
# without ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests_final.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python profiles/tools/flip_regimes.py > gpurun_out/r2_flip_regimes.txt 2>&1; echo "flip rc=$?"
GBN_X1=2 python profiles/tools/flip_regimes.py > gpurun_out/r2_flip_regimes_pair.txt 2>&1
python profiles/tools/x_track_stats.py > gpurun_out/r2_x_track.txt 2>&1
python profiles/tools/small_phases.py > gpurun_out/r2_small_phases.txt 2>&1
python profiles/tools/c4_phases.py > gpurun_out/r2_c4_phases.txt 2>&1
C4_CONTRASTS=64 python profiles/tools/c4_phases.py >> gpurun_out/r2_c4_phases.txt 2>&1
bash profiles/tools/r2_env_ab.sh GBN_S_V=1 - > gpurun_out/r2_s_kernels_ab.txt 2>&1
SHORT="python bench.py --steps 1 --warmup 1 --sweeps-per-step 5 --e2e-steps 1 --roofline-steps 1 --no-cpu-baseline --no-secondary"
$SHORT > gpurun_out/plain_launchlist_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $SHORT > gpurun_out/ncu_launchlist_r2.log 2>&1
bash profiles/tools/profile_r2.sh r2_final 24

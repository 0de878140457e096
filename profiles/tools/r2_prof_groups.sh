bash profiles/tools/profile_r2.sh r2g 24
SHORT="python bench.py --steps 6 --warmup 3 --e2e-steps 2 --roofline-steps 2 --no-cpu-baseline --no-secondary"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["roofline"]["kernel_ms_per_sweep"])'
for g in 2 3 4 5 6; do GBN_STREAM_GROUPS=$g $SHORT 2>/dev/null | python -c "$pick" groups-$g; done
python profiles/tools/c4_phases.py

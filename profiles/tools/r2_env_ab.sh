#!/bin/bash
# one short bench line per environment setting given as "VAR=val,VAR2=val2" arguments ("-" = defaults)
SHORT="python bench.py --steps 6 --warmup 3 --e2e-steps 2 --roofline-steps 2 --no-cpu-baseline --no-secondary"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["roofline"]["kernel_ms_per_sweep"].items()})'
for s in "$@"; do
  if [ "$s" = "-" ]; then $SHORT 2>/dev/null | python -c "$pick" defaults
  else env $(echo $s | tr ',' ' ') $SHORT 2>/dev/null | python -c "$pick" "$s"; fi
done

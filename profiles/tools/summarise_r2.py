#!/usr/bin/env python
"""`ncu -i <rep> --page raw --csv` -> profiles/<tag>_kernels_ncu_full.csv (one row per captured launch, the metrics
the roofline discussion uses) and profiles/<tag>_traffic.json (DRAM bytes read + written per launch, by kernel).

  ncu -i gpurun_out/r2c_full.ncu-rep --page raw --csv > /tmp/raw.csv
  python profiles/tools/summarise_r2.py /tmp/raw.csv r2
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
STALLS = ['barrier', 'branch_resolving', 'dispatch_stall', 'drain', 'lg_throttle', 'long_scoreboard', 'math_pipe_throttle',
          'membar', 'mio_throttle', 'misc', 'no_instruction', 'not_selected', 'selected', 'short_scoreboard', 'sleeping',
          'tex_throttle', 'wait']
SCALE = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}

rows = list(csv.reader(open(sys.argv[1])))
tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
out, traffic = [], {}
for r in rows[2:]:
    m = re.search(r"(k_\w+)(<[^>]*>)?", r[col['Kernel Name']])
    name = (m.group(1) + (m.group(2) or "")) if m else r[col['Kernel Name']][:40]
    d = {'kernel': name}
    for w in WANT:
        if w in col:
            d[f"{w} [{units[col[w]]}]"] = r[col[w]]
    for s in STALLS:
        k = f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio'
        if k in col:
            d['stall_' + s] = r[col[k]]
    out.append(d)
    b = sum(float(r[col[k]]) * SCALE[units[col[k]]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    traffic[m.group(1) if m else name] = b
keys = list(out[0].keys())
with open(os.path.join(ROOT, "profiles", f"{tag}_kernels_ncu_full.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(keys)
    for d in out:
        w.writerow([d.get(k, '') for k in keys])
json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
print(json.dumps(traffic))

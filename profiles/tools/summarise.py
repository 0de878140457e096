"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/:
  r1_launch_summary.csv   per-kernel launch counts / device time / share from the launch list
  r1_kernels_ncu_full.csv selected `ncu --set full` metrics + stall ratios of the sweep kernels
  traffic.json            dram bytes per launch (read by bench.py for roofline.traffic)
Run here (no GPU needed): python profiles/tools/summarise.py
"""
import collections
import csv
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")



def short_name(raw):
    """k_fast_s<0 1 0>: kernel name with its template arguments, no commas (this is a CSV field)."""
    m = re.search(r"(k_\w+)(<[^>]*>)?", raw)
    if not m:
        return raw.replace(",", " ")
    targs = (m.group(2) or "").replace("(int)", "").replace("(bool)", "").replace(",", "")
    return m.group(1) + targs


rows = [r for r in csv.reader(open(os.path.join(GP, "r1_launches.csv"))) if r and not r[0].startswith("==")]
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = short_name(r[idx["Kernel Name"]])
    v, u = float(r[idx["Metric Value"]]), r[idx["Metric Unit"]]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(os.path.join(OUT, "r1_launch_summary.csv"), "w") as f:
    f.write("kernel,launches,total_ms,avg_ms,share\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k},{n},{ms:.4f},{ms / n:.4f},{ms / tot:.4f}\n")

raw = subprocess.run(["ncu", "-i", os.path.join(GP, "r1_full.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
traffic = {}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
with open(os.path.join(OUT, "r1_kernels_ncu_full.csv"), "w") as f:
    f.write("kernel," + ",".join(f"{w} [{units[idx[w]]}]" if w in idx else w for w in want) + "," +
            ",".join(s.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "") for s in stalls) + "\n")
    for r in rows[2:]:
        name = short_name(r[idx["Kernel Name"]])
        f.write(name + "," + ",".join(r[idx[w]].replace(",", "") if w in idx else "" for w in want) + "," +
                ",".join(r[idx[s]] for s in stalls) + "\n")
        b = sum(float(r[idx[k]].replace(",", "")) * mult.get(units[idx[k]], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        key = next((k for k in ("k_fast_x", "k_fast_s", "k_fast_fx", "k_fast_tl", "k_fast_t") if k in name), name)
        traffic[key] = b
traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, kernels serialised "
                    "(GBN_STREAM_GROUPS=1), C3 x 256 chains; see profiles/r1_kernels_ncu_full.csv")
json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
print(open(os.path.join(OUT, "r1_launch_summary.csv")).read())
print(open(os.path.join(OUT, "r1_kernels_ncu_full.csv")).read())
print(traffic)

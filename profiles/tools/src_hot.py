#!/usr/bin/env python
"""Hot spots of one kernel from `ncu --page source --csv` (SASS view): executed instructions, stall
samples and shared-memory wavefronts by address range, plus the top individual instructions.

  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:k_fast_s > src.csv
  python profiles/tools/src_hot.py src.csv [n_top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
col = {h: i for i, h in enumerate(hdr)}


def f(r, name):
    try:
        return float(r[col[name]])
    except (ValueError, KeyError):
        return 0.0


tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
tot_wave = sum(f(r, "L1 Wavefronts Shared") for r in data)
print(f"instructions executed {tot_inst:.4g}   stall samples {tot_samp:.4g}   shared wavefronts {tot_wave:.4g} "
      f"(ideal {sum(f(r, 'L1 Wavefronts Shared Ideal') for r in data):.4g})")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stall mix (all samples):", ", ".join(f"{h[6:]} {100 * sum(f(r, h) for r in data) / max(tot_samp, 1):.1f}%"
                                          for h in sorted(stalls, key=lambda h: -sum(f(r, h) for r in data))[:8]))
print(f"\ntop {ntop} instructions by stall samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:ntop]:
    top = max(stalls, key=lambda h: f(r, h))
    print(f"  {r[col['Address']][-5:]}  {100 * f(r, '# Samples') / tot_samp:5.2f}%  exec {f(r, 'Instructions Executed'):.3g}  "
          f"{top[6:]:<14} wf {f(r, 'L1 Wavefronts Shared'):.3g}  {r[col['Source']][:90]}")
print(f"\ntop {ntop} instructions by shared wavefronts:")
for r in sorted(data, key=lambda r: -f(r, "L1 Wavefronts Shared"))[:ntop]:
    if f(r, "L1 Wavefronts Shared") == 0:
        break
    print(f"  {r[col['Address']][-5:]}  wf {f(r, 'L1 Wavefronts Shared'):.3g} (ideal {f(r, 'L1 Wavefronts Shared Ideal'):.3g})  "
          f"exec {f(r, 'Instructions Executed'):.3g}  {r[col['Source']][:90]}")

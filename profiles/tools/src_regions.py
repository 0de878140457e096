#!/usr/bin/env python
"""Executed-instruction profile of one kernel from `ncu -i <rep> --page source --csv --kernel-name regex:<k>`:
instructions executed and stall samples per region between the given SASS offsets (hex, relative to the kernel start),
and the hottest lines.

  python profiles/tools/src_regions.py /tmp/src.csv [warp_steps] [off1 off2 ...]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
base = int(body[0][col["Address"]], 16)
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.
cuts = sorted(int(x, 16) for x in sys.argv[3:])
tot = sum(int(r[col["Instructions Executed"]]) for r in body)
samp = sum(int(r[col["# Samples"]]) for r in body)
print(f"instructions executed {tot}  ({tot / units:.1f} per unit), samples {samp}")
reg, cur = [], [0, 0, 0]
ci = 0
start = 0
for r in body:
    off = int(r[col["Address"]], 16) - base
    while ci < len(cuts) and off >= cuts[ci]:
        reg.append((start, cuts[ci], *cur)); start = cuts[ci]; cur = [0, 0, 0]; ci += 1
    cur[0] += int(r[col["Instructions Executed"]]); cur[1] += int(r[col["# Samples"]]); cur[2] += 1
reg.append((start, off + 16, *cur))
for a, b, n, s, k in reg:
    print(f"  0x{a:04x}..0x{b:04x} ({k:4d} instr)  executed {n:12d} = {n / units:7.2f} per unit ({100. * n / tot:5.1f} %)   samples {100. * s / max(samp, 1):5.1f} %")
if "--lines" in sys.argv or True:
    pass

#!/usr/bin/env python
"""Executed-instruction profile of one kernel from the source page of an ncu report:

  ncu -i <rep> --page source --csv --kernel-name regex:<kernel> > /tmp/src.csv
  python profiles/tools/src_lines.py /tmp/src.csv <units> [--lines]

<units> = the number of work units one launch processes (the S kernel: warp-steps = chains x sum of group widths).
Prints the instructions executed per unit, then runs of consecutive SASS lines that execute equally often (loop
bodies, branches) with their share of the instructions and of the stall samples; --lines lists every SASS line.
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ie, sm = col["Instructions Executed"], col["# Samples"]
body, seen = [], set()
for r in rows[2:]:                       # (the page lists every SASS line twice with the same counts)
    if len(r) > ie and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0])
        body.append(r)
base = int(body[0][0], 16)
U = float(sys.argv[2])
L = [(int(r[0], 16) - base, int(r[ie]) / U, int(r[sm]), r[1].strip()) for r in body]
tot = sum(l[1] for l in L)
samp = sum(l[2] for l in L)
print(f"{len(L)} SASS lines, {tot * U:.0f} warp instructions executed = {tot:.2f} per unit, {samp} stall samples")
if "--lines" in sys.argv:
    for off, n, s, txt in L:
        print(f"{off:05x} {n:7.3f} {s:6d}  {txt[:90]}")
    sys.exit(0)
runs, cur = [], None
for off, n, s, txt in L:
    if cur and abs(n - cur["n"]) <= 0.12 * max(n, cur["n"], 0.02):
        cur["sum"] += n; cur["s"] += s; cur["k"] += 1; cur["end"] = off
    else:
        if cur:
            runs.append(cur)
        cur = {"start": off, "end": off, "n": n, "sum": n, "s": s, "k": 1}
runs.append(cur)
for r in runs:
    if r["sum"] > 0.003 * tot:
        print(f"  0x{r['start']:05x}..0x{r['end']:05x} {r['k']:4d} lines x {r['n']:6.3f} per unit = {r['sum']:7.2f} "
              f"({100 * r['sum'] / tot:4.1f} % of the instructions, {100 * r['s'] / max(samp, 1):4.1f} % of the samples)")

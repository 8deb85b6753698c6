#!/usr/bin/env python3
"""Per-source-line share of executed warp instructions / stall samples from
`ncu -i REP --page source --csv --print-source cuda,sass > file.csv` (lines with per-line totals)."""
import csv
import sys

rows, cur, hdr = [], None, None
tot_inst = tot_samp = 0
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0]:
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            inst, samp = int(r[i_inst] or 0), int(r[i_samp] or 0)
        except ValueError:
            continue
        rows.append((cur, int(r[0]), inst, samp, r[1]))
        tot_inst += inst
        tot_samp += samp
rows.sort(key=lambda x: -x[2])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
print(f"# total warp instructions {tot_inst}, stall samples {tot_samp}")
for f, l, i, s, src in rows[:n]:
    print(f"{f}:{l:4d} inst={100*i/tot_inst:5.2f}% samp={100*s/tot_samp:5.2f}%  {src.strip()[:110]}")

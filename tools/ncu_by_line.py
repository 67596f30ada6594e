#!/usr/bin/env python3
"""Aggregate an ncu source-page export (--page source --csv --print-source cuda,sass) by source line of one file.
Usage: ncu_by_line.py src.csv engine_core.h [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want, top = sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur, hdr, agg = None, None, {}
tot_s = tot_i = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 10:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if r[2] != "-":     # SASS rows carry an address; cuda rows ("-") have the line totals
        continue
    s, i = int(r[hdr.index("# Samples")] or 0), int(r[hdr.index("Instructions Executed")] or 0)
    tot_s += s; tot_i += i
    if cur.endswith(want):
        agg[line] = (s, i, r[1])
print(f"total samples {tot_s} instructions {tot_i}")
for line, (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{line:5d} {100.0 * s / max(tot_s, 1):6.2f}% smp {100.0 * i / max(tot_i, 1):6.2f}% inst  {src.strip()[:150]}")

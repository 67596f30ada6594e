#!/usr/bin/env python3
"""Aggregate an ncu SASS-level source export by out-of-line device function.
Usage: ncu -i rep --page source --csv --print-source sass > sass.csv ; ncu_by_function.py sass.csv obj kernel"""
import csv, re, subprocess, sys
sass_csv, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
funcs = []
for l in txt.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", l)
    if m and kern in m.group(3):
        funcs.append((int(m.group(1), 16), int(m.group(2), 16), re.sub(r"^_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}\d+|^_Z\d+", "", m.group(4))[:28]))
funcs.sort()
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, isamp, iinst, ith = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
base = None
agg = {}
for r in rows[2:]:
    if len(r) <= ith or not r[ia]:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    off = a - base
    name = "<kernel body>"
    for fo, fs, fn in funcs:
        if fo <= off < fo + fs:
            name = fn; break
    d = agg.setdefault(name, {"samples": 0, "inst": 0, "thr": 0, **{k: 0 for k in stall_cols}})
    d["samples"] += int(r[isamp] or 0); d["inst"] += int(r[iinst] or 0); d["thr"] += int(r[ith] or 0)
    for k, i in stall_cols.items():
        d[k] += int(r[i] or 0)
ts, ti = sum(d["samples"] for d in agg.values()), sum(d["inst"] for d in agg.values())
print(f"{'function':30s} {'samples%':>8s} {'inst%':>7s} {'lanes':>6s}  top stalls")
for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted(((v, k) for k, v in d.items() if k.startswith("stall_")), reverse=True)[:3]
    print(f"{name:30s} {100 * d['samples'] / ts:8.1f} {100 * d['inst'] / ti:7.1f} {d['thr'] / max(d['inst'], 1):6.1f}  " + ", ".join(f"{k[6:]} {100 * v / max(d['samples'], 1):.0f}%" for v, k in st))
print("total samples", ts, "instructions", ti)

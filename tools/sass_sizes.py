#!/usr/bin/env python3
"""Per-function SASS size of a kernel (bytes), from `cuobjdump -elf`.  Usage: sass_sizes.py <obj or so> [kernel substring]"""
import re, subprocess, sys
obj = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "k_step"
txt = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
rows = []
for l in txt.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", l)
    if m and kern in m.group(3):
        rows.append((int(m.group(2), 16), m.group(4)))
rows.sort()
tot = 0
for sz, name in rows:
    tot += sz
    print(f"{sz:8d}  {name[:70]}")
print(f"{tot:8d}  total of out-of-line functions in {kern}")

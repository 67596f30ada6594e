#!/usr/bin/env python3
"""Per-env shared-memory workspace layout of the engine for a robot / terrain (host emulation build; no GPU needed)."""
import ctypes, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from cosim_b200.config import make_config
from cosim_b200.model import build_model
from tests.hostsim import hostsim as H
cfg = bench.workload_config() if len(sys.argv) < 3 else make_config(sys.argv[1], sys.argv[2])
m = build_model(cfg)
h = H.HostSim(m, 1)
names = re.search(r"enum WsField \{(.*?)W__COUNT", open(os.path.join(os.path.dirname(H.__file__), "../../cosim_b200/csrc/engine_core.h")).read(), re.S).group(1)
names = [n.strip() for n in re.sub(r"//.*", "", names).replace("\n", " ").split(",") if n.strip()]
out = (ctypes.c_int * 128)()
L = H.lib() if hasattr(H, "lib") else H._lib
n = L.hs_layout(h.h, out, 128)
off = list(out[:n - 3]); ws, shared, arena = out[n - 3], out[n - 2], out[n - 1]
order = sorted(range(len(off)), key=lambda i: off[i])
print(f"ws_floats {ws} ({4 * ws} B/env), model+arena floats {shared} ({4 * shared} B), arena {arena} B")
for k, i in enumerate(order):
    nxt = min([off[j] for j in order[k + 1:] if off[j] > off[i]] + [ws])
    print(f"  {names[i]:12s} off {off[i]:5d}  span {nxt - off[i]:5d} floats")

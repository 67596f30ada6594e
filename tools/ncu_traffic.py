#!/usr/bin/env python3
"""One kernel launch of an `ncu --set full` report -> the metrics bench.py quotes (profiles/r02_traffic.json) and a readable raw
metric page (profiles/<tag>_raw_metrics.txt).
Usage: ncu_traffic.py report.ncu-rep name envs [tag] [note]     (name: bench config name, or k_policy_mlp)"""
import csv, json, os, subprocess, sys
rep, name, envs = sys.argv[1], sys.argv[2], int(sys.argv[3])
tag = sys.argv[4] if len(sys.argv) > 4 else "r02_" + name
note = sys.argv[5] if len(sys.argv) > 5 else ""
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
def f(k, d=None):
    try:
        return float(M[k][0].replace(",", ""))
    except Exception:
        return d
def to_bytes(k):
    v, u = f(k), M.get(k, ("", ""))[1]
    if v is None:
        return None
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def to_ms(k):
    v, u = f(k), M.get(k, ("", ""))[1]
    return None if v is None else v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warp_latency_per_inst_issued.ratio"]
WANT += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", tag + "_raw_metrics.txt"), "w") as o:
    o.write(f"# {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else name}: ncu --set full --clock-control none, one launch ({os.path.basename(rep)}), {envs} envs. {note}\n")
    for k in WANT:
        if k in M:
            o.write(f"{k:90s} {M[k][0]:>18s} {M[k][1]}\n")
e = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name, "envs": envs,
     "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"), "l2_bytes": to_bytes("lts__t_bytes.sum"),
     "gpu_time_ms_under_ncu": to_ms("gpu__time_duration.sum"), "warp_instructions_per_launch": f("smsp__inst_executed.sum"),
     "issue_slots_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps_active_pct_of_peak": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
     "fp32_pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
     "local_load_inst": f("sass__inst_executed_local_loads"), "local_store_inst": f("sass__inst_executed_local_stores"),
     "registers_per_thread": f("launch__registers_per_thread"), "block_size": int(f("launch__block_size", 0)), "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}; {note}"}
if e["dram_bytes_read"] is not None and e["dram_bytes_write"] is not None:
    e["traffic_bytes_per_launch"] = e["dram_bytes_read"] + e["dram_bytes_write"]
path = os.path.join(ROOT, "profiles", "r02_traffic.json")
allj = json.load(open(path)) if os.path.exists(path) else {}
allj[name] = e
json.dump(allj, open(path, "w"), indent=1)
print(json.dumps(e, indent=1))

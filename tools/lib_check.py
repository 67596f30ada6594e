#!/usr/bin/env python3
"""Two builds of the library against each other in one process: same seeds, same actions -> which steps / envs differ, and by how much.
lib_check.py libA.so libB.so [N] [steps] [robot terrain]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cosim_b200.lib as L
from cosim_b200.config import make_config, RANDOM_FULL
from cosim_b200.envs import BatchedEnv
la, lb = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
cfg = bench.workload_config() if len(sys.argv) < 7 else make_config(sys.argv[5], sys.argv[6], random=RANDOM_FULL, engine={"auto_reset": True})
os.environ["COSIM_LIB_PATH"] = "x"          # no rebuild
envs = []
for path in (la, lb):
    L._lib = None; L.LIB_PATH = os.path.abspath(path)
    envs.append(BatchedEnv(cfg, N, seed=0xC051))
torch.manual_seed(1)
cmd = torch.rand((N, envs[0].command_dim), device="cuda") * 3 - 1.5
outs = []
for e in envs:
    e.receive_user_command(cmd)
    outs.append(e.reset()[0].clone())
print("reset identical:", torch.equal(outs[0], outs[1]))
for k in range(steps):
    a = torch.rand((N, envs[0].action_dim), device="cuda") * 2 - 1
    for f in ("qpos", "qvel", "qacc_warmstart"):          # teacher-forced: env B starts every step from env A's state
        envs[1].set(f, envs[0].get(f))
    res = [e.step(a) for e in envs]
    d = (res[0][0] - res[1][0]).abs()
    bad = (d > 0).any(dim=1)
    qa, qb = envs[0].get("qvel"), envs[1].get("qvel")
    dq = (qa - qb).abs().max(dim=1).values
    nca, ncb = envs[0].get("counters")[:, 7], envs[1].get("counters")[:, 7]
    print(f"step {k:3d}: envs with different states {int(bad.sum()):6d} / {N}, max |dqvel| {float(dq.max()):.2e}, envs with |dqvel| > 1e-3: {int((dq > 1e-3).sum())}, different contact counts: {int((nca != ncb).sum())}", flush=True)

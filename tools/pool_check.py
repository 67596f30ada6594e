#!/usr/bin/env python3
"""Pooled step kernel (k_step_pool) against the lock-step kernel (k_step): same seeds, same actions -> bit-identical states,
done flags and statistics.  pool_check.py [N] [steps] [robot terrain]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cosim_b200.config import make_config, RANDOM_FULL
from cosim_b200.envs import BatchedEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = bench.workload_config() if len(sys.argv) < 5 else make_config(sys.argv[3], sys.argv[4], random=RANDOM_FULL, engine={"auto_reset": True})
envs = []
for r in ("0", os.environ.get("COSIM_POOL_R", "3")):
    os.environ["COSIM_POOL_R"] = r
    envs.append(BatchedEnv(cfg, N, seed=0xC051))
torch.manual_seed(1)
cmd = torch.rand((N, envs[0].command_dim), device="cuda") * 3 - 1.5
outs = []
for e in envs:
    e.receive_user_command(cmd)
    outs.append(e.reset()[0].clone())
assert torch.equal(outs[0], outs[1]), "reset states differ"
worst = 0.0
for k in range(steps):
    a = torch.rand((N, envs[0].action_dim), device="cuda") * 2 - 1
    res = [e.step(a) for e in envs]
    s0, s1 = res[0][0], res[1][0]
    same = torch.equal(s0, s1) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    if not same:
        d = (s0 - s1).abs(); bad = (d > 0).any(dim=1).nonzero().flatten()
        print(f"step {k}: {len(bad)} envs differ, max |diff| {float(d.max()):.3e}, first {bad[:8].tolist()}")
        worst = max(worst, float(d.max()))
        sys.exit(1)
for f in ("qpos", "qvel", "counters", "stats"):
    assert torch.equal(envs[0].get(f), envs[1].get(f)), f
print(f"pooled == lock-step over {steps} steps of {N} envs (states, done flags, qpos, qvel, counters, statistics bit-identical); episodes {envs[0].stats()['episodes']:.0f}")

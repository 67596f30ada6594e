#!/usr/bin/env python3
"""Tiny workload: reset + a few steps + one raw sub-step on every robot from perturbed joint poses (self-collision on)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cosim_b200.config import make_config, RANDOM_FULL
from cosim_b200.envs import BatchedEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cases = [("flamingo_p_v3", "rocky_hard"), ("humanoid_p_v0", "slope_hard"), ("flamingo_light_v1", "flat"), ("w4_p_v2", "stairs_up_hard")]
if len(sys.argv) > 3:
    cases = cases[:int(sys.argv[3])]
for robot, terrain in cases:
    env = BatchedEnv(make_config(robot, terrain, random=RANDOM_FULL, engine={"auto_reset": True}), N, seed=3)
    s, _ = env.reset()
    q = env.get("qpos"); q[:, 7:] += 0.3 * torch.randn_like(q[:, 7:]); env.set("qpos", q)       # provoke self contacts
    for _ in range(steps):
        s, _, _, _ = env.step(torch.rand((N, env.action_dim), device="cuda") * 2 - 1)
    env.substep()
    torch.cuda.synchronize()
    print(robot, terrain, "ok", bool(torch.isfinite(s).all()), flush=True)
    env.close()

#!/usr/bin/env python3
"""Small fixed workload for ncu captures: policy + step on N envs of a bench workload.  prof_run.py [N] [steps] [config name]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cosim_b200.envs import BatchedEnv
from cosim_b200.policy import MLPPolicy, synthetic_mlp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
name = sys.argv[3] if len(sys.argv) > 3 else "flamingo_rocky"
env = BatchedEnv(bench.workload_config(name), N, seed=0xC051)
pol = MLPPolicy(synthetic_mlp(env.state_dim, env.action_dim), "elu")
env.receive_user_command(torch.rand((N, env.command_dim), device="cuda") * 3 - 1.5)
s, _ = env.reset()
for _ in range(steps):
    s, _, _, _ = env.step(pol.get_action(s))
torch.cuda.synchronize()
print("done", float(s.abs().sum()))

#!/usr/bin/env python3
"""Device-timed env-steps/s of the bench workload (policy + step) for tuning sweeps:  quick_rate.py [N] [steps] [warmup]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cosim_b200.envs import BatchedEnv
from cosim_b200.policy import MLPPolicy, synthetic_mlp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 5
env = BatchedEnv(bench.workload_config(), N, seed=0xC051)
pol = MLPPolicy(synthetic_mlp(env.state_dim, env.action_dim), "elu")
env.receive_user_command(torch.rand((N, env.command_dim), device="cuda") * 3 - 1.5)
s, _ = env.reset()
for _ in range(warm):
    s, _, _, _ = env.step(pol.get_action(s))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(steps):
    s, _, _, _ = env.step(pol.get_action(s))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"wpb {env.warps_per_block if hasattr(env, 'warps_per_block') else '?'}  {N * 1e3 / ms / 1e6:.3f} M env-steps/s  {ms:.3f} ms/step", flush=True)

#!/usr/bin/env python3
"""Device-timed env-steps/s of the bench workload (policy + step) for tuning sweeps:  quick_rate.py [N] [steps] [warmup]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cosim_b200.lib as _L
if os.environ.get("COSIM_LIB_PATH"):
    _L.LIB_PATH = os.environ["COSIM_LIB_PATH"]
from cosim_b200.envs import BatchedEnv
from cosim_b200.policy import MLPPolicy, synthetic_mlp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = bench.workload_config()
if os.environ.get("COSIM_SELFCOL") == "0":
    cfg["engine"]["self_collision"] = False
env = BatchedEnv(cfg, N, seed=0xC051)
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
print(f"wpb {env._L.cosim_warps_per_block(env._h)} smem/env {env._L.cosim_smem_bytes_per_env(env._h)} npair {env.model.dim('npair')}  {N * 1e3 / ms / 1e6:.3f} M env-steps/s  {ms:.3f} ms/step", flush=True)

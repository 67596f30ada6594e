#!/usr/bin/env python3
"""Sustained run of the bench workload (or robot / terrain given on the command line): throughput per block of steps and
the reporter statistics that would reveal trouble (NaN resets, dropped contacts).  soak.py [steps] [N] [robot terrain]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cosim_b200.lib as _L
if os.environ.get("COSIM_LIB_PATH"):
    _L.LIB_PATH = os.environ["COSIM_LIB_PATH"]
from cosim_b200.config import make_config, RANDOM_FULL
from cosim_b200.envs import BatchedEnv
from cosim_b200.policy import MLPPolicy, synthetic_mlp
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
cfg = bench.workload_config() if len(sys.argv) < 5 else make_config(sys.argv[3], sys.argv[4], random=RANDOM_FULL, engine=dict({"auto_reset": True}, **({"ncon_max": int(os.environ["COSIM_NCON"])} if "COSIM_NCON" in os.environ else {})))
env = BatchedEnv(cfg, N, seed=0xC051)
pol = MLPPolicy(synthetic_mlp(env.state_dim, env.action_dim), "elu")
env.receive_user_command(torch.rand((N, env.command_dim), device="cuda") * 3 - 1.5)
s, _ = env.reset()
block = int(os.environ.get("COSIM_SOAK_BLOCK", "50"))
for b0 in range(0, steps, block):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(min(block, steps - b0)):
        s, _, _, _ = env.step(pol.get_action(s))
    e1.record(); torch.cuda.synchronize()
    n = min(block, steps - b0)
    st = env.stats()
    print(f"steps {b0:4d}-{b0 + n:4d}: {N * n / (e0.elapsed_time(e1) * 1e-3) / 1e6:.3f} M env-steps/s  finite {bool(torch.isfinite(s).all())}  "
          + "  ".join(f"{k} {st[k]:.4g}" for k in ("episodes", "termination_rate", "mean_contacts", "mean_solver_iters_per_step", "nan_resets", "ncon_dropped") if k in st), flush=True)

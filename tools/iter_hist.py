#!/usr/bin/env python3
"""Distribution of per-env solver iterations / contacts per control step (imbalance behind the phase-barrier waits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from cosim_b200.envs import BatchedEnv
from cosim_b200.policy import MLPPolicy, synthetic_mlp
N = 16384
env = BatchedEnv(bench.workload_config(), N, seed=0xC051, debug=True)
pol = MLPPolicy(synthetic_mlp(env.state_dim, env.action_dim), "elu")
env.receive_user_command(torch.rand((N, env.command_dim), device="cuda") * 3 - 1.5)
s, _ = env.reset()
for k in range(12):
    s, _, _, _ = env.step(pol.get_action(s))
it = env.get("iters")[:, 0].cpu().numpy(); nc = env.get("counters")[:, 7].cpu().numpy()
print("newton iterations per control step (4 sub-steps): mean %.1f  p50 %d  p90 %d  p99 %d  max %d" % (it.mean(), *np.percentile(it, [50, 90, 99]).astype(int), it.max()))
print("contacts in the last sub-step: mean %.2f  p90 %d  p99 %d  max %d" % (nc.mean(), *np.percentile(nc, [90, 99]).astype(int), nc.max()))
g = it[: (N // 19) * 19].reshape(-1, 19)
print("per CTA of 19 envs: mean of max %.1f vs mean %.1f  (ratio %.2f)" % (g.max(axis=1).mean(), g.mean(), g.max(axis=1).mean() / g.mean()))

#!/usr/bin/env python3
"""Two settings of a create-time switch against each other in one process; the results must be bit-identical (states, done flags, qpos,
qvel, counters, statistics).  Default switch: COSIM_BSYNC_MASK (barriers only change WHEN code runs); COSIM_CHECK_VAR names another one,
e.g. COSIM_CHECK_VAR=COSIM_HF_FINE mask_check.py 0 1 (the two instances of the terrain pass only differ in the work they skip).
mask_check.py A B [N] [steps] [robot terrain]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cosim_b200.config import make_config, RANDOM_FULL
from cosim_b200.envs import BatchedEnv
ma, mb = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cfg = bench.workload_config() if len(sys.argv) < 7 else make_config(sys.argv[5], sys.argv[6], random=RANDOM_FULL, engine={"auto_reset": True})
os.environ["COSIM_POOL_R"] = "0"
envs = []
for r in (ma, mb):
    os.environ[os.environ.get("COSIM_CHECK_VAR", "COSIM_BSYNC_MASK")] = r
    envs.append(BatchedEnv(cfg, N, seed=0xC051))
torch.manual_seed(1)
cmd = torch.rand((N, envs[0].command_dim), device="cuda") * 3 - 1.5
outs = []
for e in envs:
    e.receive_user_command(cmd)
    outs.append(e.reset()[0].clone())
assert torch.equal(outs[0], outs[1]), "reset states differ"
for k in range(steps):
    a = torch.rand((N, envs[0].action_dim), device="cuda") * 2 - 1
    res = [e.step(a) for e in envs]
    s0, s1 = res[0][0], res[1][0]
    if not (torch.equal(s0, s1) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])):
        d = (s0 - s1).abs(); bad = (d > 0).any(dim=1).nonzero().flatten()
        print(f"step {k}: {len(bad)} envs differ, max |diff| {float(d.max()):.3e}, first {bad[:8].tolist()}")
        sys.exit(1)
for f in ("qpos", "qvel", "counters", "stats"):
    assert torch.equal(envs[0].get(f), envs[1].get(f)), f
print(f"{os.environ.get('COSIM_CHECK_VAR', 'mask')} {ma} == {mb} over {steps} steps of {N} envs (bit-identical); episodes {envs[0].stats()['episodes']:.0f}")

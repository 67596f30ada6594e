#!/usr/bin/env python3
"""Quick GPU-vs-oracle comparison printed as a table (development aid; the asserts live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cosim_b200.config import make_config, RANDOM_NONE, RANDOM_FULL
from cosim_b200.model import build_model
from cosim_b200.envs import BatchedEnv
from oracle.oracle import Oracle

def run(rid, terr, N=8, steps=12, random=RANDOM_NONE, hm=False):
    kw = {}
    if hm:
        from cosim_b200.config import load_tables
        et, _ = load_tables()
        kw["non_stacked_obs_order"] = list(et[rid]["non_stacked_obs_order"]) + ["height_map"]
    cfg = make_config(rid, terr, random=random, **kw)
    env = BatchedEnv(cfg, N, seed=1, debug=True)
    o = Oracle(env.model, N, seed=1)
    so = o.reset(); sg, _ = env.reset()
    print(rid, terr, "smem/env", env._L.cosim_smem_bytes_per_env(env._h), "wpb", env._L.cosim_warps_per_block(env._h),
          "reset state maxdiff %.2e" % np.abs(so - sg.cpu().numpy()).max())
    rng = np.random.default_rng(0)
    for i in range(steps):
        a = rng.uniform(-1, 1, (N, env.action_dim))
        for k in ["qpos", "qvel", "qacc_warmstart"]:
            env.set(k, o.get(k))
        so, te, tr = o.step(a); sg, tg, trg, info = env.step(a)
        qo, qg = o.get("qvel"), env.get("qvel").cpu().numpy()
        err = np.abs(qo - qg).max(axis=1)
        print("  step %2d qvel maxerr med %.2e max %.2e  state maxdiff %.2e  ncon o %s g %s" % (
            i, np.median(err), err.max(), np.abs(so - sg.cpu().numpy()).max(), o.get("ncon")[:, 0].astype(int), env.get("counters")[:, 7].cpu().numpy()))
    env.close()

def perf(rid, terr, N, steps=20, random=RANDOM_FULL):
    cfg = make_config(rid, terr, random=random, engine={"auto_reset": True})
    env = BatchedEnv(cfg, N, seed=1)
    env.reset()
    a = torch.rand((N, env.action_dim), device="cuda") * 2 - 1
    for _ in range(3):
        env.step(a)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(steps):
        env.step(a)
    torch.cuda.synchronize(); dt = time.time() - t
    print("perf", rid, terr, "N", N, "%.3f ms/step  %.3f M env-steps/s" % (dt / steps * 1e3, N * steps / dt / 1e6), env.stats())
    env.close()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run("flamingo_p_v3", "rocky_hard")
    run("flamingo_light_v1", "flat")
    run("w4_p_v2", "stairs_up_hard", steps=6)
    run("humanoid_p_v0", "slope_hard", steps=6)
    run("flamingo_p_v3", "rocky_hard", hm=True, steps=3)
    for N in (4096, 65536):
        perf("flamingo_p_v3", "rocky_hard", N)
    perf("w4_p_v2", "stairs_up_hard", 16384, steps=5)
    perf("humanoid_p_v0", "slope_hard", 16384, steps=5)
    perf("flamingo_light_v1", "flat", 65536)

#!/usr/bin/env python3
"""Diagnostic: general constraint path (engine options on the command line as k=v) on the GPU against the fp64 AND fp32 oracle,
teacher-forced sub-steps; prints the error bands of same-geometry sub-steps and the worst cases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cosim_b200.config import make_config, RANDOM_FULL
from cosim_b200.envs import BatchedEnv
from oracle.oracle import Oracle
from tests.parity_util import geometry_gap, DEPTH_SAME, NORMAL_SAME
eng = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    eng[k] = v if not v.replace(".", "").isdigit() else (float(v) if "." in v else int(v))
N = 64
env = BatchedEnv(make_config("flamingo_p_v3", "rocky_hard", random=RANDOM_FULL, engine=dict(eng)), N, seed=1, debug=True)
o, f = Oracle(env.model, N, seed=1), Oracle(env.model, N, seed=1, use_float=True)
o.reset(); f.reset(); env.reset()
rng = np.random.default_rng(11)
cap = env.model.dim("ncon_max")
eg, ef, tag = [], [], []
for i in range(8):
    a = rng.uniform(-1, 1, (N, env.action_dim))
    for k in ("qpos", "qvel", "qacc_warmstart"):
        f.set(k, o.get(k))
    o.step(a); f.step(a)
    for s in range(2):
        for k in ("qpos", "qvel", "qacc_warmstart", "torque"):
            env.set(k, o.get(k))
        for k in ("qpos", "qvel", "qacc_warmstart"):
            f.set(k, o.get(k))
        o.substep(); f.substep(); env.substep()
        nco, ncg, ncf = o.get("ncon")[:, 0].astype(int), env.get("counters")[:, 7].cpu().numpy(), f.get("ncon")[:, 0].astype(int)
        errg = np.abs(o.get("qvel") - env.get("qvel").cpu().numpy()).max(axis=1); errf = np.abs(o.get("qvel") - f.get("qvel")).max(axis=1)
        cg_all = env.get("contacts").cpu().numpy().reshape(N, cap, 10)
        it = env.get("solver_iters").cpu().numpy() if False else None
        for e in np.nonzero(nco == ncg)[0]:
            co = o.contacts(int(e), cap); dd, dn = geometry_gap(co, cg_all[e, :len(co)])
            if dd < DEPTH_SAME and dn < NORMAL_SAME:
                eg.append(errg[e]); ef.append(errf[e] if ncf[e] == nco[e] else np.nan); tag.append((i, s, int(e), int(nco[e]), float(np.abs(o.get("qvel")[e]).max())))
eg, ef = np.array(eg), np.array(ef)
print(eng, len(eg), "engine: max %.2e q99 %.2e q999 %.2e | fp32 oracle: max %.2e q99 %.2e" % (eg.max(), np.quantile(eg, .99), np.quantile(eg, .999), np.nanmax(ef), np.nanquantile(ef, .99)))
for j in np.argsort(eg)[-6:]:
    print("  step %d sub %d env %d ncon %d max|qvel| %.2f: engine %.2e  fp32 oracle %.2e" % (*tag[j], eg[j], ef[j]))

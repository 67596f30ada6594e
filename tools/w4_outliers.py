#!/usr/bin/env python3
"""How often does a teacher-forced sub-step with the fp64 oracle's contact geometry stray beyond SAME_GEOMETRY_TOL, and what does the solver
do there?  w4_outliers.py [robot terrain] [nseeds]   (uses COSIM_LIB_PATH if set)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_gpu_parity import _env, _oracle, geometry_gap, compare_contact_lists, DEPTH_SAME, NORMAL_SAME, SAME_GEOMETRY_TOL
robot, terrain = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("w4_p_v2", "stairs_up_hard")
nseeds = int(sys.argv[3]) if len(sys.argv) > 3 else 4
N, steps = 64, 8
for seed in range(nseeds):
    env = _env(robot, terrain, N); orc = _oracle(env, N)
    orc.reset(); env.reset()
    rng = np.random.default_rng(seed)
    cap = env.model.dim("ncon_max"); gtype = env.model.sections["geom_type"]
    same_geo, out = [], []
    for i in range(steps):
        a = rng.uniform(-1, 1, (N, env.action_dim))
        for k in ("qpos", "qvel", "qacc_warmstart", "torque"):
            env.set(k, orc.get(k))
        orc.step(a); env.step(a)
        for s in range(2):
            for k in ("qpos", "qvel", "qacc_warmstart", "torque"):
                env.set(k, orc.get(k))
            orc.substep(); env.substep()
            err = np.abs(orc.get("qvel") - env.get("qvel").cpu().numpy()).max(axis=1)
            cg_all = env.get("contacts").cpu().numpy().reshape(N, cap, 10); ncg = env.get("counters")[:, 7].cpu().numpy()
            try:
                it_g = env.get("iters").cpu().numpy().reshape(N, -1)[:, 0]
            except Exception:
                it_g = None
            try:
                it_o = orc.get("solver_iter")[:, 0]
            except Exception:
                it_o = None
            for e in range(N):
                co = orc.contacts(e, cap); cg = cg_all[e, :ncg[e]]
                same, _ = compare_contact_lists(co, cg, gtype)
                if same:
                    dd, dn = geometry_gap(co, cg)
                    if dd < DEPTH_SAME and dn < NORMAL_SAME:
                        same_geo.append(err[e])
                        if err[e] > SAME_GEOMETRY_TOL:
                            out.append((i, s, e, float(err[e]), len(co), None if it_g is None else int(it_g[e]), None if it_o is None else int(it_o[e])))
    same_geo = np.array(same_geo)
    print(f"seed {seed}: same-geometry sub-steps {len(same_geo)}, worst {same_geo.max():.1e}, 99.8 % {np.quantile(same_geo, 0.998):.1e}, above tol: {len(out)}  {out}", flush=True)
    env.close()

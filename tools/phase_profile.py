#!/usr/bin/env python3
"""Per-phase cycle breakdown of the step kernel (profiling build with -DCOSIM_PHASE_TIMING).

Build:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DCOSIM_PHASE_TIMING -shared \
        -Xcompiler -fPIC -o cosim_b200/csrc/_build_prof/libcosim_b200_prof.so cosim_b200/csrc/engine.cu cosim_b200/csrc/policy.cu
Run (GPU box):  python tools/phase_profile.py [robot terrain N steps]
"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosim_b200 import lib as L
L.LIB_PATH = os.path.join(L.CSRC, "_build_prof", "libcosim_b200_prof.so")
L._stale = lambda: False
import torch
from cosim_b200.config import make_config, RANDOM_FULL
from cosim_b200.envs import BatchedEnv

NAMES = ["kin+crb+chol", "collide", "constraint", "smooth", "newton", "integrate", "obs", "io", "newton_iters", "ls_evals", "pair_mpr_calls", "mpr_calls",
         "wait after kin", "wait after collide", "wait after smooth", "wait after newton"]

def run(robot, terrain, N, steps):
    cfg = make_config(robot, terrain, random=RANDOM_FULL, engine={"auto_reset": True})
    env = BatchedEnv(cfg, N, seed=1)
    lib = L.lib(); lib.cosim_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    env.reset()
    a = torch.rand((N, env.action_dim), device="cuda") * 2 - 1
    for _ in range(3):
        env.step(a)
    buf = (ctypes.c_ulonglong * 144)()
    lib.cosim_phase_cycles(env._h, buf, 1)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(steps):
        env.step(a)
    torch.cuda.synchronize(); dt = time.time() - t
    lib.cosim_phase_cycles(env._h, buf, 1)
    v = list(buf)
    tot = sum(v[:8]) + sum(v[12:16])
    nsub = N * steps * 4
    print(f"{robot} {terrain} N={N}: {dt / steps * 1e3:.2f} ms/step, {N * steps / dt / 1e6:.3f} M env-steps/s; smem/env {lib.cosim_smem_bytes_per_env(env._h)} wpb {lib.cosim_warps_per_block(env._h)}")
    for i in range(8):
        print(f"  {NAMES[i]:14s} {100.0 * v[i] / max(tot, 1):5.1f} %   {v[i] / nsub:10.0f} cycles/sub-step")
    for i in range(12, 16):
        print(f"  {NAMES[i]:18s} {100.0 * v[i] / max(tot, 1):5.1f} %   {v[i] / nsub:10.0f} cycles/sub-step")
    for i in range(8, 12):
        print(f"  {NAMES[i]:14s} {v[i] / nsub:8.2f} per sub-step")
    for name, base in (("collide", 16), ("newton", 80)):
        h = v[base:base + 64]; tot_h = max(sum(h), 1); acc = 0; line = []
        for i, c in enumerate(h):
            acc += c
            if c: line.append(f"{8 * i}K:{100.0 * c / tot_h:.1f}%")
        print(f"  {name} time histogram (8 K-cycle bins): " + " ".join(line))
    print("  stats:", {k: round(x, 3) for k, x in env.stats().items() if k in ("mean_contacts", "mean_solver_iters_per_step", "termination_rate", "episodes")})
    env.close()

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    else:
        run("flamingo_p_v3", "rocky_hard", 16384, 10)
        run("flamingo_light_v1", "flat", 16384, 10)
        run("humanoid_p_v0", "slope_hard", 8192, 5)
        run("w4_p_v2", "stairs_up_hard", 4096, 3)

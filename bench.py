#!/usr/bin/env python3
"""bench.py -- env-steps/s of the batched stepping engine (BASELINE.json metric).

A "step" = one control step (50 Hz) of every environment: policy MLP forward (tcgen05) + cosim_step
(delay, PD, frame_skip x rigid-body sub-steps with collision and the Newton contact solve, sensors,
noise, height-map rays, state build).  Workload: flamingo_p_v3 on rocky_hard, height map in the
observation, full randomization, 65 536 envs per GPU (weak scaling), synthetic policy + commands.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...           (one rank per GPU)

`--impl reference` times the CPU implementation of the same path on the host cores: the oracle port
(oracle/, "CPU restatement, not MuJoCo" -- the reference's own arithmetic lives in the un-vendored mujoco
wheel, which is not installable here), on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "env-steps/s"
ROBOT, TERRAIN = "flamingo_p_v3", "rocky_hard"
B_ALG = 1832      # algorithmic HBM bytes per env-step, flamingo_p_v3 with height map


def workload_config():
    from cosim_b200.config import make_config, RANDOM_FULL, load_tables
    et, _ = load_tables()
    non_stacked = list(et[ROBOT]["non_stacked_obs_order"]) + ["height_map"]
    return make_config(ROBOT, TERRAIN, random=RANDOM_FULL, non_stacked_obs_order=non_stacked,
                       engine={"auto_reset": True, "seed": 0xC051})


def _ncu_capture(envs):
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        return t if int(t["envs"]) == int(envs) else None
    except Exception:
        return None


def measured_traffic(envs):
    """DRAM bytes per k_step launch from the committed ncu capture (profiles/r01_traffic.json), if it matches this size."""
    t = _ncu_capture(envs)
    return int(t["traffic_bytes_per_launch"]) if t else None


def issue_metrics(envs):
    """What actually bounds k_step (same ncu capture): warp instructions per launch and the share of issue slots in use."""
    t = _ncu_capture(envs)
    if not t or "issue_slots_active_pct" not in t:
        return None
    return {"warp_instructions_per_env_step": t["warp_instructions_per_launch"] / float(envs), "issue_slots_active_pct": t["issue_slots_active_pct"],
            "registers_per_thread": t.get("registers_per_thread"), "env_warps_per_sm": t.get("block_size", 0) // 32, "source": "profiles/r01_traffic.json (ncu --set full)"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_reference(cfg, n_envs, steps, warmup, threads=0):
    """The CPU port of the path (oracle) on `n_envs` envs, all host threads.  Returns env-steps/s and details."""
    from cosim_b200.model import build_model
    from oracle.oracle import Oracle, lib as olib
    model = build_model(cfg)
    # all host threads this process may use; torchrun exports OMP_NUM_THREADS=1, which would silently make this a 1-core run
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cores = avail if threads <= 0 else threads
    olib()
    orc = Oracle(model, n_envs, seed=0xC051)
    rng = np.random.default_rng(7)
    cmd = rng.uniform(-1.5, 1.5, (n_envs, model.dim("command_dim")))
    orc.reset(command=cmd)
    act = rng.uniform(-1, 1, (n_envs, model.dim("nu")))
    for _ in range(warmup):
        _, term, trunc = orc.step(act, cmd, nthreads=cores)
        done = term | trunc
        if done.any():
            orc.reset(mask=done, command=cmd)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, term, trunc = orc.step(act, cmd, nthreads=cores)
        done = term | trunc
        if done.any():
            orc.reset(mask=done, command=cmd)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config()
    n = args.ref_envs
    value, cores, dt = cpu_reference(cfg, n, args.steps, args.warmup)
    sample = f"{n} envs x {args.steps} control steps of the same workload (fp64 C++ restatement, OpenMP over envs)"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{ROBOT} on {TERRAIN}, height map 12x12, full randomization, {n} envs (bounded CPU sample)"},
           "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _RESULT_OUT.write(json.dumps(out) + "\n"); _RESULT_OUT.flush()


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=65536, help="environments per GPU")
    ap.add_argument("--impl", default="cosim_b200", choices=["cosim_b200", "reference"])
    ap.add_argument("--ref-envs", type=int, default=1024)
    ap.add_argument("--cpu-envs", type=int, default=512)
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner under torchrun) are sent to
    # stderr, and the result is written to a private duplicate of the original stdout
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cosim_b200.envs import BatchedEnv
    from cosim_b200.policy import MLPPolicy, synthetic_mlp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU path); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N = args.envs
    cfg = workload_config()
    env = BatchedEnv(cfg, N, device=dev, seed=0xC051, env_offset=rank * N)     # RNG substream = global env id
    pol = MLPPolicy(synthetic_mlp(env.state_dim, env.action_dim), "elu", dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
    cmd = (torch.rand((N, env.command_dim), device=dev, generator=gen) * 3.0 - 1.5)
    env.receive_user_command(cmd)
    state, _ = env.reset()

    def one_step(s):
        a = pol.get_action(s)
        s2, _, _, _ = env.step(a)
        return s2

    for _ in range(max(args.warmup, 3)):
        state = one_step(state)
    # ---------------- device-resident timing (value)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local); clocks.start()
    l0 = env.launch_count + pol.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for i in range(args.steps):
        a = pol.get_action(state)
        kev[i][0].record()
        state, _, _, _ = env.step(a)
        kev[i][1].record()
    ev[1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = env.launch_count + pol.launch_count - l0
    elapsed_ms = ev[0].elapsed_time(ev[1])
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    clk = clocks.stop()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * N * args.steps / (elapsed_ms * 1e-3)

    # ---------------- end-to-end through the public API with HOST buffers.  Every step: H2D of that step's inputs (the user
    # commands, pinned), policy on the device-resident state, cosim_step, D2H of the results (state + done flags, pinned), sync.
    nu, sd, cd = env.action_dim, env.state_dim, env.command_dim
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    h_state, h_cmd = pin((N, sd), torch.float32), pin((N, cd), torch.float32)
    h_term, h_trunc = pin((N,), torch.uint8), pin((N,), torch.uint8)
    h_cmd.copy_(cmd.cpu())
    assert h_state.is_pinned() and h_cmd.is_pinned()
    # Same simulated interval as the device-timed leg: reset, the same warm-up steps, then the same K steps (the cost of
    # a step depends on what the robots are doing, so a leg timed later in the episodes would measure a different workload)
    e2e_steps = args.steps
    env.reset()
    for _ in range(max(args.warmup, 3)):
        env.step_policy_host(pol, h_cmd, h_state, h_term, h_trunc)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        env.step_policy_host(pol, h_cmd, h_state, h_term, h_trunc)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * N * e2e_steps / (float(t.item()) * 1e-3)
    h2d = N * 4 * cd
    d2h = N * (4 * sd + 2)

    stats = env.stats(all_reduce=True)          # the one collective of this path: NCCL all-reduce of reporter statistics
    peak, peak_src = measured_peaks()
    achieved = B_ALG * N / (kernel_ms * 1e-3) / 1e9
    out = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 (physics), bf16 x bf16 -> f32 (policy MLP)", "data": "synthetic",
           "config": {"workload": f"{ROBOT} on {TERRAIN}, height map 12x12 in the observation, full randomization (friction, mass noise, load, "
                                  f"action delay, Kp/Kd), {N} envs per GPU, medium precision (4 sub-steps of 5 ms per control step), "
                                  "synthetic MLP policy state->512->256->128->8 + per-env velocity commands, auto-reset on termination",
                      "envs_per_gpu": N, "sub_steps_per_s": value * 4,
                      "l2": "per-env state, parameter and observation arrays total > 126 MB L2 at 65 536 envs (inputs larger than L2); no explicit flush"},
           "clocks": clk, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                   "path": "BatchedEnv.step_policy_host: pinned host commands -> device, tcgen05 policy on the device state, cosim_step, "
                           "state + done flags -> pinned host, stream sync (per step)"},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(N),
                        "kernel": "k_step", "kernel_ms": kernel_ms, "bytes_per_env_step": B_ALG, "peak_source": peak_src,
                        "issue": issue_metrics(N),
                        "note": "the step is bound by instruction issue / dependent-latency stalls of the warp-per-env solver and by phase-barrier waits, not by HBM (DESIGN.md section 3.2)"},
           "reporter_stats": {k: stats[k] for k in ("steps", "episodes", "success_rate", "termination_rate", "mean_abs_err_lin_vel_x",
                                                    "mean_abs_err_ang_vel_yaw", "mean_contacts", "mean_solver_iters_per_step", "nan_resets", "ncon_dropped")}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, dt = cpu_reference(cfg, args.cpu_envs, args.cpu_steps, 2)
        out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                               "sample": f"{args.cpu_envs} envs x {args.cpu_steps} control steps of the same workload after 2 warm-up steps, {dt:.1f} s wall on {cores} threads "
                                         "(fp64 C++ restatement of the reference path, OpenMP over envs; not MuJoCo itself)"}
    if rank == 0:
        _RESULT_OUT.write(json.dumps(out) + "\n"); _RESULT_OUT.flush()
    env.close(); pol.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

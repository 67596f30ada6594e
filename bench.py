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

# The five BASELINE.json configs as bench workloads.  b_alg = algorithmic HBM bytes per env-step (fp32 SoA
# rows read + written once per control step); envs = environments per GPU of the config's full size.
WORKLOADS = {
    # headline (BASELINE.json metric): flamingo_p_v3 on rocky_hard, height map, full randomization, 65 536 envs per GPU
    "flamingo_rocky": dict(robot="flamingo_p_v3", terrain="rocky_hard", hm=True, random="full", envs=65536, b_alg=1832),
    # configs[0]: the reference's own CPU-runnable case (1 env in the reference; batched here)
    "light_flat": dict(robot="flamingo_light_v1", terrain="flat", hm=False, random="defaults", envs=65536, b_alg=1048),
    # configs[1]: correctness config, no randomization, 4096 envs
    "flamingo_rocky_norand": dict(robot="flamingo_p_v3", terrain="rocky_hard", hm=True, random="none", envs=4096, b_alg=1832),
    # configs[2]
    "w4_stairs": dict(robot="w4_p_v2", terrain="stairs_up_hard", hm=False, random="full", envs=65536, b_alg=2140),
    # configs[3]: position commands (command_dim 2), 1 M envs over 8 GPUs = 131 072 per GPU
    "humanoid_slope": dict(robot="humanoid_p_v0", terrain="slope_hard", hm=False, random="full", envs=131072, b_alg=3100, position=True),
}
ROBOT, TERRAIN = "flamingo_p_v3", "rocky_hard"
B_ALG = 1832      # algorithmic HBM bytes per env-step, flamingo_p_v3 with height map


def workload_config(name="flamingo_rocky", spawn_spread=0.0):
    from cosim_b200.config import make_config, RANDOM_FULL, RANDOM_NONE, RANDOM_DEFAULTS, load_tables
    w = WORKLOADS[name]
    et, _ = load_tables()
    kw = {}
    if w["hm"]:
        kw["non_stacked_obs_order"] = list(et[w["robot"]]["non_stacked_obs_order"]) + ["height_map"]
    if w.get("position"):
        kw["position_command"] = True; kw["command_dim"] = 2
    rnd = {"full": RANDOM_FULL, "none": RANDOM_NONE, "defaults": RANDOM_DEFAULTS}[w["random"]]
    eng = {"auto_reset": True, "seed": 0xC051}
    if spawn_spread > 0:
        eng["spawn_spread"] = float(spawn_spread)
    return make_config(w["robot"], w["terrain"], random=rnd, engine=eng, **kw)


def workload_text(name, N):
    w = WORKLOADS[name]
    return (f"{w['robot']} on {w['terrain']}" + (", height map in the observation" if w["hm"] else "") +
            {"full": ", full randomization (friction, mass noise, load, action delay, Kp/Kd)", "none": ", no randomization", "defaults": ", GUI-default randomization"}[w["random"]] +
            f", {N} envs per GPU, medium precision (4 sub-steps of 5 ms per control step), synthetic MLP policy state->512->256->128->nu, " +
            ("per-env position targets" if w.get("position") else "per-env velocity commands") + ", auto-reset on termination")


def _ncu_capture(envs, name="flamingo_rocky"):
    """Metrics of the committed `ncu --set full` capture of k_step for this workload (profiles/r02_traffic.json), per launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f).get(name)
        return t if t and int(t["envs"]) == int(envs) else None
    except Exception:
        return None


def measured_traffic(envs, name="flamingo_rocky"):
    t = _ncu_capture(envs, name)
    return int(t["traffic_bytes_per_launch"]) if t else None


def issue_metrics(envs, name="flamingo_rocky"):
    """What actually bounds k_step (same ncu capture): warp instructions per launch, issue-slot and pipe utilisation."""
    t = _ncu_capture(envs, name)
    if not t:
        return None
    keys = ("issue_slots_active_pct", "fp32_pipe_fma_pct", "alu_pipe_pct", "lsu_pipe_pct", "warps_active_pct_of_peak", "registers_per_thread",
            "local_load_inst", "local_store_inst", "dram_bytes_read", "dram_bytes_write", "l2_bytes")
    out = {k: t[k] for k in keys if k in t}
    out["warp_instructions_per_env_step"] = t["warp_instructions_per_launch"] / float(envs)
    out["env_warps_per_sm"] = t.get("block_size", 0) // 32
    out["source"] = "profiles/r02_traffic.json (ncu --set full, one k_step launch)"
    return out


def policy_metrics():
    """Tensor-pipe utilisation of k_policy_mlp from its committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get("k_policy_mlp")
    except Exception:
        return None


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


def cpu_reference(cfg, n_envs, min_seconds=1.0, warmup=2, threads=0, max_steps=400):
    """The CPU port of the path (oracle) on `n_envs` envs, all host threads, stepped for at least `min_seconds` (BASELINE.md
    section 3: >= 1 s of stepping per measurement).  Returns env-steps/s, cores, seconds, control steps."""
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

    def one():
        _, term, trunc = orc.step(act, cmd, nthreads=cores)
        done = term | trunc
        if done.any():
            orc.reset(mask=done, command=cmd)
    for _ in range(warmup):
        one()
    steps, t0 = 0, time.perf_counter()
    while True:
        one(); steps += 1
        dt = time.perf_counter() - t0
        if (dt >= min_seconds and steps >= 5) or steps >= max_steps:
            break
    return n_envs * steps / dt, cores, dt, steps


PY_OVERHEAD_NOTE = ("CPU restatement (not MuJoCo): the reference's own Python adds about 2.7 ms per env-step on top of mj_step "
                    "(scipy truncnorm noise 1.45 ms + height-map loop 1.2 ms, BASELINE.md section 2), i.e. the true reference loop is "
                    "bounded above at ~350 env-steps/s per core")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(args.config, args.spawn_spread)
    n = args.cpu_envs
    # `--steps K --warmup W` of the driver apply to the GPU arm's step; the CPU arm steps a bounded sample for >= 1 s per "step" budget
    value, cores, dt, steps = cpu_reference(cfg, n, min_seconds=max(1.0, args.cpu_seconds), warmup=2)
    sample = (f"{n} envs x {steps} control steps of the same workload, {dt:.1f} s wall on {cores} threads (fp64 C++ restatement, "
              f"OpenMP over envs); {value / cores:.0f} env-steps/s per core")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": steps,
           "warmup": 2, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_text(args.config, n) + " (bounded CPU sample of the same workload)", "name": args.config},
           "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "per_core": value / cores, "kind": "port", "sample": sample, "note": PY_OVERHEAD_NOTE},
           "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _RESULT_OUT.write(json.dumps(out) + "\n"); _RESULT_OUT.flush()


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="flamingo_rocky", choices=sorted(WORKLOADS), help="BASELINE.json workload (default: the headline)")
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (default: the config's full size)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: --envs (or the config's size) is the TOTAL over all GPUs")
    ap.add_argument("--impl", default="cosim_b200", choices=["cosim_b200", "reference"])
    ap.add_argument("--cpu-envs", type=int, default=1024, help="environments of the CPU arm (cpu_baseline leg and --impl reference use the same count)")
    ap.add_argument("--cpu-seconds", type=float, default=2.0, help="minimum wall time of CPU stepping (>= 1 s)")
    ap.add_argument("--steady-steps", type=int, default=200, help="length of the steady-state window (0 = skip)")
    ap.add_argument("--steady-seconds", type=float, default=40.0, help="wall-clock cap of the steady-state warm-up and of its window")
    ap.add_argument("--policy", default="mlp", choices=["mlp", "zero"], help="mlp: synthetic random MLP (headline); zero: PD hold of the reset pose (robots stay on the terrain)")
    ap.add_argument("--spawn-spread", type=float, default=0.0, help="variant: per-env spawn offsets of up to this many metres over the terrain (0 = the reference's spawn at the origin, the headline)")
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
    from cosim_b200.envs import BatchedEnv, shard_envs
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
    W = WORKLOADS[args.config]
    size = args.envs if args.envs > 0 else W["envs"]
    if args.scaling == "strong":
        N, offset = shard_envs(size, world, rank)          # total work fixed: contiguous shards of the same global env ids
        total_envs = size
    else:
        N, offset, total_envs = size, rank * size, world * size
    cfg = workload_config(args.config, args.spawn_spread)
    env = BatchedEnv(cfg, N, device=dev, seed=0xC051, env_offset=offset)     # RNG substream = global env id
    pol = MLPPolicy(synthetic_mlp(env.state_dim, env.action_dim), "elu", dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
    span = 5.0 if W.get("position") else 1.5
    cmd = (torch.rand((N, env.command_dim), device=dev, generator=gen) * 2.0 - 1.0) * span
    env.receive_user_command(cmd)
    state, _ = env.reset()
    zero_action = torch.zeros((N, env.action_dim), device=dev)

    def act(s):
        return pol.get_action(s) if args.policy == "mlp" else zero_action

    warm = max(args.warmup, 3)
    for _ in range(warm):
        state, _, _, _ = env.step(act(state))
    # ---------------- device-resident timing (value): the K steps the driver asks for, right after the warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local); clocks.start()
    l0 = env.launch_count + pol.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for i in range(args.steps):
        a = act(state)
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
    value = total_envs * args.steps / (elapsed_ms * 1e-3)
    window_stats = env.stats(all_reduce=True, clear=True)

    # ---------------- steady state: the same loop continued until the episodes are at their stationary age mix (the first
    # steps after reset() run with every robot still upright on the spawn patch), then a window of `steady_steps` steps.
    steady = None
    if args.steady_steps > 0:
        pre = max(0, 150 - warm - args.steps)            # with the synthetic policy an episode lasts ~135 control steps
        budget = args.steady_seconds                     # wall-clock cap of each of the two loops (slow workloads run fewer steps)
        tw, ran_pre = time.perf_counter(), 0
        for i in range(pre):
            state, _, _, _ = env.step(act(state)); ran_pre += 1
            if i % 8 == 7:
                torch.cuda.synchronize()
                if time.perf_counter() - tw > budget:
                    break
        env.stats(all_reduce=False, clear=True)
        # every rank runs the same number of steps: agree on it from the measured step time of the slowest rank
        torch.cuda.synchronize()
        per_step = torch.tensor([(time.perf_counter() - tw) / max(ran_pre, 1) if ran_pre else elapsed_ms * 1e-3 / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(per_step, op=dist.ReduceOp.MAX)
        nsteady = int(max(8, min(args.steady_steps, budget / max(float(per_step.item()), 1e-6))))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        se = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        se[0].record()
        for _ in range(nsteady):
            state, _, _, _ = env.step(act(state))
        se[1].record()
        torch.cuda.synchronize()
        t = torch.tensor([se[0].elapsed_time(se[1])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st = env.stats(all_reduce=True, clear=True)
        sms = float(t.item())
        # env-steps that ran physics (a finished env spends its next step in reset_env only: reported, and excluded here)
        physics_steps = st["steps"]
        steady = {"value": total_envs * nsteady / (sms * 1e-3), "unit": "env-steps/s", "steps": nsteady, "after_steps": warm + args.steps + ran_pre,
                  "ms_per_step": sms / nsteady, "physics_env_steps_per_s": physics_steps / (sms * 1e-3),
                  "reset_only_env_steps": total_envs * nsteady - physics_steps,
                  "reporter_stats": {k: st[k] for k in ("episodes", "success_rate", "termination_rate", "mean_contacts", "mean_solver_iters_per_step", "nan_resets", "ncon_dropped")}}

    # ---------------- end-to-end through the public API with HOST buffers.  Every step: H2D of that step's inputs (the user
    # commands, pinned), policy on the device-resident state, cosim_step, D2H of the results (state + done flags) into pinned
    # host memory; the D2H of step k overlaps the compute of step k + 1 (double-buffered device state, copy stream, one event
    # wait per step: BatchedEnv.step_policy_pipelined).  Every step's results are consumed on the host inside the timed region.
    nu, sd, cd = env.action_dim, env.state_dim, env.command_dim
    h_cmd = torch.empty((N, cd), dtype=torch.float32, pin_memory=True)
    h_cmd.copy_(cmd.cpu())
    assert h_cmd.is_pinned()
    # Same simulated interval as the device-timed leg: reset, the same warm-up steps, then the same K steps (the cost of
    # a step depends on what the robots are doing, so a leg timed later in the episodes would measure a different workload)
    e2e_steps = args.steps
    env.reset()

    class _HostPolicy:             # zero-action variant for the pipelined call
        def get_action(self, s):
            return act(s)
    hp = _HostPolicy()
    for _ in range(warm):
        env.step_policy_pipelined(hp, h_cmd)
    env.flush_pipelined()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    done_seen = 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = env.step_policy_pipelined(hp, h_cmd)
        if r is not None:
            done_seen += int(r[1].sum()) + int(r[2].sum())          # the host reads the previous step's flags
    r = env.flush_pipelined()
    done_seen += int(r[1].sum()) + int(r[2].sum())
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_envs * e2e_steps / (float(t.item()) * 1e-3)
    h2d = N * 4 * cd
    d2h = N * (4 * sd + 2)
    e2e_stats = env.stats(all_reduce=True)          # the one collective of this path: NCCL all-reduce of reporter statistics

    dropped = window_stats["ncon_dropped"] + e2e_stats["ncon_dropped"] + (steady["reporter_stats"]["ncon_dropped"] if steady else 0)
    assert dropped == 0, f"{dropped} contacts were dropped: the contact capacity of the model is too small for this workload"
    peak, peak_src = measured_peaks()
    b_alg = W["b_alg"]
    achieved = b_alg * N / (kernel_ms * 1e-3) / 1e9
    out = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
           "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
           "dtype": "f32 (physics), bf16 x bf16 -> f32 (policy MLP)", "data": "synthetic",
           "config": {"workload": workload_text(args.config, N) + (f", spawn offsets up to {args.spawn_spread:g} m over the terrain (variant, not the reference's spawn)" if args.spawn_spread > 0 else ""), "name": args.config, "envs_per_gpu": N, "total_envs": total_envs, "policy": args.policy, "spawn_spread_m": args.spawn_spread,
                      "sub_steps_per_s": value * 4,
                      "l2": "per-env state, parameter and observation arrays total > 126 MB L2 at 65 536 envs (inputs larger than L2); no explicit flush"},
           "clocks": clk, "gpu_launches": launches,
           "steady_state": steady,
           "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "done_flags_read_on_host": done_seen,
                   "path": "BatchedEnv.step_policy_pipelined: pinned host commands -> device, tcgen05 policy on the device state, cosim_step, state + done flags -> "
                           "pinned host on a copy stream (double-buffered, overlaps the next step), one event wait per step"},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(N, args.config),
                        "kernel": "k_step_pool" if getattr(env, "pooled", False) else "k_step", "kernel_ms": kernel_ms, "bytes_per_env_step": b_alg, "peak_source": peak_src,
                        "issue": issue_metrics(N, args.config), "policy_kernel": policy_metrics(),
                        "fp32_pipe_frac": (lambda t: None if not t or t.get("fp32_pipe_fma_pct") is None else t["fp32_pipe_fma_pct"] / 100.0)(issue_metrics(N, args.config)),
                        "tensor_pipe_frac": (lambda t: None if not t or t.get("tensor_pipe_active_pct") is None else t["tensor_pipe_active_pct"] / 100.0)(policy_metrics()),
                        "note": "the step is bound by instruction issue / dependent-latency stalls of the warp-per-env solver and by phase-barrier waits, not by HBM (DESIGN.md section 3.2)"},
           "reporter_stats": {k: window_stats[k] for k in ("steps", "episodes", "success_rate", "termination_rate", "mean_abs_err_lin_vel_x",
                                                           "mean_abs_err_ang_vel_yaw", "mean_contacts", "mean_solver_iters_per_step", "nan_resets", "ncon_dropped")}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, dt, steps = cpu_reference(cfg, args.cpu_envs, min_seconds=max(1.0, args.cpu_seconds), warmup=2)
        out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "per_core": v / cores, "kind": "port",
                               "sample": f"{args.cpu_envs} envs x {steps} control steps of the same workload after 2 warm-up steps, {dt:.1f} s wall on {cores} threads "
                                         "(fp64 C++ restatement of the reference path, OpenMP over envs; not MuJoCo itself)", "note": PY_OVERHEAD_NOTE}
    if rank == 0:
        _RESULT_OUT.write(json.dumps(out) + "\n"); _RESULT_OUT.flush()
    env.close(); pol.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Batched drop-in for the reference env stack CommandWrapper(TimeLimitWrapper(StateBuildWrapper(Robot))).

Reference interface mirrored here (file:line under /root/reference):
  build_env(config)                     envs/build.py:8-24
  reset() / step(action)                envs/wrappers.py:385-405 (-> 303-320 -> 245-269 -> <Robot>.reset_model / step)
  receive_user_command(cmd)             envs/wrappers.py:349-375
  event("push", v)                      envs/flamingo_p_v3/flamingo_p_v3.py:257-266
  get_data()                            envs/wrappers.py:410-411
  info keys                             envs/flamingo_p_v3/flamingo_p_v3.py:209-219, envs/wrappers.py:399-400

`BatchedEnv` steps N independent instances on one GPU through the C ABI of libcosim_b200.so
(include/cosim_b200.h).  All arrays are torch CUDA tensors with a leading env dimension.
`build_env(config)` with the default num_envs=1 returns `SingleEnv`, whose signatures, dtypes and
error behaviour are the reference's (numpy in, numpy out, python bools).

There is no CPU implementation: constructing an env without a CUDA device raises RuntimeError.
"""
import ctypes
import warnings
from collections.abc import Mapping

import numpy as np
import torch

from . import lib as _libmod
from .model import build_model

STAT_NAMES = ["steps", "episodes", "success", "terminated", "err_vx", "err_vy", "err_wz", "action_diff_rmse",
              "abs_torque", "sq_torque", "max_torque", "ncon", "solver_iters", "ncon_dropped", "nan_resets", "_pad"]
_ST_MAX = STAT_NAMES.index("max_torque")


class Data:
    """Stand-in for mjData as used through get_data() (wrappers.py:360-365: qpos, qvel)."""

    def __init__(self, env):
        self._env = env

    @property
    def qpos(self):
        return self._env.get("qpos")

    @property
    def qvel(self):
        return self._env.get("qvel")


class Info(Mapping):
    """The per-step `info` dict of the reference, fetched lazily from device fields (one row per env)."""

    def __init__(self, env, user_command):
        self._env = env
        self._keys = ["dt", "action", "action_diff_RMSE", "torque", "lin_vel_x", "lin_vel_y", "ang_vel_yaw",
                      "set_points", "state"] + [f"user_command_{i}" for i in range(env.command_dim)]
        self._uc = user_command
        self._cache = {}

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)

    def __getitem__(self, k):
        if k in self._cache:
            return self._cache[k]
        e = self._env
        if k == "dt":
            v = e.dt
        elif k == "action":
            v = e.get("last_action")
        elif k == "torque":
            v = e.get("torque")
        elif k == "set_points":                      # action * action_scaler (flamingo_p_v3.py:217)
            v = e.get("last_action") * e._action_scaler
        elif k in ("action_diff_RMSE", "lin_vel_x", "lin_vel_y", "ang_vel_yaw"):
            v = e.get("info")[:, ["action_diff_RMSE", "lin_vel_x", "lin_vel_y", "ang_vel_yaw"].index(k)]
        elif k == "state":                           # actuated joint positions, wheels as velocities (:218)
            qpos, qvel = e.get("qpos"), e.get("qvel")
            v = torch.cat([qpos[:, e._state_pos_qadr], qvel[:, e._state_vel_dadr]], dim=1)
        elif k.startswith("user_command_"):
            v = self._uc[:, int(k.rsplit("_", 1)[1])]
        else:
            raise KeyError(k)
        self._cache[k] = v
        return v


class BatchedEnv:
    def __init__(self, config, num_envs=1, device="cuda:0", seed=None, env_offset=0, debug=False):
        if not torch.cuda.is_available():
            raise RuntimeError("cosim_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.config = config
        self.model = build_model(config)
        m = self.model
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        self.id = config["env"]["id"]
        self.action_dim = m.dim("nu")
        self.state_dim = m.dim("state_dim")
        self.command_dim = m.dim("command_dim")
        self.cmd_slices = m.meta["cmd_slices"]
        self.control_freq = m.meta["control_freq"]
        self.obs_to_dim = m.meta["obs_to_dim"]
        self.dt = m.meta["dt"]
        self.max_sim_step = m.dim("max_episode_steps")
        eng = config.get("engine", {}) or {}
        self.seed = int(eng.get("seed", 0) if seed is None else seed)
        self._L = _libmod.lib()
        self._h = ctypes.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(idx):
            rc = self._L.cosim_create(m.blob, len(m.blob), self.num_envs, idx, self.seed, int(env_offset), ctypes.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"cosim_create failed with code {rc}")
        N, dev = self.num_envs, self.device
        self._state = torch.zeros((N, self.state_dim), dtype=torch.float32, device=dev)
        self._term = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(N, dtype=torch.uint8, device=dev)
        cd = max(1, self.command_dim)
        self.user_command = torch.zeros((N, cd), dtype=torch.float32, device=dev)
        self.applied_command = torch.zeros((N, cd), dtype=torch.float32, device=dev)
        self._scales = torch.tensor([float(config["observation"]["command_scales"][str(i)]) for i in range(self.command_dim)] or [1.0],
                                    dtype=torch.float32, device=dev)
        self._action_scaler = torch.tensor(m.meta["action_scaler"], dtype=torch.float32, device=dev)
        self._state_pos_qadr = torch.tensor(m.sections["state_pos_qadr"], dtype=torch.long, device=dev)
        self._state_vel_dadr = torch.tensor(m.sections["state_vel_dadr"], dtype=torch.long, device=dev)
        self._stats = torch.zeros(16, dtype=torch.float64, device=dev)
        self.reset_flag = False
        if debug:
            self.set_debug(True)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._L.cosim_last_error(self._h).decode()}")

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _as(self, x, cols, dtype=torch.float32):
        t = torch.as_tensor(x, dtype=dtype, device=self.device)
        if t.dim() == 1:
            t = t.unsqueeze(0).expand(self.num_envs, -1)
        if tuple(t.shape) != (self.num_envs, cols):
            raise ValueError(f"expected shape ({self.num_envs}, {cols}), got {tuple(t.shape)}")
        return t.contiguous()

    def set_debug(self, enable=True):
        self._check(self._L.cosim_set_debug(self._h, int(bool(enable))), "cosim_set_debug")

    def get(self, field):
        """Copy of a per-env device field ([N, dim]; see include/cosim_b200.h for the names)."""
        d = self._L.cosim_field_dim(self._h, field.encode())
        if d < 0:
            raise KeyError(f"{field}: unknown field, or a debug field while set_debug(False)")
        is_int = self._L.cosim_field_is_int(self._h, field.encode())
        out = torch.empty((self.num_envs, d), dtype=torch.int32 if is_int else torch.float32, device=self.device)
        self._check(self._L.cosim_get(self._h, field.encode(), self._ptr(out), self._stream()), "cosim_get")
        return out

    def set(self, field, value):
        d = self._L.cosim_field_dim(self._h, field.encode())
        if d < 0:
            raise KeyError(field)
        v = self._as(value, d)
        self._check(self._L.cosim_set(self._h, field.encode(), self._ptr(v), self._stream()), "cosim_set")

    def rng_probe(self, stream, step, nidx):
        out = torch.empty((self.num_envs, nidx), dtype=torch.int32, device=self.device)
        self._check(self._L.cosim_rng_probe(self._h, int(stream), int(step), int(nidx), self._ptr(out), self._stream()), "cosim_rng_probe")
        return out

    @property
    def launch_count(self):
        return self._L.cosim_launch_count(self._h)

    # ------------------------------------------------------------------ reference API
    def receive_user_command(self, user_command):
        """CommandWrapper.receive_user_command (wrappers.py:349-375), one row per env."""
        if self.command_dim == 0:
            return
        uc = torch.as_tensor(user_command, dtype=torch.float32, device=self.device)
        if uc.dim() == 1:
            uc = uc.unsqueeze(0).expand(self.num_envs, -1)
        uc = uc[:, :self.command_dim].contiguous()
        self.user_command = uc
        if self.config["env"]["position_command"] is False:
            self.applied_command = (uc * self._scales).contiguous()
        else:
            assert self.command_dim == 2, f"Currently, position command only support 2 dimenstion, but got {self.command_dim}."
            warnings.warn("For position commands, 'command_scales' is always treated as 1.0.")
            qpos = self.get("qpos").double()
            dx, dy = uc[:, 0].double() - qpos[:, 0], uc[:, 1].double() - qpos[:, 1]
            w, x, y, z = qpos[:, 3], qpos[:, 4], qpos[:, 5], qpos[:, 6]
            yaw = torch.atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
            cosy, siny = torch.cos(-yaw), torch.sin(-yaw)
            self.applied_command = torch.stack([cosy * dx - siny * dy, siny * dx + cosy * dy], dim=1).float().contiguous()

    def reset(self, mask=None):
        """-> (state [N, state_dim] float32, info).  mask: optional bool[N], reset only those envs."""
        mk = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        cmd = self.applied_command if self.command_dim > 0 else None
        self._check(self._L.cosim_reset(self._h, self._ptr(mk), self._ptr(cmd), self._ptr(self._state), self._stream()), "cosim_reset")
        self.reset_flag = True
        return self._state, {}

    def step(self, action):
        """-> (next_state [N, state_dim], terminated [N] bool, truncated [N] bool, info)."""
        assert self.reset_flag is True, "Call 'reset()' before calling 'step()'."
        if self.command_dim < 0 or self.command_dim > 6:
            raise ValueError(f"Invalid 'command_dim': expected 0> or <7; but got {self.command_dim}.")
        a = self._as(action, self.action_dim)
        cmd = self.applied_command if self.command_dim > 0 else None
        uc = self.user_command if self.command_dim > 0 else None
        self._check(self._L.cosim_step(self._h, self._ptr(a), self._ptr(cmd), self._ptr(uc), self._ptr(self._state),
                                       self._ptr(self._term), self._ptr(self._trunc), self._stream()), "cosim_step")
        return self._state, self._term.bool(), self._trunc.bool(), Info(self, self.user_command)

    def step_host(self, action, command, state_out, terminated_out, truncated_out):
        """cosim_step_host: numpy (pinned or pageable) in / out, H2D + step + D2H + sync inside the call."""
        assert self.reset_flag is True, "Call 'reset()' before calling 'step()'."
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        self._check(self._L.cosim_step_host(self._h, p(action), p(command), p(state_out), p(terminated_out), p(truncated_out)), "cosim_step_host")

    def step_policy_host(self, policy, user_command_host, state_out_host, terminated_out_host, truncated_out_host):
        """One control step of the whole loop body of core/tester.py:66-97 with HOST I/O: the user commands of this step come
        from (pinned) host memory, the policy runs on the device-resident state of the previous step, the new state and
        the done flags are copied back to (pinned) host memory, and the call returns after a stream sync."""
        assert self.reset_flag is True, "Call 'reset()' before calling 'step()'."
        if self.command_dim > 0:
            self._uc_stage = getattr(self, "_uc_stage", None)
            if self._uc_stage is None:
                self._uc_stage = torch.empty((self.num_envs, self.command_dim), dtype=torch.float32, device=self.device)
            self._uc_stage.copy_(_as_cpu_tensor(user_command_host), non_blocking=True)          # H2D: this step's inputs
            self.receive_user_command(self._uc_stage)
        state, term, trunc, _ = self.step(policy.get_action(self._state))
        _as_cpu_tensor(state_out_host).copy_(state, non_blocking=True)                        # D2H: this step's results
        _as_cpu_tensor(terminated_out_host).copy_(self._term, non_blocking=True)
        _as_cpu_tensor(truncated_out_host).copy_(self._trunc, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    def step_policy_pipelined(self, policy, user_command_host):
        """The same loop body with the host transfers OVERLAPPED with the next step: the device state is double-buffered, the
        device->host copy of step k runs on a copy stream while the policy and k_step of step k + 1 run on the compute stream
        (the policy only needs the device copy of the state), and the host waits for one event per step instead of
        draining the stream.  Returns the results of the PREVIOUS call as pinned host tensors (state, terminated,
        truncated) -- valid until the call after next -- or None on the first call; `flush_pipelined()` completes and
        returns the last step.  One cudaMemcpyAsync per buffer, as in the synchronous call."""
        assert self.reset_flag is True, "Call 'reset()' before calling 'step()'."
        dev, N = self.device, self.num_envs
        P = getattr(self, "_pipe", None)
        if P is None:
            pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
            P = self._pipe = {
                "copy": torch.cuda.Stream(device=dev), "k": 0, "pending": None,
                "d_state": [self._state, torch.zeros_like(self._state)], "d_term": [self._term, torch.zeros_like(self._term)],
                "d_trunc": [self._trunc, torch.zeros_like(self._trunc)],
                "h_state": [pin((N, self.state_dim), torch.float32) for _ in range(2)], "h_term": [pin((N,), torch.uint8) for _ in range(2)],
                "h_trunc": [pin((N,), torch.uint8) for _ in range(2)],
                "computed": [torch.cuda.Event() for _ in range(2)], "copied": [torch.cuda.Event() for _ in range(2)],
                "uc": torch.empty((N, max(1, self.command_dim)), dtype=torch.float32, device=dev), "cur": 0}
        cs = torch.cuda.current_stream(dev)
        cur, nxt = P["cur"], 1 - P["cur"]
        if self.command_dim > 0:
            P["uc"].copy_(_as_cpu_tensor(user_command_host), non_blocking=True)                 # H2D: this step's inputs
            self.receive_user_command(P["uc"])
        action = policy.get_action(P["d_state"][cur])                                           # reads the state of the previous step
        if P["k"] >= 2:
            cs.wait_event(P["copied"][nxt])           # the buffer about to be overwritten has reached the host
        self._state, self._term, self._trunc = P["d_state"][nxt], P["d_term"][nxt], P["d_trunc"][nxt]
        self.step(action)
        P["computed"][nxt].record(cs)
        with torch.cuda.stream(P["copy"]):                                                      # D2H: this step's results, off the compute stream
            P["copy"].wait_event(P["computed"][nxt])
            P["h_state"][nxt].copy_(P["d_state"][nxt], non_blocking=True)
            P["h_term"][nxt].copy_(P["d_term"][nxt], non_blocking=True)
            P["h_trunc"][nxt].copy_(P["d_trunc"][nxt], non_blocking=True)
            P["copied"][nxt].record(P["copy"])
        out = None
        if P["pending"] is not None:
            P["copied"][P["pending"]].synchronize()                                              # one event wait per step
            out = (P["h_state"][P["pending"]], P["h_term"][P["pending"]], P["h_trunc"][P["pending"]])
        P["pending"], P["cur"], P["k"] = nxt, nxt, P["k"] + 1
        return out

    def flush_pipelined(self):
        """Completes the last pipelined step and returns its host results (or None)."""
        P = getattr(self, "_pipe", None)
        if P is None or P["pending"] is None:
            return None
        P["copied"][P["pending"]].synchronize()
        out = (P["h_state"][P["pending"]], P["h_term"][P["pending"]], P["h_trunc"][P["pending"]])
        P["pending"] = None
        return out

    def substep(self):
        """One raw physics sub-step with ctrl = last applied torque (parity aid, see cosim_substep)."""
        self._check(self._L.cosim_substep(self._h, self._stream()), "cosim_substep")

    def event(self, event, value, mask=None):
        if event == "push":
            v = self._as(value, 3)
            mk = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            self._check(self._L.cosim_push(self._h, self._ptr(mk), self._ptr(v), self._stream()), "cosim_push")
        else:
            raise NotImplementedError(f"event:{event} is not supported.")

    def get_data(self):
        return Data(self)

    def render(self):
        pass

    @property
    def pooled(self):
        """True when the model is stepped by the pooled kernel (k_step_pool: stage queues over a pool of envs per CTA)."""
        return self._L.cosim_pool_size(self._h) > 0

    @property
    def general_path(self):
        """True when the model's friction-cone / solver options select the general constraint path (engine_general.h)."""
        return self._L.cosim_general_path(self._h) > 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cosim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reporter statistics
    def stats(self, all_reduce=True, clear=False):
        """Episode statistics summed over envs (and over ranks when torch.distributed is initialised).

        Reporter semantics per SURVEY.md C-17 over the reference's info keys (core/reporter.py:210-218,
        429-581): success = truncated without termination; tracking error = mean |lin_vel_x - cmd0| etc."""
        self._check(self._L.cosim_stats_reduce(self._h, self._ptr(self._stats), self._stream()), "cosim_stats_reduce")
        s = self._stats.clone()
        if all_reduce:
            s = all_reduce_stats(s)
        if clear:
            self._check(self._L.cosim_stats_clear(self._h, self._stream()), "cosim_stats_clear")
        return derive_stats(s.cpu().numpy(), self.action_dim)


def _as_cpu_tensor(x):
    """numpy array or CPU tensor (ideally pinned) -> CPU tensor sharing the memory."""
    return x if torch.is_tensor(x) else torch.from_numpy(x)


def all_reduce_stats(s):
    """Sum the raw statistics vector over ranks (max for the max-torque slot).  The only collective of the
    path: NCCL on GPU tensors, gloo on CPU tensors (tests); no-op without a process group."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        mx = s[_ST_MAX].clone()
        torch.distributed.all_reduce(s, op=torch.distributed.ReduceOp.SUM)
        torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
        s[_ST_MAX] = mx
    return s


def shard_envs(total_envs, world_size, rank):
    """Contiguous env ranges per rank: (num_envs, env_offset).  Global env id = offset + local id = RNG substream."""
    base, rem = divmod(int(total_envs), int(world_size))
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return n, off


def derive_stats(raw, action_dim):
    d = {n: float(v) for n, v in zip(STAT_NAMES, raw) if not n.startswith("_")}
    steps, eps = max(d["steps"], 1.0), max(d["episodes"], 1.0)
    d["success_rate"] = d["success"] / eps
    d["termination_rate"] = d["terminated"] / eps
    d["mean_abs_err_lin_vel_x"] = d["err_vx"] / steps
    d["mean_abs_err_lin_vel_y"] = d["err_vy"] / steps
    d["mean_abs_err_ang_vel_yaw"] = d["err_wz"] / steps
    d["mean_action_diff_rmse"] = d["action_diff_rmse"] / steps
    d["mean_abs_torque"] = d["abs_torque"] / (steps * action_dim)
    d["rms_torque"] = (d["sq_torque"] / (steps * action_dim)) ** 0.5
    d["mean_contacts"] = d["ncon"] / steps
    d["mean_solver_iters_per_step"] = d["solver_iters"] / steps
    return d


class SingleEnv:
    """num_envs = 1 adapter with the reference's exact signatures (numpy arrays, python bools, dict info)."""

    def __init__(self, benv):
        self.env = benv
        self.config = benv.config
        self.id, self.action_dim, self.state_dim = benv.id, benv.action_dim, benv.state_dim
        self.command_dim, self.cmd_slices = benv.command_dim, benv.cmd_slices
        self.control_freq, self.obs_to_dim = benv.control_freq, benv.obs_to_dim
        self.user_command = np.zeros(self.config["observation"]["command_dim"])
        self.reset_flag = False

    def receive_user_command(self, user_command):
        self.user_command = user_command[:self.command_dim]
        self.env.receive_user_command(np.asarray(user_command, dtype=np.float64)[None, :])

    def reset(self):
        self.reset_flag = True
        state, info = self.env.reset()
        return state[0].cpu().numpy(), info

    def step(self, action):
        assert self.reset_flag is True, "Call 'reset()' before calling 'step()'."
        if torch.is_tensor(action):                 # the engine's policies return device tensors; the reference's return numpy
            action = action.detach().reshape(1, -1)
        else:
            action = np.asarray(action, dtype=np.float32)[None, :]
        state, term, trunc, info = self.env.step(action)
        out = {}
        for k in info:
            v = info[k]
            if k.startswith("user_command_"):
                out[k] = self.user_command[int(k.rsplit("_", 1)[1])]
            elif torch.is_tensor(v):
                v = v[0].cpu().numpy()
                out[k] = v if v.ndim else v.item()
            else:
                out[k] = v
        terminated, truncated = bool(term[0].item()), bool(trunc[0].item())
        if terminated or truncated:
            self.reset_flag = False
            self.env.reset_flag = False
        return state[0].cpu().numpy(), terminated, truncated, out

    def event(self, event, value):
        return self.env.event(event, np.asarray(value, dtype=np.float32)[None, :])

    def get_data(self):
        d = self.env.get_data()

        class _D:
            qpos = d.qpos[0].double().cpu().numpy()
            qvel = d.qvel[0].double().cpu().numpy()
        return _D

    def render(self):
        self.env.render()

    def close(self):
        self.env.close()


def build_env(config, num_envs=None, device="cuda:0", seed=None, env_offset=0, debug=False):
    """envs/build.py:8-24.  num_envs=None -> the reference's single-env API; an int -> BatchedEnv."""
    if num_envs is None:
        return SingleEnv(BatchedEnv(config, 1, device, seed, env_offset, debug))
    return BatchedEnv(config, num_envs, device, seed, env_offset, debug)

"""Evaluation driver with the reference's interface (`core/tester.py:11-121`) on top of the batched engine.

The public methods carry the reference's names so that its UI code can drive either implementation: `load_config`,
`load_policy`, `init_user_command`, `receive_user_command`, `update_command`, `activate_push_event`,
`deactivate_push_event`, `test`, `stop`, `close`.  Qt is not involved: the `stepFinished` / `finished` signals are plain
callbacks (`on_step`, `on_finished`).

Two modes:
* `Tester()` -- one environment, numpy state, python bools, every step's `info` goes to the report (what the reference does);
* `Tester(num_envs=N)` -- N domain-randomised environments advance together until each has completed an episode; environment
  `trace_env` supplies the time series of the report and `BatchedEnv.stats()` (all-reduced over ranks when
  torch.distributed is initialised) its population page.
"""
import os

import numpy as np

from .envs import build_env
from .policy import build_policy
from .reporter import Reporter

_POLICY_ERROR = ("Failed to run inference with the selected ONNX policy: {path}.\n\nThe state has {n} entries; the policy may expect a "
                 "different input length.\n")


class Tester:
    def __init__(self, num_envs=None, device="cuda:0", seed=None, trace_env=0, on_step=None, on_finished=None):
        self.num_envs, self.device, self.seed, self.trace_env = num_envs, device, seed, trace_env
        self.on_step, self.on_finished = on_step, on_finished
        self.config, self.policy_path = None, None
        self.env = self.policy = self.reporter = None
        self.user_command = None          # [command_dim], or [num_envs, command_dim] for per-environment commands
        self._push_vel = None             # world-frame velocity kick applied before every step while set
        self._stop = self._had_error = False

    # ---- configuration ------------------------------------------------------------------------------------------------
    def load_config(self, config):
        self.config = config

    def load_policy(self, policy_path):
        self.policy_path = policy_path

    def _command_dim(self):
        return int(self.config["observation"]["command_dim"])

    def init_user_command(self):
        self.user_command = np.zeros(self._command_dim())

    def update_command(self, index, value):
        """Set command slot `index` (a UI slider in the reference); out-of-range slots are ignored as there."""
        if self.user_command is None:
            self.init_user_command()
        if 0 <= index < self._command_dim():
            self.user_command[..., index] = value

    def receive_user_command(self):
        """Hand the current command to the environment (called once per step by `test`)."""
        if self.user_command is None:
            self.init_user_command()
        self.env.receive_user_command(self.user_command)

    def activate_push_event(self, push_vel):
        self._push_vel = push_vel

    def deactivate_push_event(self):
        self._push_vel = None

    # ---- the loop -----------------------------------------------------------------------------------------------------
    def _act(self, state):
        try:
            return self.policy.get_action(state)
        except Exception as err:
            self._had_error = True
            self.close()
            raise RuntimeError(_POLICY_ERROR.format(path=self.policy_path, n=state.shape[-1])) from err

    def test(self, report_path=None):
        if report_path is None:
            report_path = os.path.join(os.path.dirname(self.policy_path) if self.policy_path else os.getcwd(), "report.pdf")
        self.reporter = Reporter(report_path=report_path, config=self.config)
        self.env = build_env(self.config, num_envs=self.num_envs, device=self.device, seed=self.seed)
        self.policy = build_policy(self.config, self.policy_path, state_dim=self.env.state_dim, action_dim=self.env.action_dim,
                                   device=self.device)
        batched = self.num_envs is not None
        finished = None                                   # batched mode: which environments have completed their episode
        state, _ = self.env.reset()
        while not self._stop:
            self.receive_user_command()
            action = self._act(state)
            if self._push_vel is not None:
                self.env.event(event="push", value=self._push_vel)
            self.env.render()
            state, terminated, truncated, info = self.env.step(action)
            if batched:
                ended = terminated | truncated
                if hasattr(self.policy, "reset") and bool(ended.any()):
                    self.policy.reset(mask=ended)         # recurrent state must not leak into the auto-reset episode (reference: a fresh policy per run)
                traced_running = finished is None or not bool(finished[self.trace_env])
                finished = ended.clone() if finished is None else finished | ended
                if traced_running:
                    self.reporter.write_info(info, env_index=self.trace_env)
                over = bool(finished.all())
            else:
                self.reporter.write_info(info)
                over = bool(terminated or truncated)
            if self.on_step is not None:
                self.on_step()
            if over:
                break
        if not self._had_error:
            if batched:
                self.reporter.write_population(self.env.stats())
            self.reporter.generate_report()
        self.close()
        if self.on_finished is not None:
            self.on_finished()
        return report_path

    def stop(self):
        self._stop = True

    def close(self):
        env, self.env = self.env, None
        if env is not None:
            try:
                env.close()
            except Exception:
                pass

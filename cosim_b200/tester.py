"""The reference's evaluation driver on the batched engine: `core/tester.py:11-121` without the Qt plumbing.

Same methods and call order as the reference `Tester` (`load_config`, `load_policy`, `init_user_command`,
`receive_user_command`, `update_command`, `activate_push_event`, `deactivate_push_event`, `test`, `stop`, `close`); the Qt
signals `stepFinished` / `finished` become optional callbacks.  `num_envs=None` runs the reference's single-environment loop
unchanged (numpy state, python bools, report from every step's `info`).  With `num_envs=N` the same loop advances N
domain-randomised environments at once: it runs until every environment has finished an episode, traces environment
`trace_env` into the report and adds the population statistics (all-reduced over ranks when torch.distributed is up).
"""
import os

import numpy as np

from .envs import build_env
from .policy import build_policy
from .reporter import Reporter


class Tester:
    def __init__(self, num_envs=None, device="cuda:0", seed=None, trace_env=0, on_step=None, on_finished=None):
        self.user_command = None
        self._push_event = False
        self._stop = False
        self._had_error = False
        self.num_envs, self.device, self.seed, self.trace_env = num_envs, device, seed, trace_env
        self.on_step, self.on_finished = on_step, on_finished
        self.env = self.policy = self.reporter = None
        self.policy_path = None

    def load_config(self, config):
        self.config = config

    def load_policy(self, policy_path):
        self.policy_path = policy_path

    def init_user_command(self):
        """Initialize the user command array before starting the test."""
        self.user_command = np.zeros(self.config["observation"]["command_dim"])

    def receive_user_command(self):
        """Send the current user command value to the environment."""
        if self.user_command is None:
            self.init_user_command()
        self.env.receive_user_command(self.user_command)

    def update_command(self, index, value):
        """Update one slot of the command (the UI sliders of the reference); a [num_envs, command_dim] array may be assigned to
        `user_command` directly for per-environment commands."""
        if self.user_command is None:
            self.init_user_command()
        if index < self.config["observation"]["command_dim"]:
            self.user_command[..., index] = value

    def activate_push_event(self, push_vel):
        self._push_event = True
        self._push_vel = push_vel

    def deactivate_push_event(self):
        self._push_event = False

    def test(self, report_path=None):
        if report_path is None:
            base = os.path.dirname(self.policy_path) if self.policy_path else os.getcwd()
            report_path = os.path.join(base, "report.pdf")
        self.reporter = Reporter(report_path=report_path, config=self.config)
        self.env = build_env(self.config, num_envs=self.num_envs, device=self.device, seed=self.seed)
        self.policy = build_policy(self.config, self.policy_path, state_dim=self.env.state_dim, action_dim=self.env.action_dim,
                                   device=self.device)
        batched = self.num_envs is not None
        state, info = self.env.reset()
        done = False
        finished = None
        while not done and not self._stop:
            self.receive_user_command()
            try:
                action = self.policy.get_action(state)
            except Exception as e:
                self.close()
                self._had_error = True
                raise RuntimeError(f"Failed to run inference with the selected ONNX policy: {self.policy_path}."
                                   f"\n\nThe current state length (={state.shape[-1]}) may not match the input length expected by the "
                                   "ONNX policy, which could have caused this error.\n") from e
            if self._push_event:
                self.env.event(event="push", value=self._push_vel)
            self.env.render()
            assert self.user_command is not None, "user_command must not be None."
            next_state, terminated, truncated, info = self.env.step(action)
            if batched:
                ended = (terminated | truncated)
                finished = ended.clone() if finished is None else (finished | ended)
                if not bool(finished[self.trace_env]) or bool(ended[self.trace_env]):
                    self.reporter.write_info(info, env_index=self.trace_env)      # the traced environment's own episode
                done = bool(finished.all())
            else:
                self.reporter.write_info(info)
                done = terminated or truncated
            if self.on_step is not None:
                self.on_step()
            state = next_state
        if not self._had_error:
            if batched:
                self.reporter.write_population(self.env.stats())
            self.reporter.generate_report()
        self.close()
        if self.on_finished is not None:
            self.on_finished()
        return report_path

    def stop(self):
        """Stop the test loop."""
        self._stop = True

    def close(self):
        """Attempt to close the environment."""
        if self.env is not None:
            try:
                self.env.close()
            except Exception:
                pass

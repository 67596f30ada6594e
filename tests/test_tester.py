"""The reference's evaluation driver (core/tester.py) on the engine: an ONNX policy file, the env, the reporter and the
PDF, for the single-env API and for a batched run."""
import os

import numpy as np
import pytest

from cosim_b200.config import make_config, RANDOM_DEFAULTS
from tests.test_onnx_reader import _model, _node, _tensor

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _write_policy(path, state_dim, action_dim, seed=0):
    rng = np.random.default_rng(seed)
    dims = [state_dim, 64, 32, action_dim]
    nodes, inits = [], []
    x = "x"
    for i in range(3):
        w = (rng.standard_normal((dims[i + 1], dims[i])) * 0.1).astype(np.float32)
        b = np.zeros(dims[i + 1], np.float32)
        nodes.append(_node("Gemm", [x, f"w{i}", f"b{i}"], [f"h{i}"], {"transB": 1}))
        inits += [_tensor(f"w{i}", w), _tensor(f"b{i}", b)]
        x = f"h{i}"
        if i < 2:
            nodes.append(_node("Elu", [x], [f"a{i}"])); x = f"a{i}"
    path.write_bytes(_model(nodes, inits))


def test_tester_single_env_and_batched(tmp_path):
    from cosim_b200.tester import Tester
    from cosim_b200.envs import build_env
    cfg = make_config("flamingo_p_v3", "rocky_hard", random=RANDOM_DEFAULTS, max_duration=0.4)       # 20 control steps
    probe = build_env(cfg, num_envs=1)
    sd, ad = probe.state_dim, probe.action_dim
    probe.close()
    pol = tmp_path / "policy.onnx"
    _write_policy(pol, sd, ad)
    # --- the reference's single-environment loop
    steps = []
    t = Tester(on_step=lambda: steps.append(1))
    t.load_config(cfg); t.load_policy(str(pol)); t.init_user_command(); t.update_command(0, 0.5)
    t.activate_push_event([0.3, 0.0, 0.0]); t.deactivate_push_event()
    report = t.test()
    assert report == os.path.join(str(tmp_path), "report.pdf") and os.path.getsize(report) > 2000
    assert 1 <= t.reporter.timesteps == len(steps) <= 20
    assert set(("dt", "torque", "set_points", "state", "user_command_0")) <= set(t.reporter.history)
    assert np.asarray(t.reporter.history["torque"][0]).shape == (ad,)
    # --- batched: runs until every env has finished its episode, traces env 3, adds population statistics
    tb = Tester(num_envs=96, seed=5, trace_env=3)
    tb.load_config(cfg); tb.load_policy(str(pol))
    tb.user_command = np.zeros((96, cfg["observation"]["command_dim"]), np.float32)        # per-env commands
    tb.user_command[:, 0] = np.linspace(-1.0, 1.0, 96)
    out = tb.test(report_path=str(tmp_path / "batched.pdf"))
    assert os.path.getsize(out) > 2000 and tb.reporter.population["episodes"] == 96
    assert 1 <= tb.reporter.timesteps <= 20 and np.asarray(tb.reporter.history["torque"][0]).shape == (ad,)
    # a state-length mismatch surfaces as the reference's error
    bad = tmp_path / "bad.onnx"
    _write_policy(bad, sd + 3, ad)
    t2 = Tester(); t2.load_config(cfg); t2.load_policy(str(bad)); t2.init_user_command()
    with pytest.raises(RuntimeError, match="Failed to run inference"):
        t2.test(report_path=str(tmp_path / "x.pdf"))

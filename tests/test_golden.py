"""Golden vectors recorded by running the reference's OWN Python env stack (build_env -> <Robot> env, managers,
wrappers; unmodified sources under /root/reference) on top of the oracle's physics (tools/gen_golden.py, shims for
the uninstallable mujoco / gymnasium packages).  Replaying the same actions and commands through the oracle's
restated env layer must reproduce the recorded states, torques, flags and info values: this pins PD control, gear
ratios, clips, observation assembly, frequency gating, stacking, command slots and lag, truncation and termination
against the reference code itself.  (MuJoCo's own numerics stay unpinned -- DESIGN.md section 6.)"""
import glob
import json
import os

import numpy as np
import pytest

from cosim_b200.model import build_model
from cosim_b200.robots import ROBOTS
from oracle.oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    z = np.load(path, allow_pickle=False)
    cfg = json.loads(str(z["config_json"]))
    cfg["random"]["sensor_noise"] = "zero"          # the reference's `none` level adds <= 1e-8 noise (quirk C-13)
    return z, cfg


def push_oracle(o, vel):
    """The reference's push event (<robot>.py event(): robot-frame xy, world-frame z -- quirk C-12) applied to the oracle state."""
    q, v = o.get("qpos"), o.get("qvel")
    w, x, y, z = q[0, 3:7]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    local = R.T @ np.asarray(vel, dtype=np.float64)
    v[0, 0:2] = local[0:2]; v[0, 2] = vel[2]
    o.set("qvel", v)


def test_fixtures_present():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_env_layer_reproduces_reference_python(path):
    z, cfg = _load(path)
    m = build_model(cfg)
    assert m.dim("state_dim") == int(z["state_dim"])
    assert [[s.start, s.stop] for s in m.meta["cmd_slices"]] == z["cmd_slices"].tolist()
    o = Oracle(m, 1, seed=0)
    s0 = o.reset()
    np.testing.assert_allclose(s0[0], z["reset_state"], atol=1e-6)
    n = len(z["states"])
    pushes = dict(zip(z["push_steps"].tolist(), z["push_vels"])) if "push_steps" in z.files else {}
    for k in range(n):
        if k in pushes:
            push_oracle(o, pushes[k])
        s, term, trunc = o.step(z["actions"][k][None, :], z["applied"][k][None, :])
        np.testing.assert_allclose(s[0], z["states"][k], atol=3e-6, err_msg=f"state at step {k}")
        np.testing.assert_allclose(o.get("torque")[0], z["torque"][k], atol=1e-9, rtol=1e-12, err_msg=f"torque at step {k}")
        assert bool(term[0]) == bool(z["terminated"][k]) and bool(trunc[0]) == bool(z["truncated"][k])
        info = o.get("info")[0]
        np.testing.assert_allclose(info[0], z["action_diff_RMSE"][k], atol=1e-12)
        np.testing.assert_allclose(info[1], z["lin_vel_x"][k], atol=1e-6)
        np.testing.assert_allclose(info[3], z["ang_vel_yaw"][k], atol=1e-9)
        np.testing.assert_allclose(o.get("qpos")[0], z["qpos"][k], atol=1e-12)
        np.testing.assert_allclose(z["actions"][k] * m.meta["action_scaler"], z["set_points"][k], atol=1e-12)
        st = np.concatenate([o.get("qpos")[0][m.sections["state_pos_qadr"]], o.get("qvel")[0][m.sections["state_vel_dadr"]]])
        np.testing.assert_allclose(st, z["info_state"][k], atol=1e-12)
    if "short" in path:
        assert z["truncated"][-1] and n == 10


def test_randomization_targets_match_xml_manager():
    """Which bodies / geoms / joint classes XMLManager.get_model_path edits (recorded from the reference) vs the
    flags the model builder derives (quirks C-1 .. C-5)."""
    sem = json.load(open(os.path.join(HERE, "golden", "xml_semantics.json")))
    from cosim_b200.config import make_config, RANDOM_NONE
    for robot, rec in sem.items():
        spec = ROBOTS[robot]
        assert sorted(spec.mass_noise_bodies) == rec["mass_changed"], robot
        assert spec.base_body == rec["load_body"], robot
        m = build_model(make_config(robot, "rocky_hard", random=RANDOM_NONE))
        names = m.meta["body_names"]
        fr_bodies = sorted({names[int(m.sections["geom_body"][g])] for g in np.nonzero(m.sections["geom_fr_random"])[0]})
        assert fr_bodies == rec["friction_bodies"], (robot, fr_bodies, rec["friction_bodies"])
        assert rec["ground_friction"].startswith("0.33") and m.sections["ground_friction"][3] == 1.0       # ground always replaced (C-3)
        assert rec["ground_type"] == "hfield" and (rec["timestep"], rec["iterations"]) == ("0.005", "50")
        # friction loss reaches exactly the dofs whose joint class is one of the classes the reference edited
        assert set(rec["frictionloss_classes"]) <= {"joints", "wheels"}
        assert (int(m.sections["dof_fl_random"].sum()) > 0) == (len(rec["frictionloss_classes"]) > 0)

#!/usr/bin/env python3
"""Golden vectors produced by running the REFERENCE'S OWN PYTHON on top of the oracle's physics.

Runs ONLY in the build container (needs /root/reference).  The reference's env stack
(`envs/build.py:build_env` -> `<Robot>` env + ControlManager + XMLManager + MuJoCoUtils + MathUtils +
StateBuildWrapper / TimeLimitWrapper / CommandWrapper) is imported UNMODIFIED; the three packages it
needs that are not installable here (`mujoco`, `gymnasium`, `glfw`) are replaced by thin shims whose
physics calls (`mj_step`, `mj_forward`, `mj_resetData`, `mj_rnePostConstraint`, `mj_rayHfield`, sensors,
`cfrc_ext`) are served by the CPU oracle (oracle/).  Everything the reference computes ITSELF on the hot
path therefore runs for real: PD control + gear ratios + torque clips, action delay bookkeeping,
observation assembly, projected gravity via scipy, the height-map ray loop, frequency gating, frame
stacking, command scaling and slot overwrite, the one-step command lag, time-limit truncation,
termination on cfrc_ext, the info dict.

The recorded inputs/outputs are committed under tests/golden/*.npz; tests replay the same inputs through the
oracle's restated env layer (and, on the GPU box, through the CUDA engine) and compare.
What this does NOT pin: MuJoCo's own numerics (the shim's physics IS the oracle) -- see DESIGN.md section 6.

Usage: python tools/gen_golden.py
"""
import json
import os
import sys
import types
import warnings
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from cosim_b200.config import make_config, load_tables, RANDOM_NONE  # noqa: E402
from cosim_b200.model import build_model  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CAPTURED_XML = {}


# --------------------------------------------------------------------------------------- shims
class _Sensor:
    def __init__(self, data):
        self.data = data


class _Data:
    def __init__(self, env):
        self._env = env
        m = env._cosim_model
        self.qpos = np.zeros(m.dim("nq"))
        self.qvel = np.zeros(m.dim("nv"))
        self.ctrl = np.zeros(m.dim("nu"))
        self.cfrc_ext = np.zeros((m.dim("nbody"), 6))
        self.site_xpos = np.zeros((4096, 3))
        self.ncon = 0
        self.contact = []
        self._sens = {"angular-velocity": np.zeros(3), "linear-velocity": np.zeros(3), "orientation": np.array([1.0, 0, 0, 0])}

    def sensor(self, name):
        return _Sensor(self._sens[name])


class _Model:
    def __init__(self, cm):
        self.nq, self.nv, self.nu = cm.dim("nq"), cm.dim("nv"), cm.dim("nu")
        self.jnt_qposadr = cm.sections["jnt_qposadr"]
        self.jnt_dofadr = cm.sections["jnt_dofadr"]
        self.geom_bodyid = np.concatenate([[0], cm.sections["geom_body"]])     # MuJoCo geom 0 = the ground geom of the world body
        self.site_size = np.zeros((4096, 3))
        self.site_rgba = np.zeros((4096, 4))
        self._names = {"joint": cm.meta["jnt_names"], "body": cm.meta["body_names"]}


def _pull(env):
    o, d = env._oracle, env.data
    d.qpos[:] = o.get("qpos")[0]
    d.qvel[:] = o.get("qvel")[0]
    d._sens["angular-velocity"][:] = o.get("sens_gyro")[0]
    d._sens["linear-velocity"][:] = o.get("sens_vel")[0]
    d._sens["orientation"][:] = o.get("sens_quat")[0]
    d.ncon = int(o.get("ncon")[0, 0])
    d.contact = [types.SimpleNamespace(geom1=0, geom2=int(c[7]) + 1) for c in o.contacts(0)]


def _push(env):
    env._oracle.set("qpos", env.data.qpos[None, :])
    env._oracle.set("qvel", env.data.qvel[None, :])


class MujocoEnv:
    """Stand-in for gymnasium.envs.mujoco.MujocoEnv (gymnasium 1.0.0): the four calls the reference relies on."""

    def __init__(self, model_path, frame_skip, observation_space, render_mode=None, **kw):
        cfg = json.loads(json.dumps(self.config))
        cfg["random"]["sensor_noise"] = "zero" if cfg["random"]["sensor_noise"] == "none" else cfg["random"]["sensor_noise"]
        self._cosim_model = build_model(cfg)
        self._oracle = Oracle(self._cosim_model, 1, seed=0)
        self.model = _Model(self._cosim_model)
        self.data = _Data(self)
        self.frame_skip = frame_skip
        _ENVS.append(self)

    def do_simulation(self, ctrl, n_frames):          # gymnasium: data.ctrl[:] = ctrl; mj_step(nstep); mj_rnePostConstraint
        self.data.ctrl[:] = ctrl
        _push(self)
        self._oracle.set("ctrl", np.asarray(ctrl, dtype=np.float64)[None, :])
        for _ in range(n_frames):
            self._oracle.substep()
        self._oracle.rne_post()
        _pull(self)
        self.data.cfrc_ext[:] = self._oracle.get("cfrc_ext")[0].reshape(-1, 6)

    def reset(self, seed=None, options=None):         # gymnasium: mj_resetData; ob = reset_model(); info = _get_reset_info()
        mj_resetData(self.model, self.data)
        ob = self.reset_model()
        return ob, self._get_reset_info()

    def render(self):
        pass

    def close(self):
        pass


_ENVS = []


def _env_of(data):
    return data._env


def mj_resetData(model, data):
    env = _env_of(data)
    data.qpos[:] = env._cosim_model.sections["qpos0"]
    data.qvel[:] = 0
    data.ctrl[:] = 0
    env._oracle.set("qacc_warmstart", np.zeros((1, model.nv)))
    env._oracle.set("ctrl", np.zeros((1, model.nu)))


def mj_forward(model, data):
    env = _env_of(data)
    _push(env)
    env._oracle.forward()
    _pull(env)


class mjtObj:
    mjOBJ_BODY, mjOBJ_JOINT, mjOBJ_GEOM, mjOBJ_SITE = 1, 3, 5, 6


def mj_name2id(model, objtype, name):
    if objtype == mjtObj.mjOBJ_JOINT:
        return model._names["joint"].index(name) if name in model._names["joint"] else -1
    if objtype == mjtObj.mjOBJ_BODY:
        return model._names["body"].index(name) if name in model._names["body"] else -1
    if objtype == mjtObj.mjOBJ_GEOM:
        return 0 if name == "ground" else -1
    if objtype == mjtObj.mjOBJ_SITE and name.startswith("heightmap_site_"):
        _, _, i, j = name.split("_")
        return int(i) * 64 + int(j)
    return -1


def mj_rayHfield(model, data, geomid, pnt, vec):
    env = _env_of(data)
    z = env._oracle.ray_hfield(float(pnt[0, 0]), float(pnt[1, 0]))
    if z != z:
        return -1.0
    dist = float(pnt[2, 0]) - z
    return dist if dist >= 0 else -1.0


def install_shims():
    mj = types.ModuleType("mujoco")
    mj.mj_resetData, mj.mj_forward, mj.mj_name2id, mj.mj_rayHfield, mj.mjtObj = mj_resetData, mj_forward, mj_name2id, mj_rayHfield, mjtObj
    gym = types.ModuleType("gymnasium")
    gym_utils = types.ModuleType("gymnasium.utils")

    class EzPickle:
        def __init__(self, *a, **k):
            pass
    gym_utils.EzPickle = EzPickle
    gym_envs = types.ModuleType("gymnasium.envs")
    gym_mj = types.ModuleType("gymnasium.envs.mujoco")
    gym_mj.MujocoEnv = MujocoEnv
    gym_spaces = types.ModuleType("gymnasium.spaces")
    gym_spaces.Box = lambda **kw: None
    gym.utils, gym.envs, gym.spaces = gym_utils, gym_envs, gym_spaces
    gym_envs.mujoco = gym_mj
    glfw = types.ModuleType("glfw")
    for name, mod in [("mujoco", mj), ("gymnasium", gym), ("gymnasium.utils", gym_utils), ("gymnasium.envs", gym_envs),
                      ("gymnasium.envs.mujoco", gym_mj), ("gymnasium.spaces", gym_spaces), ("glfw", glfw)]:
        sys.modules[name] = mod
    # XMLManager writes applied_*.xml into the (read-only) reference tree: capture the tree instead
    def _write(self, path, *a, **k):
        CAPTURED_XML[os.path.basename(path)] = self.getroot()
    ET.ElementTree.write = _write
    sys.path.insert(0, REF)


# --------------------------------------------------------------------------------------- recording
def record(robot, terrain, steps, seed, height_map=False, position_command=False, max_duration=120.0, freq_override=None, push=None):
    from envs.build import build_env          # the reference's factory, unmodified
    et, _ = load_tables()
    kw = {}
    if height_map:
        kw["non_stacked_obs_order"] = list(et[robot]["non_stacked_obs_order"]) + ["height_map"]
    if position_command:
        kw["command_dim"] = 2
    cfg = make_config(robot, terrain, random=dict(RANDOM_NONE, sensor_noise="none"), max_duration=max_duration,
                      position_command=position_command, **kw)
    if freq_override:
        for k, f in freq_override.items():
            cfg["observation"][k]["freq"] = f
    env = build_env(cfg)
    inner = _ENVS[-1]
    rng = np.random.default_rng(seed)
    nu, cd = env.action_dim, env.command_dim
    out = {"states": [], "actions": [], "commands": [], "applied": [], "torque": [], "terminated": [], "truncated": [],
           "lin_vel_x": [], "ang_vel_yaw": [], "action_diff_RMSE": [], "set_points": [], "info_state": [], "qpos": []}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        state, info = env.reset()
        out["reset_state"] = np.array(state, dtype=np.float32)
        user_cmd = np.zeros(cfg["observation"]["command_dim"])
        for k in range(steps):
            if k % 7 == 0:
                user_cmd = rng.uniform(-1.0, 1.0, cfg["observation"]["command_dim"]) * (3.0 if position_command else 1.0)
            env.receive_user_command(user_cmd.copy())
            if push and k in push:                       # the reference's own event code (<robot>.py event())
                env.event("push", push[k])
            action = np.clip(rng.normal(0.0, 0.6, nu), -1, 1)
            state, term, trunc, info = env.step(action)
            out["states"].append(np.array(state, dtype=np.float32)); out["actions"].append(action); out["commands"].append(user_cmd.copy())
            out["applied"].append(np.array(env.applied_command, dtype=np.float64)); out["torque"].append(np.array(info["torque"], dtype=np.float64))
            out["terminated"].append(bool(term)); out["truncated"].append(bool(trunc))
            out["lin_vel_x"].append(float(info["lin_vel_x"])); out["ang_vel_yaw"].append(float(info["ang_vel_yaw"]))
            out["action_diff_RMSE"].append(float(info["action_diff_RMSE"])); out["set_points"].append(np.array(info["set_points"], dtype=np.float64))
            out["info_state"].append(np.array(info["state"], dtype=np.float64)); out["qpos"].append(inner.data.qpos.copy())
            if term or trunc:
                break
    res = {k: np.array(v) for k, v in out.items()}
    res["config_json"] = np.array(json.dumps(cfg))
    res["state_dim"] = np.array(env.state_dim)
    res["cmd_slices"] = np.array([[s.start, s.stop] for s in env.cmd_slices])
    res["push_steps"] = np.array(sorted(push) if push else [], dtype=np.int64)
    res["push_vels"] = np.array([push[k] for k in sorted(push)] if push else np.zeros((0, 3)), dtype=np.float64).reshape(-1, 3)
    return res


def record_xml_semantics():
    """What XMLManager.get_model_path edits (masses, frictions, frictionloss) for a config with every knob away from
    its XML default -> which bodies / geoms / dofs the randomization reaches (quirks C-1 .. C-4)."""
    out = {}
    for robot in ("flamingo_light_v1", "flamingo_p_v3", "w4_p_v2", "humanoid_p_v0"):
        mod = __import__(f"envs.{robot}.manager.xml_manager", fromlist=["XMLManager"])
        cfg = make_config(robot, "rocky_hard", random=dict(RANDOM_NONE, sensor_noise="none", mass_noise=0.25, load=3.0,
                                                            sliding_friction=0.33, torsional_friction=0.044, rolling_friction=0.0055, friction_loss=0.77))
        np.random.seed(123)
        CAPTURED_XML.clear()
        mod.XMLManager(cfg).get_model_path()
        root = list(CAPTURED_XML.values())[0]
        orig = ET.parse(os.path.join(REF, "envs", robot, "assets", "xml", f"{robot}.xml")).getroot()
        mass_changed, load_body = [], None
        om = {b.get("name"): float(b.find("inertial").get("mass")) for b in orig.iter("body") if b.find("inertial") is not None}
        for b in root.iter("body"):
            i = b.find("inertial")
            if i is not None and abs(float(i.get("mass")) - om[b.get("name")]) > 1e-12:
                mass_changed.append(b.get("name"))
                if abs(float(i.get("mass")) - om[b.get("name")]) > 0.25 * om[b.get("name")] + 1e-9:
                    load_body = b.get("name")
        fr_bodies = []
        for b in root.iter("body"):
            for g in b.findall("geom"):
                if g.get("friction", "").startswith("0.33"):
                    fr_bodies.append(b.get("name"))
        ground = [g for g in root.iter("geom") if g.get("name") == "ground"][0]
        fl_classes = [d.get("class") for d in root.iter("default") for j in d.findall("joint") if j.get("frictionloss") == "0.77"]
        out[robot] = dict(mass_changed=sorted(set(mass_changed)), load_body=load_body, friction_bodies=sorted(set(fr_bodies)),
                          ground_friction=ground.get("friction"), ground_type=ground.get("type"), frictionloss_classes=sorted(set(fl_classes)),
                          timestep=root.find("option").get("timestep"), iterations=root.find("option").get("iterations"))
    return out


def main():
    install_shims()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    cases = [("flamingo_p_v3", "rocky_hard", 40, 1, dict(height_map=True)),
             ("flamingo_p_v3", "flat", 12, 2, dict(max_duration=0.2)),                      # truncation at 10 control steps
             ("flamingo_light_v1", "flat", 30, 3, dict(freq_override={"dof_vel": 25, "ang_vel": 10})),
             ("w4_p_v2", "stairs_up_hard", 16, 4, {}),
             ("humanoid_p_v0", "slope_hard", 25, 5, dict(position_command=True)),
             ("flamingo_p_v3", "rocky_hard", 24, 6, dict(push={4: [0.6, -0.4, 0.3], 15: [-0.5, 0.2, 0.0]}))]
    for robot, terrain, steps, seed, kw in cases:
        res = record(robot, terrain, steps, seed, **kw)
        tag = f"{robot}__{terrain}" + ("__hm" if kw.get("height_map") else "") + ("__poscmd" if kw.get("position_command") else "") + \
              ("__short" if kw.get("max_duration") else "") + ("__freq" if kw.get("freq_override") else "") + ("__push" if kw.get("push") else "")
        np.savez_compressed(os.path.join(outdir, tag + ".npz"), **res)
        print(tag, "steps", len(res["states"]), "state_dim", int(res["state_dim"]), "terminated", bool(res["terminated"][-1]), "truncated", bool(res["truncated"][-1]))
    with open(os.path.join(outdir, "xml_semantics.json"), "w") as f:
        json.dump(record_xml_semantics(), f, indent=1, sort_keys=True)
    print("xml semantics written")


if __name__ == "__main__":
    main()

"""Config tables and the `config` dict schema consumed by `build_env`.

Mirrors what the reference's GUI assembles in `ui/main_window.py:_gather_config`
(/root/reference/ui/main_window.py:709-792) and `_make_observation_defaults`
(:100-152): same keys, same defaults (GUI slider defaults at :504-519), so a config
dict produced by the reference GUI can be handed to `cosim_b200.envs.build_env`
unchanged.

Extensions (not in the reference, all optional):
  config["random"][knob] may be a 2-list [lo, hi] -> drawn per env ~ U(lo, hi)
  config["random"]["kp_scale"] / ["kd_scale"]: [lo, hi] per-env gain multipliers
  config["random"]["sensor_noise"] == "zero": a true-zero noise level (SURVEY cfg 2)
  config["engine"]: {"seed", "ncon_max", "auto_reset", "self_collision", "condim", "cone", "solver", "impratio", "iterations", "spawn_spread", "spawn_radius"}
"""
import copy
import os

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))

OBS_TYPES = ["dof_pos", "dof_vel", "ang_vel", "lin_vel", "projected_gravity", "last_action", "height_map"]

# GUI slider defaults, /root/reference/ui/main_window.py:504-519
RANDOM_DEFAULTS = dict(precision="medium", sensor_noise="low", init_noise=0.05, sliding_friction=0.8,
                       torsional_friction=0.02, rolling_friction=0.01, friction_loss=0.10,
                       action_delay_prob=0.05, mass_noise=0.05, load=0.0)
RANDOM_NONE = dict(precision="medium", sensor_noise="zero", init_noise=0.0, sliding_friction=0.8,
                   torsional_friction=0.02, rolling_friction=0.01, friction_loss=0.10,
                   action_delay_prob=0.0, mass_noise=0.0, load=0.0)
# full randomization grid (cfg 3); ranges = GUI slider ranges
RANDOM_FULL = dict(precision="medium", sensor_noise="low", init_noise=0.05, sliding_friction=[0.2, 1.0],
                   torsional_friction=[0.0, 0.1], rolling_friction=[0.0, 0.1], friction_loss=[0.0, 1.0],
                   action_delay_prob=[0.0, 0.5], mass_noise=0.05, load=[0.0, 20.0],
                   kp_scale=[0.8, 1.2], kd_scale=[0.8, 1.2])


def load_tables():
    with open(os.path.join(_HERE, "config", "env_table.yaml")) as f:
        env_table = yaml.safe_load(f)
    with open(os.path.join(_HERE, "config", "random_table.yaml")) as f:
        random_table = yaml.safe_load(f)["random_table"]
    return env_table, random_table


def _to_float(v):
    try:
        return float(v)
    except (TypeError, ValueError):
        return v


def make_observation_defaults(env_cfg, stacked=None, non_stacked=None):
    """/root/reference/ui/main_window.py:100-152."""
    stacked_list = list(stacked if stacked is not None else env_cfg.get("stacked_obs_order", []))
    non_stacked_list = list(non_stacked if non_stacked is not None else env_cfg.get("non_stacked_obs_order", []))
    obs_scales = env_cfg.get("obs_scales", {}) or {}
    obs = {}
    for name in stacked_list + non_stacked_list:
        if name != "command":
            obs[name] = {"freq": 50, "scale": float(obs_scales.get(name, 1.0))}
    for name in OBS_TYPES:
        obs.setdefault(name, None)
    cmd_dim = int((env_cfg.get("command", {}) or {}).get("command_dim", 6))
    scales_cfg = {str(k): float(v) for k, v in (env_cfg.get("command_scales", {}) or {}).items()}
    command_scales = {str(i): scales_cfg.get(str(i), 1.0) for i in range(cmd_dim)}
    if "height_map" in stacked_list or "height_map" in non_stacked_list:
        hm = env_cfg.get("height_map", {}) or {}
        height_map = {"size_x": float(hm.get("size_x", 1.0)), "size_y": float(hm.get("size_y", 0.6)),
                      "res_x": int(hm.get("res_x", 15)), "res_y": int(hm.get("res_y", 9)), "freq": 50, "scale": 1.0}
    else:
        height_map = None
    out = {"stacked_obs_order": stacked_list, "non_stacked_obs_order": non_stacked_list,
           "stack_size": int(env_cfg.get("stack_size", 3)), "command_dim": cmd_dim,
           "command_scales": command_scales}
    out.update(obs)
    out["height_map"] = height_map
    return out


def make_config(env_id, terrain="flat", random=None, max_duration=120.0, position_command=False,
                stacked_obs_order=None, non_stacked_obs_order=None, command_dim=None, engine=None,
                policy=None):
    """Build the reference `config` dict (schema: SURVEY.md A.6) from the YAML tables."""
    env_table, random_table = load_tables()
    if env_id not in env_table:
        raise NameError(f"Please select a valid environment id. Received '{env_id}'.")
    env_cfg = env_table[env_id]
    observation = make_observation_defaults(env_cfg, stacked_obs_order, non_stacked_obs_order)
    if command_dim is not None:
        observation["command_dim"] = int(command_dim)
        observation["command_scales"] = {str(i): observation["command_scales"].get(str(i), 1.0)
                                         for i in range(int(command_dim))}
    hardware = {}
    for k, v in env_cfg["hardware"].items():
        hardware[k] = {kk: _to_float(vv) for kk, vv in v.items()} if isinstance(v, dict) else _to_float(v)
    rnd = dict(RANDOM_DEFAULTS)
    if random:
        rnd.update(random)
    cfg = {
        "env": {"id": env_id, "terrain": terrain, "max_duration": float(max_duration),
                "position_command": bool(position_command)},
        "observation": observation,
        "policy": policy or {"use_lstm": False, "h_in_dim": 256, "c_in_dim": 256, "onnx_file": ""},
        "random": rnd,
        "hardware": hardware,
        "random_table": copy.deepcopy(random_table),
    }
    if engine:
        cfg["engine"] = dict(engine)
    return cfg

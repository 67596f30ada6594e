"""Per-robot control / observation tables.

Each entry restates, as data, what the reference hard-codes in
`envs/<robot>/<robot>.py` (joint orders, PD groups, torque clips, reset pose,
termination bodies).  The batched kernels are generic over these tables.

Actuator modes (reference `ControlManager.pd_controller`, kp*(tq-q)+kd*(td-d),
/root/reference/envs/flamingo_p_v3/manager/control_manager.py:13-15):
  POS : tau = (kp*(a*scale - q*posfac) + kd*(0.0 - qd*posfac)) * gamma, clip
  VEL : tau = 0.0*(0.0-0.0) + kd*(a*scale - qd), clip
"""
from dataclasses import dataclass, field
from typing import List

POS, VEL = 0, 1


@dataclass
class ActGroup:
    joints: List[str]
    mode: int
    kp: str          # hardware key or "" (VEL mode)
    kd: str
    scale: str       # key into hardware["action_scales"]
    clip: str        # hardware key of the torque clip
    geared: bool = False   # legs on p_v3 / w4: q*gear_ratio, qd*gear_ratio, tau*gamma


@dataclass
class RobotSpec:
    id: str
    z0: float
    groups: List[ActGroup]
    dof_pos_joints: List[str]
    dof_vel_joints: List[str]
    geared_joints: List[str]
    init_noise_joints: List[str]          # [] -> all actuated/hinge qpos after the free joint
    state_pos_joints: List[str]           # info["state"]: positions of these, then velocities of state_vel_joints
    state_vel_joints: List[str]
    term_bodies: List[str] = field(default_factory=list)   # cfrc_ext > 1.0 test (p_v3 only)
    term_threshold: float = 1.0
    hm_zmin: float = -1.0
    lin_vel_f32: bool = True
    mass_noise_bodies: List[str] = field(default_factory=list)
    base_body: str = "base_link"
    wheel_bodies: List[str] = field(default_factory=list)  # bodies whose geoms get the random friction (if attr exists)


def _pairs(prefixes, name):
    return [f"{p}_{name}_joint" for p in prefixes]


# ---------------------------------------------------------------- flamingo_light_v1
# /root/reference/envs/flamingo_light_v1/flamingo_light_v1.py:20-31,95-98,131-164,169-187,224-232
LIGHT = RobotSpec(
    id="flamingo_light_v1", z0=0.13,
    groups=[ActGroup(["left_shoulder_joint", "right_shoulder_joint"], POS, "Kp_shoulder", "Kd_shoulder", "shoulder", "leg_max_torque"),
            ActGroup(["left_wheel_joint", "right_wheel_joint"], VEL, "", "Kd_wheel", "wheel", "wheel_max_torque")],
    dof_pos_joints=["left_shoulder_joint", "right_shoulder_joint"],
    dof_vel_joints=["left_shoulder_joint", "right_shoulder_joint", "left_wheel_joint", "right_wheel_joint"],
    geared_joints=[],
    init_noise_joints=["left_shoulder_joint", "right_shoulder_joint", "left_wheel_joint", "right_wheel_joint"],
    state_pos_joints=["left_shoulder_joint", "right_shoulder_joint"],
    state_vel_joints=["left_wheel_joint", "right_wheel_joint"],
    mass_noise_bodies=["base_link", "left_shoulder_link", "right_shoulder_link", "left_wheel_link", "right_wheel_link"],
    wheel_bodies=["left_wheel_link", "right_wheel_link"],
)

# ---------------------------------------------------------------- flamingo_p_v3
# /root/reference/envs/flamingo_p_v3/flamingo_p_v3.py:23-45,108-113,150-188,201-255
_P3_POSJ = ["left_hip_joint", "right_hip_joint", "left_shoulder_joint", "right_shoulder_joint",
            "left_leg_joint", "right_leg_joint"]
_P3_WH = ["left_wheel_joint", "right_wheel_joint"]
P_V3 = RobotSpec(
    id="flamingo_p_v3", z0=0.61282,
    groups=[ActGroup(_P3_POSJ[0:2], POS, "Kp_hip", "Kd_hip", "hip", "leg_max_torque"),          # sic: hips clip with leg_max_torque (:183)
            ActGroup(_P3_POSJ[2:4], POS, "Kp_shoulder", "Kd_shoulder", "shoulder", "leg_max_torque"),  # sic (:184)
            ActGroup(_P3_POSJ[4:6], POS, "Kp_leg", "Kd_leg", "leg", "leg_max_torque", geared=True),
            ActGroup(_P3_WH, VEL, "", "Kd_wheel", "wheel", "wheel_max_torque")],
    dof_pos_joints=_P3_POSJ, dof_vel_joints=_P3_POSJ + _P3_WH, geared_joints=_P3_POSJ[4:6],
    init_noise_joints=[], state_pos_joints=_P3_POSJ, state_vel_joints=_P3_WH,
    term_bodies=["base_link", "left_hip_link", "right_hip_link", "left_shoulder_link", "right_shoulder_link"],
    mass_noise_bodies=["base_link", "left_hip_link", "right_hip_link", "left_shoulder_link", "right_shoulder_link",
                       "left_leg_link", "right_leg_link", "left_wheel_link", "right_wheel_link"],
    wheel_bodies=["left_wheel_link", "right_wheel_link"],
)

# ---------------------------------------------------------------- w4_p_v2
# /root/reference/envs/w4_p_v2/w4_p_v2.py:22-45,109-110,151-198,225-249
_W4 = ["FL", "FR", "RL", "RR"]
_W4_POSJ = _pairs(_W4, "hip") + _pairs(_W4, "shoulder") + _pairs(_W4, "leg")
_W4_WH = _pairs(_W4, "wheel")
W4 = RobotSpec(
    id="w4_p_v2", z0=0.47957,
    groups=[ActGroup(_pairs(_W4, "hip"), POS, "Kp_hip", "Kd_hip", "hip", "hip_max_torque"),
            ActGroup(_pairs(_W4, "shoulder"), POS, "Kp_shoulder", "Kd_shoulder", "shoulder", "shoulder_max_torque"),
            ActGroup(_pairs(_W4, "leg"), POS, "Kp_leg", "Kd_leg", "leg", "leg_max_torque", geared=True),
            ActGroup(_W4_WH, VEL, "", "Kd_wheel", "wheel", "wheel_max_torque")],
    dof_pos_joints=_W4_POSJ, dof_vel_joints=_W4_POSJ + _W4_WH, geared_joints=_pairs(_W4, "leg"),
    init_noise_joints=[], state_pos_joints=_W4_POSJ, state_vel_joints=_W4_WH,
    lin_vel_f32=False,
    mass_noise_bodies=["base_link"] + [f"{p}_{n}_link" for n in ("hip", "shoulder", "leg", "wheel") for p in _W4],
    wheel_bodies=[f"{p}_wheel_link" for p in _W4],
)

# ---------------------------------------------------------------- humanoid_p_v0
# /root/reference/envs/humanoid_p_v0/humanoid_p_v0.py:22-110,139-150,186-262,305-312
_H_GROUPS = [("hip_pitch", 2), ("torso", 1), ("hip_roll", 2), ("shoulder_pitch", 2), ("hip_yaw", 2),
             ("shoulder_roll", 2), ("knee", 2), ("shoulder_yaw", 2), ("ankle_pitch", 2), ("elbow_pitch", 2),
             ("ankle_roll", 2), ("elbow_yaw", 2)]


def _hj(name, n):
    return [f"{name}_joint"] if n == 1 else [f"left_{name}_joint", f"right_{name}_joint"]


_H_JOINTS = [j for name, n in _H_GROUPS for j in _hj(name, n)]
HUMANOID = RobotSpec(
    id="humanoid_p_v0", z0=1.105,
    groups=[ActGroup(_hj(name, n), POS, f"Kp_{name}", f"Kd_{name}", name, f"{name}_joint_max_torque")
            for name, n in _H_GROUPS],
    dof_pos_joints=_H_JOINTS, dof_vel_joints=_H_JOINTS, geared_joints=[],
    init_noise_joints=[], state_pos_joints=_H_JOINTS, state_vel_joints=[],
    hm_zmin=-5.0, base_body="pelvis_link",
    mass_noise_bodies=["pelvis_link", "torso_link",
                       "left_shoulder_pitch_link", "left_shoulder_roll_link", "left_shoulder_yaw_link",
                       "left_elbow_pitch_link", "left_elbow_yaw_link",
                       "right_shoulder_pitch_link", "right_shoulder_roll_link", "right_shoulder_yaw_link",
                       "right_elbow_pitch_link", "right_elbow_yaw_link",
                       "left_hip_pitch_link", "left_hip_roll_link", "left_hip_yaw_link",
                       "left_knee_link", "left_ankle_pitch_link", "left_ankle_roll_link",
                       "right_hip_pitch_link", "right_hip_roll_link", "right_hip_yaw_link",
                       "right_knee_link", "right_ankle_pitch_link", "right_ankle_roll_link"],
    wheel_bodies=["left_ankle_roll_link", "right_ankle_roll_link"],
)

ROBOTS = {r.id: r for r in (LIGHT, P_V3, W4, HUMANOID)}

OBS_DIMS = {  # obs_to_dim, e.g. /root/reference/envs/flamingo_p_v3/flamingo_p_v3.py:81-90
    "flamingo_light_v1": (2, 4), "flamingo_p_v3": (6, 8), "w4_p_v2": (12, 16), "humanoid_p_v0": (23, 23)}

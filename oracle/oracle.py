"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package `cosim_b200`.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def _stale():
    if not os.path.exists(_LIB):
        return True
    t = os.path.getmtime(_LIB)
    srcs = [os.path.join(_HERE, f) for f in ("oracle.hpp", "oracle_math.hpp", "oracle_capi.cpp", "Makefile")] + \
           [os.path.join(_HERE, "..", "include", "cosim_blob.h")]
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in srcs)


def build(force=False):
    if force or _stale():
        subprocess.check_call(["make", "-B", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_reset.argtypes = [ctypes.c_void_p] * 4
        L.orc_step.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int]
        L.orc_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        L.orc_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        L.orc_forward.argtypes = [ctypes.c_void_p]
        L.orc_substep.argtypes = [ctypes.c_void_p]
        L.orc_contacts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.orc_contact_forces.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.orc_philox.restype = ctypes.c_uint32
        L.orc_philox.argtypes = [ctypes.c_uint64] + [ctypes.c_uint32] * 4
        L.orc_philox_block.argtypes = [ctypes.c_uint32] * 6 + [ctypes.c_void_p]
        L.orc_rne_post.argtypes = [ctypes.c_void_p]
        L.orc_ray_hfield.restype = ctypes.c_double
        L.orc_ray_hfield.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        L.orc_convex_pair.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_box_box.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_norm_ppf.restype = ctypes.c_double
        L.orc_norm_ppf.argtypes = [ctypes.c_double]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """N independent scalar environments stepped on the CPU (fp64 by default)."""

    def __init__(self, model, num_envs=1, seed=0, env_offset=0, use_float=False):
        self.model = model
        self.N = int(num_envs)
        self._blob = ctypes.create_string_buffer(model.blob, len(model.blob))
        self.h = lib().orc_create(ctypes.cast(self._blob, ctypes.c_void_p), len(model.blob), self.N, int(seed),
                                  int(env_offset), int(use_float))
        self.nq, self.nv, self.nu = model.dim("nq"), model.dim("nv"), model.dim("nu")
        self.nbody, self.ngeom = model.dim("nbody"), model.dim("ngeom")
        self.state_dim, self.command_dim = model.dim("state_dim"), model.dim("command_dim")
        self._sizes = {
            "qpos": self.nq, "qvel": self.nv, "ctrl": self.nu, "qacc_warmstart": self.nv, "qacc": self.nv,
            "qacc_smooth": self.nv, "qfrc_bias": self.nv, "qfrc_smooth": self.nv, "qfrc_constraint": self.nv,
            "M": self.nv * self.nv, "xpos": 3 * self.nbody, "xquat": 4 * self.nbody, "xipos": 3 * self.nbody,
            "geom_xpos": 3 * self.ngeom, "cfrc_ext": 6 * self.nbody, "body_mass": self.nbody,
            "geom_friction": 3 * self.ngeom, "dof_frictionloss": self.nv, "kp": self.nu, "kd": self.nu,
            "dof_invweight0": self.nv, "body_invweight0": 2 * self.nbody, "torque": self.nu, "action": self.nu,
            "heightmap": max(1, model.dim("hm_res_x") * model.dim("hm_res_y")),
            "hm_cell": max(1, model.dim("hm_res_x") * model.dim("hm_res_y")),
            "obs_buffer": model.dim("stack_size") * model.dim("stacked_dim"),
            "subtree_com": 3, "ground_friction": 3, "delay_prob": 1, "meaninertia": 1, "ncon": 1, "nefc": 1,
            "solver_iter": 1, "ls_evals": 1, "ncon_dropped": 1, "sim_step": 1, "nan_count": 1, "sens_gyro": 3, "sens_vel": 3,
            "sens_quat": 4, "info": 4,
        }

    def __del__(self):
        try:
            if self.h:
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def get(self, name):
        out = np.zeros((self.N, self._sizes[name]), dtype=np.float64)
        rc = lib().orc_get(self.h, name.encode(), _p(out))
        if rc != 0:
            raise KeyError(name)
        return out

    def set(self, name, value):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (self.N, self._sizes[name])))
        if lib().orc_set(self.h, name.encode(), _p(v)) != 0:
            raise KeyError(name)

    def _cmd(self, command):
        if self.command_dim == 0:
            return None
        if command is None:
            return np.zeros((self.N, self.command_dim))
        return np.ascontiguousarray(np.broadcast_to(np.asarray(command, dtype=np.float64), (self.N, self.command_dim)))

    def reset(self, mask=None, command=None):
        state = np.zeros((self.N, self.state_dim), dtype=np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        c = self._cmd(command)
        lib().orc_reset(self.h, _p(m), _p(c), _p(state))
        return state

    def step(self, action, command=None, nthreads=0):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(action, dtype=np.float64), (self.N, self.nu)))
        c = self._cmd(command)
        state = np.zeros((self.N, self.state_dim), dtype=np.float32)
        term = np.zeros(self.N, dtype=np.uint8)
        trunc = np.zeros(self.N, dtype=np.uint8)
        lib().orc_step(self.h, _p(a), _p(c), _p(state), _p(term), _p(trunc), int(nthreads))
        return state, term.astype(bool), trunc.astype(bool)

    def forward(self):
        lib().orc_forward(self.h)

    def substep(self):
        lib().orc_substep(self.h)

    def rne_post(self):
        """mj_rnePostConstraint: fills cfrc_ext from the last solve."""
        lib().orc_rne_post(self.h)

    def ray_hfield(self, x, y):
        """Terrain height under (x, y) as mj_rayHfield sees it for a vertical ray; NaN on a miss."""
        return float(lib().orc_ray_hfield(self.h, float(x), float(y)))

    def convex_pair(self, g1, g2):
        """MPR penetration query between two primitive geoms, each (type, size[3], pos[3], mat[3, 3]); -> (hit, depth, normal, pos).
        The normal points from g1 to g2 (mjc_Convex convention)."""
        buf = np.zeros(32)
        for i, (t, size, pos, mat) in enumerate((g1, g2)):
            buf[16 * i] = t
            buf[16 * i + 1:16 * i + 4] = size
            buf[16 * i + 4:16 * i + 7] = pos
            buf[16 * i + 7:16 * i + 16] = np.asarray(mat, dtype=np.float64).reshape(9)
        out = np.zeros(7)
        r = lib().orc_convex_pair(self.h, _p(buf), _p(out))
        return r == 0, float(out[0]), out[1:4].copy(), out[4:7].copy()

    def box_box(self, b1, b2):
        """mjc_BoxBox restatement between two boxes, each (size[3], pos[3], mat[3, 3]); -> array [n, 7] of pos(3), normal(3), dist."""
        buf = np.zeros(32)
        for i, (size, pos, mat) in enumerate((b1, b2)):
            buf[16 * i] = 6
            buf[16 * i + 1:16 * i + 4] = size
            buf[16 * i + 4:16 * i + 7] = pos
            buf[16 * i + 7:16 * i + 16] = np.asarray(mat, dtype=np.float64).reshape(9)
        out = np.zeros((8, 7))
        n = lib().orc_box_box(self.h, _p(buf), _p(out))
        return out[:n].copy()

    def contacts(self, env=0, cap=256):
        out = np.zeros((cap, 10), dtype=np.float64)
        n = lib().orc_contacts(self.h, int(env), _p(out), cap)
        return out[:min(n, cap)]

    def contact_forces(self, env=0, cap=256):
        """Per contact: force in the contact frame (normal, t1, t2, torsion, roll1, roll2), friction[5], condim."""
        out = np.zeros((cap, 12), dtype=np.float64)
        n = lib().orc_contact_forces(self.h, int(env), _p(out), cap)
        return out[:min(n, cap)]


def philox_block(ctr, key):
    out = (ctypes.c_uint32 * 4)()
    lib().orc_philox_block(*[int(c) for c in ctr], int(key[0]), int(key[1]), out)
    return list(out)


def philox(seed, env, stream, step, idx):
    return int(lib().orc_philox(int(seed), int(env), int(stream), int(step), int(idx)))

"""ctypes binding of tests/hostsim (single-lane host emulation of the CUDA engine's per-env code).

TEST INFRASTRUCTURE ONLY -- see hostsim.cpp.  Never imported by the cosim_b200 package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build_hostsim.so")
_SRC = [os.path.join(_HERE, "hostsim.cpp")] + [os.path.join(_HERE, "..", "..", "cosim_b200", "csrc", f)
                                                for f in ("engine_core.h", "engine_env.h", "engine_setup.h")]
_lib = None


def build(force=False):
    stale = force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in _SRC)
    if stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-o", _LIB, _SRC[0]])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.hs_create.restype = ctypes.c_void_p
        L.hs_create.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32]
        L.hs_destroy.argtypes = [ctypes.c_void_p]
        L.hs_reset.argtypes = [ctypes.c_void_p] * 4
        L.hs_step.argtypes = [ctypes.c_void_p] * 6
        L.hs_push.argtypes = [ctypes.c_void_p] * 3
        L.hs_substep.argtypes = [ctypes.c_void_p]
        L.hs_field_dim.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.hs_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        L.hs_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        L.hs_ws_floats.argtypes = [ctypes.c_void_p]
        L.hs_chol.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.hs_philox.restype = ctypes.c_uint32
        L.hs_philox.argtypes = [ctypes.c_void_p] + [ctypes.c_uint32] * 4
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class HostSim:
    def __init__(self, model, num_envs=1, seed=0, env_offset=0):
        self.model, self.N = model, int(num_envs)
        self._blob = ctypes.create_string_buffer(model.blob, len(model.blob))
        self.h = lib().hs_create(ctypes.cast(self._blob, ctypes.c_void_p), len(model.blob), self.N, int(seed), int(env_offset))
        if not self.h:
            raise RuntimeError("hostsim: create failed")
        self.nu, self.state_dim, self.command_dim = model.dim("nu"), model.dim("state_dim"), model.dim("command_dim")

    def __del__(self):
        try:
            if self.h:
                lib().hs_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _cmd(self, command):
        if self.command_dim == 0 or command is None:
            return None
        return np.ascontiguousarray(np.broadcast_to(np.asarray(command, dtype=np.float32), (self.N, self.command_dim)))

    def reset(self, mask=None, command=None):
        state = np.zeros((self.N, self.state_dim), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().hs_reset(self.h, _p(m), _p(self._cmd(command)), _p(state))
        return state

    def step(self, action, command=None):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(action, dtype=np.float32), (self.N, self.nu)))
        state = np.zeros((self.N, self.state_dim), np.float32)
        term = np.zeros(self.N, np.uint8)
        trunc = np.zeros(self.N, np.uint8)
        lib().hs_step(self.h, _p(a), _p(self._cmd(command)), _p(state), _p(term), _p(trunc))
        return state, term.astype(bool), trunc.astype(bool)

    def substep(self):
        lib().hs_substep(self.h)

    def chol_solve(self, A, b, sparse):
        """x = A^-1 b through the engine's chol_factor / chol_solve (sparse: the tree-sparse pair lists)."""
        nv = self.model.dim("nv")
        A = np.ascontiguousarray(A, dtype=np.float32).reshape(nv, nv); b = np.ascontiguousarray(b, dtype=np.float32)
        x = np.zeros(nv, np.float32)
        lib().hs_chol(self.h, _p(A), _p(b), int(bool(sparse)), _p(x))
        return x

    def push(self, vel, mask=None):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(vel, dtype=np.float32), (self.N, 3)))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().hs_push(self.h, _p(m), _p(v))

    def get(self, name):
        d = lib().hs_field_dim(self.h, name.encode())
        if d < 0:
            raise KeyError(name)
        buf = np.zeros((self.N, d), np.float32)
        is_int = lib().hs_get(self.h, name.encode(), _p(buf))
        return buf.view(np.int32) if is_int else buf

    def set(self, name, value):
        d = lib().hs_field_dim(self.h, name.encode())
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float32), (self.N, d)))
        if lib().hs_set(self.h, name.encode(), _p(v)) != 0:
            raise KeyError(name)

    def philox(self, env, stream, step, idx):
        return int(lib().hs_philox(self.h, env, stream, step, idx))

    @property
    def ws_bytes(self):
        return 4 * lib().hs_ws_floats(self.h)

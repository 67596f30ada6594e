"""Loader for libcosim_b200.so (the C ABI declared in include/cosim_b200.h).

The library is hand-written CUDA for sm_100a (cosim_b200/csrc/engine.cu, policy.cu) and is built
in-tree with nvcc.  There is no CPU path: creating an engine without a CUDA device raises.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
BUILD_DIR = os.path.join(CSRC, "_build")
LIB_PATH = os.environ.get("COSIM_LIB_PATH") or os.path.join(BUILD_DIR, "libcosim_b200.so")      # COSIM_LIB_PATH: a library built elsewhere (tools/build_ab.sh, same-box A / B runs)
SOURCES = ["engine.cu", "engine_gen.cu", "engine_w24.cu", "engine_w12.cu", "policy.cu"]
HEADERS = ["engine_core.h", "engine_env.h", "engine_setup.h", "engine_general.h", "engine_kernels.cuh", os.path.join("..", "..", "include", "cosim_b200.h"),
           os.path.join("..", "..", "include", "cosim_blob.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]

_lib = None


def _stale():
    if os.environ.get("COSIM_LIB_PATH"):
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    files = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))] + \
            [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(f) > t for f in files)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> cosim_b200/csrc/_build/libcosim_b200.so"""
    if not (force or _stale()):
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    # one builder at a time (torchrun starts one process per GPU; all of them may find the library stale): the others wait on the
    # lock and then find a fresh library.  Objects and the library are written under private names and renamed into place.
    import fcntl
    lock = open(os.path.join(BUILD_DIR, ".build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not (force or _stale()):
            return LIB_PATH
        return _build_locked(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    for s in srcs:   # one nvcc per translation unit, in parallel
        o = os.path.join(BUILD_DIR, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", "-o", o, s]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    tmp = LIB_PATH + f".{os.getpid()}.tmp"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


def lib():
    """ctypes handle of libcosim_b200.so with argtypes set; builds it if the sources are newer."""
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            build()
        elif not os.path.exists(LIB_PATH):
            raise RuntimeError("libcosim_b200.so is not built and nvcc is unavailable; run __graft_entry__.build()")
    L = ctypes.CDLL(LIB_PATH)
    vp, cp, i32, u32, u64, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t
    L.cosim_create.argtypes = [vp, sz, i32, i32, u64, u32, ctypes.POINTER(vp)]
    L.cosim_destroy.argtypes = [vp]
    L.cosim_destroy.restype = None
    L.cosim_last_error.argtypes = [vp]
    L.cosim_last_error.restype = cp
    L.cosim_reset.argtypes = [vp, vp, vp, vp, vp]
    L.cosim_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.cosim_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.cosim_push.argtypes = [vp, vp, vp, vp]
    L.cosim_substep.argtypes = [vp, vp]
    L.cosim_field_dim.argtypes = [vp, cp]
    L.cosim_field_is_int.argtypes = [vp, cp]
    L.cosim_get.argtypes = [vp, cp, vp, vp]
    L.cosim_set.argtypes = [vp, cp, vp, vp]
    L.cosim_set_debug.argtypes = [vp, i32]
    L.cosim_stats_reduce.argtypes = [vp, vp, vp]
    L.cosim_stats_clear.argtypes = [vp, vp]
    L.cosim_rng_probe.argtypes = [vp, u32, u32, i32, vp, vp]
    L.cosim_num_envs.argtypes = [vp]
    L.cosim_dim.argtypes = [vp, cp]
    L.cosim_launch_count.argtypes = [vp]
    L.cosim_smem_bytes_per_env.argtypes = [vp]
    L.cosim_warps_per_block.argtypes = [vp]
    L.cosim_pool_size.argtypes = [vp]
    L.cosim_general_path.argtypes = [vp]
    if hasattr(L, "cosim_policy_create"):
        L.cosim_policy_create.argtypes = [i32, i32, vp, vp, vp, i32, ctypes.POINTER(vp)]
        L.cosim_policy_destroy.argtypes = [vp]
        L.cosim_policy_destroy.restype = None
        L.cosim_policy_forward.argtypes = [vp, vp, i32, vp, vp]
        L.cosim_policy_launch_count.argtypes = [vp]
        L.cosim_lstm_cell.argtypes = [vp, vp, vp, i32, i32, vp]
    _lib = L
    return L


EXPORTS = ["cosim_create", "cosim_destroy", "cosim_last_error", "cosim_reset", "cosim_step", "cosim_step_host",
           "cosim_push", "cosim_substep", "cosim_field_dim", "cosim_field_is_int", "cosim_get", "cosim_set", "cosim_set_debug",
           "cosim_stats_reduce", "cosim_stats_clear", "cosim_rng_probe", "cosim_num_envs", "cosim_dim",
           "cosim_launch_count", "cosim_smem_bytes_per_env", "cosim_warps_per_block", "cosim_pool_size", "cosim_general_path",
           "cosim_policy_create", "cosim_policy_destroy", "cosim_policy_forward", "cosim_policy_launch_count", "cosim_lstm_cell"]

"""CPU tests: the C-ABI library loads and exports every symbol include/cosim_b200.h declares (no compute calls
without a GPU), the product path fails loudly without CUDA, and the N > 1 host logic (env sharding + the one
collective: all-reduce of reporter statistics) works on a world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cosim_b200 import lib
    path = lib.build()
    L = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "cosim_b200.h")).read()
    declared = set(re.findall(r"\b(cosim_\w+)\s*\(", hdr))
    declared -= {"cosim_handle", "cosim_policy"}
    assert len(declared) >= 24
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(lib.EXPORTS) <= declared


def test_sass_has_tcgen05_and_tma():
    """The policy kernel really is tcgen05/TMEM/TMA code (SASS mnemonics per B200_PROFILING.md)."""
    from cosim_b200 import lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib.build()], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnem in sass, mnem


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cosim_b200.config import make_config
    from cosim_b200.envs import build_env
    from cosim_b200.policy import MLPPolicy, synthetic_mlp
    with pytest.raises(RuntimeError):
        build_env(make_config("flamingo_p_v3", "flat"))
    with pytest.raises(RuntimeError):
        MLPPolicy(synthetic_mlp(8, 2))
    # the C ABI itself refuses too
    from cosim_b200 import lib
    from cosim_b200.model import build_model
    m = build_model(make_config("flamingo_p_v3", "flat"))
    h = ctypes.c_void_p()
    assert lib.lib().cosim_create(m.blob, len(m.blob), 4, 0, 0, 0, ctypes.byref(h)) != 0 and not h


def test_product_package_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "cosim_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "hostsim" not in src.replace("tests/hostsim", ""), f


def test_shard_envs():
    from cosim_b200.envs import shard_envs
    for total, world in [(65536, 8), (1048576, 8), (10, 3), (7, 8)]:
        parts = [shard_envs(total, world, r) for r in range(world)]
        assert sum(n for n, _ in parts) == total
        off = 0
        for n, o in parts:
            assert o == off
            off += n


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from cosim_b200.envs import all_reduce_stats, derive_stats, STAT_NAMES
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", rank=rank, world_size=2)
s = torch.zeros(16, dtype=torch.float64)
s[STAT_NAMES.index("steps")] = 100 * (rank + 1)
s[STAT_NAMES.index("episodes")] = 4 + rank
s[STAT_NAMES.index("success")] = 2 + rank
s[STAT_NAMES.index("err_vx")] = 10.0 * (rank + 1)
s[STAT_NAMES.index("max_torque")] = 30.0 + 7 * rank
s = all_reduce_stats(s)
d = derive_stats(s.numpy(), 8)
assert d["steps"] == 300 and d["episodes"] == 9 and d["max_torque"] == 37.0, d
assert abs(d["success_rate"] - 5 / 9) < 1e-12 and abs(d["mean_abs_err_lin_vel_x"] - 0.1) < 1e-12
dist.destroy_process_group()
print("ok", rank)
"""


def test_stats_all_reduce_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out

"""The engine's per-env code (host emulation, tests/hostsim) under AddressSanitizer: every robot through reset, control steps,
raw sub-steps, self-collision and the far-from-origin case without an out-of-bounds access on the model tables, the per-env
arrays or the workspace end.  (compute-sanitizer is not available on the GPU pool; this is the CPU stand-in.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_emulation_is_asan_clean(tmp_path):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    lib = str(tmp_path / "hostsim_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                           "-fsanitize=address", "-fno-omit-frame-pointer", "-o", lib, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    code = (
        "import numpy as np, tests.hostsim.hostsim as H\n"
        f"H._LIB = {lib!r}; H.build = lambda force=False: H._LIB\n"
        "from cosim_b200.config import make_config, RANDOM_FULL\n"
        "from cosim_b200.model import build_model\n"
        "rng = np.random.default_rng(0)\n"
        "for robot, terrain in [('flamingo_p_v3', 'rocky_hard'), ('humanoid_p_v0', 'slope_hard'), ('flamingo_light_v1', 'flat'), ('w4_p_v2', 'stairs_up_hard')]:\n"
        "    m = build_model(make_config(robot, terrain, random=RANDOM_FULL))\n"
        "    h = H.HostSim(m, 3, seed=2); h.reset()\n"
        "    q = h.get('qpos'); q[:, 7:] += rng.uniform(-0.4, 0.4, q[:, 7:].shape); h.set('qpos', q)\n"
        "    for _ in range(3):\n"
        "        h.step(rng.uniform(-1, 1, (3, m.dim('nu'))))\n"
        "    h.substep()\n"
        "print('asan-clean')\n")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0 and "asan-clean" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])

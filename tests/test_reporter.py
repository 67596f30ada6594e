"""Reporter: reference interface (core/reporter.py:197-218) + a dependency-free PDF."""
import re
import zlib

import numpy as np

from cosim_b200.config import make_config
from cosim_b200.reporter import Reporter


def _fake_info(t, nu=8):
    rng = np.random.default_rng(t)
    return {"dt": 0.02, "action": rng.uniform(-1, 1, nu), "action_diff_RMSE": float(rng.uniform(0, 0.2)),
            "torque": rng.normal(0, 5, nu), "lin_vel_x": np.float32(np.sin(0.05 * t)), "lin_vel_y": 0.0, "ang_vel_yaw": 0.1,
            "set_points": rng.uniform(-1, 1, nu), "state": rng.uniform(-1, 1, nu),
            "user_command_0": 0.5, "user_command_1": 0.0, "user_command_2": -0.2}


def _parse(path):
    data = open(path, "rb").read()
    assert data.startswith(b"%PDF-1.4") and data.rstrip().endswith(b"%%EOF")
    xref = int(re.search(rb"startxref\s+(\d+)", data).group(1))
    assert data[xref:xref + 4] == b"xref"
    n = int(re.search(rb"xref\s+0 (\d+)", data[xref:]).group(1))
    offs = [int(m) for m in re.findall(rb"(\d{10}) 00000 n", data[xref:])]
    assert len(offs) == n - 1
    for i, o in enumerate(offs):                                   # every xref entry points at its object
        assert data[o:].startswith(b"%d 0 obj" % (i + 1))
    pages = int(re.search(rb"/Type /Pages /Count (\d+)", data).group(1))
    text = b""
    for m in re.finditer(rb"/Length (\d+) /Filter /FlateDecode >>\nstream\n", data):
        text += zlib.decompress(data[m.end():m.end() + int(m.group(1))])
    return pages, text


def test_history_bookkeeping_matches_reference():
    r = Reporter("/tmp/unused.pdf", {})
    for t in range(3):
        r.write_info(_fake_info(t))
    assert r.timesteps == 3 and len(r.history["torque"]) == 3 and set(r.history) == set(_fake_info(0))
    rows = r._build_config_rows({"env": {"id": "x", "list": [1, 2]}, "seed": 3})
    assert rows == [["env", ""], ["    id", "x"], ["    list", "1, 2"], ["seed", "3"]]


def test_pdf_sections(tmp_path):
    cfg = make_config("flamingo_p_v3", "rocky_hard")
    r = Reporter(str(tmp_path / "report.pdf"), cfg)
    for t in range(400):
        r.write_info(_fake_info(t))
    r.write_population({"episodes": 1024.0, "termination_rate": 0.03, "mean_contacts": 1.2})
    path = r.generate_report()
    pages, text = _parse(path)
    for title in (b"Set Points vs. States", b"Command Inputs vs. Measured Outputs", b"Action Oscillation and Applied Torques",
                  b"Torque Distribution of All Joints", b"Population Statistics", b"Configuration", b"flamingo_p_v3"):
        assert title in text, title
    assert pages >= 6


def test_batched_info_rows_and_empty_history(tmp_path):
    r = Reporter(str(tmp_path / "b.pdf"), {"env": {"id": "w4"}})
    for t in range(5):
        info = {k: np.stack([np.asarray(v, dtype=float)] * 4) for k, v in _fake_info(t).items() if k != "dt"}
        info["dt"] = 0.02
        r.write_info(info, env_index=2)
    assert np.asarray(r.history["torque"][0]).shape == (8,)
    pages, _ = _parse(r.generate_report())
    assert pages >= 4
    empty = Reporter(str(tmp_path / "e.pdf"), {})
    pages, _ = _parse(empty.generate_report())          # a report with no steps is just the cover
    assert pages == 1

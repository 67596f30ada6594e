"""Direct comparison with MuJoCo itself -- auto-skipped where it cannot run.

`mujoco` (the reference pins 3.2.7, requirements.txt:19) is not installable in the build container or on the GPU box, and
the reference's MJCF / STL assets only exist under /root/reference.  The moment both are available this test pins the
oracle against `mj_step` on the one robot whose meshes are all present in the checkout (humanoid_p_v0): 50 sub-steps of a
contact-free drop, qpos/qvel within 1e-5 relative per step (north_star)."""
import os

import numpy as np
import pytest

mujoco = pytest.importorskip("mujoco")
XML = "/root/reference/envs/humanoid_p_v0/assets/xml/humanoid_p_v0.xml"
pytestmark = pytest.mark.skipif(not os.path.exists(XML), reason="reference assets not present")


def test_contact_free_drop_matches_mj_step():
    from cosim_b200.config import make_config, RANDOM_NONE
    from cosim_b200.model import build_model
    from oracle.oracle import Oracle
    import xml.etree.ElementTree as ET
    root = ET.parse(XML).getroot()
    for geom in root.iter("geom"):                       # `flat` terrain as XMLManager sets it (xml_manager.py:26-29)
        if geom.get("name") == "ground":
            geom.set("type", "plane"); geom.attrib.pop("hfield", None); geom.set("size", "100 100 0.1")
    # meshdir etc. are relative to the XML: write next to a symlinked asset tree
    os.makedirs("/tmp/humanoid_assets/xml", exist_ok=True)
    for d in ("mesh", "terrain"):
        dst = f"/tmp/humanoid_assets/{d}"
        if not os.path.exists(dst):
            os.symlink(os.path.join(os.path.dirname(XML), "..", d), dst)
    tmp = "/tmp/humanoid_assets/xml/humanoid_flat.xml"
    ET.ElementTree(root).write(tmp)
    mm = mujoco.MjModel.from_xml_path(tmp)
    md = mujoco.MjData(mm)
    m = build_model(make_config("humanoid_p_v0", "flat", random=RANDOM_NONE))
    o = Oracle(m, 1)
    o.reset()
    q = o.get("qpos"); q[0, 2] = 3.0
    o.set("qpos", q); o.set("qvel", np.zeros((1, m.dim("nv")))); o.set("ctrl", np.zeros((1, m.dim("nu"))))
    mujoco.mj_resetData(mm, md)
    md.qpos[:] = q[0]; md.qvel[:] = 0
    for k in range(50):
        mujoco.mj_step(mm, md); o.substep()
        for ref, got in ((md.qpos, o.get("qpos")[0]), (md.qvel, o.get("qvel")[0])):
            rel = np.abs(ref - got).max() / max(1.0, np.abs(ref).max())
            assert rel <= 1e-5, f"sub-step {k}: relative error {rel:.2e}"

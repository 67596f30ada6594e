"""CPU tests: pin the oracle (test infrastructure) with known-answer values derived from the reference's
DATA, published test vectors and analytic physics.  No GPU needed."""
import hashlib
import math

import numpy as np
import pytest

from cosim_b200.config import make_config, RANDOM_NONE, RANDOM_DEFAULTS, RANDOM_FULL
from cosim_b200.model import build_model, load_robot, load_terrain_raster, fk_qpos0, hfield_from_raster
from oracle.oracle import Oracle, philox, lib as olib

ROBOTS = {"flamingo_light_v1": dict(mass=4.94905, nq=19, nv=18, nu=4, nbody=13, state_dim=52, z0=0.13),
          "flamingo_p_v3": dict(mass=16.51937, nq=15, nv=14, nu=8, nbody=9, state_dim=88, z0=0.61282),
          "w4_p_v2": dict(mass=36.20476, nq=23, nv=22, nu=16, nbody=17, state_dim=153, z0=0.47957),
          "humanoid_p_v0": dict(mass=61.80264, nq=30, nv=29, nu=23, nbody=25, state_dim=238, z0=1.105)}


@pytest.mark.parametrize("robot", list(ROBOTS))
def test_model_dimensions_and_mass(robot):
    k = ROBOTS[robot]
    m = build_model(make_config(robot, "flat", random=RANDOM_NONE))
    assert (m.dim("nq"), m.dim("nv"), m.dim("nu"), m.dim("nbody") - 1, m.dim("state_dim")) == (k["nq"], k["nv"], k["nu"], k["nbody"], k["state_dim"])
    assert abs(m.sections["body_mass"].sum() - k["mass"]) < 1e-5
    assert abs(m.opt("z0") - k["z0"]) < 1e-12


def test_fk_known_answers():
    """World positions of body frames at the reset pose (joints 0, unit quaternion, base at z0)."""
    want = {"flamingo_p_v3": {"left_wheel_link": (0.030892, 0.1624, 0.102468), "right_wheel_link": (0.030892, -0.1624, 0.102468)},
            "w4_p_v2": {"FL_wheel_link": (0.286498, 0.1624, 0.10357), "RR_wheel_link": (-0.286499, -0.157614, 0.102378)},
            "flamingo_light_v1": {"left_wheel_link": (-0.04458, 0.162, 0.073179)},
            "humanoid_p_v0": {"left_ankle_roll_link": (0.007871, 0.123403, 0.19069), "right_ankle_roll_link": (0.007893, -0.134697, 0.190695)}}
    com = {"flamingo_p_v3": (0.031932, -0.00025, 0.454436), "w4_p_v2": (-0.014233, -0.000617, 0.374889),
           "flamingo_light_v1": (-0.045053, -0.000651, 0.085868), "humanoid_p_v0": (0.002123, -0.000873, 0.98741)}
    for robot, bodies in want.items():
        m = build_model(make_config(robot, "flat", random=RANDOM_NONE))
        o = Oracle(m, 1)
        o.reset()
        xpos = o.get("xpos")[0].reshape(-1, 3)
        names = m.meta["body_names"]
        for b, p in bodies.items():
            np.testing.assert_allclose(xpos[names.index(b)], p, atol=(5e-5 if robot == "flamingo_light_v1" else 2e-6), err_msg=f"{robot}:{b}")   # Appendix D quotes the light wheels as +-0.162
        np.testing.assert_allclose(o.get("subtree_com")[0], com[robot], atol=2e-6)
        rb = load_robot(robot)       # numpy FK of the model builder agrees with the oracle's kinematics
        xp, _ = fk_qpos0(rb, base_pos=np.array([0, 0, ROBOTS[robot]["z0"]]))
        np.testing.assert_allclose(xp[1:], xpos[1:], atol=1e-9)


def test_terrain_rasters():
    r = load_terrain_raster("rocky_hard")
    assert r.shape == (512, 512) and hashlib.md5(np.ascontiguousarray(r).tobytes()).hexdigest() == hashlib.md5(load_terrain_raster("rocky_easy").tobytes()).hexdigest()
    assert r.min() == 1 and r.max() == 101 and (r[253:259, 253:259] == 1).all()
    s = load_terrain_raster("stairs_up_hard")
    assert s.shape == (1024, 1024) and sorted(np.unique(s)) == [0, 51, 102, 153, 204, 255]
    row = s[512].astype(int)
    assert list(np.nonzero(np.diff(row))[0]) == [275, 309, 342, 375, 409, 613, 647, 680, 713, 747]
    h = hfield_from_raster(r)
    assert h.dtype == np.float32 and h.min() == 0.0 and h.max() == 1.0 and (h[0] == ((r[-1].astype(np.float32) - 1) / 100)).all()
    m = build_model(make_config("flamingo_p_v3", "rocky_hard", random=RANDOM_NONE))
    assert (m.dim("hf_nrow"), m.dim("hf_ncol")) == (512, 512) and m.opt("hf_sx") == 140 and m.opt("hf_sz") == 0.25
    assert build_model(make_config("w4_p_v2", "stairs_up_hard", random=RANDOM_NONE)).opt("hf_sz") == 1.5


def test_hull_sizes():
    for robot in ("flamingo_p_v3", "w4_p_v2"):
        rb = load_robot(robot)
        bodies = [str(n) for n in rb["body_names"]]           # geom_body is 1-based (0 = world)
        wheels = [g for g in range(len(rb["geom_body"])) if "wheel" in bodies[int(rb["geom_body"][g]) - 1]]
        assert wheels and all(int(rb["geom_vnum"][g]) == 696 for g in wheels)


def test_philox_published_vectors():
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11)."""
    from oracle.oracle import philox_block
    assert philox_block((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox_block((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox_block((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # the per-draw accessor indexes the same blocks: word idx & 3 of counter (env, stream, step, idx >> 2)
    assert [philox(0, 0, 0, 0, i) for i in range(4)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


def test_norm_ppf_and_truncated_normal():
    from scipy import stats
    L = olib()
    for p in [1e-9, 1e-4, 0.01, 0.3, 0.5, 0.77, 0.999, 1 - 1e-7]:
        assert abs(L.orc_norm_ppf(p) - stats.norm.ppf(p)) < 1e-9 * max(1, abs(stats.norm.ppf(p)))
    # the oracle's sensor noise is the inverse-CDF truncated normal of scipy.stats.truncnorm (noise_generator_utils.py:22-28)
    m = build_model(make_config("flamingo_p_v3", "flat", random=dict(RANDOM_NONE, sensor_noise="extreme")))
    nz = m.sections["noise"].reshape(6, 6)[1]       # dof_vel: mean, std, lower, upper, cdf(a), cdf(b)
    a, b = (nz[2] - nz[0]) / nz[1], (nz[3] - nz[0]) / nz[1]
    u = (np.arange(2000) + 0.5) / 2000
    x = np.array([nz[0] + nz[1] * L.orc_norm_ppf(nz[4] + ui * (nz[5] - nz[4])) for ui in u])
    np.testing.assert_allclose(x, stats.truncnorm.ppf(u, a, b, loc=nz[0], scale=nz[1]), atol=1e-9)


def test_free_fall_and_momentum():
    """humanoid (free joint without friction loss) dropped from 3 m: base acceleration = g, then analytic z(t)."""
    m = build_model(make_config("humanoid_p_v0", "flat", random=RANDOM_NONE))
    o = Oracle(m, 1)
    o.reset()
    q = o.get("qpos"); q[0, 2] = 3.0; o.set("qpos", q)
    o.set("qvel", np.zeros((1, m.dim("nv")))); o.set("ctrl", np.zeros((1, m.dim("nu"))))
    dt, g = m.opt("timestep"), m.opt("gz")
    mass = m.sections["body_mass"].sum()
    z0 = None
    for k in range(1, 21):
        o.substep()
        M = o.get("M")[0].reshape(m.dim("nv"), -1)
        # total linear momentum = row 0..2 of M times qvel (free-joint translational dofs are world-aligned)
        p = M[:3] @ o.get("qvel")[0]
        np.testing.assert_allclose(p, [0, 0, mass * g * dt * k], rtol=1e-9, atol=1e-9)
    assert o.get("ncon")[0, 0] == 0


def test_newton_kkt_residual():
    """At the solver's answer: M (qacc - qacc_smooth) = J^T f, contact forces non-negative, friction-loss bounded."""
    for robot, terrain in [("flamingo_p_v3", "rocky_hard"), ("flamingo_light_v1", "flat"), ("humanoid_p_v0", "slope_hard")]:
        m = build_model(make_config(robot, terrain, random=RANDOM_NONE))
        o = Oracle(m, 4, seed=5)
        o.reset()
        rng = np.random.default_rng(1)
        for _ in range(15):
            o.step(rng.uniform(-1, 1, (4, m.dim("nu"))))
        nv = m.dim("nv")
        for e in range(4):
            M = o.get("M")[e].reshape(nv, nv)
            lhs = M @ (o.get("qacc")[e] - o.get("qacc_smooth")[e])
            rhs = o.get("qfrc_constraint")[e]
            scale = 1.0 / (o.get("meaninertia")[e, 0] * nv)
            assert scale * np.linalg.norm(lhs - rhs) < 1e-5, f"{robot}: KKT residual {scale * np.linalg.norm(lhs - rhs):.2e}"


def test_resting_contact_supports_weight():
    """flamingo_light_v1 standing on the plane with zero action: after settling, cfrc_ext normal forces sum to m g."""
    m = build_model(make_config("flamingo_light_v1", "flat", random=RANDOM_NONE))
    o = Oracle(m, 1)
    o.reset()
    for _ in range(150):
        o.step(np.zeros((1, 4)))
    f = o.get("cfrc_ext")[0].reshape(-1, 6)
    mg = m.sections["body_mass"].sum() * abs(m.opt("gz"))
    assert abs(f[:, 5].sum() - mg) / mg < 0.02, f"normal force {f[:, 5].sum():.3f} vs weight {mg:.3f}"
    assert abs(o.get("qvel")[0, :3]).max() < 0.05


def test_reference_quirks():
    """C-2/C-4 (inert knobs), C-7 (first state has zero command), C-13 (zero-noise extension), wheel mu = max(ground, 1)."""
    m = build_model(make_config("flamingo_p_v3", "rocky_hard", random=dict(RANDOM_DEFAULTS, sliding_friction=0.3)))
    assert int(m.sections["geom_fr_random"].sum()) == 0                       # wheel meshes carry no friction attr
    assert int(m.sections["dof_fl_random"].sum()) == 2                        # only the `wheels` class joints
    o = Oracle(m, 2, seed=3)
    assert (o.get("ground_friction")[:, 0] == np.float64(0.3)).all()
    o.reset()
    for _ in range(10):
        o.step(np.zeros((2, 8)))
    cons = o.contacts(0)
    assert len(cons) == 0 or (cons[:, 9] == 1.0).all()                          # mu = max(0.3, default 1.0)
    light = build_model(make_config("flamingo_light_v1", "flat", random=dict(RANDOM_DEFAULTS, sliding_friction=0.3)))
    assert int(light.sections["geom_fr_random"].sum()) == 2                   # cylinders have the attr
    with pytest.raises(ValueError):                                            # mj_rayHfield on a plane is fatal in the reference
        from cosim_b200.config import load_tables
        et, _ = load_tables()
        build_model(make_config("flamingo_p_v3", "flat", non_stacked_obs_order=list(et["flamingo_p_v3"]["non_stacked_obs_order"]) + ["height_map"]))


def test_blob_header_in_sync():
    import re
    from cosim_b200.model import DIMS, OPTS
    src = open("include/cosim_blob.h").read()
    assert re.findall(r"CD_(\w+) = \d+", src) == DIMS and re.findall(r"CO_(\w+) = \d+", src) == OPTS


def test_every_robot_terrain_and_precision_builds():
    """All 4 robots x their terrains (xml_manager.py:21-32) x precision presets (random_table.yaml:2-22) build a model."""
    from cosim_b200.config import load_tables
    et, rt = load_tables()
    from cosim_b200.model import load_robot
    for robot in ROBOTS:
        rb = load_robot(robot)
        terrains = ["flat"] + [str(n) for n in rb["hfield_names"] if str(n) != "flat"]
        assert len(terrains) >= 8
        for t in terrains:
            m = build_model(make_config(robot, t, random=RANDOM_NONE))
            assert m.dim("ground_type") == (0 if t == "flat" else 1)
    for prec, (dt, it, fs) in {"low": (0.01, 50, 2), "medium": (0.005, 50, 4), "high": (0.0025, 75, 8), "ultra": (0.00125, 75, 16), "extreme": (0.000625, 100, 32)}.items():
        m = build_model(make_config("flamingo_p_v3", "rocky_easy", random=dict(RANDOM_NONE, precision=prec)))
        assert (m.opt("timestep"), m.dim("iterations"), m.dim("frame_skip")) == (dt, it, fs)


def test_self_collision_pair_table():
    """Broad-phase candidates [upstream filterBodyPair + contype/conaffinity + <exclude>; SURVEY.md 8f row 3]."""
    from cosim_b200.model import self_collision_pairs, load_robot
    counts = {}
    for robot in ("flamingo_light_v1", "flamingo_p_v3", "w4_p_v2", "humanoid_p_v0"):
        rb = load_robot(robot)
        pairs = self_collision_pairs(rb)
        counts[robot] = len(pairs)
        par = np.concatenate([[0], rb["body_parent"]])
        excl = {tuple(sorted(x)) for x in rb["exclude_body"].tolist()}
        seen = set()
        for g1, g2 in pairs.tolist():
            b1, b2 = int(rb["geom_body"][g1]), int(rb["geom_body"][g2])
            assert b1 != b2 and par[b1] != b2 and par[b2] != b1            # one joint per body here: weld parent = parent
            assert tuple(sorted((b1, b2))) not in excl
            assert rb["geom_type"][g1] <= rb["geom_type"][g2]
            assert not rb["geom_proxy"][g1] and not rb["geom_proxy"][g2]
            assert (g1, g2) not in seen and (g2, g1) not in seen
            seen.add((g1, g2))
        bp = [tuple(sorted((int(rb["geom_body"][a]), int(rb["geom_body"][b])))) for a, b in pairs.tolist()]
        assert bp == sorted(bp)                                            # (body1, body2) order of mj_collision
    # flamingo_light: every robot geom is contype 1 / conaffinity 2 -> no robot-robot pair at all (flamingo_light_v1.xml:17)
    assert counts == {"flamingo_light_v1": 0, "flamingo_p_v3": 4, "w4_p_v2": 39, "humanoid_p_v0": 212}
    m = build_model(make_config("humanoid_p_v0", "flat", engine={"self_collision": False}))
    assert m.dim("npair") == 0


def test_convex_pair_known_answers():
    """mjc_Convex restatement on primitives with closed-form penetration: normal from geom1 to geom2, pos = midpoint."""
    o = Oracle(build_model(make_config("flamingo_p_v3", "flat")), 1)
    I = np.eye(3)
    hit, depth, n, p = o.convex_pair((2, [0.1, 0, 0], [0, 0, 0], I), (2, [0.2, 0, 0], [0.25, 0, 0], I))
    assert hit and abs(depth - 0.05) < 1e-9
    np.testing.assert_allclose(n, [1, 0, 0], atol=1e-9); np.testing.assert_allclose(p, [0.075, 0, 0], atol=1e-9)
    hit, depth, n, p = o.convex_pair((6, [0.1, 0.1, 0.1], [0, 0, 0], I), (6, [0.1, 0.1, 0.1], [0.01, 0.02, 0.18], I))
    assert hit and abs(depth - 0.02) < 1e-9 and abs(p[2] - 0.09) < 1e-9
    np.testing.assert_allclose(n, [0, 0, 1], atol=1e-9)
    # parallel cylinders side by side, rotated frame: depth = 2 r - gap along the line of centres
    c, s = np.cos(0.7), np.sin(0.7)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    d = np.array([0.06, 0.05, 0.0]); d = d / np.linalg.norm(d) * 0.08
    hit, depth, n, p = o.convex_pair((5, [0.05, 0.2, 0], [1, 2, 3], R), (5, [0.05, 0.2, 0], np.array([1, 2, 3.01]) + d, R))
    assert hit and abs(depth - 0.02) < 1e-5
    np.testing.assert_allclose(n, d / 0.08, atol=1e-9)          # mjc_fixNormal: radial direction of the cylinder walls, exact
    # a cylinder standing on a box: the point is on the cylinder's cap -> no normal from the cylinder, MPR's is kept
    hit, depth, n, p = o.convex_pair((5, [0.05, 0.1, 0], [0, 0, 0.19], I), (6, [0.3, 0.3, 0.1], [0, 0, 0], I))
    assert hit and abs(depth - 0.01) < 1e-6
    np.testing.assert_allclose(n, [0, 0, -1], atol=1e-6)
    # a sphere resting in a box corner region: normal = centre -> contact point
    hit, depth, n, p = o.convex_pair((2, [0.1, 0, 0], [0, 0, 0.19], I), (6, [0.3, 0.3, 0.1], [0, 0, 0], I))
    assert hit and abs(depth - 0.01) < 1e-6
    np.testing.assert_allclose(n, [0, 0, -1], atol=1e-4)         # through MPR's contact point (accurate to the ccd tolerance)
    hit, *_ = o.convex_pair((5, [0.05, 0.2, 0], [0, 0, 0], I), (5, [0.05, 0.2, 0], [0, 0.11, 0.01], I))
    assert not hit


def test_self_contact_is_action_reaction():
    """A geom-geom contact pushes its two bodies with equal and opposite force: with the legs pressed into each other in
    mid-air the humanoid's self-contact rows must leave the momentum balance of free fall untouched."""
    m = build_model(make_config("humanoid_p_v0", "flat", random=RANDOM_NONE))
    o = Oracle(m, 1, seed=1)
    o.reset()
    q = o.get("qpos"); q[:, 2] += 5.0          # lift off the ground: only self-contacts remain
    q[:, 19] -= 0.4                            # swing one leg into the other
    o.set("qpos", q); o.set("qvel", np.zeros_like(o.get("qvel")))
    o.forward()
    con = o.contacts(0)
    assert len(con) >= 1 and (con[:, 8] <= -2).all()
    o.rne_post()
    cf = o.get("cfrc_ext").reshape(-1, 6)
    assert np.abs(cf).max() > 1e-3
    np.testing.assert_allclose(cf.sum(axis=0), 0, atol=1e-9)
    # internal forces cannot move the centre of mass: no net generalized force on the floating base's translation
    qc = o.get("qfrc_constraint")[0]
    assert np.abs(qc).max() > 1e-3
    np.testing.assert_allclose(qc[:3], 0, atol=1e-9)


def test_terrain_authoring_from_png(tmp_path):
    """A user's own height image: rows flipped, normalised to [0, 1], scaled by the MJCF size; the
    robot comes to rest on the authored surface."""
    from PIL import Image
    yy, xx = np.mgrid[0:64, 0:96]
    img = (40 + 60 * (xx / 95.0) + 20 * np.sin(yy / 6.0)).astype(np.uint8)         # ramp along x with ripples along y
    path = str(tmp_path / "ramp.png")
    Image.fromarray(img, mode="L").save(path)
    size = [12.0, 8.0, 0.6, 0.1]
    cfg = make_config("flamingo_p_v3", {"png": path, "size": size, "name": "ramp"}, random=RANDOM_NONE)
    m = build_model(cfg)
    assert (m.dim("hf_nrow"), m.dim("hf_ncol")) == (64, 96) and m.dim("ground_type") == 1
    hf = m.sections["hfield_data"].reshape(64, 96)
    ref = img[::-1].astype(np.float32); ref = (ref - ref.min()) / (ref.max() - ref.min())
    np.testing.assert_array_equal(hf, ref)
    o = Oracle(m, 1, seed=1)
    o.reset()
    # mj_rayHfield: height at the centre of the map = bilinear surface of the authored raster
    z_mid = o.ray_hfield(0.0, 0.0)
    assert abs(z_mid - size[2] * float(ref[31:33, 47:49].mean())) < 0.02
    for _ in range(150):
        o.step(np.zeros((1, m.dim("nu"))))
    q = o.get("qpos")[0]
    ground = o.ray_hfield(q[0], q[1])
    assert 0.05 < q[2] - ground < 0.6 and o.get("ncon")[0, 0] >= 1            # resting on the authored terrain
    same = build_model(make_config("flamingo_p_v3", {"raster": img, "size": size}, random=RANDOM_NONE))
    np.testing.assert_array_equal(same.sections["hfield_data"], m.sections["hfield_data"])
    with pytest.raises(ValueError):
        build_model(make_config("flamingo_p_v3", {"raster": img.astype(np.float32), "size": size}))


def test_box_box_known_answers():
    """mjc_BoxBox restatement [upstream engine_collision_box.c]: face contacts give the clipped incident polygon (4 corners
    when one face lies inside the other, 8 for two squares turned by 45 degrees), crossed edges give one point; position
    midway between the surfaces, normal from box 1 to box 2, dist = -depth.  (humanoid_p_v0.xml:33,40,110,139: its box geoms.)"""
    o = Oracle(build_model(make_config("flamingo_p_v3", "flat")), 1)
    I = np.eye(3)
    c = o.box_box(([0.5, 0.5, 0.1], [0, 0, 0], I), ([0.1, 0.2, 0.1], [0.05, 0.02, 0.19], I))      # small box resting on a big one, 1 cm deep
    assert len(c) == 4
    np.testing.assert_allclose(c[:, 6], -0.01, atol=1e-12); np.testing.assert_allclose(c[:, 3:6], [[0, 0, 1]] * 4, atol=1e-12)
    np.testing.assert_allclose(c[:, 2], 0.095, atol=1e-12)
    assert sorted(map(tuple, np.round(c[:, :2], 9))) == sorted([(-0.05, -0.18), (-0.05, 0.22), (0.15, 0.22), (0.15, -0.18)])
    np.testing.assert_allclose(c[:, :2].mean(axis=0), [0.05, 0.02], atol=1e-12)      # centred under the small box: equal normal forces at rest
    c = o.box_box(([0.1, 0.2, 0.1], [0, 0, 0], I), ([0.5, 0.5, 0.1], [0.05, 0.02, 0.19], I))      # the big one on top: corners of the small one
    assert len(c) == 4 and sorted(map(tuple, np.round(c[:, :2], 9))) == sorted([(0.1, 0.2), (0.1, -0.2), (-0.1, -0.2), (-0.1, 0.2)])
    a = np.pi / 4
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    c = o.box_box(([0.2, 0.2, 0.1], [0, 0, 0], I), ([0.2, 0.2, 0.1], [0, 0, 0.19], Rz))            # two squares at 45 degrees: an octagon
    assert len(c) == 8
    np.testing.assert_allclose(np.sort(np.abs(c[:, :2]), axis=1), [[0.2 * (np.sqrt(2) - 1), 0.2]] * 8, atol=1e-9)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]]); Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    c = o.box_box(([0.5, 0.1, 0.1], [0, 0, 0], Rx), ([0.1, 0.5, 0.1], [0, 0, 0.27], Ry))            # crossed edges
    assert len(c) == 1
    np.testing.assert_allclose(c[0], [0, 0, 0.135, 0, 0, 1, -(0.2 * np.sqrt(2) - 0.27)], atol=1e-6)      # the edge test pads its radii by 1e-6 per term against parallel edges
    assert len(o.box_box(([0.1, 0.1, 0.1], [0, 0, 0], I), ([0.1, 0.1, 0.1], [0.3, 0, 0], I))) == 0
    # through the collision pipeline: the normal forces of a symmetric 4-point support are equal -- checked on the solver level by
    # tests/test_hostsim_vs_oracle.py::test_box_box_contacts_match_oracle (humanoid box geoms pressed flat on each other)

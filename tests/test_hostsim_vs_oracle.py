"""CPU tests of the CUDA engine's per-env LOGIC: tests/hostsim compiles cosim_b200/csrc/engine_core.h +
engine_env.h in a single-lane host emulation (LANES = 1) and is compared with the fp64 oracle.  This checks
layouts, indexing, solver control flow and the env layer without a GPU; the real kernels are tested by
tests/test_gpu_parity.py on the B200 (-m gpu).  hostsim is test infrastructure, never a product path."""
import numpy as np
import pytest

from cosim_b200.config import make_config, load_tables, RANDOM_NONE, RANDOM_FULL
from cosim_b200.model import build_model
from oracle.oracle import Oracle
from tests.hostsim.hostsim import HostSim

CASES = [("flamingo_p_v3", "rocky_hard"), ("flamingo_light_v1", "flat"), ("w4_p_v2", "stairs_up_hard"), ("humanoid_p_v0", "slope_hard")]


def geometry_gap(co, cg):
    """Largest difference between two contact lists [dist, pos(3), normal(3), geom, cell, mu] of equal length: (depth, normal)."""
    if not len(co):
        return 0.0, 0.0
    return float(np.abs(co[:, 0] - cg[:, 0]).max()), float(np.abs(co[:, 4:7] - cg[:, 4:7]).max())


@pytest.mark.parametrize("robot,terrain", CASES)
def test_reset_and_substeps(robot, terrain):
    """Single sub-steps teacher-forced from the fp64 oracle.  Contact COUNTS, geoms and height-field cells must be identical
    (the engine's bounding-shape culls are conservative, capacity = what the model can generate).  Velocities: when the fp32
    MPR returns the same contact geometry as the fp64 one (depth within 2e-6, normal within 1e-4) the sub-step must agree
    to 1e-3; an fp32 MPR that stops on a different portal of a finely tessellated hull against a 1 cm prism returns a
    different penetration direction (normals off by 0.1 .. 0.9 rad on the w4 stairs case, in the fp32 build of the oracle
    as well), and only those sub-steps may be far off."""
    m = build_model(make_config(robot, terrain, random=RANDOM_NONE))
    N = 4
    o, h = Oracle(m, N, seed=1), HostSim(m, N, seed=1)
    np.testing.assert_allclose(h.reset(), o.reset(), atol=1e-6)
    rng = np.random.default_rng(0)
    errs, same_geo_errs, n_far, n_far_explained = [], [], 0, 0
    cap = m.dim("ncon_max")
    for i in range(5):
        a = rng.uniform(-1, 1, (N, m.dim("nu")))
        o.step(a)
        for s in range(2):
            for k in ("qpos", "qvel", "qacc_warmstart", "torque"):
                h.set(k, o.get(k))
            o.substep(); h.substep()
            assert (o.get("ncon")[:, 0].astype(int) == h.get("counters")[:, 7]).all()
            assert (o.get("ncon_dropped") == 0).all()
            err = np.abs(o.get("qvel") - h.get("qvel")).max(axis=1)
            errs.append(err)
            cg_all = h.get("contacts").reshape(N, cap, 10)
            for e in range(N):
                co = o.contacts(e, cap); cg = cg_all[e, :len(co)]
                assert (co[:, 7].astype(int) == cg[:, 7].astype(int)).all() and (co[:, 8].astype(int) == cg[:, 8].astype(int)).all()
                dd, dn = geometry_gap(co, cg)
                if dd < 2e-6 and dn < 1e-4:
                    same_geo_errs.append(err[e])
                if err[e] > 1e-2:
                    n_far += 1; n_far_explained += int(dd > 2e-6 or dn > 1e-4)
    errs = np.concatenate(errs); same_geo_errs = np.array(same_geo_errs)
    print(f"{robot}/{terrain}: qvel error histogram (decades 1e-7..1e0):", np.histogram(np.log10(np.maximum(errs, 1e-7)), bins=np.arange(-7, 1.5))[0].tolist(),
          f"; same-geometry sub-steps {len(same_geo_errs)}/{len(errs)}, worst {same_geo_errs.max() if len(same_geo_errs) else 0:.1e}; far off {n_far}, explained by MPR geometry {n_far_explained}")
    assert np.median(errs) < 2e-4
    assert len(same_geo_errs) >= 0.3 * len(errs) and same_geo_errs.max() < 1e-3
    assert n_far == n_far_explained, "a sub-step is far off although the contact geometry agrees"


def test_env_layer_state_and_flags():
    """Whole control steps incl. PD, delay, obs build, frequency gating, stack, command slots, truncation."""
    et, _ = load_tables()
    cfg = make_config("flamingo_p_v3", "rocky_hard", random=dict(RANDOM_NONE, action_delay_prob=0.3, init_noise=0.05),
                      non_stacked_obs_order=list(et["flamingo_p_v3"]["non_stacked_obs_order"]) + ["height_map"], max_duration=0.2)
    cfg["observation"]["dof_vel"]["freq"] = 25          # refresh every 2nd control step
    m = build_model(cfg)
    N = 3
    o, h = Oracle(m, N, seed=4), HostSim(m, N, seed=4)
    cmd = np.array([[0.5, 0.1, -0.2, 0.0]] * N)
    np.testing.assert_allclose(h.reset(command=cmd), o.reset(command=cmd), atol=1e-6)
    rng = np.random.default_rng(2)
    for i in range(10):
        a = rng.uniform(-1, 1, (N, 8))
        for k in ("qpos", "qvel", "qacc_warmstart"):
            h.set(k, o.get(k))
        so, to, tro = o.step(a, cmd)
        sh, th, trh = h.step(a, cmd)
        np.testing.assert_allclose(h.get("torque"), o.get("torque"), atol=2e-4)
        assert (tro == trh).all() and tro.all() == (i == 9)
        ok = np.abs(o.get("qvel") - h.get("qvel")).max(axis=1) < 1e-3
        np.testing.assert_allclose(sh[ok], so[ok], atol=2e-3)
        assert (sh[:, 84:88] == cmd.astype(np.float32)).all()          # command slots carry the applied command


def test_randomization_draws_match_fp32_oracle():
    m = build_model(make_config("humanoid_p_v0", "slope_hard", random=RANDOM_FULL))
    N = 6
    f, h = Oracle(m, N, seed=0xC051, use_float=True), HostSim(m, N, seed=0xC051)
    for g, k in [("body_mass", "body_mass"), ("frictionloss", "dof_frictionloss"), ("kp", "kp"), ("kd", "kd")]:
        assert (h.get(g) == f.get(k).astype(np.float32)).all(), g
    np.testing.assert_allclose(h.get("invweight_dof"), f.get("dof_invweight0"), rtol=5e-4)
    from oracle.oracle import philox
    assert h.philox(3, 2, 17, 5) == philox(0xC051, 3, 2, 17, 5)


def test_push_event():
    m = build_model(make_config("flamingo_p_v3", "flat", random=RANDOM_NONE))
    h = HostSim(m, 2, seed=1)
    h.reset()
    q = h.get("qpos"); q[:, 3:7] = [np.cos(0.4), 0, 0, np.sin(0.4)]     # yaw 0.8 rad
    h.set("qpos", q)
    h.push([1.0, 0.0, 0.3])
    v = h.get("qvel")[0, :3]
    np.testing.assert_allclose(v, [np.cos(0.8), -np.sin(0.8), 0.3], atol=1e-6)   # robot-frame xy, world z (quirk C-12)


def test_far_from_origin_contacts():
    """+-130 m from the world origin fp32 world coordinates resolve ~1e-5 m; the engine runs hull / prism queries in a
    local frame, so contact counts still match the fp64 oracle and depths agree to ~1e-6."""
    m = build_model(make_config("flamingo_p_v3", "rocky_hard", random=RANDOM_NONE))
    N = 8
    o, h = Oracle(m, N, seed=1), HostSim(m, N, seed=1)
    o.reset(); h.reset()
    rng = np.random.default_rng(3)
    q = o.get("qpos")
    q[:, 0] = rng.uniform(-130, 130, N); q[:, 1] = rng.uniform(-130, 130, N); q[:, 2] += 0.3
    o.set("qpos", q); h.set("qpos", q)
    depth_err, ncontacts = [], 0
    for i in range(25):
        o.step(rng.uniform(-1, 1, (N, 8)))
        for k in ("qpos", "qvel", "qacc_warmstart", "torque"):
            h.set(k, o.get(k))
        o.substep(); h.substep()
        nco, nch = o.get("ncon")[:, 0].astype(int), h.get("counters")[:, 7]
        assert (nco == nch).all()
        for e in range(N):
            if nco[e]:
                co = o.contacts(e); ch = h.get("contacts")[e].reshape(-1, 10)[:len(co)]
                assert (co[:, 8].astype(int) == ch[:, 8].astype(int)).all()
                depth_err.append(np.abs(co[:, 0] - ch[:, 0])); ncontacts += len(co)
    depth_err = np.concatenate(depth_err)
    assert ncontacts > 50 and np.median(depth_err) < 5e-6 and (depth_err > 5e-5).mean() < 0.05


def test_self_collision_contacts_match_oracle():
    """Geom-geom contacts (humanoid limbs pressed into each other / into the torso by random joint offsets): same pairs in
    the same order; depths, normals (after mjc_fixNormal) and the accelerations from the two-body rows within fp32 of the oracle.

    MPR on flat-faced primitives (cylinder caps, boxes) has tied supports: in fp32 the portal may settle on a neighbouring
    face.  The oracle's own fp32 build shows the same outliers against its fp64 build, so contact geometry is compared with
    fp64 statistically (as for terrain contacts in test_gpu_parity.py) and the resulting velocities with the fp32 build."""
    m = build_model(make_config("humanoid_p_v0", "slope_hard", random=RANDOM_NONE))
    N = 16
    o, f, h = Oracle(m, N, seed=1), Oracle(m, N, seed=1, use_float=True), HostSim(m, N, seed=1)
    o.reset(); f.reset(); h.reset()
    rng = np.random.default_rng(5)
    q = o.get("qpos")
    q[:, 7:] += rng.uniform(-0.45, 0.45, q[:, 7:].shape)
    q[:, 2] += 0.05
    for x in (o, f, h):
        x.set("qpos", q)
    nself, derr, nerr, verr = 0, [], [], []
    for i in range(6):
        o.substep(); f.substep(); h.substep()
        nco, nch = o.get("ncon")[:, 0].astype(int), h.get("counters")[:, 7]
        assert (nco == nch).mean() >= 0.9
        for e in np.nonzero(nco == nch)[0]:
            co = o.contacts(int(e)); ch = h.get("contacts")[e].reshape(-1, 10)[:len(co)]
            if not (co[:, 8].astype(int) == ch[:, 8].astype(int)).all():
                continue
            assert (co[:, 7].astype(int) == ch[:, 7].astype(int)).all()
            s = co[:, 8] <= -2
            nself += int(s.sum())
            if s.any():
                derr.append(np.abs(ch[s, 0] - co[s, 0])); nerr.append(np.abs(ch[s, 4:7] - co[s, 4:7]).max(axis=1))
        verr.append(np.abs(h.get("qvel") - f.get("qvel")).max(axis=1))
        for k in ("qpos", "qvel", "qacc_warmstart"):
            h.set(k, o.get(k)); f.set(k, o.get(k))
    derr, nerr, verr = np.concatenate(derr), np.concatenate(nerr), np.concatenate(verr)
    assert nself >= 60 and np.median(derr) < 5e-6 and (derr > 1e-4).mean() <= 0.10
    assert np.median(nerr) < 1e-4 and (nerr > 1e-2).mean() <= 0.10
    assert np.median(verr) < 1e-3 and (verr > 1e-2).mean() <= 0.10


HUMANOID_BOXBOX_POSE = [0.0, 0.0, 4.105, 1.0, 0.0, 0.0, 0.0, 0.817, -0.049, -0.867, -0.311, -0.77, -0.341, 0.073, -0.736, -0.746, 0.097, 0.892, -0.46,
                        -0.272, 0.541, 0.742, 0.66, 0.686, -0.233, 0.501, -0.851, 0.43, -0.65, -0.057]


def test_box_box_contacts_match_oracle():
    """mjc_BoxBox in the engine (engine_core.h box_box, multi-contact append in collide_pairs) against the oracle: a humanoid pose
    (mid-air) in which its box geoms press into each other gives several contacts per box pair; same contacts in the same order,
    positions / depths to fp32, and the accelerations from the multi-contact rows agree."""
    m = build_model(make_config("humanoid_p_v0", "slope_hard", random=RANDOM_NONE))
    gt = m.sections["geom_type"]
    N = 2
    o, h = Oracle(m, N, seed=1), HostSim(m, N, seed=1)
    o.reset(); h.reset()
    q = np.tile(np.array(HUMANOID_BOXBOX_POSE), (N, 1)); q[1, 7:] *= 0.97
    for x in (o, h):
        x.set("qpos", q); x.set("qvel", np.zeros((N, m.dim("nv")))); x.set("qacc_warmstart", np.zeros((N, m.dim("nv"))))
    o.substep(); h.substep()
    nco, nch = o.get("ncon")[:, 0].astype(int), h.get("counters")[:, 7]
    assert (nco == nch).all() and nco.min() >= 4
    nbb = 0
    for e in range(N):
        co = o.contacts(e); ch = h.get("contacts")[e].reshape(-1, 10)[:len(co)]
        assert (co[:, 7].astype(int) == ch[:, 7].astype(int)).all() and (co[:, 8].astype(int) == ch[:, 8].astype(int)).all()
        bb = np.array([gt[int(a)] == 6 and c <= -2 and gt[int(-2 - c)] == 6 for a, c in zip(co[:, 7], co[:, 8])])
        nbb += int(bb.sum())
        np.testing.assert_allclose(ch[bb, 0], co[bb, 0], atol=2e-6); np.testing.assert_allclose(ch[bb, 1:4], co[bb, 1:4], atol=5e-6)
        np.testing.assert_allclose(ch[bb, 4:7], co[bb, 4:7], atol=1e-5)
    assert nbb >= 6, "the pose no longer produces multi-contact box pairs"
    np.testing.assert_allclose(h.get("qvel"), o.get("qvel"), atol=2e-3, rtol=2e-3)      # deep interpenetration: velocities up to 12 rad/s after one sub-step


@pytest.mark.parametrize("robot,terrain", [("flamingo_p_v3", "rocky_hard"), ("flamingo_light_v1", "flat"), ("w4_p_v2", "stairs_up_hard"), ("humanoid_p_v0", "slope_hard")])
def test_tree_sparse_cholesky_against_numpy(robot, terrain):
    """engine_core.h chol_factor / chol_solve (leaf-first elimination, [upstream mj_factorM / mj_solveM]): on a random SPD matrix with
    the sparsity of the robot's kinematic tree the tree-sparse pair lists and the dense variant give numpy's solution; on a full
    SPD matrix (a Hessian coupled across branches) the dense variant does."""
    m = build_model(make_config(robot, terrain, random=RANDOM_NONE))
    nv, dp = m.dim("nv"), m.sections["dof_parent"]
    h = HostSim(m, 1, seed=1)
    rng = np.random.default_rng(5)
    L = np.eye(nv)
    for i in range(nv):
        j = dp[i]
        while j >= 0:
            L[i, j] = rng.uniform(-0.6, 0.6); j = dp[j]
    D = rng.uniform(0.05, 3.0, nv)
    A = L.T @ np.diag(D) @ L                    # M = L^T D L has the tree pattern: (a, b) nonzero only on a common root path
    pattern = np.zeros((nv, nv), bool)
    for i in range(nv):
        j = i
        while j >= 0:
            pattern[i, j] = pattern[j, i] = True; j = dp[j]
    assert not np.any(A[~pattern] != 0.0)
    for trial in range(4):
        b = rng.normal(size=nv)
        ref = np.linalg.solve(A, b)
        scale = np.abs(ref).max()
        for sparse in (1, 0):
            x = h.chol_solve(A, b, sparse)
            assert np.abs(x - ref).max() < 2e-4 * scale * np.linalg.cond(A) ** 0.5, (robot, sparse)
            assert np.abs(A @ x - b).max() < 1e-4 * max(1.0, np.abs(b).max()) * nv
    G = rng.normal(size=(nv, nv)); F = A + 0.3 * G @ G.T          # fill outside the tree pattern: dense variant only
    b = rng.normal(size=nv)
    x = h.chol_solve(F, b, 0)
    assert np.abs(F @ x - b).max() < 1e-4 * max(1.0, np.abs(b).max()) * nv


def test_spawn_spread_reset():
    """engine.spawn_spread (not in the reference, off by default): per-env spawn spots spread over the terrain, lifted by the highest
    terrain vertex within spawn_radius.  Engine logic == oracle, spots differ per env and stay inside the spread, nothing spawns inside
    the terrain, and spread 0 is the reference's spawn at the origin."""
    et, _ = load_tables()
    kw = dict(non_stacked_obs_order=list(et["flamingo_p_v3"]["non_stacked_obs_order"]) + ["height_map"])
    m0 = build_model(make_config("flamingo_p_v3", "rocky_hard", random=RANDOM_NONE, **kw))
    m = build_model(make_config("flamingo_p_v3", "rocky_hard", random=RANDOM_NONE, engine={"spawn_spread": 25.0}, **kw))
    N = 16
    o, h, o0 = Oracle(m, N, seed=7), HostSim(m, N, seed=7), Oracle(m0, N, seed=7)
    s_o, s_h, s_0 = o.reset(), h.reset(), o0.reset()
    np.testing.assert_allclose(s_h, s_o, atol=2e-5)          # height-map rays 25 m from the origin: fp32 world coordinates resolve 2e-6 m
    q, q0 = o.get("qpos"), o0.get("qpos")
    np.testing.assert_allclose(h.get("qpos"), q, atol=1e-6)
    assert np.abs(q0[:, :2]).max() == 0.0 and np.abs(q[:, :2]).max() <= 25.0
    assert len(np.unique(np.round(q[:, 0], 3))) == N and np.abs(q[:, :2]).max() > 5.0
    assert (q[:, 2] >= q0[:, 2] - 1e-9).all() and (q[:, 2] > q0[:, 2] + 1e-3).any()        # lifted by the local terrain height
    cap = m.dim("ncon_max")
    for e in range(N):                                                                         # nothing starts inside the terrain
        c = o.contacts(e, cap)
        assert len(c) == 0 or c[:, 0].min() > -5e-3
    assert not np.allclose(s_o, s_0)                                                           # the height map sees other terrain
    a = np.zeros((N, m.dim("nu")))
    for _ in range(3):
        for k in ("qpos", "qvel", "qacc_warmstart"):
            h.set(k, o.get(k))
        so, to, _ = o.step(a); sh, th, _ = h.step(a)
        assert np.isfinite(sh).all() and (o.get("ncon_dropped") == 0).all()
    # a second reset of the same env draws a new spot (the draw is keyed by the reset counter)
    o.reset()
    assert np.abs(o.get("qpos")[:, :2] - q[:, :2]).max() > 1.0


@pytest.mark.parametrize("robot,terrain", [("flamingo_p_v3", "rocky_hard"), ("w4_p_v2", "stairs_up_hard")])
def test_terrain_pass_instances_agree(robot, terrain, monkeypatch):
    """collide_hfield_all<false> (coarse rasters: no bounding-shape culls, no block-wise sweep) and <true> only differ in the work they
    skip: forced onto the same model (COSIM_HF_FINE) they must produce the same contact lists and the same states, bit for bit."""
    m = build_model(make_config(robot, terrain, random=RANDOM_NONE))
    N = 4
    sims = []
    for fine in ("0", "1"):
        monkeypatch.setenv("COSIM_HF_FINE", fine)
        sims.append(HostSim(m, N, seed=1))
    monkeypatch.delenv("COSIM_HF_FINE")
    o = Oracle(m, N, seed=1)
    o.reset()
    for h in sims:
        h.reset()
    rng = np.random.default_rng(5)
    cap = m.dim("ncon_max")
    ncon_seen = 0
    for i in range(4):
        a = rng.uniform(-1, 1, (N, m.dim("nu")))
        o.step(a)
        for h in sims:
            for k in ("qpos", "qvel", "qacc_warmstart", "torque"):
                h.set(k, o.get(k))
            h.substep()
        c0, c1 = sims[0].get("contacts").reshape(N, cap, 10), sims[1].get("contacts").reshape(N, cap, 10)
        n0, n1 = sims[0].get("counters")[:, 7], sims[1].get("counters")[:, 7]
        assert (n0 == n1).all()
        for e in range(N):
            assert (c0[e, :n0[e]] == c1[e, :n1[e]]).all()
        assert (sims[0].get("qvel") == sims[1].get("qvel")).all() and (sims[0].get("qpos") == sims[1].get("qpos")).all()
        ncon_seen += int(n0.sum())
    assert ncon_seen > 0

"""CPU tests of the friction-cone / solver options beyond the reference's MJCF defaults (condim 1 / 4 / 6, elliptic cone, PGS):
properties that hold exactly for the convex problem MuJoCo states, checked on the oracle.  The reference's four robots use
condim 3 / pyramidal / Newton (e.g. /root/reference/envs/flamingo_p_v3/assets/xml/flamingo_p_v3.xml:3,32); the torsional and
rolling coefficients it randomizes (envs/flamingo_p_v3/manager/xml_manager.py:57-75) only act through these options."""
import numpy as np
import pytest

from cosim_b200.config import make_config, RANDOM_NONE
from cosim_b200.model import build_model
from oracle.oracle import Oracle


def _model(robot, terrain, **eng):
    return build_model(make_config(robot, terrain, random=RANDOM_NONE, engine=eng))


def _settled(m, steps=60, seed=3, actions=0.0):
    o = Oracle(m, 1, seed=seed)
    o.reset()
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        o.step(actions * rng.uniform(-1, 1, (1, m.dim("nu"))))
    return o


def _kkt(o, m):
    nv = m.dim("nv")
    M = o.get("M")[0].reshape(nv, nv)
    r = M @ (o.get("qacc")[0] - o.get("qacc_smooth")[0]) - o.get("qfrc_constraint")[0]
    return np.linalg.norm(r) / (o.get("meaninertia")[0, 0] * nv)


@pytest.mark.parametrize("condim", [3, 4, 6])
def test_elliptic_forces_lie_in_the_cone(condim):
    """Primal solution with the elliptic cone: every contact force satisfies f_n >= 0 and sum (f_j / friction_j)^2 <= f_n^2
    (on the boundary while sliding), and the stationarity condition M (qacc - qacc_smooth) = J^T f holds."""
    m = _model("flamingo_p_v3", "rocky_hard", cone="elliptic", condim=condim)
    o = _settled(m, 25, actions=1.0)
    sliding = 0
    for k in range(10):
        o.step(np.random.default_rng(k).uniform(-1, 1, (1, m.dim("nu"))))
        assert _kkt(o, m) < 1e-6
        cf = o.contact_forces(0)
        for c in cf:
            fn, fr, dim = c[0], c[6:11], int(c[11])
            assert dim == condim and fn >= -1e-12
            t = np.sqrt(sum((c[j] / fr[j - 1]) ** 2 for j in range(1, dim)))
            assert t <= fn * (1 + 1e-7) + 1e-9, f"friction outside the ellipse: {t} vs {fn}"
            sliding += fn > 1e-3 and t > 0.999 * fn
    assert sliding > 0, "no contact ever reached the cone boundary (test would be vacuous)"


@pytest.mark.parametrize("condim", [4, 6])
def test_pyramidal_forces_lie_in_the_pyramid(condim):
    """Pyramidal cone with torsional / rolling edges: sum_k |f_k| / friction_k <= f_n for every contact."""
    m = _model("flamingo_p_v3", "rocky_hard", condim=condim)
    o = _settled(m, 25, actions=1.0)
    used_extra = 0
    for k in range(10):
        o.step(np.random.default_rng(k).uniform(-1, 1, (1, m.dim("nu"))))
        assert _kkt(o, m) < 1e-6
        for c in o.contact_forces(0):
            fn, fr, dim = c[0], c[6:11], int(c[11])
            s = sum(abs(c[j]) / fr[j - 1] for j in range(1, dim))
            assert fn >= -1e-12 and s <= fn * (1 + 1e-9) + 1e-9
            used_extra += abs(c[3]) > 1e-9
    assert used_extra > 0, "the torsional edge never carried force"


def test_condim3_default_path_is_unchanged():
    """engine.condim = 3 / pyramidal / newton given explicitly is the default model (same blob options, same trajectory)."""
    a = _settled(_model("flamingo_light_v1", "flat"), 20, actions=1.0)
    b = _settled(_model("flamingo_light_v1", "flat", condim=3, cone="pyramidal", solver="newton", impratio=1.0), 20, actions=1.0)
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))


def test_condim1_is_frictionless():
    """condim 1: contacts carry no tangential force, so a robot sliding over the plane keeps its horizontal momentum."""
    m = _model("flamingo_light_v1", "flat", condim=1)
    o = _settled(m, 80)
    nv = m.dim("nv")
    v = o.get("qvel"); v[0, 0] = 0.7; v[0, 1] = -0.3; o.set("qvel", v)
    o.forward()
    p0 = (o.get("M")[0].reshape(nv, nv)[:2] @ o.get("qvel")[0]).copy()
    for _ in range(20):
        o.substep()
    o.forward()
    assert o.get("ncon")[0, 0] > 0
    for c in o.contact_forces(0):
        assert int(c[11]) == 1 and np.all(c[1:6] == 0)
    p1 = o.get("M")[0].reshape(nv, nv)[:2] @ o.get("qvel")[0]
    np.testing.assert_allclose(p1, p0, rtol=0, atol=1e-9)


def test_torsional_friction_opposes_spin():
    """condim 4, elliptic: a robot spinning about the vertical on the plane gets a torsional contact moment that opposes the spin
    and is bounded by friction_torsion * f_n; with condim 3 there is none and the spin about the vertical is not braked by it."""
    res = {}
    for condim in (3, 4):
        m = _model("flamingo_light_v1", "flat", cone="elliptic", condim=condim)
        o = _settled(m, 80)
        v = np.zeros_like(o.get("qvel")); v[0, 5] = 3.0; o.set("qvel", v)      # yaw rate of the floating base
        o.forward()
        cf = o.contact_forces(0)
        assert len(cf) > 0
        res[condim] = cf
    assert np.all(res[3][:, 3] == 0)
    tor, fn, mu_t = res[4][:, 3], res[4][:, 0], res[4][:, 8]
    assert np.all(np.abs(tor) <= mu_t * fn * (1 + 1e-7) + 1e-12)
    load = fn > 1e-3
    assert load.any() and np.all(tor[load] * 3.0 < 0) or np.all(tor[load] * 3.0 > 0)     # one sign for all loaded contacts ...
    # ... and that sign brakes: the contact normal is +z, a positive yaw rate must meet a negative moment about the normal
    assert np.all(tor[load] < 0)


def test_rolling_friction_needs_condim6():
    m = _model("flamingo_light_v1", "flat", cone="elliptic", condim=6)
    o = _settled(m, 80)
    v = np.zeros_like(o.get("qvel")); v[0, 0] = 1.0; o.set("qvel", v)
    o.forward()
    cf = o.contact_forces(0)
    fn, roll, mu_r = cf[:, 0], cf[:, 4:6], cf[:, 9]
    assert np.all(np.linalg.norm(roll, axis=1) <= mu_r * fn * (1 + 1e-7) + 1e-12)
    assert np.abs(roll).max() > 0


@pytest.mark.parametrize("cone,condim", [("pyramidal", 3), ("pyramidal", 6), ("elliptic", 3), ("elliptic", 4)])
def test_pgs_reaches_the_newton_solution(cone, condim):
    """PGS (dual) and Newton (primal) minimise the same convex cost: same accelerations and contact forces."""
    mn = _model("flamingo_p_v3", "rocky_hard", cone=cone, condim=condim)
    mp = _model("flamingo_p_v3", "rocky_hard", cone=cone, condim=condim, solver="pgs", iterations=4000)
    on = _settled(mn, 30, actions=1.0)
    op = Oracle(mp, 1, seed=3); op.reset()
    for k in ("qpos", "qvel", "qacc_warmstart", "ctrl"):
        op.set(k, on.get(k))
    on.forward(); op.forward()
    assert on.get("ncon")[0, 0] == op.get("ncon")[0, 0] > 0
    scale = np.abs(on.get("qacc")).max()
    np.testing.assert_allclose(op.get("qacc"), on.get("qacc"), atol=2e-4 * scale)
    assert _kkt(op, mp) < 1e-9          # PGS builds qacc from its forces: stationarity holds by construction
    fn_n, fn_p = on.contact_forces(0)[:, 0], op.contact_forces(0)[:, 0]
    np.testing.assert_allclose(fn_p, fn_n, atol=2e-3 * max(1.0, np.abs(fn_n).max()))


@pytest.mark.parametrize("eng", [dict(cone="elliptic"), dict(solver="pgs", iterations=500), dict(cone="elliptic", condim=6, impratio=4.0)])
def test_resting_contact_supports_weight_general(eng):
    m = _model("flamingo_light_v1", "flat", **eng)
    o = _settled(m, 150)
    f = o.get("cfrc_ext")[0].reshape(-1, 6)
    mg = m.sections["body_mass"].sum() * abs(m.opt("gz"))
    assert abs(f[:, 5].sum() - mg) / mg < 0.02
    assert abs(o.get("qvel")[0, :3]).max() < 0.05


# ---- the engine's general constraint path (cosim_b200/csrc/engine_general.h) in the single-lane host emulation against the oracle
GENERAL = [dict(condim=1), dict(condim=4), dict(condim=6), dict(cone="elliptic"), dict(cone="elliptic", condim=4),
           dict(cone="elliptic", condim=6, impratio=3.0), dict(solver="pgs", iterations=200), dict(solver="pgs", cone="elliptic", condim=4, iterations=200)]


@pytest.mark.parametrize("eng", GENERAL, ids=lambda e: "-".join(f"{k}{v}" for k, v in e.items()))
@pytest.mark.parametrize("robot,terrain", [("flamingo_p_v3", "rocky_hard"), ("flamingo_light_v1", "flat")])
def test_engine_general_path_matches_oracle(robot, terrain, eng):
    """Teacher-forced sub-steps: the engine's general rows / cone / PGS code (fp32) against the fp64 oracle with the same options."""
    from tests.hostsim.hostsim import HostSim
    m = _model(robot, terrain, **eng)
    N = 3
    o, h = Oracle(m, N, seed=2), HostSim(m, N, seed=2)
    np.testing.assert_allclose(h.reset(), o.reset(), atol=2e-5)
    rng = np.random.default_rng(1)
    errs, same_geo, ncon_seen, cap = [], [], 0, m.dim("ncon_max")
    from tests.test_hostsim_vs_oracle import geometry_gap
    for i in range(6):
        o.step(rng.uniform(-1, 1, (N, m.dim("nu"))))
        for s in range(2):
            for k in ("qpos", "qvel", "qacc_warmstart", "torque"):
                h.set(k, o.get(k))
            o.substep(); h.substep()
            assert (o.get("ncon")[:, 0].astype(int) == h.get("counters")[:, 7]).all()
            ncon_seen += int(o.get("ncon").sum())
            err = np.abs(o.get("qvel") - h.get("qvel")).max(axis=1)
            errs.append(err)
            cg_all = h.get("contacts").reshape(N, cap, 10)
            for e in range(N):          # a sub-step may be far off only where the fp32 MPR returned another contact geometry (as in test_hostsim_vs_oracle)
                co = o.contacts(e, cap); dd, dn = geometry_gap(co, cg_all[e, :len(co)])
                if dd < 2e-6 and dn < 1e-4:
                    same_geo.append(err[e])
    errs, same_geo = np.concatenate(errs), np.array(same_geo)
    assert ncon_seen > 0
    tol = 5e-3 if eng.get("solver") == "pgs" else 1e-3          # PGS stops on a cost decrease of 1e-8: its answer is looser in fp32
    assert np.median(errs) < tol / 5 and len(same_geo) >= 0.3 * len(errs) and same_geo.max() < 2 * tol, \
        f"qvel error median {np.median(errs):.1e}, same-geometry sub-steps {len(same_geo)}/{len(errs)} worst {same_geo.max():.1e}"


def test_engine_general_cfrc_ext_and_friction_draws():
    """cfrc_ext (used by the termination test) includes the torsional / rolling moments; the per-env torsional and rolling
    coefficients come from the same counter-based draws as the oracle's randomization."""
    from tests.hostsim.hostsim import HostSim
    from cosim_b200.config import RANDOM_FULL
    m = build_model(make_config("flamingo_light_v1", "flat", random=RANDOM_FULL, engine=dict(cone="elliptic", condim=6)))
    N = 4
    o, h = Oracle(m, N, seed=9), HostSim(m, N, seed=9)
    o.reset(); h.reset()
    rng = np.random.default_rng(4)
    for i in range(8):
        o.step(rng.uniform(-1, 1, (N, m.dim("nu"))))
    v = o.get("qvel"); v[:, 5] += 2.0; o.set("qvel", v)         # spin: torsional rows carry force
    for k in ("qpos", "qvel", "qacc_warmstart", "torque"):
        h.set(k, o.get(k))
    o.substep(); h.substep(); o.rne_post()
    assert (o.get("ncon")[:, 0].astype(int) == h.get("counters")[:, 7]).all() and o.get("ncon").sum() > 0
    cf_o, cf_h = o.get("cfrc_ext"), h.get("cfrc_ext")
    np.testing.assert_allclose(cf_h, cf_o, atol=2e-3 * max(1.0, np.abs(cf_o).max()))

"""Helpers shared by the hostsim (CPU) and GPU parity tests: comparison of contact lists and height-map cells with the
classification of every mismatch (test infrastructure)."""
import numpy as np

# contact list rows: [dist, pos(3), normal(3), geom, cell, mu]
DEPTH_SAME, NORMAL_SAME = 2e-6, 1e-4          # "the fp32 MPR returned the same geometry as the fp64 one"
BORDERLINE_DEPTH = 2e-5                       # a contact present on one side only must be a grazing one


def geometry_gap(co, cg):
    """Largest difference between two contact lists of equal length and identical (geom, cell) columns: (depth, normal)."""
    if not len(co):
        return 0.0, 0.0
    return float(np.abs(co[:, 0] - cg[:, 0]).max()), float(np.abs(co[:, 4:7] - cg[:, 4:7]).max())


def compare_contact_lists(co, cg, geom_type=None):
    """-> (identical keys in identical order, worst |dist| of the UNEXPLAINED contacts present on one side only).
    Keys = (geom, cell).  A contact found by one side only is explained when
      * it is grazing: the query of the other precision returned 'separated' for a penetration below BORDERLINE_DEPTH, or
      * its geom sits at MuJoCo's cap of 50 contacts per geom pair on either side (an earlier grazing contact shifts which
        prisms make it under the cap), or
      * (geom_type given) its geom is a mesh or a box: MPR on flat faces meets tied support vertices, and which of them is
        returned decides whether libccd's portal refinement finds the origin or gives up (the reference's own tie-break is the
        path-dependent hill climb of mjc_support, so no tie-break rule is 'the' reference); such a query can miss a
        centimetre-deep contact on either side.  Counted and bounded by the callers."""
    ko = [(int(r[7]), int(r[8])) for r in co]
    kg = [(int(r[7]), int(r[8])) for r in cg]
    if ko == kg:
        return True, 0.0
    so, sg = set(ko), set(kg)
    cnt = {}
    for k in ko + kg:
        cnt[k[0]] = cnt.get(k[0], 0) + 1
    worst = 0.0
    for rows, other in ((co, sg), (cg, so)):
        n_geom = {}
        for r in rows:
            n_geom[int(r[7])] = n_geom.get(int(r[7]), 0) + 1
        for r in rows:
            g = int(r[7])
            if (g, int(r[8])) in other:
                continue
            capped = max(sum(1 for k in ko if k[0] == g), sum(1 for k in kg if k[0] == g)) >= 50
            flat = geom_type is not None and int(geom_type[g]) in (6, 7)
            if not capped and not flat:
                worst = max(worst, abs(float(r[0])))
    return False, worst


def hm_ray_margin(model, qpos, ray_index):
    """Distance [m] of height-map ray `ray_index` of one env from the nearest height-field cell edge or cell diagonal (where the
    cell / triangle index of mj_rayHfield flips), computed in fp64 from qpos."""
    rx, ry = model.dim("hm_res_x"), model.dim("hm_res_y")
    sxm, sym = model.opt("hm_size_x"), model.opt("hm_size_y")
    i, j = divmod(int(ray_index), rx)
    xr = -sxm * 0.5 + sxm * j / (rx - 1) if rx > 1 else -sxm * 0.5
    yr = -sym * 0.5 + sym * i / (ry - 1) if ry > 1 else -sym * 0.5
    w, x, y, z = qpos[3:7] / np.linalg.norm(qpos[3:7])
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    px, py = qpos[0] + R[0, 0] * xr + R[0, 1] * yr, qpos[1] + R[1, 0] * xr + R[1, 1] * yr
    sx, sy = model.opt("hf_sx"), model.opt("hf_sy")
    ncol, nrow = model.dim("hf_ncol"), model.dim("hf_nrow")
    dx, dy = 2 * sx / (ncol - 1), 2 * sy / (nrow - 1)
    u, v = (px + sx) / dx, (py + sy) / dy
    fu, fv = u - np.floor(u), v - np.floor(v)
    return float(min(min(fu, 1 - fu) * dx, min(fv, 1 - fv) * dy, abs(fu - fv) * min(dx, dy) / np.sqrt(2.0),
                     abs(px - sx), abs(px + sx), abs(py - sy), abs(py + sy)))


def decade_histogram(errs):
    return np.histogram(np.log10(np.maximum(np.asarray(errs, dtype=np.float64), 1e-7)), bins=np.arange(-7, 1.5))[0].tolist()

"""Model builder: (config dict, baked assets) -> flat model blob.

Replaces, for the batched engine, what the reference does by rewriting the MJCF and
letting MuJoCo recompile it per env instance:
  * XMLManager.get_model_path  (/root/reference/envs/flamingo_p_v3/manager/xml_manager.py:16-121):
    terrain select, precision (timestep / iterations), which bodies get mass noise and load,
    which geoms / dofs receive the random friction / frictionloss  -> *flags and ranges*
    in the blob; the per-env draws themselves happen on the device (counter-based RNG).
  * MjModel.from_xml_path (gymnasium MujocoEnv.__init__, flamingo_p_v3.py:94-100):
    hfield PNG normalisation + row flip [upstream MuJoCo, SURVEY.md A.4], qpos0, connect anchors.
  * the per-robot constants of envs/<robot>/<robot>.py (tables in cosim_b200/robots.py).

Blob layout ("CSB1"): u32 magic, u32 nsections, then nsections directory entries
{char name[24]; u32 dtype (0=i32,1=f32,2=f64); u32 pad; u64 count; u64 offset}, then 16-byte
aligned payloads.  `include/cosim_blob.h` is the C view of the same contract.
"""
import math
import os
import struct

import numpy as np

from .robots import ROBOTS, OBS_DIMS, POS, VEL

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

# ---- dims / opts indices (keep in sync with include/cosim_blob.h) ----
DIMS = ["nq", "nv", "nu", "nbody", "njnt", "ngeom", "nhullvert", "neq", "ground_type", "hf_nrow", "hf_ncol",
        "frame_skip", "iterations", "ls_iterations", "ccd_iterations", "ncon_max", "hm_res_x", "hm_res_y",
        "state_dim", "stack_size", "stacked_dim", "nonstacked_dim", "command_dim", "n_term_body", "n_dofpos",
        "n_dofvel", "n_initnoise", "max_episode_steps", "lin_vel_f32", "n_sobs", "n_nobs", "cache_dim",
        "n_state_pos", "n_state_vel", "position_command", "nefc_max", "imu_body", "n_massnoise", "base_body",
        "zero_noise", "auto_reset", "nfl", "nlimit_max", "npair", "condim", "cone", "solver"]
OPTS = ["timestep", "gx", "gy", "gz", "tolerance", "ls_tolerance", "ccd_tolerance", "hf_sx", "hf_sy", "hf_sz",
        "hf_base", "z0", "hm_size_x", "hm_size_y", "hm_zmin", "term_threshold", "init_noise",
        "solref0", "solref1", "solimp0", "solimp1", "solimp2", "solimp3", "solimp4",
        "slide_lo", "slide_hi", "tors_lo", "tors_hi", "roll_lo", "roll_hi", "floss_lo", "floss_hi",
        "delay_lo", "delay_hi", "mass_noise", "load_lo", "load_hi", "kp_lo", "kp_hi", "kd_lo", "kd_hi",
        "plane_sx", "plane_sy", "impratio", "spawn_spread", "spawn_radius"]
DIM = {k: i for i, k in enumerate(DIMS)}
OPT = {k: i for i, k in enumerate(OPTS)}

OBS_KINDS = {"dof_pos": 0, "dof_vel": 1, "ang_vel": 2, "lin_vel": 3, "projected_gravity": 4,
             "last_action": 5, "height_map": 6, "command": 7}
NOISE_ORDER = ["dof_pos", "dof_vel", "ang_vel", "lin_vel", "projected_gravity", "height_map"]


def self_collision_pairs(rb):
    """Geom pairs the broad phase of MuJoCo 3.2.7 hands to the narrow phase [upstream mj_collision / filterBodyPair /
    mj_collideGeoms; SURVEY.md section 8f row 3]: geoms of different bodies, excluding (a) bodies welded together, (b)
    parent-child weld pairs unless one of them is the world (filterparent), (c) <contact><exclude> body pairs, and (d)
    pairs whose contype / conaffinity masks do not meet.  Body pairs are visited in (body1, body2) order, and the two geoms
    of a pair are ordered by type as the narrow-phase function table requires.  Body ids are 1-based (0 = world).
    Deviation: geoms that stand in for an STL missing from the reference checkout (inertia-box proxies, geom_proxy = 1) take
    no part in self-collision: the proxy only approximates the link for ground contact, and neighbouring proxies overlap at
    rest (flamingo_p_v3's two hip boxes by 27 mm), which the real meshes do not."""
    gt, gb = rb["geom_type"], rb["geom_body"]
    if "geom_contype" not in rb:
        return np.zeros((0, 2), np.int32)
    par = np.concatenate([[0], rb["body_parent"]]).astype(int)
    dofnum = np.concatenate([[0], rb["body_dofnum"]]).astype(int)
    nb = len(par)
    weld = np.zeros(nb, int)
    for b in range(1, nb):
        weld[b] = b if dofnum[b] > 0 else weld[par[b]]
    excl = {tuple(sorted(x)) for x in np.asarray(rb["exclude_body"]).reshape(-1, 2).tolist()}
    ct, ca = rb["geom_contype"], rb["geom_conaffinity"]
    out = []
    for i in range(len(gt)):
        for j in range(i + 1, len(gt)):
            b1, b2 = int(gb[i]), int(gb[j])
            if b1 == b2:
                continue
            w1, w2 = weld[b1], weld[b2]
            if w1 == w2:
                continue
            if w1 != 0 and w2 != 0 and (w1 == weld[par[w2]] or w2 == weld[par[w1]]):
                continue
            if tuple(sorted((b1, b2))) in excl:
                continue
            if not ((int(ct[i]) & int(ca[j])) or (int(ct[j]) & int(ca[i]))):
                continue
            if int(rb["geom_proxy"][i]) or int(rb["geom_proxy"][j]):
                continue
            g1, g2 = (i, j) if b1 < b2 else (j, i)
            if gt[g1] > gt[g2]:
                g1, g2 = g2, g1
            out.append((min(b1, b2), max(b1, b2), g1, g2))
    out.sort(key=lambda t: (t[0], t[1]))
    return np.array([(t[2], t[3]) for t in out], np.int32).reshape(-1, 2)


def load_robot(robot_id):
    z = np.load(os.path.join(_ASSETS, f"robot_{robot_id}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def load_terrain_raster(name):
    z = np.load(os.path.join(_ASSETS, f"terrain_{name}.npz"), allow_pickle=False)
    if "alias" in z.files:
        z = np.load(os.path.join(_ASSETS, f"terrain_{str(z['alias'])}.npz"), allow_pickle=False)
    return z["raster"]


def load_png_raster(path):
    """8-bit grey raster of a PNG, as MuJoCo reads hfield images [upstream mjCHField::LoadPNG: lodepng decode to LCT_GREY, 8 bit]."""
    from PIL import Image
    im = Image.open(path)
    if im.mode in ("I;16", "I;16B", "I"):
        return (np.asarray(im).astype(np.uint32) >> 8).astype(np.uint8)
    return np.asarray(im.convert("L"), dtype=np.uint8)


def hfield_from_raster(raster_u8):
    """PNG raster -> MuJoCo hfield_data (float32, row 0 = -y edge, normalised to [0,1]).

    [upstream MuJoCo 3.2.7 user_objects: LoadPNG flips rows so image-top is +y; Compile subtracts
    the minimum and divides by the range]  (confidence M)."""
    data = raster_u8[::-1, :].astype(np.float32)
    emin, emax = float(data.min()), float(data.max())
    data = data - np.float32(emin)
    if emax - emin > 1e-15:
        data = data / np.float32(emax - emin)
    return np.ascontiguousarray(data, dtype=np.float32)


# ------------------------------------------------------------------ support maps for convex hulls
SUP_G = 8          # cube-map grid per face: 6 * G * G direction buckets per mesh


def support_bucket(d, G=SUP_G):
    """Direction -> bucket id (cube map).  Must match `support_bucket` in csrc/engine_core.h bit for bit."""
    d = np.asarray(d, dtype=np.float32)
    a = np.abs(d)
    ax = 0 if (a[0] >= a[1] and a[0] >= a[2]) else (1 if a[1] >= a[2] else 2)
    face = 2 * ax + (1 if d[ax] < 0 else 0)
    inv = np.float32(1.0) / np.maximum(a[ax], np.float32(1e-30))
    u = d[(ax + 1) % 3] * inv
    v = d[(ax + 2) % 3] * inv
    iu = int(min(G - 1, max(0, np.floor((u + np.float32(1.0)) * np.float32(0.5 * G)))))
    iv = int(min(G - 1, max(0, np.floor((v + np.float32(1.0)) * np.float32(0.5 * G)))))
    return (face * G + iu) * G + iv


def support_buckets(D, G=SUP_G):
    """Vectorised `support_bucket` for an array of directions [n, 3] (float32 arithmetic, same result)."""
    D = np.asarray(D, dtype=np.float32)
    a = np.abs(D)
    ax = np.where((a[:, 0] >= a[:, 1]) & (a[:, 0] >= a[:, 2]), 0, np.where(a[:, 1] >= a[:, 2], 1, 2))
    r = np.arange(len(D))
    major = D[r, ax]
    face = 2 * ax + (major < 0)
    inv = np.float32(1.0) / np.maximum(a[r, ax], np.float32(1e-30))
    u = D[r, (ax + 1) % 3] * inv
    v = D[r, (ax + 2) % 3] * inv
    iu = np.clip(np.floor((u + np.float32(1.0)) * np.float32(0.5 * G)), 0, G - 1).astype(np.int64)
    iv = np.clip(np.floor((v + np.float32(1.0)) * np.float32(0.5 * G)), 0, G - 1).astype(np.int64)
    return (face * G + iu) * G + iv


def build_support_map(verts, G=SUP_G, depth=7, checks=20000, seed=0):
    """Per direction bucket, the hull vertices that can be the support point for a direction in that bucket.

    The engine's hull support (`support_lane()` in csrc/engine_core.h) scans only the bucket's candidates instead of all
    vertices (696 for a wheel).  The set of directions for which a vertex is the support point is a convex cone (its normal
    cone), and a bucket -- a rectangle on a face of the direction cube -- is the convex cone of its four corner rays: if
    ONE vertex is the arg-max at all four corners it is the arg-max on the whole rectangle.  Every bucket is therefore
    subdivided (quad-tree) until the four corners of a cell agree, to a depth of `depth` levels (cells of 1/128 of a
    bucket, ~0.1 degrees); leaf cells that still disagree contribute their corner arg-max vertices (ties included); a
    verification pass on random directions patches what is left.  A round-1 version sampled 9 x 9 directions per bucket and missed vertices whose normal
    cone is a thin sliver between two samples (humanoid foot mesh: a contact of 1 cm depth went undetected); the fp64 oracle's
    brute-force scan and tests/test_model.py::test_support_map_is_complete pin this construction.
    Returns (offsets[6*G*G + 1], indices)."""
    from scipy.spatial import ConvexHull
    V = np.asarray(verts, dtype=np.float64)
    n = len(V)
    nb = 6 * G * G
    if n <= 32:                                   # tiny hulls (proxy boxes): every bucket lists every vertex
        return np.arange(nb + 1, dtype=np.int32) * n, np.tile(np.arange(n, dtype=np.int32), nb)
    hull = ConvexHull(V)
    nbr = [set() for _ in range(n)]
    for tri in hull.simplices:
        for a in tri:
            nbr[a].update(int(b) for b in tri if b != a)
    cmat = np.zeros((nb, n), dtype=bool)          # candidate matrix [bucket, vertex]

    def argmax_sets(face, uv):
        """For direction samples (u, v) on a cube face: arg-max vertex and the mask of (near-)tied vertices."""
        ax, sgn = face // 2, (-1.0 if face % 2 else 1.0)
        D = np.zeros((len(uv), 3))
        D[:, ax] = sgn; D[:, (ax + 1) % 3] = uv[:, 0]; D[:, (ax + 2) % 3] = uv[:, 1]
        P = D @ V.T
        mx = P.max(axis=1, keepdims=True)
        # near-ties count as ties: the engine scans in float32 and breaks ties towards the lowest vertex index, like the oracle's
        # scan over all vertices -- every vertex that can be the float32 maximum has to be listed (flat mesh faces!)
        return np.argmax(P, axis=1), P >= mx - 1e-6 * (1.0 + np.abs(mx))

    for face in range(6):
        # level-0 cells = the buckets of this face; cells: [bucket, u0, v0, size]
        iu, iv = np.meshgrid(np.arange(G), np.arange(G), indexing="ij")
        cells = np.stack([((face * G + iu) * G + iv).ravel().astype(np.float64), -1.0 + 2.0 * iu.ravel() / G, -1.0 + 2.0 * iv.ravel() / G,
                          np.full(G * G, 2.0 / G)], axis=1)
        for level in range(depth + 1):
            if not len(cells):
                break
            corners = np.concatenate([cells[:, 1:3] + cells[:, 3:4] * np.array(o) for o in ((0, 0), (1, 0), (0, 1), (1, 1))], axis=0)
            am, tie = argmax_sets(face, corners)
            am = am.reshape(4, -1); tie = tie.reshape(4, len(cells), n).any(axis=0)
            agree = (am[0] == am[1]) & (am[0] == am[2]) & (am[0] == am[3])
            done = agree | (level == depth)
            kk, vv = np.nonzero(tie[done])
            cmat[cells[done][kk, 0].astype(np.int64), vv] = True
            split = cells[~done]
            if len(split):
                h = split[:, 3:4] * 0.5
                cells = np.concatenate([np.concatenate([split[:, 0:1], split[:, 1:2] + ou * h, split[:, 2:3] + ov * h, h], axis=1)
                                        for ou, ov in ((0, 0), (1, 0), (0, 1), (1, 1))], axis=0)
            else:
                cells = split
    # Gauss-map vertices: the unit normal of every hull triangle is a direction in which its three corners tie, and around it
    # the normal cones of all the vertices of a (near-)flat face meet in wedges thinner than any cell -- list the corners in
    # the bucket(s) of the normal and of its close neighbourhood (bucket borders, float32 binning)
    tn = hull.equations[:, :3]
    e1 = np.cross(tn, np.array([1.0, 0.0, 0.0])); bad1 = np.linalg.norm(e1, axis=1) < 0.1
    e1[bad1] = np.cross(tn[bad1], np.array([0.0, 1.0, 0.0]))
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(tn, e1)
    for a, b in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (1, -1), (-1, 1), (-1, -1)):
        dirs = (tn + 4e-3 * (a * e1 + b * e2)).astype(np.float32)
        bk = support_buckets(dirs, G)
        for c in range(3):
            cmat[bk, hull.simplices[:, c]] = True
    cand = [set(np.nonzero(cmat[k])[0].tolist()) for k in range(nb)]
    rng = np.random.default_rng(seed)
    for _ in range(8):                            # verify on random directions (float32, as the engine computes the bucket) + patch
        D = rng.standard_normal((checks, 3)).astype(np.float32)
        best = np.argmax(D.astype(np.float64) @ V.T, axis=1)
        bad = 0
        for k, b in zip(support_buckets(D, G), best):
            if int(b) not in cand[int(k)]:
                cand[int(k)].add(int(b)); cand[int(k)].update(nbr[int(b)]); bad += 1
        if bad == 0:
            break
    # a direction on a bucket border may be binned to either side in float32: a bucket also lists what its 8 neighbours on the
    # same face list for their border cells -- covered by the closed corner samples above (borders are shared corners)
    off = np.zeros(nb + 1, dtype=np.int32)
    idx = []
    for k in range(nb):
        lst = sorted(cand[k])
        idx.extend(lst)
        off[k + 1] = len(idx)
    return off, np.array(idx, dtype=np.int32)


# ------------------------------------------------------------------ small numpy FK (qpos0 only)
def _quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def _quat_mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def fk_qpos0(rb, base_pos=None):
    """World pose of every body at the reference pose (joints at 0). Returns (xpos[nb,3], xquat[nb,4])."""
    nb = len(rb["body_mass"])
    xpos = np.zeros((nb + 1, 3))
    xquat = np.zeros((nb + 1, 4))
    xquat[0, 0] = 1
    for i in range(nb):
        p = int(rb["body_parent"][i])
        R = _quat_mat(xquat[p])
        xpos[i + 1] = xpos[p] + R @ rb["body_pos"][i]
        xquat[i + 1] = _quat_mul(xquat[p], rb["body_quat"][i])
        if i == 0 and base_pos is not None:
            xpos[1] = base_pos
    return xpos, xquat


def _ncdf(x):
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def _rng_pair(v):
    if isinstance(v, (list, tuple)):
        return float(v[0]), float(v[1])
    return float(v), float(v)


_SUP_CACHE = {}


def _support_map_cached(verts):
    """Support maps are a function of the hull vertices only: kept in-process and as baked assets
    (cosim_b200/assets/supmaps/<sha1>.npz, committed for the shipped meshes; rebuilt and stored for new ones)."""
    import hashlib
    key = verts.tobytes()
    if key not in _SUP_CACHE:
        tag = hashlib.sha1(key + f"G{SUP_G}d7t6n".encode()).hexdigest()[:20]
        path = os.path.join(_ASSETS, "supmaps", tag + ".npz")
        if os.path.exists(path):
            z = np.load(path, allow_pickle=False)
            _SUP_CACHE[key] = (z["off"], z["idx"])
        else:
            _SUP_CACHE[key] = build_support_map(verts)
            try:
                os.makedirs(os.path.dirname(path), exist_ok=True)
                tmp = path + f".{os.getpid()}.tmp.npz"
                np.savez_compressed(tmp, off=_SUP_CACHE[key][0], idx=_SUP_CACHE[key][1])
                os.replace(tmp, path)
            except OSError:
                pass
    return _SUP_CACHE[key]


class Model:
    """Holds the named sections and the packed blob."""

    def __init__(self, sections, meta):
        self.sections = sections
        self.meta = meta
        self.blob = pack_blob(sections)

    def dim(self, name):
        return int(self.sections["dims"][DIM[name]])

    def opt(self, name):
        return float(self.sections["opts"][OPT[name]])


def pack_blob(sections):
    names = list(sections)
    hdr = 8 + 48 * len(names)
    off = (hdr + 15) // 16 * 16
    entries, payload = [], []
    for n in names:
        a = np.ascontiguousarray(sections[n])
        dt = {np.dtype(np.int32): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}[a.dtype]
        raw = a.tobytes()
        entries.append(struct.pack("<24sIIQQ", n.encode()[:23], dt, 0, a.size, off))
        pad = (-len(raw)) % 16
        payload.append(raw + b"\0" * pad)
        off += len(raw) + pad
    head = struct.pack("<4sI", b"CSB1", len(names)) + b"".join(entries)
    head += b"\0" * ((-len(head)) % 16)
    return head + b"".join(payload)


def build_model(config, ncon_max=None, auto_reset=False):
    env_id = config["env"]["id"]
    if env_id not in ROBOTS:
        raise NameError(f"Please select a valid environment id. Received '{env_id}'.")
    spec = ROBOTS[env_id]
    rb = load_robot(env_id)
    obs_cfg = config["observation"]
    rnd = config["random"]
    hw = config["hardware"]
    rtab = config["random_table"]
    eng = config.get("engine", {}) or {}

    nb = len(rb["body_mass"]) + 1           # incl. world
    njnt = len(rb["jnt_type"])
    nv = len(rb["dof_body"])
    nq = int(sum(7 if t == 0 else 1 for t in rb["jnt_type"]))
    nu = len(rb["act_joint"])
    ngeom = len(rb["geom_type"])
    jname = {str(n): i for i, n in enumerate(rb["jnt_names"])}
    bname = {str(n): i + 1 for i, n in enumerate(rb["body_names"])}

    # ---- precision (xml_manager.py:34-41, flamingo_p_v3.py:48-55)
    prec = rtab["precision"][rnd["precision"]]
    timestep, iterations, frame_skip = float(prec["timestep"]), int(prec["iterations"]), int(prec["frame_skip"])
    if eng.get("iterations") is not None:        # solver iteration budget override (PGS needs many more sweeps than Newton needs iterations)
        iterations = int(eng["iterations"])
    control_freq = 1.0 / (timestep * frame_skip)
    assert control_freq == 50, "Currently, only control frequency of 50 is supported."

    # ---- terrain (xml_manager.py:21-32)
    terrain = config["env"]["terrain"]
    if terrain == "flat":
        ground_type, hf = 0, np.zeros((1, 1), np.float32)
        hf_size = np.zeros(4)
    elif isinstance(terrain, dict):
        # terrain authoring: the user's own height image instead of a baked reference terrain --
        # what adding an <hfield file=... size=.../> asset to the MJCF does in the reference
        ground_type = 1
        raster = load_png_raster(terrain["png"]) if "png" in terrain else np.asarray(terrain["raster"])
        if raster.ndim != 2 or raster.dtype != np.uint8 or min(raster.shape) < 2:
            raise ValueError("terrain raster must be a 2-D uint8 image of at least 2 x 2 pixels")
        hf = hfield_from_raster(raster)
        hf_size = np.asarray(terrain["size"], dtype=np.float64).reshape(4)      # MJCF hfield size: x, y half extents, z top, base
        if not (hf_size > 0).all():
            raise ValueError("terrain size = [radius_x, radius_y, elevation_z, base_z] must be positive")
        terrain = str(terrain.get("name", "custom"))
    else:
        names = [str(n) for n in rb["hfield_names"]]
        if terrain not in names:
            raise ValueError(f"unknown terrain '{terrain}' for {env_id}")
        ground_type = 1
        hf = hfield_from_raster(load_terrain_raster(terrain))
        hf_size = rb["hfield_size"][names.index(terrain)].astype(np.float64)

    # ---- height map
    hm = obs_cfg.get("height_map")
    if hm is not None and ground_type == 0:
        raise ValueError("height_map needs an hfield terrain: mj_rayHfield on a plane geom is a fatal error "
                         "in the reference (SURVEY.md B.11)")
    res_x, res_y = (int(hm["res_x"]), int(hm["res_y"])) if hm is not None else (0, 0)

    # ---- observation layout (wrappers.py:93-121,160-202)
    n_dofpos, n_dofvel = OBS_DIMS[env_id]
    command_dim = int(obs_cfg["command_dim"])
    obs_to_dim = {"dof_pos": n_dofpos, "dof_vel": n_dofvel, "ang_vel": 3, "lin_vel": 3, "projected_gravity": 3,
                  "last_action": nu, "height_map": res_x * res_y, "command": command_dim}
    cache_off, cache_dim = {}, 0

    def items(order):
        nonlocal cache_dim
        kinds, dims, scales, intervals, offs = [], [], [], [], []
        for n in order:
            if n not in obs_to_dim:
                raise KeyError(n)
            kinds.append(OBS_KINDS[n])
            dims.append(obs_to_dim[n])
            if n == "command":
                scales.append(1.0)
                intervals.append(1)
                offs.append(-1)
                continue
            ncfg = obs_cfg[n]
            freq = float(ncfg["freq"])
            if freq <= 0:
                raise ValueError(f"Invalid observation update frequency for '{n}': {freq}. Must be > 0.")
            scales.append(float(ncfg["scale"]))
            intervals.append(max(1, int(round(control_freq / freq))))
            if n not in cache_off:
                cache_off[n] = cache_dim
                cache_dim += obs_to_dim[n]
            offs.append(cache_off[n])
        return (np.array(kinds, np.int32), np.array(dims, np.int32), np.array(scales, np.float64),
                np.array(intervals, np.int32), np.array(offs, np.int32))

    s_kind, s_dim, s_scale, s_int, s_off = items(list(obs_cfg["stacked_obs_order"]))
    n_kind, n_dim, n_scale, n_int, n_off = items(list(obs_cfg["non_stacked_obs_order"]))
    stack_size = int(obs_cfg["stack_size"])
    stacked_dim, nonstacked_dim = int(s_dim.sum()), int(n_dim.sum())
    state_dim = stack_size * stacked_dim + nonstacked_dim

    # ---- actuators / PD tables
    act_dof = np.array([rb["jnt_dofadr"][j] for j in rb["act_joint"]], np.int32)
    act_qadr = np.array([rb["jnt_qposadr"][j] for j in rb["act_joint"]], np.int32)
    act_jnt_name = [str(rb["jnt_names"][j]) for j in rb["act_joint"]]
    kp = np.zeros(nu); kd = np.zeros(nu); sc = np.zeros(nu); posfac = np.ones(nu); gam = np.ones(nu)
    clip = np.zeros(nu); mode = np.zeros(nu, np.int32)
    k = 0
    for g in spec.groups:
        for jn in g.joints:
            assert act_jnt_name[k] == jn, f"actuator order mismatch {act_jnt_name[k]} vs {jn}"
            mode[k] = g.mode
            kp[k] = float(hw[g.kp]) if g.mode == POS else 0.0
            kd[k] = float(hw[g.kd])
            sc[k] = float(hw["action_scales"][g.scale])
            clip[k] = float(hw[g.clip])
            if g.geared:
                posfac[k] = float(hw["gear_ratio"])
                gam[k] = float(hw["gamma"])
            k += 1
    assert k == nu
    gear_of = {jn: float(hw["gear_ratio"]) for jn in spec.geared_joints}

    dofpos_qadr = np.array([rb["jnt_qposadr"][jname[j]] for j in spec.dof_pos_joints], np.int32)
    dofpos_fac = np.array([gear_of.get(j, 1.0) for j in spec.dof_pos_joints])
    dofvel_dadr = np.array([rb["jnt_dofadr"][jname[j]] for j in spec.dof_vel_joints], np.int32)
    dofvel_fac = np.array([gear_of.get(j, 1.0) for j in spec.dof_vel_joints])
    if spec.init_noise_joints:
        initnoise_qadr = np.array([rb["jnt_qposadr"][jname[j]] for j in spec.init_noise_joints], np.int32)
    else:
        initnoise_qadr = np.arange(7, nq, dtype=np.int32)
    state_pos_qadr = np.array([rb["jnt_qposadr"][jname[j]] for j in spec.state_pos_joints], np.int32)
    state_vel_dadr = np.array([rb["jnt_dofadr"][jname[j]] for j in spec.state_vel_joints], np.int32)
    term_body = np.array([bname[b] for b in spec.term_bodies], np.int32)
    massnoise_body = np.array(sorted(bname[b] for b in spec.mass_noise_bodies), np.int32)  # document order

    # ---- randomisation flags (xml_manager.py:57-87)
    wheel_ids = {bname[b] for b in spec.wheel_bodies}
    geom_fr_random = np.array([int(rb["geom_body"][g] in wheel_ids and rb["geom_has_friction_attr"][g])
                               for g in range(ngeom)], np.int32)
    dof_fl_random = np.array([int(str(rb["jnt_class"][rb["dof_jnt"][d]]) in ("joints", "wheels"))
                              for d in range(nv)], np.int32)
    nfl_upper = int(np.sum((rb["dof_frictionloss"] > 0) | (dof_fl_random > 0)))
    nlimit_max = int(np.sum(rb["jnt_limited"]))

    # ---- sensor noise table (random_table.yaml; "zero" = true-zero extension)
    level = rnd["sensor_noise"]
    zero_noise = int(level == "zero")
    noise = np.zeros((6, 6))
    if not zero_noise:
        nm = rtab["sensor_noise"][level]
        for i, n in enumerate(NOISE_ORDER):
            mean, std, lo, hi = (float(nm[n][q]) for q in ("mean", "std", "lower", "upper"))
            a, b = (lo - mean) / std, (hi - mean) / std
            noise[i] = [mean, std, lo, hi, _ncdf(a), _ncdf(b)]

    # ---- qpos0 and connect anchors
    qpos0 = np.zeros(nq)
    qpos0[0:3] = rb["body_pos"][0]
    qpos0[3:7] = rb["body_quat"][0]
    xpos0, xquat0 = fk_qpos0(rb)
    neq = len(rb["eq_body1"])
    eq_anchor2 = np.zeros((neq, 3))
    for e in range(neq):
        b1, b2 = int(rb["eq_body1"][e]), int(rb["eq_body2"][e])
        pw = xpos0[b1] + _quat_mat(xquat0[b1]) @ rb["eq_anchor"][e]
        eq_anchor2[e] = _quat_mat(xquat0[b2]).T @ (pw - xpos0[b2])

    pair_geom = self_collision_pairs(rb) if eng.get("self_collision", True) else np.zeros((0, 2), np.int32)

    # Contact capacity of one env.  MuJoCo has no per-env cap, only mjMAXCONPAIR = 50 contacts per geom pair: the default
    # capacity is what the model can generate (50 per geom against a height field, 5 against the plane, 1 per geom-geom
    # candidate pair), bounded by 1024 records.  The engine keeps the first few records of an env in shared memory and
    # the rest in a global-memory slot (engine_setup.h), so a large capacity costs memory, not occupancy.
    if ncon_max is None:
        ncon_max = eng.get("ncon_max")
    if ncon_max is None:
        n_boxbox = int(sum(1 for a, b in pair_geom if rb["geom_type"][a] == 6 and rb["geom_type"][b] == 6))      # mjc_BoxBox: up to 8 contacts
        ncon_max = min(1024, (50 if ground_type == 1 else 5) * ngeom + len(pair_geom) + 7 * n_boxbox)
    ncon_max = int(ncon_max)
    # Friction-cone / solver options of the general constraint path (MuJoCo <option cone= solver= impratio=>, geom condim).  The
    # reference's four MJCF files all say condim 3 / pyramidal / Newton / impratio 1, which is the specialised fast path.
    condim = int(eng.get("condim", 3))
    if condim not in (1, 3, 4, 6):
        raise ValueError("engine.condim must be 1, 3, 4 or 6")
    cone = {"pyramidal": 0, "elliptic": 1}[str(eng.get("cone", "pyramidal")).lower()]
    solver = {"newton": 0, "pgs": 1}[str(eng.get("solver", "newton")).lower()]
    impratio = float(eng.get("impratio", 1.0))
    if impratio <= 0:
        raise ValueError("engine.impratio must be positive")
    # Not in the reference (it spawns every robot at the origin, on the flat centre patch of its rasters): an optional per-env spawn
    # offset, uniform in [-spawn_spread, spawn_spread]^2 m, lifted by the highest terrain vertex within spawn_radius of the spot
    # (bench.py --spawn-spread: robots that stand on rough cells from the first step on).  0 = the reference's spawn.
    spawn_spread, spawn_radius = float(eng.get("spawn_spread", 0.0)), float(eng.get("spawn_radius", 0.6))
    if spawn_spread < 0 or spawn_radius < 0:
        raise ValueError("engine.spawn_spread / spawn_radius must not be negative")
    if ground_type == 1 and spawn_spread + spawn_radius >= min(hf_size[0], hf_size[1]):
        raise ValueError("engine.spawn_spread reaches beyond the terrain")
    rows_per_contact = 1 if condim == 1 else (condim if cone == 1 else 2 * (condim - 1))
    nefc_max = 3 * neq + nfl_upper + nlimit_max + rows_per_contact * ncon_max

    dims = np.zeros(64, np.int32)
    opts = np.zeros(64, np.float64)

    def setd(**kw):
        for a, b in kw.items():
            dims[DIM[a]] = int(b)

    def seto(**kw):
        for a, b in kw.items():
            opts[OPT[a]] = float(b)

    max_steps = int(config["env"]["max_duration"] * control_freq)
    setd(nq=nq, nv=nv, nu=nu, nbody=nb, njnt=njnt, ngeom=ngeom, nhullvert=len(rb["hull_verts"]), neq=neq,
         ground_type=ground_type, hf_nrow=hf.shape[0], hf_ncol=hf.shape[1], frame_skip=frame_skip,
         iterations=iterations, ls_iterations=50, ccd_iterations=50, ncon_max=ncon_max, hm_res_x=res_x,
         hm_res_y=res_y, state_dim=state_dim, stack_size=stack_size, stacked_dim=stacked_dim,
         nonstacked_dim=nonstacked_dim, command_dim=command_dim, n_term_body=len(term_body), n_dofpos=n_dofpos,
         n_dofvel=n_dofvel, n_initnoise=len(initnoise_qadr), max_episode_steps=max_steps,
         lin_vel_f32=int(spec.lin_vel_f32), n_sobs=len(s_kind), n_nobs=len(n_kind), cache_dim=cache_dim,
         n_state_pos=len(state_pos_qadr), n_state_vel=len(state_vel_dadr),
         position_command=int(bool(config["env"]["position_command"])), nefc_max=nefc_max,
         imu_body=int(rb["imu_body"]), n_massnoise=len(massnoise_body), base_body=bname[spec.base_body],
         zero_noise=zero_noise, auto_reset=int(bool(eng.get("auto_reset", auto_reset))), nfl=nfl_upper,
         nlimit_max=nlimit_max, npair=len(pair_geom), condim=condim, cone=cone, solver=solver)
    g = rb["gravity"]
    sl, tl, rl = _rng_pair(rnd["sliding_friction"]), _rng_pair(rnd["torsional_friction"]), _rng_pair(rnd["rolling_friction"])
    fl, dl, ld = _rng_pair(rnd["friction_loss"]), _rng_pair(rnd["action_delay_prob"]), _rng_pair(rnd["load"])
    kpr, kdr = _rng_pair(rnd.get("kp_scale", 1.0)), _rng_pair(rnd.get("kd_scale", 1.0))
    seto(timestep=timestep, gx=g[0], gy=g[1], gz=g[2], tolerance=1e-8, ls_tolerance=0.01, ccd_tolerance=1e-6,
         hf_sx=hf_size[0], hf_sy=hf_size[1], hf_sz=hf_size[2], hf_base=hf_size[3], z0=spec.z0,
         hm_size_x=(hm["size_x"] if hm else 0.0), hm_size_y=(hm["size_y"] if hm else 0.0), hm_zmin=spec.hm_zmin,
         term_threshold=spec.term_threshold, init_noise=float(rnd["init_noise"]),
         solref0=0.02, solref1=1.0, solimp0=0.9, solimp1=0.95, solimp2=0.001, solimp3=0.5, solimp4=2.0,
         slide_lo=sl[0], slide_hi=sl[1], tors_lo=tl[0], tors_hi=tl[1], roll_lo=rl[0], roll_hi=rl[1],
         floss_lo=fl[0], floss_hi=fl[1], delay_lo=dl[0], delay_hi=dl[1], mass_noise=float(rnd["mass_noise"]),
         load_lo=ld[0], load_hi=ld[1], kp_lo=kpr[0], kp_hi=kpr[1], kd_lo=kdr[0], kd_hi=kdr[1],
         plane_sx=100.0, plane_sy=100.0, impratio=impratio, spawn_spread=spawn_spread, spawn_radius=spawn_radius)

    def w0(a, fill=0.0):  # prepend the world body row
        a = np.asarray(a)
        z = np.full((1,) + a.shape[1:], fill, dtype=a.dtype)
        return np.concatenate([z, a])

    wq = np.zeros((1, 4)); wq[0, 0] = 1.0
    S = {}
    S["dims"], S["opts"] = dims, opts
    S["body_parent"] = w0(rb["body_parent"]).astype(np.int32)
    S["body_pos"] = w0(rb["body_pos"]).astype(np.float64)
    S["body_quat"] = np.concatenate([wq, rb["body_quat"]]).astype(np.float64)
    S["body_mass"] = w0(rb["body_mass"]).astype(np.float64)
    S["body_ipos"] = w0(rb["body_ipos"]).astype(np.float64)
    S["body_inertia"] = w0(rb["body_inertia"]).astype(np.float64)
    S["body_jntadr"] = w0(rb["body_jntadr"]).astype(np.int32)
    S["body_jntnum"] = w0(rb["body_jntnum"]).astype(np.int32)
    S["body_dofadr"] = w0(rb["body_dofadr"]).astype(np.int32)
    S["body_dofnum"] = w0(rb["body_dofnum"]).astype(np.int32)
    for key in ("jnt_type", "jnt_body", "jnt_qposadr", "jnt_dofadr", "jnt_limited", "jnt_actfrclimited",
                "dof_body", "dof_jnt", "dof_parent", "act_ctrllimited", "geom_type", "geom_body", "geom_vadr",
                "geom_vnum"):
        S[key] = rb[key].astype(np.int32)
    for key in ("jnt_pos", "jnt_axis", "jnt_range", "jnt_actfrcrange", "dof_armature", "dof_damping",
                "dof_frictionloss", "act_gear", "act_ctrlrange", "geom_size", "geom_pos", "geom_quat",
                "geom_friction", "geom_center", "geom_rbound"):
        S[key] = rb[key].astype(np.float64)
    S["dof_fl_random"], S["geom_fr_random"] = dof_fl_random, geom_fr_random
    S["qpos0"] = qpos0
    S["act_dof"], S["act_qadr"], S["act_mode"] = act_dof, act_qadr, mode
    S["act_kp"], S["act_kd"], S["act_scale"], S["act_posfac"], S["act_gamma"], S["act_clip"] = kp, kd, sc, posfac, gam, clip
    S["hull_verts"] = rb["hull_verts"].astype(np.float32).reshape(-1)
    # support maps (engine only; the oracle scans all hull vertices): per mesh geom, base of its bucket table
    hv = rb["hull_verts"].astype(np.float32).reshape(-1, 3)
    sup_adr = np.full(ngeom, -1, np.int32)
    sup_off, sup_idx, cache = [np.zeros(1, np.int32)], [], {}
    nidx = 0
    for g in range(ngeom):
        if int(rb["geom_type"][g]) != 7:
            continue
        key = (int(rb["geom_vadr"][g]), int(rb["geom_vnum"][g]))
        if key not in cache:
            off, idx = _support_map_cached(hv[key[0]:key[0] + key[1]])
            cache[key] = sum(len(o) for o in sup_off) - 1
            sup_off.append(off[1:] + nidx)
            sup_idx.append(idx)
            nidx += len(idx)
        sup_adr[g] = cache[key]
    S["geom_supadr"] = sup_adr
    S["pair_geom"] = (pair_geom if len(pair_geom) else np.zeros((1, 2), np.int32)).astype(np.int32).reshape(-1)
    S["sup_off"] = np.concatenate(sup_off).astype(np.int32)
    S["sup_idx"] = (np.concatenate(sup_idx) if sup_idx else np.zeros(1)).astype(np.int32)
    gf = np.concatenate([rb["ground_friction"].astype(np.float64), [float(rb["ground_has_friction_attr"])]])
    S["ground_friction"] = gf
    S["hfield_data"] = hf.reshape(-1)
    S["eq_body1"], S["eq_body2"] = rb["eq_body1"].astype(np.int32), rb["eq_body2"].astype(np.int32)
    S["eq_anchor1"], S["eq_anchor2"] = rb["eq_anchor"].astype(np.float64), eq_anchor2
    S["eq_solref"], S["eq_solimp"] = rb["eq_solref"].astype(np.float64), rb["eq_solimp"].astype(np.float64)
    S["imu_pos"], S["imu_quat"] = rb["imu_pos"].astype(np.float64), rb["imu_quat"].astype(np.float64)
    S["dofpos_qadr"], S["dofpos_fac"] = dofpos_qadr, dofpos_fac
    S["dofvel_dadr"], S["dofvel_fac"] = dofvel_dadr, dofvel_fac
    S["initnoise_qadr"], S["term_body"] = initnoise_qadr, term_body
    S["state_pos_qadr"], S["state_vel_dadr"], S["massnoise_body"] = state_pos_qadr, state_vel_dadr, massnoise_body
    S["sobs_kind"], S["sobs_dim"], S["sobs_scale"], S["sobs_interval"], S["sobs_off"] = s_kind, s_dim, s_scale, s_int, s_off
    S["nobs_kind"], S["nobs_dim"], S["nobs_scale"], S["nobs_interval"], S["nobs_off"] = n_kind, n_dim, n_scale, n_int, n_off
    S["noise"] = noise.reshape(-1)
    S["command_scales"] = np.array([float(obs_cfg["command_scales"][str(i)]) for i in range(command_dim)], np.float64)
    # empty arrays still need a dtype the packer knows
    for k_, v_ in list(S.items()):
        if v_.dtype not in (np.int32, np.float32, np.float64):
            S[k_] = v_.astype(np.float64)

    meta = dict(robot=env_id, terrain=terrain, nq=nq, nv=nv, nu=nu, nbody=nb, ngeom=ngeom, state_dim=state_dim,
                command_dim=command_dim, control_freq=control_freq, dt=timestep * frame_skip,
                frame_skip=frame_skip, max_episode_steps=max_steps, body_names=["world"] + [str(n) for n in rb["body_names"]],
                jnt_names=[str(n) for n in rb["jnt_names"]], geom_names=[str(n) for n in rb["geom_names"]],
                obs_to_dim=obs_to_dim, action_scaler=sc.copy(), notes=[str(n) for n in rb["notes"]],
                cmd_slices=_cmd_slices(obs_cfg, obs_to_dim, stack_size, stacked_dim, command_dim))
    return Model(S, meta)


def _cmd_slices(obs_cfg, obs_to_dim, stack_size, stacked_dim, command_dim):
    """StateBuildWrapper._get_cmd_index_cache, /root/reference/envs/wrappers.py:123-158."""
    out = []
    if command_dim <= 0:
        return out
    off, starts = 0, []
    for n in obs_cfg["stacked_obs_order"]:
        if n == "command":
            starts.append(off)
        off += obs_to_dim[n]
    for k in range(stack_size):
        for s in starts:
            out.append(slice(k * stacked_dim + s, k * stacked_dim + s + command_dim))
    base, off = stack_size * stacked_dim, 0
    for n in obs_cfg["non_stacked_obs_order"]:
        if n == "command":
            out.append(slice(base + off, base + off + command_dim))
        off += obs_to_dim[n]
    return out

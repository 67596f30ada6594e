#!/usr/bin/env python3
"""Asset baker: reference MJCF / STL / PNG  ->  cosim_b200/assets/*.npz

Runs ONLY in the build container (it reads /root/reference, which does not exist
on the GPU box).  Its outputs are committed, so nothing at run time touches the
reference tree.  No reference *source* is copied: only model data (numbers parsed
from the MJCF files, convex-hull vertices of the STL collision meshes, terrain
rasters) is extracted.

What it restates of MuJoCo's model compiler (un-vendored `mujoco==3.2.7`,
/root/reference/requirements.txt:19; semantics per SURVEY.md Appendix B.14):
  * <default> class inheritance, joint/geom/motor attribute resolution
  * body tree in depth-first document order (= MuJoCo body ids)
  * fullinertia -> eigen-decomposition (+ balanceinertia) -> body-frame inertia
  * joint `limited` auto rule (lo < hi), ctrllimited auto rule
  * STL -> welded vertices -> convex hull (qhull via scipy) -> hull vertices,
    hull volume centroid (MuJoCo recentres mesh geoms on the mesh COM), rbound
  * PNG -> uint8 raster (row flip / normalisation are applied at model build
    time in cosim_b200/model.py so they are visible and testable)
  * missing STL blobs (/root/reference/.MISSING_LARGE_BLOBS) -> documented proxy:
    the "equivalent inertia box" of the body's <inertial>

Usage: python tools/bake_assets.py [--ref /root/reference] [--out cosim_b200/assets]
"""
import argparse
import hashlib
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

ROBOTS = {
    "flamingo_light_v1": "flamingo_light_v1.xml",
    "flamingo_p_v3": "flamingo_p_v3.xml",
    "w4_p_v2": "w4_p_v2.xml",
    "humanoid_p_v0": "humanoid_p_v0.xml",
}

GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4,
              "cylinder": 5, "box": 6, "mesh": 7}

JOINT_DEFAULTS = dict(type="hinge", pos="0 0 0", axis="0 0 1", damping="0", stiffness="0",
                      frictionloss="0", armature="0", range="0 0", limited="auto",
                      actuatorfrcrange="0 0", actuatorfrclimited="auto")
GEOM_DEFAULTS = dict(type="sphere", pos="0 0 0", quat="1 0 0 0", size="0 0 0",
                     friction="1 0.005 0.0001", contype="1", conaffinity="1", condim="3",
                     priority="0", margin="0", gap="0")
MOTOR_DEFAULTS = dict(gear="1", ctrllimited="auto", ctrlrange="0 0")


def fvec(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size < n:
        v = np.concatenate([v, np.zeros(n - v.size)])
    return v


def quat_to_mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def mat_to_quat(R):
    # standard branch-stable conversion
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


# --------------------------------------------------------------------------- defaults
def collect_defaults(root):
    """class name -> {tag: attrib dict}; 'main' is the unnamed top-level class."""
    classes = {}

    def walk(node, name, inherited):
        cur = {k: dict(v) for k, v in inherited.items()}
        for child in node:
            if child.tag != "default":
                cur.setdefault(child.tag, {}).update(child.attrib)
        classes[name] = cur
        for child in node:
            if child.tag == "default":
                walk(child, child.attrib["class"], cur)

    top = root.find("default")
    if top is None:
        classes["main"] = {}
    else:
        walk(top, top.attrib.get("class", "main"), {})
    return classes


def resolve(elem, tag, classes, base):
    cls = elem.attrib.get("class", "main")
    out = dict(base)
    out.update(classes.get(cls, classes["main"]).get(tag, {}))
    out.update({k: v for k, v in elem.attrib.items() if k != "class"})
    return out


# --------------------------------------------------------------------------- meshes
def read_stl(path):
    with open(path, "rb") as f:
        data = f.read()
    ntri = struct.unpack_from("<I", data, 80)[0]
    assert len(data) >= 84 + 50 * ntri, f"not a binary STL: {path}"
    rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                        count=ntri, offset=84)
    return rec["v"].reshape(-1, 3).astype(np.float32)


def hull_of(points):
    """points f32[n,3] -> (hull vertices f32[h,3] in first-occurrence order, centroid f64[3])."""
    from scipy.spatial import ConvexHull
    pts64 = points.astype(np.float64)
    # weld duplicates, keep first-occurrence order (MuJoCo welds STL vertices)
    _, first = np.unique(points.view([("", points.dtype)] * 3), return_index=True)
    first.sort()
    uniq = pts64[first]
    hull = ConvexHull(uniq)
    vid = np.sort(hull.vertices)
    verts = uniq[vid]
    # volume centroid of the hull from signed tetrahedra against an interior point
    c0 = verts.mean(axis=0)
    vol = 0.0
    cen = np.zeros(3)
    for simplex, eq in zip(hull.simplices, hull.equations):
        a, b, c = uniq[simplex] - c0
        v = np.dot(a, np.cross(b, c)) / 6.0
        v = abs(v)
        vol += v
        cen += v * (a + b + c) / 4.0
    centroid = c0 + cen / vol
    return verts.astype(np.float32), centroid, vol


def inertia_box_vertices(mass, ipos, I_body):
    """Equivalent-inertia box of an <inertial> (proxy for a missing collision STL)."""
    w, V = np.linalg.eigh(I_body)
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    # Ixx = m/3 (b^2 + c^2) with half sizes a,b,c
    half2 = 3.0 * (w.sum() - 2.0 * w) / (2.0 * mass)
    half = np.sqrt(np.maximum(half2, 1e-8))
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64)
    verts = ipos[None, :] + (corners * half[None, :]) @ V.T
    return verts.astype(np.float32), np.asarray(ipos, dtype=np.float64)


# --------------------------------------------------------------------------- compile
def compile_robot(ref, robot, xml_name):
    xml_dir = os.path.join(ref, "envs", robot, "assets", "xml")
    root = ET.parse(os.path.join(xml_dir, xml_name)).getroot()

    compiler = {}
    for c in root.findall("compiler"):
        compiler.update(c.attrib)
    assert compiler.get("angle", "degree") == "radian"
    balance = compiler.get("balanceinertia", "false") == "true"
    option = root.find("option").attrib
    gravity = fvec(option.get("gravity", "0 0 -9.81"))
    classes = collect_defaults(root)

    mesh_files = {m.attrib["name"]: os.path.normpath(os.path.join(xml_dir, m.attrib["file"]))
                  for m in root.find("asset").findall("mesh")}
    hfields = {h.attrib["name"]: dict(size=fvec(h.attrib["size"]), file=os.path.basename(h.attrib["file"]))
               for h in root.find("asset").findall("hfield")}

    world = root.find("worldbody")
    ground = [g for g in world.findall("geom") if g.attrib.get("name") == "ground"][0]
    gnd = resolve(ground, "geom", classes, GEOM_DEFAULTS)
    gnd_contype, gnd_conaff = int(gnd["contype"]), int(gnd["conaffinity"])

    bodies, joints, dofs, geoms, sites = [], [], [], [], {}
    hull_verts, hull_adr = [], 0
    notes = []

    def add_body(elem, parent):
        bid = len(bodies) + 1  # world = 0
        name = elem.attrib["name"]
        pos = fvec(elem.attrib.get("pos", "0 0 0"))
        quat = fvec(elem.attrib.get("quat", "1 0 0 0"))
        quat = quat / np.linalg.norm(quat)
        inert = elem.find("inertial")
        assert inert is not None, f"{robot}:{name} has no <inertial>"
        mass = float(inert.attrib["mass"])
        ipos = fvec(inert.attrib.get("pos", "0 0 0"))
        if "fullinertia" in inert.attrib:
            ixx, iyy, izz, ixy, ixz, iyz = fvec(inert.attrib["fullinertia"])
            I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
            w, V = np.linalg.eigh(I)
        else:
            w = fvec(inert.attrib["diaginertia"])
            V = quat_to_mat(fvec(inert.attrib.get("quat", "1 0 0 0")))
        assert w.min() > 0, f"{robot}:{name} inertia not positive"
        srt = np.sort(w)
        if srt[0] + srt[1] < srt[2]:
            assert balance, f"{robot}:{name} violates A+B>=C without balanceinertia"
            w = np.full(3, w.mean())
            notes.append(f"balanceinertia applied to {name}")
        I_body = V @ np.diag(w) @ V.T
        body = dict(name=name, parent=parent, pos=pos, quat=quat, mass=mass, ipos=ipos,
                    inertia=np.array([I_body[0, 0], I_body[1, 1], I_body[2, 2], I_body[0, 1], I_body[0, 2], I_body[1, 2]]),
                    jntadr=len(joints), jntnum=0, dofadr=len(dofs), dofnum=0)
        bodies.append(body)
        nonlocal hull_adr
        for j in elem.findall("joint"):
            ja = resolve(j, "joint", classes, JOINT_DEFAULTS)
            jtype = ja["type"]
            rng = fvec(ja["range"])
            lim = ja["limited"]
            limited = (lim == "true") or (lim == "auto" and rng[0] < rng[1])
            afr = fvec(ja["actuatorfrcrange"])
            al = ja["actuatorfrclimited"]
            afl = (al == "true") or (al == "auto" and afr[0] < afr[1])
            axis = fvec(ja["axis"])
            axis = axis / np.linalg.norm(axis)
            jid = len(joints)
            ndof = 6 if jtype == "free" else 1
            assert jtype in ("free", "hinge")
            qposadr = sum(7 if jj["type"] == 0 else 1 for jj in joints)
            joints.append(dict(name=ja.get("name", ""), type=0 if jtype == "free" else 3, body=bid,
                               qposadr=qposadr, dofadr=len(dofs), pos=fvec(ja["pos"]), axis=axis,
                               limited=int(limited and jtype == "hinge"), range=rng,
                               actfrclimited=int(afl), actfrcrange=afr, cls=j.attrib.get("class", "main")))
            for k in range(ndof):
                # parent dof: previous dof of the same body, else last dof of the nearest ancestor with dofs
                if k > 0 or body["dofnum"] > 0:
                    par = len(dofs) - 1
                else:
                    par = -1
                    p = parent
                    while p > 0:
                        pb = bodies[p - 1]
                        if pb["dofnum"] > 0:
                            par = pb["dofadr"] + pb["dofnum"] - 1
                            break
                        p = pb["parent"]
                dofs.append(dict(body=bid, jnt=jid, parent=par, armature=float(ja["armature"]),
                                 damping=float(ja["damping"]), frictionloss=float(ja["frictionloss"])))
                body["dofnum"] += 1
            body["jntnum"] += 1
        for s in elem.findall("site"):
            sites[s.attrib.get("name", "")] = dict(body=bid, pos=fvec(s.attrib.get("pos", "0 0 0")),
                                                   quat=fvec(s.attrib.get("quat", "1 0 0 0")))
        for g in elem.findall("geom"):
            ga = resolve(g, "geom", classes, GEOM_DEFAULTS)
            ct, ca = int(ga["contype"]), int(ga["conaffinity"])
            if not ((gnd_contype & ca) or (ct & gnd_conaff)):
                continue  # cannot touch the ground (visual copies, disabled caster meshes)
            gtype = GEOM_TYPES[ga["type"]]
            size = fvec(ga["size"], 3)
            gpos, gquat = fvec(ga["pos"]), fvec(ga["quat"])
            gquat = gquat / np.linalg.norm(gquat)
            vadr = vnum = 0
            center = np.zeros(3)
            proxy = 0
            if gtype == 7:
                path = mesh_files[ga["mesh"]]
                if os.path.exists(path):
                    verts, center, _ = hull_of(read_stl(path))
                else:
                    verts, center = inertia_box_vertices(mass, ipos, I_body)
                    proxy = 1
                    notes.append(f"missing STL {os.path.basename(path)} -> inertia-box proxy on {name}")
                vadr, vnum = hull_adr, len(verts)
                hull_verts.append(verts)
                hull_adr += vnum
                rbound = float(np.linalg.norm(verts.astype(np.float64) - center[None], axis=1).max())
            elif gtype == 2:
                rbound = size[0]
            elif gtype == 5:
                rbound = float(np.hypot(size[0], size[1]))
            elif gtype == 6:
                rbound = float(np.linalg.norm(size))
            else:
                raise NotImplementedError(ga["type"])
            geoms.append(dict(name=ga.get("name", ""), type=gtype, body=bid, size=size, pos=gpos, quat=gquat,
                              friction=fvec(ga["friction"], 3), has_friction_attr=int("friction" in g.attrib),
                              condim=int(ga["condim"]), vadr=vadr, vnum=vnum, center=center, rbound=rbound,
                              proxy=proxy, contype=ct, conaffinity=ca))
        for child in elem.findall("body"):
            add_body(child, bid)

    for b in world.findall("body"):
        add_body(b, 0)

    # actuators
    jname = {j["name"]: i for i, j in enumerate(joints)}
    acts = []
    for m in root.find("actuator").findall("motor"):
        ma = resolve(m, "motor", classes, MOTOR_DEFAULTS)
        cr = fvec(ma["ctrlrange"])
        cl = ma["ctrllimited"]
        acts.append(dict(joint=jname[ma["joint"]], gear=float(ma["gear"].split()[0]),
                         ctrllimited=int(cl == "true" or (cl == "auto" and cr[0] < cr[1])), ctrlrange=cr))

    # equality (connect only)
    bname = {b["name"]: i + 1 for i, b in enumerate(bodies)}
    eqs = []
    eq_root = root.find("equality")
    if eq_root is not None:
        for c in eq_root.findall("connect"):
            solimp = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
            if "solimp" in c.attrib:
                v = fvec(c.attrib["solimp"])
                solimp[:v.size] = v
            solref = fvec(c.attrib.get("solref", "0.02 1"))
            eqs.append(dict(body1=bname[c.attrib["body1"]], body2=bname[c.attrib["body2"]],
                            anchor=fvec(c.attrib["anchor"]), solref=solref, solimp=solimp))

    # <contact><exclude body1 body2/>: body pairs whose geoms never collide
    excl = []
    for ct_root in root.findall("contact"):
        for x in ct_root.findall("exclude"):
            excl.append((bname[x.attrib["body1"]], bname[x.attrib["body2"]]))

    imu_name = "imu" if "imu" in sites else "imu_site"
    imu = sites[imu_name]

    out = dict(
        robot=np.array(robot),
        gravity=gravity,
        timestep=np.array(float(option["timestep"])), iterations=np.array(int(option["iterations"])),
        body_names=np.array([b["name"] for b in bodies]),
        body_parent=np.array([b["parent"] for b in bodies], dtype=np.int32),
        body_pos=np.stack([b["pos"] for b in bodies]), body_quat=np.stack([b["quat"] for b in bodies]),
        body_mass=np.array([b["mass"] for b in bodies]), body_ipos=np.stack([b["ipos"] for b in bodies]),
        body_inertia=np.stack([b["inertia"] for b in bodies]),
        body_jntadr=np.array([b["jntadr"] for b in bodies], dtype=np.int32),
        body_jntnum=np.array([b["jntnum"] for b in bodies], dtype=np.int32),
        body_dofadr=np.array([b["dofadr"] for b in bodies], dtype=np.int32),
        body_dofnum=np.array([b["dofnum"] for b in bodies], dtype=np.int32),
        jnt_names=np.array([j["name"] for j in joints]), jnt_class=np.array([j["cls"] for j in joints]),
        jnt_type=np.array([j["type"] for j in joints], dtype=np.int32),
        jnt_body=np.array([j["body"] for j in joints], dtype=np.int32),
        jnt_qposadr=np.array([j["qposadr"] for j in joints], dtype=np.int32),
        jnt_dofadr=np.array([j["dofadr"] for j in joints], dtype=np.int32),
        jnt_pos=np.stack([j["pos"] for j in joints]), jnt_axis=np.stack([j["axis"] for j in joints]),
        jnt_limited=np.array([j["limited"] for j in joints], dtype=np.int32),
        jnt_range=np.stack([j["range"] for j in joints]),
        jnt_actfrclimited=np.array([j["actfrclimited"] for j in joints], dtype=np.int32),
        jnt_actfrcrange=np.stack([j["actfrcrange"] for j in joints]),
        dof_body=np.array([d["body"] for d in dofs], dtype=np.int32),
        dof_jnt=np.array([d["jnt"] for d in dofs], dtype=np.int32),
        dof_parent=np.array([d["parent"] for d in dofs], dtype=np.int32),
        dof_armature=np.array([d["armature"] for d in dofs]),
        dof_damping=np.array([d["damping"] for d in dofs]),
        dof_frictionloss=np.array([d["frictionloss"] for d in dofs]),
        act_joint=np.array([a["joint"] for a in acts], dtype=np.int32),
        act_gear=np.array([a["gear"] for a in acts]),
        act_ctrllimited=np.array([a["ctrllimited"] for a in acts], dtype=np.int32),
        act_ctrlrange=np.stack([a["ctrlrange"] for a in acts]),
        geom_names=np.array([g["name"] for g in geoms]),
        geom_type=np.array([g["type"] for g in geoms], dtype=np.int32),
        geom_body=np.array([g["body"] for g in geoms], dtype=np.int32),
        geom_size=np.stack([g["size"] for g in geoms]), geom_pos=np.stack([g["pos"] for g in geoms]),
        geom_quat=np.stack([g["quat"] for g in geoms]), geom_friction=np.stack([g["friction"] for g in geoms]),
        geom_has_friction_attr=np.array([g["has_friction_attr"] for g in geoms], dtype=np.int32),
        geom_condim=np.array([g["condim"] for g in geoms], dtype=np.int32),
        geom_vadr=np.array([g["vadr"] for g in geoms], dtype=np.int32),
        geom_vnum=np.array([g["vnum"] for g in geoms], dtype=np.int32),
        geom_center=np.stack([g["center"] for g in geoms]),
        geom_rbound=np.array([g["rbound"] for g in geoms]),
        geom_proxy=np.array([g["proxy"] for g in geoms], dtype=np.int32),
        geom_contype=np.array([g["contype"] for g in geoms], dtype=np.int32),
        geom_conaffinity=np.array([g["conaffinity"] for g in geoms], dtype=np.int32),
        exclude_body=np.array(excl, dtype=np.int32).reshape(-1, 2),
        hull_verts=(np.concatenate(hull_verts) if hull_verts else np.zeros((0, 3), np.float32)),
        ground_friction=fvec(gnd["friction"], 3), ground_condim=np.array(int(gnd["condim"])),
        ground_has_friction_attr=np.array(int("friction" in ground.attrib)),
        hfield_names=np.array(sorted(hfields)), hfield_size=np.stack([hfields[k]["size"] for k in sorted(hfields)]),
        hfield_file=np.array([hfields[k]["file"] for k in sorted(hfields)]),
        imu_body=np.array(imu["body"]), imu_pos=imu["pos"], imu_quat=imu["quat"],
        eq_body1=np.array([e["body1"] for e in eqs], dtype=np.int32),
        eq_body2=np.array([e["body2"] for e in eqs], dtype=np.int32),
        eq_anchor=(np.stack([e["anchor"] for e in eqs]) if eqs else np.zeros((0, 3))),
        eq_solref=(np.stack([e["solref"] for e in eqs]) if eqs else np.zeros((0, 2))),
        eq_solimp=(np.stack([e["solimp"] for e in eqs]) if eqs else np.zeros((0, 5))),
        notes=np.array(notes),
    )
    return out


def bake_terrains(ref, out_dir):
    from PIL import Image
    tdir = os.path.join(ref, "envs", "flamingo_p_v3", "assets", "terrain")
    seen = {}
    for fn in sorted(os.listdir(tdir)):
        if not fn.endswith(".png") or fn == "flat.png":
            continue
        raw = open(os.path.join(tdir, fn), "rb").read()
        md5 = hashlib.md5(raw).hexdigest()
        im = Image.open(os.path.join(tdir, fn))
        assert im.mode == "L"
        a = np.array(im, dtype=np.uint8)
        name = fn[:-4]
        if md5 in seen:  # rocky_easy == rocky_hard etc: store once, alias by name
            np.savez_compressed(os.path.join(out_dir, f"terrain_{name}.npz"), alias=np.array(seen[md5]), md5=np.array(md5))
        else:
            seen[md5] = name
            np.savez_compressed(os.path.join(out_dir, f"terrain_{name}.npz"), raster=a, md5=np.array(md5))
        print(f"terrain {name}: {a.shape} md5 {md5[:8]}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "cosim_b200", "assets"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for robot, xml in ROBOTS.items():
        m = compile_robot(args.ref, robot, xml)
        np.savez_compressed(os.path.join(args.out, f"robot_{robot}.npz"), **m)
        nv = len(m["dof_body"])
        print(f"{robot}: nbody {len(m['body_mass'])} nv {nv} nu {len(m['act_joint'])} ngeom {len(m['geom_type'])} "
              f"hullverts {len(m['hull_verts'])} mass {m['body_mass'].sum():.5f}")
        for n in m["notes"]:
            print("   note:", n)
    bake_terrains(args.ref, args.out)


if __name__ == "__main__":
    main()

"""MJCF mini-compiler for the myCobot 280 model tree of the reference.

Subsystem (1) of the north star: turn the reference's MJCF include tree
(`mycobotgym/envs/assets/mycobot280.xml:1-10` -> `mycobot280_main.xml:1-270`,
`joint_actuators.xml:1-23`) into a flat, mjModel-like table of numpy arrays
("FlatModel").  It restates what MuJoCo 2.3.2's compiler + `mj_setConst` do for the
elements this model uses (defaults/childclass, local coordinates, radians,
inertia-from-geom for bodies without <inertial>, legacy mesh inertia, connect
anchor2, body/dof invweight0, meaninertia).  MuJoCo itself is an un-vendored
dependency of the reference (`requirements.txt:4`), so this is a restatement from
its published algorithm; see DESIGN.md "parity unpinned".

The compiler runs only where the reference assets are present (this container).
Its output is committed as `mycobotgym_b200/assets/*.json` so that the GPU box,
which has no `/root/reference`, loads the compiled model.  When `mujoco` is
importable `flatmodel_from_mjmodel()` fills the same table from a live mjModel.

Mesh geoms are used for inertia only (bodies `flange`, `gripper_base`); their
convex hulls are NOT emitted as collision geoms in this round.
`base_link.STL` is absent from the reference mount; it sits on a static body and
only affects collision.
"""
from __future__ import annotations

import json
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

mjMINVAL = 1e-15

# ----------------------------------------------------------------------------------
# small quaternion / rotation helpers (w, x, y, z)


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat2mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat2quat(R):
    """Rotation matrix -> unit quaternion (w>=0 branch selection by largest diagonal)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    return q


def euler2quat_xyz(e):
    """MJCF `euler` with the default eulerseq "xyz" (intrinsic)."""
    q = np.array([1.0, 0, 0, 0])
    for ax, ang in zip(range(3), e):
        h = 0.5 * ang
        r = np.array([np.cos(h), 0, 0, 0])
        r[1 + ax] = np.sin(h)
        q = quat_mul(q, r)
    return q


def axisangle2quat(axis, angle):
    h = 0.5 * angle
    s = np.sin(h)
    return np.array([np.cos(h), axis[0] * s, axis[1] * s, axis[2] * s])


def principal_axes(I):
    """Symmetric 3x3 inertia -> (iquat, diag) with a right-handed eigenbasis,
    eigenvalues sorted in decreasing order (MuJoCo's convention)."""
    w, V = np.linalg.eigh(I)
    order = np.argsort(-w)
    w = w[order]
    V = V[:, order]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    return mat2quat(V), w


# ----------------------------------------------------------------------------------
# STL + legacy mesh inertia  (MuJoCo 2.3.2 user_mesh.cc mjCMesh::Process, exactmeshinertia=false)


def read_stl(path):
    with open(path, "rb") as f:
        raw = f.read()
    n = struct.unpack_from("<I", raw, 80)[0]
    if len(raw) != 84 + 50 * n:
        raise ValueError(f"{path}: not a binary STL")
    rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
    return rec["v"].astype(np.float64)  # (F,3,3); MuJoCo keeps float32 vertices


def _area_normal(tri):
    e1 = tri[:, 1] - tri[:, 0]
    e2 = tri[:, 2] - tri[:, 0]
    nrm = np.cross(e1, e2)
    ln = np.linalg.norm(nrm, axis=1)
    ok = ln > mjMINVAL
    nrm = np.where(ok[:, None], nrm / np.where(ok, ln, 1.0)[:, None], 0.0)
    return 0.5 * ln, nrm, ok


def mesh_legacy_inertia(tri):
    """Returns (volume, com[3], inertia 3x3 about com in mesh file axes, per unit density)."""
    area, nrm, ok = _area_normal(tri)
    tri = tri[ok]
    area = area[ok]
    nrm = nrm[ok]
    cen = tri.mean(axis=1)
    facecen = (area[:, None] * cen).sum(0) / area.sum()
    vol_f = np.abs(((cen - facecen) * nrm).sum(1) * area / 3.0)
    vol = vol_f.sum()
    com = (vol_f[:, None] * (0.75 * cen + 0.25 * facecen)).sum(0) / vol
    tri = tri - com
    cen = tri.mean(axis=1)
    vol_f = np.abs((cen * nrm).sum(1) * area / 3.0)
    vol = vol_f.sum()
    D, E, F = tri[:, 0], tri[:, 1], tri[:, 2]
    P = np.zeros((3, 3))
    for a in range(3):
        for b in range(a, 3):
            val = (vol_f / 20.0 * (
                2.0 * (D[:, a] * D[:, b] + E[:, a] * E[:, b] + F[:, a] * F[:, b])
                + D[:, a] * E[:, b] + D[:, b] * E[:, a]
                + D[:, a] * F[:, b] + D[:, b] * F[:, a]
                + E[:, a] * F[:, b] + E[:, b] * F[:, a])).sum()
            P[a, b] = P[b, a] = val
    I = np.array([
        [P[1, 1] + P[2, 2], -P[0, 1], -P[0, 2]],
        [-P[0, 1], P[0, 0] + P[2, 2], -P[1, 2]],
        [-P[0, 2], -P[1, 2], P[0, 0] + P[1, 1]],
    ])
    return vol, com, I


# ----------------------------------------------------------------------------------
# XML loading with <include> and <default> classes

GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4, "cylinder": 5, "box": 6, "mesh": 7}
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
EQ_CONNECT, EQ_WELD, EQ_JOINT = 0, 1, 2

BUILTIN_DEFAULTS = {
    "joint": dict(type="hinge", pos="0 0 0", axis="0 0 1", armature="0", damping="0", limited="false",
                  range="0 0", margin="0", solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2",
                  stiffness="0", frictionloss="0"),
    "geom": dict(type="sphere", contype="1", conaffinity="1", condim="3", friction="1 0.005 0.0001",
                 solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", solmix="1", margin="0", gap="0",
                 density="1000", pos="0 0 0", size="0 0 0", priority="0", group="0"),
    "site": dict(pos="0 0 0", size="0.005 0.005 0.005", type="sphere"),
    "general": dict(ctrllimited="false", forcelimited="false", ctrlrange="0 0", forcerange="0 0", gear="1 0 0 0 0 0",
                    dyntype="none", gaintype="fixed", biastype="none", gainprm="1 0 0", biasprm="0 0 0"),
    "equality": dict(solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", active="true"),
}


def _load_xml(path, _top=True):
    root = ET.parse(path).getroot()
    base = os.path.dirname(path)
    out = ET.Element("mujoco")
    for child in list(root):
        if child.tag == "include":
            inc = _load_xml(os.path.join(base, child.get("file")), _top=False)
            out.extend(list(inc))
        else:
            out.append(child)
    if _top:
        # MuJoCo merges repeated top-level sections (worldbody, equality, actuator, ...) in document order
        merged = ET.Element("mujoco")
        seen = {}
        for child in list(out):
            if child.tag in ("worldbody", "equality", "actuator", "contact", "tendon", "keyframe", "asset") and child.tag in seen:
                seen[child.tag].extend(list(child))
            else:
                seen.setdefault(child.tag, child)
                merged.append(child)
        return merged
    return out


def _vec(s, n=None, pad=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and len(v) < n:
        v = np.concatenate([v, np.asarray(pad[len(v):n], dtype=np.float64)])
    return v


def _bool(s):
    return s.strip().lower() == "true"


class _Defaults:
    def __init__(self, root):
        self.classes = {"main": {k: dict(v) for k, v in BUILTIN_DEFAULTS.items()}}
        for d in root.findall("default"):
            self._walk(d, "main", top=True)

    def _walk(self, node, parent, top=False):
        name = node.get("class", "main" if top else None)
        if name is None:
            raise ValueError("nested default without class")
        if name != "main":
            self.classes[name] = {k: dict(v) for k, v in self.classes[parent].items()}
        cur = self.classes[name]
        for el in node:
            if el.tag == "default":
                continue
            cur.setdefault(el.tag, {}).update(el.attrib)
        for el in node.findall("default"):
            self._walk(el, name)

    def resolve(self, el, tag, childclass):
        cls = el.get("class", childclass or "main")
        attrs = dict(self.classes[cls].get(tag, {}))
        attrs.update({k: v for k, v in el.attrib.items() if k != "class"})
        return attrs


def _orient(el):
    if el.get("quat") is not None:
        q = _vec(el.get("quat"))
        return q / np.linalg.norm(q)
    if el.get("euler") is not None:
        return euler2quat_xyz(_vec(el.get("euler")))
    return np.array([1.0, 0, 0, 0])


# ----------------------------------------------------------------------------------


HULL_SIDE_KEYS = ("hull_vert",)
HULL_SIDE_FILE = "mycobot280_hulls.npz"


class FlatModel(dict):
    """dict of numpy arrays / scalars with attribute access; JSON round-trips exactly."""

    __getattr__ = dict.__getitem__

    def to_json(self, path):
        def enc(v):
            if isinstance(v, np.ndarray):
                return {"dtype": str(v.dtype), "shape": list(v.shape), "data": v.ravel().tolist()}
            if isinstance(v, (np.integer,)):
                return int(v)
            if isinstance(v, (np.floating,)):
                return float(v)
            return v

        side = {k: v for k, v in self.items() if k in HULL_SIDE_KEYS}
        if side:                                                # the hull vertex table is binary and shared by the model variants
            np.savez_compressed(os.path.join(os.path.dirname(path), HULL_SIDE_FILE), **side)
        with open(path, "w") as f:
            json.dump({k: enc(v) for k, v in self.items() if k not in HULL_SIDE_KEYS}, f, indent=0, sort_keys=True)

    @staticmethod
    def from_json(path):
        with open(path) as f:
            raw = json.load(f)
        m = FlatModel()
        for k, v in raw.items():
            if isinstance(v, dict) and "dtype" in v:
                m[k] = np.array(v["data"], dtype=v["dtype"]).reshape(v["shape"])
            else:
                m[k] = v
        side = os.path.join(os.path.dirname(path), HULL_SIDE_FILE)
        if m.get("nhull") and os.path.exists(side):
            z = np.load(side)
            for k in HULL_SIDE_KEYS:
                m[k] = z[k]
        return m


def compile_mjcf(xml_path, log=None):
    """Compile the reference MJCF (joint variant) into a FlatModel."""
    log = log if log is not None else []
    root = _load_xml(xml_path)
    base = os.path.dirname(xml_path)
    comp = root.find("compiler")
    assert comp.get("angle") == "radian" and comp.get("coordinate", "local") == "local"
    meshdir = os.path.join(base, comp.get("meshdir", ""))
    defs = _Defaults(root)
    opt = root.find("option")
    timestep = float(opt.get("timestep", "0.002"))

    meshes = {}
    for me in root.find("asset").findall("mesh"):
        meshes[me.get("name")] = os.path.join(meshdir, me.get("file"))
    mesh_cache = {}

    def mesh_props(name):
        if name not in mesh_cache:
            p = meshes[name]
            if not os.path.exists(p):
                log.append(f"mesh {name}: file missing ({os.path.basename(p)}); geom dropped")
                mesh_cache[name] = None
            else:
                mesh_cache[name] = mesh_legacy_inertia(read_stl(p))
        return mesh_cache[name]

    bodies, joints, geoms, sites = [], [], [], []

    def add_body(el, parent, childclass):
        bid = len(bodies)
        if el is None:  # world
            b = dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), mocap=False)
            children = root.find("worldbody")
        else:
            childclass = el.get("childclass", childclass)
            b = dict(name=el.get("name"), parent=parent, pos=_vec(el.get("pos", "0 0 0")), quat=_orient(el),
                     mocap=_bool(el.get("mocap", "false")))
            children = el
        b.update(jntadr=-1, jntnum=0, inertial=None, geom_ids=[])
        bodies.append(b)
        for ch in children:
            if ch.tag == "inertial":
                di = _vec(ch.get("diaginertia"))
                b["inertial"] = dict(pos=_vec(ch.get("pos")), quat=_orient(ch), mass=float(ch.get("mass")), diag=di)
            elif ch.tag == "joint":
                a = defs.resolve(ch, "joint", childclass)
                jt = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[a["type"]]
                if b["jntnum"] == 0:
                    b["jntadr"] = len(joints)
                b["jntnum"] += 1
                ax = _vec(a["axis"])
                joints.append(dict(name=ch.get("name"), type=jt, body=bid, pos=_vec(a["pos"]),
                                   axis=ax / max(np.linalg.norm(ax), mjMINVAL),
                                   armature=float(a["armature"]), damping=float(a["damping"]),
                                   limited=_bool(a["limited"]), range=_vec(a["range"]), margin=float(a["margin"]),
                                   solref=_vec(a["solreflimit"]),
                                   solimp=_vec(a["solimplimit"], 5, [0.9, 0.95, 0.001, 0.5, 2])))
            elif ch.tag == "geom":
                a = defs.resolve(ch, "geom", childclass)
                gt = GEOM_TYPES[a["type"]]
                g = dict(name=ch.get("name"), type=gt, body=bid, pos=_vec(a["pos"]), quat=_orient(ch),
                         size=_vec(a["size"], 3, [0, 0, 0]), contype=int(a["contype"]), conaffinity=int(a["conaffinity"]),
                         condim=int(a["condim"]), friction=_vec(a["friction"], 3, [1, 0.005, 0.0001]),
                         solref=_vec(a["solref"]), solimp=_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2]),
                         solmix=float(a["solmix"]), margin=float(a["margin"]), gap=float(a["gap"]),
                         density=float(a["density"]), mass=(float(a["mass"]) if "mass" in a else None),
                         mesh=a.get("mesh"))
                b["geom_ids"].append(len(geoms))
                geoms.append(g)
            elif ch.tag == "site":
                a = defs.resolve(ch, "site", childclass)
                sites.append(dict(name=ch.get("name"), body=bid, pos=_vec(a["pos"]), quat=_orient(ch)))
            elif ch.tag == "body":
                add_body(ch, bid, childclass)

    add_body(None, 0, None)
    nbody, njnt = len(bodies), len(joints)

    # ---- joints -> qpos / dof addresses
    qadr, dadr = 0, 0
    dof_body, dof_jnt, dof_parent, dof_armature, dof_damping = [], [], [], [], []
    body_lastdof = [-1] * nbody
    for b in bodies:
        b["dofadr"], b["dofnum"] = -1, 0
    for bid, b in enumerate(bodies):
        last = body_lastdof[b["parent"]] if bid else -1
        for j in range(b["jntadr"], b["jntadr"] + b["jntnum"]) if b["jntnum"] else []:
            jn = joints[j]
            jn["qposadr"], jn["dofadr"] = qadr, dadr
            nq_j, nv_j = {JNT_FREE: (7, 6), JNT_BALL: (4, 3), JNT_SLIDE: (1, 1), JNT_HINGE: (1, 1)}[jn["type"]]
            if b["dofadr"] < 0:
                b["dofadr"] = dadr
            for k in range(nv_j):
                dof_body.append(bid)
                dof_jnt.append(j)
                dof_parent.append(last)
                dof_armature.append(jn["armature"])
                dof_damping.append(jn["damping"])
                last = dadr + k
            b["dofnum"] += nv_j
            qadr += nq_j
            dadr += nv_j
        body_lastdof[bid] = last
    nq, nv = qadr, dadr

    # ---- body inertial properties
    for bid, b in enumerate(bodies):
        if b["inertial"] is not None:
            it = b["inertial"]
            b["mass"], b["ipos"], b["iquat"], b["inertia"] = it["mass"], it["pos"], it["quat"], it["diag"]
            continue
        parts = []  # (mass, pos, I 3x3 in body axes about geom/mesh com)
        for gid in b["geom_ids"]:
            g = geoms[gid]
            R = quat2mat(g["quat"])
            if g["type"] == GEOM_TYPES["mesh"]:
                mp = mesh_props(g["mesh"])
                if mp is None:
                    continue
                vol, com, I = mp
                m = g["density"] * vol if g["mass"] is None else g["mass"]
                if m <= 0:
                    continue
                parts.append((m, g["pos"] + R @ com, R @ (I * (m / vol)) @ R.T))
            elif g["type"] == GEOM_TYPES["box"]:
                s = g["size"]
                vol = 8 * s[0] * s[1] * s[2]
                m = g["density"] * vol if g["mass"] is None else g["mass"]
                if m <= 0:
                    continue
                I = np.diag([m / 3 * (s[1] ** 2 + s[2] ** 2), m / 3 * (s[0] ** 2 + s[2] ** 2), m / 3 * (s[0] ** 2 + s[1] ** 2)])
                parts.append((m, g["pos"].copy(), R @ I @ R.T))
            elif g["type"] == GEOM_TYPES["plane"]:
                continue
            else:
                raise NotImplementedError(g["type"])
        if not parts:
            b["mass"], b["ipos"], b["iquat"], b["inertia"] = 0.0, np.zeros(3), np.array([1.0, 0, 0, 0]), np.zeros(3)
            continue
        mtot = sum(p[0] for p in parts)
        com = sum(p[0] * p[1] for p in parts) / mtot
        I = np.zeros((3, 3))
        for m, p, Ig in parts:
            d = p - com
            I += Ig + m * (d @ d * np.eye(3) - np.outer(d, d))
        iq, diag = principal_axes(I)
        b["mass"], b["ipos"], b["iquat"], b["inertia"] = mtot, com, iq, diag

    # ---- weld / root ids, subtree mass
    for bid, b in enumerate(bodies):
        if bid == 0:
            b["weld"], b["root"] = 0, 0
        else:
            p = bodies[b["parent"]]
            b["weld"] = bid if b["jntnum"] else p["weld"]
            b["root"] = bid if b["parent"] == 0 else p["root"]
    subtreemass = np.array([b["mass"] for b in bodies])
    for bid in range(nbody - 1, 0, -1):
        subtreemass[bodies[bid]["parent"]] += subtreemass[bid]

    name2body = {b["name"]: i for i, b in enumerate(bodies)}
    name2jnt = {j["name"]: i for i, j in enumerate(joints)}

    m = FlatModel()
    m["nq"], m["nv"], m["nbody"], m["njnt"] = nq, nv, nbody, njnt
    m["timestep"] = timestep
    m["gravity"] = np.array([0.0, 0.0, -9.81])
    m["tolerance"], m["iterations"], m["ls_iterations"], m["ls_tolerance"], m["impratio"] = 1e-8, 100, 50, 0.01, 1.0
    m["body_names"] = [b["name"] for b in bodies]
    m["body_parentid"] = np.array([b["parent"] for b in bodies], dtype=np.int32)
    m["body_rootid"] = np.array([b["root"] for b in bodies], dtype=np.int32)
    m["body_weldid"] = np.array([b["weld"] for b in bodies], dtype=np.int32)
    m["body_jntnum"] = np.array([b["jntnum"] for b in bodies], dtype=np.int32)
    m["body_jntadr"] = np.array([b["jntadr"] for b in bodies], dtype=np.int32)
    m["body_dofnum"] = np.array([b["dofnum"] for b in bodies], dtype=np.int32)
    m["body_dofadr"] = np.array([b["dofadr"] for b in bodies], dtype=np.int32)
    m["body_pos"] = np.array([b["pos"] for b in bodies])
    m["body_quat"] = np.array([b["quat"] for b in bodies])
    m["body_ipos"] = np.array([b["ipos"] for b in bodies])
    m["body_iquat"] = np.array([b["iquat"] for b in bodies])
    m["body_mass"] = np.array([b["mass"] for b in bodies])
    m["body_inertia"] = np.array([b["inertia"] for b in bodies])
    m["body_subtreemass"] = subtreemass
    mocapid, nmocap = [], 0
    for b in bodies:
        if b["mocap"]:
            assert b["jntnum"] == 0 and b["parent"] == 0, "mocap bodies are static children of the world"
            mocapid.append(nmocap)
            nmocap += 1
        else:
            mocapid.append(-1)
    m["nmocap"] = nmocap
    m["body_mocapid"] = np.array(mocapid, dtype=np.int32)
    m["jnt_names"] = [j["name"] for j in joints]
    m["jnt_type"] = np.array([j["type"] for j in joints], dtype=np.int32)
    m["jnt_qposadr"] = np.array([j["qposadr"] for j in joints], dtype=np.int32)
    m["jnt_dofadr"] = np.array([j["dofadr"] for j in joints], dtype=np.int32)
    m["jnt_bodyid"] = np.array([j["body"] for j in joints], dtype=np.int32)
    m["jnt_pos"] = np.array([j["pos"] for j in joints])
    m["jnt_axis"] = np.array([j["axis"] for j in joints])
    m["jnt_limited"] = np.array([int(j["limited"]) for j in joints], dtype=np.int32)
    m["jnt_range"] = np.array([j["range"] for j in joints])
    m["jnt_margin"] = np.array([j["margin"] for j in joints])
    m["jnt_solref"] = np.array([j["solref"] for j in joints])
    m["jnt_solimp"] = np.array([j["solimp"] for j in joints])
    m["dof_bodyid"] = np.array(dof_body, dtype=np.int32)
    m["dof_jntid"] = np.array(dof_jnt, dtype=np.int32)
    m["dof_parentid"] = np.array(dof_parent, dtype=np.int32)
    m["dof_armature"] = np.array(dof_armature)
    m["dof_damping"] = np.array(dof_damping)
    madr, a = [], 0
    for i in range(nv):
        madr.append(a)
        k = i
        while k >= 0:
            a += 1
            k = dof_parent[k]
    m["dof_Madr"] = np.array(madr, dtype=np.int32)
    m["nM"] = a

    # qpos0
    qpos0 = np.zeros(nq)
    for j in joints:
        if j["type"] == JNT_FREE:
            b = bodies[j["body"]]
            qpos0[j["qposadr"]:j["qposadr"] + 3] = b["pos"]
            qpos0[j["qposadr"] + 3:j["qposadr"] + 7] = b["quat"]
    m["qpos0"] = qpos0

    # ---- collision geoms: primitives only (plane, box); mesh hulls are a documented gap
    cg = [g for g in geoms if g["type"] in (GEOM_TYPES["plane"], GEOM_TYPES["box"]) and (g["contype"] or g["conaffinity"])]
    # ---- mesh geoms collide as convex hulls (MuJoCo: qhull at compile time, mjc_Convex / mjc_PlaneConvex at run time).  Identical
    # mesh geoms of one body (every robot body carries the same mesh twice: a group-1 density-0 copy and a default one, both with
    # contype = conaffinity = 1) are emitted once with a multiplicity.  Hull vertices are kept in the BODY frame.
    hulls = {}
    for g in geoms:
        if g["type"] != GEOM_TYPES["mesh"] or not (g["contype"] or g["conaffinity"]):
            continue
        mp = mesh_props(g["mesh"])
        if mp is None:
            continue
        key = (g["body"], g["mesh"], tuple(g["pos"]), tuple(g["quat"]), g["condim"], tuple(g["friction"]), tuple(g["solref"]), tuple(g["solimp"]))
        if key in hulls:
            hulls[key]["mult"] += 1
            continue
        from scipy.spatial import ConvexHull

        v = read_stl(meshes[g["mesh"]]).reshape(-1, 3)
        v = np.unique(v, axis=0)
        hv = v[np.sort(ConvexHull(v).vertices)]
        R = quat2mat(g["quat"])
        hv = hv @ R.T + g["pos"]
        center = g["pos"] + R @ mp[1]                              # geom frame origin after MuJoCo recentres the mesh at its COM
        hulls[key] = dict(g=g, mult=1, vert=hv, center=center, rbound=float(np.linalg.norm(hv - center, axis=1).max()))
    hl = list(hulls.values())
    m["nhull"] = len(hl)
    m["hull_names"] = [bodies[h["g"]["body"]]["name"] for h in hl]
    m["hull_bodyid"] = np.array([h["g"]["body"] for h in hl], dtype=np.int32)
    m["hull_mult"] = np.array([h["mult"] for h in hl], dtype=np.int32)
    m["hull_vertnum"] = np.array([len(h["vert"]) for h in hl], dtype=np.int32)
    m["hull_vertadr"] = np.concatenate(([0], np.cumsum(m["hull_vertnum"])[:-1])).astype(np.int32) if hl else np.zeros(0, dtype=np.int32)
    m["hull_vert"] = np.concatenate([h["vert"] for h in hl]) if hl else np.zeros((0, 3))
    m["hull_center"] = np.array([h["center"] for h in hl]).reshape(len(hl), 3)
    m["hull_rbound"] = np.array([h["rbound"] for h in hl])
    m["hull_condim"] = np.array([h["g"]["condim"] for h in hl], dtype=np.int32)
    m["hull_friction"] = np.array([h["g"]["friction"] for h in hl]).reshape(len(hl), 3)
    m["hull_solref"] = np.array([h["g"]["solref"] for h in hl]).reshape(len(hl), 2)
    m["hull_solimp"] = np.array([h["g"]["solimp"] for h in hl]).reshape(len(hl), 5)
    m["hull_solmix"] = np.array([h["g"]["solmix"] for h in hl])
    log.append(f"{len(hl)} convex hulls ({int(m['hull_vertnum'].sum()) if hl else 0} vertices) from {sum(h['mult'] for h in hl)} mesh geoms")
    m["ngeom"] = len(cg)
    m["geom_names"] = [g["name"] or "" for g in cg]
    m["geom_type"] = np.array([g["type"] for g in cg], dtype=np.int32)
    m["geom_bodyid"] = np.array([g["body"] for g in cg], dtype=np.int32)
    m["geom_pos"] = np.array([g["pos"] for g in cg])
    m["geom_quat"] = np.array([g["quat"] for g in cg])
    m["geom_size"] = np.array([g["size"] for g in cg])
    m["geom_contype"] = np.array([g["contype"] for g in cg], dtype=np.int32)
    m["geom_conaffinity"] = np.array([g["conaffinity"] for g in cg], dtype=np.int32)
    m["geom_condim"] = np.array([g["condim"] for g in cg], dtype=np.int32)
    m["geom_friction"] = np.array([g["friction"] for g in cg])
    m["geom_solref"] = np.array([g["solref"] for g in cg])
    m["geom_solimp"] = np.array([g["solimp"] for g in cg])
    m["geom_solmix"] = np.array([g["solmix"] for g in cg])
    m["geom_margin"] = np.array([g["margin"] for g in cg])
    m["geom_gap"] = np.array([g["gap"] for g in cg])
    rb = []
    for g in cg:
        rb.append(0.0 if g["type"] == GEOM_TYPES["plane"] else float(np.linalg.norm(g["size"])))
    m["geom_rbound"] = np.array(rb)

    m["nsite"] = len(sites)
    m["site_names"] = [s["name"] for s in sites]
    m["site_bodyid"] = np.array([s["body"] for s in sites], dtype=np.int32)
    m["site_pos"] = np.array([s["pos"] for s in sites])
    m["site_quat"] = np.array([s["quat"] for s in sites])

    # ---- contact excludes (body-id pairs)
    ex = []
    con = root.find("contact")
    if con is not None:
        for e in con.findall("exclude"):
            ex.append(sorted([name2body[e.get("body1")], name2body[e.get("body2")]]))
    m["exclude"] = np.array(ex, dtype=np.int32).reshape(-1, 2)

    # ---- tendons (fixed only)
    tendons = []
    tn = root.find("tendon")
    if tn is not None:
        for t in tn.findall("fixed"):
            tendons.append(dict(name=t.get("name"),
                                jnt=[name2jnt[w.get("joint")] for w in t.findall("joint")],
                                coef=[float(w.get("coef")) for w in t.findall("joint")]))
    name2ten = {t["name"]: i for i, t in enumerate(tendons)}
    m["ntendon"] = len(tendons)
    ten_J = np.zeros((len(tendons), nv))
    for i, t in enumerate(tendons):
        for j, c in zip(t["jnt"], t["coef"]):
            ten_J[i, joints[j]["dofadr"]] = c
    m["ten_J"] = ten_J  # length = ten_J @ qpos[hinge adr]; all wrapped joints are hinges

    # ---- equality
    eqs = []
    eqn = root.find("equality")
    if eqn is not None:
        for e in eqn:
            a = dict(BUILTIN_DEFAULTS["equality"])
            a.update(e.attrib)
            d = np.zeros(11)
            if e.tag == "weld":
                d[0:3] = _vec(a.get("anchor", "0 0 0"))
                d[3:10] = _vec(a.get("relpose", "0 1 0 0 0 0 0"))
                d[10] = float(a.get("torquescale", "1"))
                eqs.append(dict(type=EQ_WELD, o1=name2body[a["body1"]], o2=name2body[a["body2"]], data=d,
                                solref=_vec(a["solref"]), solimp=_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2])))
            elif e.tag == "connect":
                d[:3] = _vec(a["anchor"])
                eqs.append(dict(type=EQ_CONNECT, o1=name2body[a["body1"]], o2=name2body[a["body2"]], data=d,
                                solref=_vec(a["solref"]), solimp=_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2])))
            elif e.tag == "joint":
                d[:5] = _vec(a["polycoef"])
                eqs.append(dict(type=EQ_JOINT, o1=name2jnt[a["joint1"]], o2=name2jnt[a["joint2"]], data=d,
                                solref=_vec(a["solref"]), solimp=_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2])))
            else:
                raise NotImplementedError(e.tag)
    m["neq"] = len(eqs)
    m["eq_type"] = np.array([e["type"] for e in eqs], dtype=np.int32)
    m["eq_obj1id"] = np.array([e["o1"] for e in eqs], dtype=np.int32)
    m["eq_obj2id"] = np.array([e["o2"] for e in eqs], dtype=np.int32)
    m["eq_data"] = np.array([e["data"] for e in eqs]).reshape(len(eqs), 11)
    m["eq_solref"] = np.array([e["solref"] for e in eqs])
    m["eq_solimp"] = np.array([e["solimp"] for e in eqs])

    # ---- actuators (general, dyntype none, gain fixed, bias affine)
    acts = []
    an = root.find("actuator")
    if an is not None:
        for e in an:
            assert e.tag == "general"
            a = defs.resolve(e, "general", None)
            assert a["dyntype"] == "none" and a.get("gaintype", "fixed") == "fixed"
            moment = np.zeros(nv)
            if "joint" in a:
                jn = joints[name2jnt[a["joint"]]]
                moment[jn["dofadr"]] = _vec(a["gear"])[0]
            else:
                moment[:] = ten_J[name2ten[a["tendon"]]] * _vec(a["gear"])[0]
            bias = _vec(a["biasprm"], 3, [0, 0, 0]) if a["biastype"] == "affine" else np.zeros(3)
            acts.append(dict(moment=moment, gain=_vec(a["gainprm"], 3, [1, 0, 0])[0], bias=bias,
                             ctrllimited=_bool(a["ctrllimited"]), ctrlrange=_vec(a["ctrlrange"]),
                             forcelimited=_bool(a["forcelimited"]), forcerange=_vec(a["forcerange"])))
    m["nu"] = len(acts)
    # actuator_length = moment @ qpos(dof-indexed hinge positions); velocity = moment @ qvel
    m["actuator_moment"] = np.array([a["moment"] for a in acts])
    m["actuator_gain"] = np.array([a["gain"] for a in acts])
    m["actuator_biasprm"] = np.array([a["bias"] for a in acts])
    m["actuator_ctrllimited"] = np.array([int(a["ctrllimited"]) for a in acts], dtype=np.int32)
    m["actuator_ctrlrange"] = np.array([a["ctrlrange"] for a in acts])
    m["actuator_forcelimited"] = np.array([int(a["forcelimited"]) for a in acts], dtype=np.int32)
    m["actuator_forcerange"] = np.array([a["forcerange"] for a in acts])

    # ---- keyframes
    keys = []
    kn = root.find("keyframe")
    if kn is not None:
        for k in kn.findall("key"):
            keys.append(dict(qpos=_vec(k.get("qpos")), qvel=_vec(k.get("qvel")),
                             ctrl=_vec(k.get("ctrl")) if k.get("ctrl") else np.zeros(len(acts)),
                             mpos=_vec(k.get("mpos")) if k.get("mpos") else None,
                             mquat=_vec(k.get("mquat")) if k.get("mquat") else None))
    m["nkey"] = len(keys)
    m["key_qpos"] = np.array([k["qpos"] for k in keys]).reshape(len(keys), nq)
    m["key_qvel"] = np.array([k["qvel"] for k in keys]).reshape(len(keys), nv)
    m["key_ctrl"] = np.array([k["ctrl"] for k in keys]).reshape(len(keys), len(acts))
    mb = [i for i, b in enumerate(bodies) if b["mocap"]]
    m["key_mpos"] = np.array([k["mpos"] if k["mpos"] is not None else np.concatenate([bodies[i]["pos"] for i in mb] or [np.zeros(0)])
                              for k in keys]).reshape(len(keys), 3 * nmocap)
    m["key_mquat"] = np.array([k["mquat"] if k["mquat"] is not None else np.concatenate([bodies[i]["quat"] for i in mb] or [np.zeros(0)])
                               for k in keys]).reshape(len(keys), 4 * nmocap)

    set_const(m)
    m["compile_log"] = list(log)
    return m


# ----------------------------------------------------------------------------------
# mj_setConst restatement (engine_setconst.c set0): runs at qpos0 in numpy, independent of oracle/


def fk_numpy(m, qpos):
    """Body frames at qpos (restates mj_kinematics). Returns xpos, xquat, xmat, xipos, ximat, xanchor, xaxis."""
    nb = m["nbody"]
    xpos = np.zeros((nb, 3))
    xquat = np.zeros((nb, 4))
    xquat[0, 0] = 1
    xanchor = np.zeros((m["njnt"], 3))
    xaxis = np.zeros((m["njnt"], 3))
    for i in range(1, nb):
        pid = m["body_parentid"][i]
        ja, jn = m["body_jntadr"][i], m["body_jntnum"][i]
        if jn == 1 and m["jnt_type"][ja] == JNT_FREE:
            qa = m["jnt_qposadr"][ja]
            xpos[i] = qpos[qa:qa + 3]
            q = qpos[qa + 3:qa + 7]
            xquat[i] = q / np.linalg.norm(q)
            xanchor[ja] = xpos[i]
            xaxis[ja] = np.array([0, 0, 1.0])
            continue
        xpos[i] = xpos[pid] + quat2mat(xquat[pid]) @ m["body_pos"][i]
        xquat[i] = quat_mul(xquat[pid], m["body_quat"][i])
        for j in range(ja, ja + jn):
            assert m["jnt_type"][j] == JNT_HINGE
            xaxis[j] = quat2mat(xquat[i]) @ m["jnt_axis"][j]
            xanchor[j] = quat2mat(xquat[i]) @ m["jnt_pos"][j] + xpos[i]
            qa = m["jnt_qposadr"][j]
            xquat[i] = quat_mul(xquat[i], axisangle2quat(m["jnt_axis"][j], qpos[qa] - m["qpos0"][qa]))
            xpos[i] = xanchor[j] - quat2mat(xquat[i]) @ m["jnt_pos"][j]
        xquat[i] /= np.linalg.norm(xquat[i])
    xmat = np.array([quat2mat(q) for q in xquat])
    xipos = np.array([xpos[i] + xmat[i] @ m["body_ipos"][i] for i in range(nb)])
    ximat = np.array([quat2mat(quat_mul(xquat[i], m["body_iquat"][i])) for i in range(nb)])
    return xpos, xquat, xmat, xipos, ximat, xanchor, xaxis


def jac_point(m, fk, body, point):
    """3 x nv translational and rotational world-frame Jacobians of `point` fixed to `body`."""
    xpos, xquat, xmat, xipos, ximat, xanchor, xaxis = fk
    nv = m["nv"]
    jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
    b = body
    while b and m["body_dofnum"][b] == 0:
        b = m["body_parentid"][b]
    if b == 0:
        return jp, jr
    d = m["body_dofadr"][b] + m["body_dofnum"][b] - 1
    while d >= 0:
        j = m["dof_jntid"][d]
        k = d - m["jnt_dofadr"][j]
        bj = m["jnt_bodyid"][j]
        if m["jnt_type"][j] == JNT_HINGE:
            jr[:, d] = xaxis[j]
            jp[:, d] = np.cross(xaxis[j], point - xanchor[j])
        elif m["jnt_type"][j] == JNT_FREE:
            if k < 3:
                jp[k, d] = 1.0
            else:
                ax = xmat[bj][:, k - 3]
                jr[:, d] = ax
                jp[:, d] = np.cross(ax, point - xpos[bj])
        d = m["dof_parentid"][d]
    return jp, jr


def mass_matrix_numpy(m, fk):
    xpos, xquat, xmat, xipos, ximat, xanchor, xaxis = fk
    nv = m["nv"]
    M = np.diag(m["dof_armature"]).astype(np.float64)
    for b in range(1, m["nbody"]):
        if m["body_mass"][b] == 0 or m["body_weldid"][b] == 0:
            continue
        jp, jr = jac_point(m, fk, b, xipos[b])
        Iw = ximat[b] @ np.diag(m["body_inertia"][b]) @ ximat[b].T
        M += m["body_mass"][b] * jp.T @ jp + jr.T @ Iw @ jr
    return M


def set_const(m):
    """Fields computed by mj_setConst at qpos0: connect anchor2, body/dof invweight0, meaninertia."""
    fk = fk_numpy(m, m["qpos0"])
    xpos, xquat, xmat, xipos, ximat, xanchor, xaxis = fk
    for e in range(m["neq"]):
        if m["eq_type"][e] == EQ_CONNECT:
            b1, b2 = m["eq_obj1id"][e], m["eq_obj2id"][e]
            g = xpos[b1] + xmat[b1] @ m["eq_data"][e, :3]
            m["eq_data"][e, 3:6] = xmat[b2].T @ (g - xpos[b2])
        elif m["eq_type"][e] == EQ_WELD:
            # engine_setconst.c: anchor (data[0:3]) is in body2's frame; data[3:6] = the same point in body1's frame;
            # data[6:10] = orientation of body2 relative to body1 at qpos0, unless the user gave a quaternion
            b1, b2 = m["eq_obj1id"][e], m["eq_obj2id"][e]
            if np.all(m["eq_data"][e, 6:10] == 0):
                g = xpos[b2] + xmat[b2] @ m["eq_data"][e, 0:3]
                m["eq_data"][e, 3:6] = xmat[b1].T @ (g - xpos[b1])
                m["eq_data"][e, 6:10] = quat_mul(quat_conj(xquat[b1]), xquat[b2])
            else:
                m["eq_data"][e, 6:10] /= np.linalg.norm(m["eq_data"][e, 6:10])
    M = mass_matrix_numpy(m, fk)
    Minv = np.linalg.inv(M)
    nv = m["nv"]
    biw = np.zeros((m["nbody"], 2))
    for b in range(1, m["nbody"]):
        if m["body_weldid"][b] == 0:
            continue
        jp, jr = jac_point(m, fk, b, xipos[b])
        biw[b, 0] = max(mjMINVAL, np.trace(jp @ Minv @ jp.T) / 3)
        biw[b, 1] = max(mjMINVAL, np.trace(jr @ Minv @ jr.T) / 3)
    diw = np.zeros(nv)
    for j in range(m["njnt"]):
        d = m["jnt_dofadr"][j]
        if m["jnt_type"][j] == JNT_FREE:
            diw[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
            diw[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        else:
            diw[d] = Minv[d, d]
    m["body_invweight0"] = biw
    m["dof_invweight0"] = diw
    m["stat_meaninertia"] = float(np.trace(M) / nv)
    m["M0"] = M
    return m


# ----------------------------------------------------------------------------------


def flatmodel_from_mjmodel(mjm, mujoco=None):
    """Fill the same table from a live `mujoco.MjModel` -- the reference's compiled model (`MujocoEnv.__init__`,
    mycobot.py:69-75).  Field names follow mjModel.  Only plane / box geoms are kept as collision geoms (mesh hulls are
    a documented gap); `set_const()` is NOT re-run: invweight0 / meaninertia / connect anchors come from MuJoCo's own
    mj_setConst, so diffing this table against `compile_mjcf()` checks the mini-compiler field by field
    (`diff_flatmodels`).  `mujoco`: the module providing mj_id2name and the mjtObj / mjtWrap / mjtEq / mjtTrn enums
    (default: `import mujoco`); tests/test_model_compiler.py executes this function against a stand-in mjModel built
    from the compiled table (tests/fake_mujoco.py), since no mujoco wheel is installable offline."""
    if mujoco is None:
        import mujoco  # pragma: no cover - needs the real package

    m = FlatModel()
    nv, nq, nbody, njnt = mjm.nv, mjm.nq, mjm.nbody, mjm.njnt
    m["nq"], m["nv"], m["nbody"], m["njnt"] = int(nq), int(nv), int(nbody), int(njnt)
    m["timestep"] = float(mjm.opt.timestep)
    m["gravity"] = np.array(mjm.opt.gravity, dtype=np.float64)
    m["tolerance"], m["iterations"] = float(mjm.opt.tolerance), int(mjm.opt.iterations)
    m["ls_iterations"], m["ls_tolerance"], m["impratio"] = int(mjm.opt.ls_iterations), float(mjm.opt.ls_tolerance), float(mjm.opt.impratio)

    def name(objtype, i):
        return mujoco.mj_id2name(mjm, objtype, i) or ""

    m["body_names"] = [name(mujoco.mjtObj.mjOBJ_BODY, i) for i in range(nbody)]
    for k in ("body_parentid", "body_rootid", "body_weldid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr", "body_mocapid"):
        m[k] = np.array(getattr(mjm, k), dtype=np.int32)
    m["nmocap"] = int(mjm.nmocap)
    for k in ("body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_subtreemass"):
        m[k] = np.array(getattr(mjm, k), dtype=np.float64)
    m["body_invweight0"] = np.array(mjm.body_invweight0, dtype=np.float64).reshape(nbody, 2)
    m["jnt_names"] = [name(mujoco.mjtObj.mjOBJ_JOINT, i) for i in range(njnt)]
    for k in ("jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited"):
        m[k] = np.array(getattr(mjm, k), dtype=np.int32)
    for k in ("jnt_pos", "jnt_axis", "jnt_range", "jnt_margin", "jnt_solref", "jnt_solimp"):
        m[k] = np.array(getattr(mjm, k), dtype=np.float64)
    for k in ("dof_bodyid", "dof_jntid", "dof_parentid", "dof_Madr"):
        m[k] = np.array(getattr(mjm, k), dtype=np.int32)
    for k in ("dof_armature", "dof_damping", "dof_invweight0"):
        m[k] = np.array(getattr(mjm, k), dtype=np.float64)
    m["nM"] = int(mjm.nM)
    m["qpos0"] = np.array(mjm.qpos0, dtype=np.float64)
    keep = [g for g in range(mjm.ngeom) if mjm.geom_type[g] in (GEOM_TYPES["plane"], GEOM_TYPES["box"])
            and (mjm.geom_contype[g] or mjm.geom_conaffinity[g])]
    m["ngeom"] = len(keep)
    m["geom_names"] = [name(mujoco.mjtObj.mjOBJ_GEOM, g) for g in keep]
    for k, dt in (("geom_type", np.int32), ("geom_bodyid", np.int32), ("geom_contype", np.int32), ("geom_conaffinity", np.int32),
                  ("geom_condim", np.int32), ("geom_pos", np.float64), ("geom_quat", np.float64), ("geom_size", np.float64),
                  ("geom_friction", np.float64), ("geom_solref", np.float64), ("geom_solimp", np.float64), ("geom_solmix", np.float64),
                  ("geom_margin", np.float64), ("geom_gap", np.float64), ("geom_rbound", np.float64)):
        m[k] = np.array(getattr(mjm, k), dtype=dt)[keep]
    # convex hulls of the collidable mesh geoms: hull vertex ids from mesh_graph (numvert, numface, vert_edgeadr[numvert],
    # vert_globalid[numvert], ...), vertices moved from the (recentred, principal-axes) mesh frame into the body frame with the
    # geom's pose; identical copies on one body fold into a multiplicity like in compile_mjcf()
    hulls = {}
    for g in range(mjm.ngeom):
        if mjm.geom_type[g] != GEOM_TYPES["mesh"] or not (mjm.geom_contype[g] or mjm.geom_conaffinity[g]):
            continue
        mid = int(mjm.geom_dataid[g])
        key = (int(mjm.geom_bodyid[g]), mid, tuple(np.round(np.array(mjm.geom_pos[g]), 12)), tuple(np.round(np.array(mjm.geom_quat[g]), 12)))
        if key in hulls:
            hulls[key]["mult"] += 1
            continue
        va, vn = int(mjm.mesh_vertadr[mid]), int(mjm.mesh_vertnum[mid])
        verts = np.array(mjm.mesh_vert[va:va + vn], dtype=np.float64).reshape(vn, 3)
        ga = int(mjm.mesh_graphadr[mid])
        if ga >= 0:
            nvh = int(mjm.mesh_graph[ga])
            ids = np.array(mjm.mesh_graph[ga + 2 + nvh:ga + 2 + 2 * nvh], dtype=np.int64)
            verts = verts[np.sort(ids)]
        R = quat2mat(np.array(mjm.geom_quat[g], dtype=np.float64))
        pos = np.array(mjm.geom_pos[g], dtype=np.float64)
        hv = verts @ R.T + pos
        hulls[key] = dict(g=g, mult=1, vert=hv, center=pos, rbound=float(np.linalg.norm(hv - pos, axis=1).max()))
    hl = list(hulls.values())
    m["nhull"] = len(hl)
    m["hull_names"] = [name(mujoco.mjtObj.mjOBJ_BODY, int(mjm.geom_bodyid[h["g"]])) for h in hl]
    m["hull_bodyid"] = np.array([mjm.geom_bodyid[h["g"]] for h in hl], dtype=np.int32)
    m["hull_mult"] = np.array([h["mult"] for h in hl], dtype=np.int32)
    m["hull_vertnum"] = np.array([len(h["vert"]) for h in hl], dtype=np.int32)
    m["hull_vertadr"] = np.concatenate(([0], np.cumsum(m["hull_vertnum"])[:-1])).astype(np.int32) if hl else np.zeros(0, dtype=np.int32)
    m["hull_vert"] = np.concatenate([h["vert"] for h in hl]) if hl else np.zeros((0, 3))
    m["hull_center"] = np.array([h["center"] for h in hl]).reshape(len(hl), 3)
    m["hull_rbound"] = np.array([h["rbound"] for h in hl])
    m["hull_condim"] = np.array([mjm.geom_condim[h["g"]] for h in hl], dtype=np.int32)
    m["hull_friction"] = np.array([mjm.geom_friction[h["g"]] for h in hl], dtype=np.float64).reshape(len(hl), 3)
    m["hull_solref"] = np.array([mjm.geom_solref[h["g"]] for h in hl], dtype=np.float64).reshape(len(hl), 2)
    m["hull_solimp"] = np.array([mjm.geom_solimp[h["g"]] for h in hl], dtype=np.float64).reshape(len(hl), 5)
    m["hull_solmix"] = np.array([mjm.geom_solmix[h["g"]] for h in hl], dtype=np.float64)
    m["nsite"] = int(mjm.nsite)
    m["site_names"] = [name(mujoco.mjtObj.mjOBJ_SITE, i) for i in range(mjm.nsite)]
    m["site_bodyid"] = np.array(mjm.site_bodyid, dtype=np.int32)
    m["site_pos"], m["site_quat"] = np.array(mjm.site_pos, dtype=np.float64), np.array(mjm.site_quat, dtype=np.float64)
    ex = [sorted((int(sig) >> 16, int(sig) & 0xFFFF)) for sig in mjm.exclude_signature]
    m["exclude"] = np.array(ex, dtype=np.int32).reshape(-1, 2)
    m["ntendon"] = int(mjm.ntendon)
    ten_J = np.zeros((mjm.ntendon, nv))
    for t in range(mjm.ntendon):
        for w in range(mjm.tendon_adr[t], mjm.tendon_adr[t] + mjm.tendon_num[t]):
            assert mjm.wrap_type[w] == mujoco.mjtWrap.mjWRAP_JOINT
            ten_J[t, mjm.jnt_dofadr[mjm.wrap_objid[w]]] = mjm.wrap_prm[w]
    m["ten_J"] = ten_J
    m["neq"] = int(mjm.neq)
    m["eq_type"] = np.array([{int(mujoco.mjtEq.mjEQ_CONNECT): EQ_CONNECT, int(mujoco.mjtEq.mjEQ_WELD): EQ_WELD,
                              int(mujoco.mjtEq.mjEQ_JOINT): EQ_JOINT}[int(t)] for t in mjm.eq_type], dtype=np.int32)
    m["eq_obj1id"], m["eq_obj2id"] = np.array(mjm.eq_obj1id, dtype=np.int32), np.array(mjm.eq_obj2id, dtype=np.int32)
    m["eq_data"] = np.array(mjm.eq_data, dtype=np.float64)[:, :11]
    m["eq_solref"], m["eq_solimp"] = np.array(mjm.eq_solref, dtype=np.float64), np.array(mjm.eq_solimp, dtype=np.float64)
    nu = mjm.nu
    m["nu"] = int(nu)
    moment = np.zeros((nu, nv))
    for a in range(nu):
        tid = int(mjm.actuator_trnid[a, 0])
        if mjm.actuator_trntype[a] == mujoco.mjtTrn.mjTRN_JOINT:
            moment[a, mjm.jnt_dofadr[tid]] = mjm.actuator_gear[a, 0]
        elif mjm.actuator_trntype[a] == mujoco.mjtTrn.mjTRN_TENDON:
            moment[a] = ten_J[tid] * mjm.actuator_gear[a, 0]
        else:
            raise NotImplementedError("actuator transmission")
    m["actuator_moment"] = moment
    m["actuator_gain"] = np.array(mjm.actuator_gainprm[:, 0], dtype=np.float64)
    m["actuator_biasprm"] = np.array(mjm.actuator_biasprm[:, :3], dtype=np.float64)
    m["actuator_ctrllimited"] = np.array(mjm.actuator_ctrllimited, dtype=np.int32)
    m["actuator_ctrlrange"] = np.array(mjm.actuator_ctrlrange, dtype=np.float64)
    m["actuator_forcelimited"] = np.array(mjm.actuator_forcelimited, dtype=np.int32)
    m["actuator_forcerange"] = np.array(mjm.actuator_forcerange, dtype=np.float64)
    m["nkey"] = int(mjm.nkey)
    m["key_qpos"] = np.array(mjm.key_qpos, dtype=np.float64).reshape(mjm.nkey, nq)
    m["key_qvel"] = np.array(mjm.key_qvel, dtype=np.float64).reshape(mjm.nkey, nv)
    m["key_ctrl"] = np.array(mjm.key_ctrl, dtype=np.float64).reshape(mjm.nkey, nu)
    m["key_mpos"] = np.array(mjm.key_mpos, dtype=np.float64).reshape(mjm.nkey, 3 * mjm.nmocap)
    m["key_mquat"] = np.array(mjm.key_mquat, dtype=np.float64).reshape(mjm.nkey, 4 * mjm.nmocap)
    m["stat_meaninertia"] = float(mjm.stat.meaninertia)
    m["M0"] = mass_matrix_numpy(m, fk_numpy(m, m["qpos0"]))
    m["compile_log"] = ["filled from a live mujoco.MjModel; mesh geoms dropped from collision"]
    return m


def diff_flatmodels(a, b, rtol=1e-9):
    """Field-by-field comparison of two FlatModels (mini-compiler vs live mjModel).  Returns {field: max abs diff}."""
    out = {}
    for k in sorted(set(a) | set(b)):
        if k not in a or k not in b:
            out[k] = "missing"
            continue
        va, vb = a[k], b[k]
        if isinstance(va, np.ndarray):
            if va.shape != np.asarray(vb).shape:
                out[k] = f"shape {va.shape} vs {np.asarray(vb).shape}"
            elif va.size and not np.allclose(va, vb, rtol=rtol, atol=1e-12):
                out[k] = float(np.abs(va - vb).max())
        elif isinstance(va, float):
            if abs(va - vb) > rtol * max(1.0, abs(va)):
                out[k] = abs(va - vb)
        elif va != vb and k != "compile_log":
            out[k] = (va, vb)
    return out


ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
COMPILED_JOINT = os.path.join(ASSET_DIR, "mycobot280_joint.json")     # mycobot280.xml (joint and IK controllers)
COMPILED_MOCAP = os.path.join(ASSET_DIR, "mycobot280_mocap.json")     # mycobot280_mocap.xml (mocap controller)


def load_compiled(path=COMPILED_JOINT):
    return FlatModel.from_json(path)

"""Stand-in for the parts of the `mujoco` package `mjcf.flatmodel_from_mjmodel` touches: an MjModel-shaped namespace built
from a compiled FlatModel (arrays laid out the way mjModel lays them out: flattened keyframes, packed exclude signatures,
tendon wrap arrays, actuator transmission tables, mesh geoms interleaved with the primitives) plus the enums and
mj_id2name.  It exists so that the live-mjModel loader path EXECUTES in CI (MuJoCo 2.3.2 is not installable offline); it
cannot and does not validate MuJoCo's own compiler."""
import types

import numpy as np


class _Enum(types.SimpleNamespace):
    pass


mjtObj = _Enum(mjOBJ_BODY=1, mjOBJ_JOINT=3, mjOBJ_GEOM=5, mjOBJ_SITE=6)
mjtWrap = _Enum(mjWRAP_JOINT=1)
mjtEq = _Enum(mjEQ_CONNECT=0, mjEQ_WELD=1, mjEQ_JOINT=2)
mjtTrn = _Enum(mjTRN_JOINT=0, mjTRN_TENDON=3)
GEOM_MESH = 7


def mj_id2name(model, objtype, i):
    table = {mjtObj.mjOBJ_BODY: model._body_names, mjtObj.mjOBJ_JOINT: model._jnt_names, mjtObj.mjOBJ_GEOM: model._geom_names,
             mjtObj.mjOBJ_SITE: model._site_names}[objtype]
    return table[i] or None


def model_from_flat(f):
    """MjModel-shaped object carrying the FlatModel's content, mesh geoms (two identical ones per robot body, like the real
    model) and their meshes included."""
    m = types.SimpleNamespace()
    for k in ("nq", "nv", "nbody", "njnt", "nsite", "ntendon", "neq", "nu", "nkey", "nmocap", "nM"):
        setattr(m, k, int(f[k]))
    m.opt = types.SimpleNamespace(timestep=f["timestep"], gravity=np.array(f["gravity"]), tolerance=f["tolerance"], iterations=f["iterations"],
                                  ls_iterations=f["ls_iterations"], ls_tolerance=f["ls_tolerance"], impratio=f["impratio"])
    m.stat = types.SimpleNamespace(meaninertia=f["stat_meaninertia"])
    for k in ("body_parentid", "body_rootid", "body_weldid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr", "body_mocapid",
              "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_subtreemass", "body_invweight0",
              "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited", "jnt_pos", "jnt_axis", "jnt_range", "jnt_margin",
              "jnt_solref", "jnt_solimp", "dof_bodyid", "dof_jntid", "dof_parentid", "dof_Madr", "dof_armature", "dof_damping",
              "dof_invweight0", "qpos0", "site_bodyid", "site_pos", "site_quat", "eq_obj1id", "eq_obj2id", "eq_solref", "eq_solimp",
              "actuator_ctrllimited", "actuator_ctrlrange", "actuator_forcelimited", "actuator_forcerange"):
        setattr(m, k, np.array(f[k]))
    m._body_names, m._jnt_names, m._site_names = list(f["body_names"]), list(f["jnt_names"]), list(f["site_names"])
    # geoms: the primitives of the table, then every hull as `mult` identical mesh geoms (type 7, collidable, no name) whose
    # mesh holds the hull vertices + three interior points the loader must discard through mesh_graph's vert_globalid
    ng, nh = int(f["ngeom"]), int(f["nhull"])
    order = [("p", g) for g in range(ng)]
    for h in range(nh):
        order += [("h", h)] * int(np.array(f["hull_mult"])[h])
    m.ngeom = len(order)
    m._geom_names = [f["geom_names"][i] if kind == "p" else "" for kind, i in order]
    hull_src = {"geom_bodyid": "hull_bodyid", "geom_condim": "hull_condim", "geom_friction": "hull_friction", "geom_solref": "hull_solref",
                "geom_solimp": "hull_solimp", "geom_solmix": "hull_solmix"}

    def geom_field(k, fill):
        src = np.array(f[k])
        out = np.zeros((len(order),) + src.shape[1:], dtype=src.dtype)
        for i, (kind, g) in enumerate(order):
            if kind == "p":
                out[i] = src[g]
            elif k in hull_src:
                out[i] = np.array(f[hull_src[k]])[g]
            elif k == "geom_pos":
                out[i] = np.array(f["hull_center"])[g]        # MuJoCo puts the mesh geom's frame at the mesh's centre
            else:
                out[i] = fill
        return out

    m.geom_type = geom_field("geom_type", GEOM_MESH)
    for k, fill in (("geom_bodyid", 4), ("geom_contype", 1), ("geom_conaffinity", 1), ("geom_condim", 3), ("geom_pos", 0.0), ("geom_quat", [1.0, 0, 0, 0]),
                    ("geom_size", 0.01), ("geom_friction", 1.0), ("geom_solref", 0.02), ("geom_solimp", 0.9), ("geom_solmix", 1.0),
                    ("geom_margin", 0.0), ("geom_gap", 0.0), ("geom_rbound", 0.05)):
        setattr(m, k, geom_field(k, fill))
    m.geom_dataid = np.array([-1 if kind == "p" else g for kind, g in order], dtype=np.int32)
    vert, vadr, vnum, graph, gadr = [], [], [], [], []
    for h in range(nh):
        a, n = int(np.array(f["hull_vertadr"])[h]), int(np.array(f["hull_vertnum"])[h])
        hv = np.array(f["hull_vert"])[a:a + n] - np.array(f["hull_center"])[h]         # mesh frame = geom frame
        inner = np.zeros((3, 3)) + hv.mean(0)
        allv = np.concatenate([inner[:1], hv, inner[1:]])                                # hull vertices are ids 1 .. n
        vadr.append(sum(len(x) for x in vert)); vnum.append(len(allv)); vert.append(allv)
        gadr.append(len(graph))
        graph += [n, 0] + [0] * n + list(range(1, n + 1))
    m.mesh_vert = np.concatenate(vert).astype(np.float64) if vert else np.zeros((0, 3))
    m.mesh_vertadr, m.mesh_vertnum = np.array(vadr, dtype=np.int32), np.array(vnum, dtype=np.int32)
    m.mesh_graph, m.mesh_graphadr = np.array(graph, dtype=np.int32), np.array(gadr, dtype=np.int32)
    m.exclude_signature = np.array([(int(a) << 16) + int(b) for a, b in np.array(f["exclude"])], dtype=np.int64)
    # fixed tendons as wrap arrays
    adr, num, wtype, wobj, wprm = [], [], [], [], []
    for t in range(m.ntendon):
        adr.append(len(wtype))
        cols = np.nonzero(np.array(f["ten_J"])[t])[0]
        for c in cols:
            wtype.append(mjtWrap.mjWRAP_JOINT)
            wobj.append(int(np.array(f["dof_jntid"])[c]))
            wprm.append(float(np.array(f["ten_J"])[t, c]))
        num.append(len(cols))
    m.tendon_adr, m.tendon_num = np.array(adr, dtype=np.int32), np.array(num, dtype=np.int32)
    m.wrap_type, m.wrap_objid, m.wrap_prm = np.array(wtype), np.array(wobj), np.array(wprm)
    m.eq_type = np.array(f["eq_type"])                                    # the table already uses mjtEq's values
    m.eq_data = np.array(f["eq_data"])
    # actuators: joint transmission where the moment row has one entry, tendon transmission otherwise; gear = 1
    nu, moment = m.nu, np.array(f["actuator_moment"])
    m.actuator_trntype, m.actuator_trnid = np.zeros(nu, dtype=np.int32), np.zeros((nu, 2), dtype=np.int32)
    m.actuator_gear = np.zeros((nu, 6)); m.actuator_gear[:, 0] = 1.0
    for a in range(nu):
        cols = np.nonzero(moment[a])[0]
        if len(cols) == 1:
            m.actuator_trntype[a], m.actuator_trnid[a, 0] = mjtTrn.mjTRN_JOINT, int(np.array(f["dof_jntid"])[cols[0]])
            m.actuator_gear[a, 0] = moment[a, cols[0]]
        else:
            t = next(t for t in range(m.ntendon) if np.array_equal(np.array(f["ten_J"])[t] != 0, moment[a] != 0))
            m.actuator_trntype[a], m.actuator_trnid[a, 0] = mjtTrn.mjTRN_TENDON, t
    m.actuator_gainprm = np.zeros((nu, 10)); m.actuator_gainprm[:, 0] = np.array(f["actuator_gain"])
    m.actuator_biasprm = np.zeros((nu, 10)); m.actuator_biasprm[:, :3] = np.array(f["actuator_biasprm"])
    for k in ("key_qpos", "key_qvel", "key_ctrl", "key_mpos", "key_mquat"):
        setattr(m, k, np.array(f[k]).ravel())
    return m

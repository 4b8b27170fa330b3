"""The oracle's convex narrow phase (MPR restatement of libccd's ccdMPRPenetration as MuJoCo 2.3.2 calls it from mjc_Convex):
analytic cases and an independent separability check (linear programme) on hull pairs of the real model."""
import ctypes as C
import itertools
import random

import numpy as np
import pytest

from mycobotgym_b200 import mjcf
from oracle.oracle import OracleEnv, OracleSim, lib


def _mpr(va, vb):
    va, vb = np.ascontiguousarray(va, dtype=np.float64), np.ascontiguousarray(vb, dtype=np.float64)
    out = np.zeros(7)
    P = C.POINTER(C.c_double)
    hit = lib().o_test_mpr(va.ctypes.data_as(P), len(va), vb.ctypes.data_as(P), len(vb), out.ctypes.data_as(P))
    return bool(hit), out[0], out[1:4].copy(), out[4:7].copy()


CUBE = np.array(list(itertools.product((-1, 1), repeat=3)), dtype=np.float64)


def test_mpr_on_boxes_matches_the_overlap():
    for shift, depth, axis in (((1.5, 0.2, 0.1), 0.5, 0), ((0.3, -1.8, 0.2), 0.2, 1), ((0.1, 0.2, 1.9), 0.1, 2)):
        hit, d, n, p = _mpr(CUBE, CUBE + np.array(shift))
        assert hit and abs(d - depth) < 2e-6                                 # mpr_tolerance 1e-6
        want = np.zeros(3); want[axis] = np.sign(shift[axis])
        np.testing.assert_allclose(n, want, atol=1e-5)                       # from A to B
        lo, hi = np.maximum(-1, np.array(shift) - 1), np.minimum(1, np.array(shift) + 1)
        assert np.all(p >= lo - 1e-6) and np.all(p <= hi + 1e-6)             # inside the overlap region
    assert not _mpr(CUBE, CUBE + np.array([2.001, 0, 0]))[0]                 # separated
    assert not _mpr(CUBE, CUBE + np.array([1.5, 1.5, 2.2]))[0]


def test_mpr_on_a_sphere_like_hull():
    rng = np.random.default_rng(0)
    v = rng.normal(size=(400, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]                                   # unit sphere, 400 vertices
    for dist in (1.2, 1.7, 1.95):
        hit, d, n, p = _mpr(v, v + np.array([dist, 0, 0]))
        assert hit and abs(d - (2 - dist)) < 0.03 and n[0] > 0.97            # faceted spheres: a few per cent
    assert not _mpr(v, v + np.array([2.05, 0, 0]))[0]


def _separable(P, Q):
    from scipy.optimize import linprog

    A = np.vstack([np.hstack([P, -np.ones((len(P), 1))]), np.hstack([-Q, np.ones((len(Q), 1))])])
    r = linprog(np.zeros(4), A_ub=A, b_ub=-np.ones(len(P) + len(Q)), bounds=[(None, None)] * 4, method="highs")
    return r.status == 0


def test_hull_contacts_of_the_model_agree_with_a_separating_plane_lp():
    """States of the mocap workload (the arm folds onto itself and the table there): every hull-hull contact MPR reports is an
    intersecting pair by the LP, and candidate pairs whose bounding spheres overlap but that MPR rejects are LP-separable."""
    fm = mjcf.load_compiled(mjcf.COMPILED_MOCAP)
    env = OracleEnv(fm, has_object=True, reward_type="sparse", controller_type="mocap")
    env.sim.om.mesh_collision = 1
    rng = np.random.default_rng(1)
    random.seed(1)
    env.reset(seed=1)
    tcp = fm["body_names"].index("gripper_tcp")
    ng, nh = fm["ngeom"], fm["nhull"]
    checked_hit = checked_miss = 0
    for t in range(60):
        a = rng.uniform(-1, 1, 8).astype(np.float32)
        a[3:7] = env.sim.xquat[tcp] + 0.1 * rng.uniform(-1, 1, 4)
        env.step(a)
        if t % 2:
            continue
        s = env.sim
        W = []
        for h in range(nh):
            b, a0, n = fm["hull_bodyid"][h], fm["hull_vertadr"][h], fm["hull_vertnum"][h]
            W.append(fm["hull_vert"][a0:a0 + n] @ s.xmat[b].T + s.xpos[b])
        hits = {(c["geom1"] - ng, c["geom2"] - ng): c for c in s.contacts() if c["geom1"] >= ng and c["geom2"] >= ng}
        for (h1, h2), c in hits.items():
            if c["dist"] < -2e-4:                                              # clear of the LP's own tolerance
                assert not _separable(W[h1], W[h2]), (t, h1, h2, c["dist"])
                checked_hit += 1
        cen = [w.mean(0) for w in W]
        for h1, h2 in itertools.combinations(range(nh), 2):
            if (h1, h2) in hits or np.linalg.norm(cen[h1] - cen[h2]) > 0.06:
                continue
            b1, b2 = fm["hull_bodyid"][h1], fm["hull_bodyid"][h2]
            w1, w2 = fm["body_weldid"][b1], fm["body_weldid"][b2]
            par = fm["body_parentid"]
            excl = any(sorted((b1, b2)) == list(e) for e in fm["exclude"])
            if w1 == w2 or excl or fm["body_weldid"][par[w1]] == w2 or fm["body_weldid"][par[w2]] == w1:
                continue
            if _separable(W[h1] * 1.0, W[h2] * 1.0):
                checked_miss += 1
            else:                                                              # LP says they touch: MPR may only have missed a graze
                d = _mpr(W[h1], W[h2])
                assert not d[0] or d[1] < 2e-4, (t, h1, h2, d[1])
    assert checked_hit >= 10 and checked_miss >= 10, (checked_hit, checked_miss)


def test_twin_mesh_geoms_scale_the_regulariser():
    """Twin mesh geoms (every robot body carries its mesh twice) make MuJoCo emit every hull contact 2x (hull x primitive) or 4x
    (hull x hull); the oracle emits ONE contact whose rows have R / mult: k identical rows with regulariser R add up to k * D in
    the dual cost, so qacc is the same.  Checked on a hull x table contact: condim 3, mu = 1, default solref / solimp."""
    fm = mjcf.load_compiled(mjcf.COMPILED_JOINT)
    s = OracleSim(fm, mesh_collision=True)
    rng = np.random.default_rng(3)
    table = [g for g in range(fm["ngeom"]) if fm["geom_type"][g] == 6 and fm["geom_bodyid"][g] == list(fm["body_names"]).index("table")][0]
    mesh = []
    for _ in range(400):                                     # seeded poses until a link hull dips into the table top
        s.qpos[:6] = rng.uniform(-2.0, 2.0, 6)
        s.forward()
        mesh = [(i, c) for i, c in enumerate(s.contacts()) if c["geom2"] >= fm["ngeom"] and c["geom1"] == table and c["dim"] == 3]
        if mesh:
            break
    assert mesh, "no hull x table contact found"
    i, c = mesh[0]
    assert fm["hull_mult"][c["geom2"] - fm["ngeom"]] == 2
    rows_before = s.nefc - sum(2 * (k["dim"] - 1) for k in s.contacts()[i:])
    R = s.efc("R")[rows_before:rows_before + 4]
    b = fm["hull_bodyid"][c["geom2"] - fm["ngeom"]]
    tran = fm["body_invweight0"][b][0]                   # the table is static
    imp = s.efc("KBIP")[rows_before, 2]
    np.testing.assert_allclose(R, 2 * 1.0 * (1 - imp) / imp * tran * 2.0 / 2, rtol=1e-12)     # 2 mu^2 R0 with R0 = (1-d)/d * tran (1 + mu^2), halved

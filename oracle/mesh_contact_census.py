"""TEST INFRASTRUCTURE (development-time analysis, needs /root/reference and scipy): how often would the reference's
30 mesh geoms (convex hulls; contype = conaffinity = 1, mycobot280_main.xml:104-250) be in contact on the benchmark's
state distribution?  The CUDA engine and the oracle collide plane / box primitives only, so this
census bounds what that omission can change.

States: oracle rollouts of the pick-and-place env under the benchmark's action distribution (uniform float32
U[-1,1]^7 joint targets, 50-step episodes).  Checks per state:
  * hull vertices of every link mesh against the table top / floor half-spaces and inside the cube box;
  * hull-vs-hull separability (LP) for every mesh pair MuJoCo's filters keep (no shared weld body, not parent-child,
    not in <contact><exclude>).
Result recorded in DESIGN.md: 0 / 2000 states with a mesh-vs-table/floor/cube contact, 0 / 80 states with a
mesh-vs-mesh contact (base_link.STL is absent from the reference checkout; it is static and sits on the table).
With the IK controller (`python oracle/mesh_contact_census.py IK`, random 7-d actions): 0 / 300 states -- the finger-layer
boxes reach the table first.  Round 2 adds the modes `push` (block_gripper), `mocap` (random mocap targets) and `grasp`
(the committed grasp state lifted with the grip closed) and prints the contacts per (mesh, partner) pair; a vertex test
against the cube box under-counts face-face hull contacts, so `grasp` is a lower bound.  Static bodies are handled by the
table / floor half-space tests, i.e. MuJoCo's "welded to world" exemption from the parent-child filter is honoured there
(link1 x table is tested); link1 x base_link cannot be: base_link.STL is absent from the reference checkout.
"""
import itertools
import random
import sys
import xml.etree.ElementTree as ET

import numpy as np

A = "/root/reference/mycobotgym/envs/assets/"


def main(episodes=40, mesh_mesh_every=5, controller="joint"):
    from scipy.optimize import linprog
    from scipy.spatial import ConvexHull

    sys.path.insert(0, __file__.rsplit("/oracle/", 1)[0])
    from mycobotgym_b200 import mjcf
    from oracle.oracle import OracleEnv

    root = ET.parse(A + "mycobot280_main.xml").getroot()
    excl = {frozenset((e.get("body1"), e.get("body2"))) for e in root.iter("exclude")}
    geoms = []

    def walk(b):
        for g in b.findall("geom"):
            if g.get("type") == "mesh" and g.get("group") != "1":          # the group-1 copy has the same shape
                geoms.append((b.get("name"), g.get("mesh")))
        for c in b.findall("body"):
            walk(c)

    for b in root.find("worldbody").findall("body"):
        walk(b)
    flat = mjcf.load_compiled()
    names = list(flat["body_names"])
    hulls = {}
    for bn, mn in geoms:
        try:
            tri = mjcf.read_stl(A + "meshes/" + mn + ".STL")
        except Exception:
            print("missing mesh file:", mn)
            continue
        v = np.asarray(tri).reshape(-1, 3)
        hulls[bn] = v[ConvexHull(v).vertices]
    par, weld = flat["body_parentid"], flat["body_weldid"]

    def filtered(b1, b2):
        i, j = names.index(b1), names.index(b2)
        wi, wj = weld[i], weld[j]
        return frozenset((b1, b2)) in excl or wi == wj or weld[par[wi]] == wj or weld[par[wj]] == wi

    pairs = [(a, b) for a, b in itertools.combinations(hulls, 2) if not filtered(a, b)]

    def intersect(P, Q):            # separable iff some (n, d) has n.p - d <= -1 and n.q - d >= 1
        Aub = np.vstack([np.hstack([P, -np.ones((len(P), 1))]), np.hstack([-Q, np.ones((len(Q), 1))])])
        r = linprog(np.zeros(4), A_ub=Aub, b_ub=-np.ones(len(P) + len(Q)), bounds=[(None, None)] * 4, method="highs")
        return r.status != 0

    kw = dict(has_object=True, reward_type="sparse")
    adim = 7
    if controller == "push":
        kw.update(block_gripper=True, target_in_the_air=False)
    elif controller == "mocap":
        flat_m = mjcf.load_compiled(mjcf.COMPILED_MOCAP)
        kw.update(controller_type="mocap")
        adim = 8
    elif controller != "grasp":
        kw.update(controller_type=controller)
    env = OracleEnv(flat_m if controller == "mocap" else flat, **kw)
    names_e = list(env.flat["body_names"])
    rng = np.random.default_rng(0)
    random.seed(0)
    cube_b = names_e.index("object0")
    n_states = n_prim = n_mm_states = n_mm = 0
    by_pair = {}
    grasp = None
    if controller == "grasp":
        import os
        grasp = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "grasp_pick_sparse.npz"))
    for ep in range(episodes):
        env.reset(seed=ep)
        if grasp is not None:      # a held cube: the committed grasp state, then lift with the grip closed (what a policy does)
            env.sim.set_state(grasp["qpos0"], grasp["qvel0"], grasp["ctrl0"], grasp["warm0"])
        for t in range(50):
            if grasp is not None:
                a = np.zeros(7, dtype=np.float32)
                a[:6] = grasp["qpos0"][:6] + 0.02 * ep * rng.uniform(-1, 1, 6)
                a[1] -= 0.004 * t                  # raise the shoulder slowly
                a[6] = 0.8
            else:
                a = rng.uniform(-1, 1, adim).astype(np.float32)
                if controller == "mocap":
                    a[3:7] = env.sim.xquat[names_e.index("gripper_tcp")] + 0.1 * rng.uniform(-1, 1, 4)
            env.step(a)
            s = env.sim
            W = {bn: hv @ s.xmat[names_e.index(bn)].T + s.xpos[names_e.index(bn)] for bn, hv in hulls.items()}
            hit = False
            for bn, w in W.items():
                on_table = (np.abs(w[:, 0]) < 0.2) & (np.abs(w[:, 1]) < 0.25) & (w[:, 2] < 0.2) & (w[:, 2] > 0.0)
                loc = (w - s.xpos[cube_b]) @ s.xmat[cube_b]
                in_cube = (np.abs(loc) < 0.01).all(axis=1).any()
                below = (w[:, 2] < 0).any()
                for kind, flag in (("table", on_table.any()), ("cube", in_cube), ("floor", below)):
                    if flag:
                        by_pair[(bn, kind)] = by_pair.get((bn, kind), 0) + 1
                hit |= bool(on_table.any() or below or in_cube)
            n_states += 1
            n_prim += hit
            if t % mesh_mesh_every == 0 and ep < 8:
                C = {bn: (w.mean(0), np.linalg.norm(w - w.mean(0), axis=1).max()) for bn, w in W.items()}
                mm = False
                for a_, b_ in pairs:
                    if np.linalg.norm(C[a_][0] - C[b_][0]) <= C[a_][1] + C[b_][1] and intersect(W[a_], W[b_]):
                        mm = True
                        by_pair[(a_, b_)] = by_pair.get((a_, b_), 0) + 1
                n_mm_states += 1
                n_mm += mm
    print(f"[{controller}] mesh vs table/floor/cube: {n_prim} / {n_states} states;  mesh vs mesh ({len(pairs)} pairs): {n_mm} / {n_mm_states} states")
    for k, v in sorted(by_pair.items(), key=lambda kv: -kv[1]):
        print(f"    {k[0]:>20s} x {k[1]:<20s} {v}")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "joint"
    main(episodes={"joint": 40, "IK": 6, "push": 12, "mocap": 12, "grasp": 6}[mode], controller=mode)

"""Generate tests/golden/*.npz from the CPU oracle (committed fixtures; regenerate only deliberately).

The reference itself (MuJoCo 2.3.2) cannot be run in this image, so these vectors pin the ORACLE's
behaviour (regression) and give the GPU tests a box-independent target; they are not MuJoCo outputs.
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from mycobotgym_b200 import mjcf  # noqa: E402
from oracle.oracle import OracleEnv  # noqa: E402

flat = mjcf.load_compiled()
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def rollout(name, has_object, block_gripper, reward_type, seed, nsteps, perturb):
    rng = np.random.default_rng(seed)
    env = OracleEnv(flat, has_object=has_object, block_gripper=block_gripper, reward_type=reward_type)
    random.seed(seed)
    env.reset(seed=seed)
    if perturb:
        env.sim.qpos[:6] = rng.uniform(-0.8, 0.8, 6)
        env.sim.qvel[:6] = rng.normal(size=6) * 0.3
        env.sim.forward()
    st0 = env.sim.get_state()
    rec = dict(qpos0=st0["qpos"], qvel0=st0["qvel"], ctrl0=st0["ctrl"], warm0=st0["qacc_warmstart"], goal=env.goal.copy(),
               actions=[], qpos=[], qvel=[], warm=[], obs=[], ag=[], reward=[], terminated=[], truncated=[], success=[])
    for t in range(nsteps):
        a = rng.uniform(-1, 1, 7).astype(np.float32)
        o, r, te, tr, info = env.step(a)
        rec["actions"].append(a)
        rec["qpos"].append(env.sim.qpos.copy())
        rec["qvel"].append(env.sim.qvel.copy())
        rec["warm"].append(env.sim.qacc_warmstart.copy())
        rec["obs"].append(o["observation"])
        rec["ag"].append(o["achieved_goal"])
        rec["reward"].append(np.float64(r))
        rec["terminated"].append(te)
        rec["truncated"].append(tr)
        rec["success"].append(info["is_success"])
    np.savez(os.path.join(OUT, name + ".npz"), has_object=has_object, block_gripper=block_gripper, reward_type=reward_type,
             **{k: np.asarray(v) for k, v in rec.items()})
    print(name, "final qpos[:6]", rec["qpos"][-1][:6])


def joint_rollouts():
    rollout("reach_dense_seed0", False, False, "dense", 0, 6, False)
    rollout("reach_dense_seed1_perturbed", False, False, "dense", 1, 6, True)
    rollout("pick_sparse_seed0", True, False, "sparse", 0, 6, False)
    rollout("pick_sparse_seed4_perturbed", True, False, "sparse", 4, 6, True)
    rollout("push_sparse_seed2", True, True, "sparse", 2, 6, False)


def grasp():
    """Cube held between the two finger-layer boxes in mid-air (BASELINE config 4 'extra contact-parity set')."""
    from mycobotgym_b200.mjcf import mat2quat
    from oracle.oracle import OracleSim

    s = OracleSim(flat)
    s.ctrl[6] = 0.5
    for _ in range(30):
        s.step(20)
        s.qpos[12:15] = [0.3, 0.3, 1.5]
        s.qpos[15:19] = [1, 0, 0, 0]
        s.qvel[12:18] = 0
    s.forward()
    gp = s.geom_xpos.copy()
    R = s._arr("geom_xmat", 45).reshape(5, 3, 3).copy()
    s.qpos[12:15] = 0.5 * (gp[2] + gp[3])
    s.qpos[15:19] = mat2quat(R[2])
    s.qvel[12:18] = 0
    s.ctrl[6] = 0.8
    s.step(60)
    st = s.get_state()
    env = OracleEnv(flat, has_object=True, reward_type="sparse")
    env.sim.set_state(**st)
    env.goal = np.array([0.05, 0.02, 0.3])
    rec = dict(qpos0=st["qpos"], qvel0=st["qvel"], ctrl0=st["ctrl"], warm0=st["qacc_warmstart"], goal=env.goal.copy(),
               actions=[], qpos=[], qvel=[], warm=[], obs=[], ag=[], reward=[], terminated=[], truncated=[], success=[], ncon=[], nefc=[])
    for t in range(3):
        a = np.array([0, 0, 0, 0, 0, 0, 0.8], dtype=np.float32)
        o, r, te, tr, info = env.step(a)
        rec["actions"].append(a); rec["qpos"].append(env.sim.qpos.copy()); rec["qvel"].append(env.sim.qvel.copy())
        rec["warm"].append(env.sim.qacc_warmstart.copy()); rec["obs"].append(o["observation"]); rec["ag"].append(o["achieved_goal"])
        rec["reward"].append(np.float64(r)); rec["terminated"].append(te); rec["truncated"].append(tr); rec["success"].append(info["is_success"])
        rec["ncon"].append(env.sim.ncon); rec["nefc"].append(env.sim.nefc)
    assert min(rec["ncon"]) >= 3, rec["ncon"]
    np.savez(os.path.join(OUT, "grasp_pick_sparse.npz"), has_object=True, block_gripper=False, reward_type="sparse",
             **{k: np.asarray(v) for k, v in rec.items()})
    print("grasp ncon", rec["ncon"], "nefc", rec["nefc"])



def controller_rollout(name, controller, fetch, seed, nsteps=3):
    """IK / mocap controllers: every step is recorded with the state it starts from, including the arm
    configuration the reference's stale frames belong to (`qprev`) and data.mocap_pos / mocap_quat, plus the oracle's own
    sensitivity to a one-ulp perturbation of the arm velocities (the bang-bang actuators make 100 substeps chaotic)."""
    fm = mjcf.load_compiled(mjcf.COMPILED_MOCAP) if controller == "mocap" else flat
    tight = mjcf.FlatModel(fm)
    tight["tolerance"] = 1e-13
    kw = dict(has_object=True, reward_type="dense", controller_type=controller, fetch_env=fetch)
    env, env2 = OracleEnv(fm, **kw), OracleEnv(tight, **kw)
    rng = np.random.default_rng(seed)
    random.seed(seed)
    env.reset(seed=seed)
    adim = 4 if fetch else (8 if controller == "mocap" else 7)
    tcp = fm["body_names"].index("gripper_tcp") if controller == "mocap" else None
    last_pre = {}

    def wrap(e, key):
        orig = e.sim.step

        def step_rec(n):
            orig(n - 1)
            last_pre[key] = e.sim.qpos.copy()
            orig(1)
        e.sim.step = step_rec

    wrap(env, "a"); wrap(env2, "b")
    rec = {k: [] for k in ["qpos0", "qvel0", "ctrl0", "warm0", "qprev0", "qstale0", "mocap0", "actions", "qpos", "qvel", "ctrl", "warm", "qprev", "mocap", "obs",
                           "reward", "sens"]}
    qprev = env.sim.qpos[:6].copy()                     # frames are fresh after reset
    qstale = env.sim.qpos.copy()                        # full previous-substep configuration: bit-exact restarts of the oracle
    for t in range(nsteps):
        s = env.sim
        rec["qpos0"].append(s.qpos.copy()); rec["qvel0"].append(s.qvel.copy()); rec["ctrl0"].append(s.ctrl.copy())
        rec["warm0"].append(s.qacc_warmstart.copy()); rec["qprev0"].append(qprev.copy()); rec["qstale0"].append(qstale.copy())
        rec["mocap0"].append(np.concatenate((s.mocap_pos, s.mocap_quat)))
        # the perturbed twin starts from the same state with stale frames re-created the same way the tests do it
        for e, scale in ((env2, 1 + 2.2e-16),):
            q_stale = qstale.copy()
            e.sim.set_state(q_stale, s.qvel, s.ctrl, s.qacc_warmstart)
            e.sim.kinematics()
            e.sim.qpos[:] = s.qpos
            e.sim.mocap_pos[:] = s.mocap_pos; e.sim.mocap_quat[:] = s.mocap_quat
            e.sim.qvel[:6] *= scale
            e.goal = env.goal.copy()
        a = rng.uniform(-1, 1, adim).astype(np.float32)
        if controller == "mocap" and not fetch and t % 2 == 0:
            a[3:7] = (s.xquat[tcp] + 0.1 * rng.uniform(-1, 1, 4)).astype(np.float32)
        o, r, te, tr, info = env.step(a)
        env2.step(a)
        qstale = last_pre["a"].copy()
        qprev = qstale[:6].copy()
        rec["actions"].append(a); rec["qpos"].append(s.qpos.copy()); rec["qvel"].append(s.qvel.copy()); rec["ctrl"].append(s.ctrl.copy())
        rec["warm"].append(s.qacc_warmstart.copy()); rec["qprev"].append(qprev.copy())
        rec["mocap"].append(np.concatenate((s.mocap_pos, s.mocap_quat)))
        rec["obs"].append(o["observation"]); rec["reward"].append(np.float64(r))
        rec["sens"].append(np.abs(env2.sim.qpos - s.qpos).max())
    np.savez(os.path.join(OUT, name + ".npz"), controller=controller, fetch=fetch, goal=env.goal.copy(),
             **{k: np.asarray(v) for k, v in rec.items()})
    print(name, "sens", rec["sens"], "final arm", rec["qpos"][-1][:6])


def controller_rollouts():
    controller_rollout("ik_pick_dense_seed3", "IK", False, 3)
    controller_rollout("ik_fetch_pick_dense_seed5", "IK", True, 5)
    controller_rollout("mocap_pick_dense_seed6", "mocap", False, 6)
    controller_rollout("mocap_fetch_pick_dense_seed7", "mocap", True, 7)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"      # all | joint | grasp | controllers
    if which in ("all", "joint"):
        joint_rollouts()
    if which in ("all", "grasp"):
        grasp()
    if which in ("all", "controllers"):
        controller_rollouts()

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


grasp()

"""GPU-side diagnosis of states kept by tests/test_gpu_rollout_parity.py: stage-level tap (mcb_debug_forward) vs the oracle's
forward pass on the same state, substep by substep (the GPU state is re-injected from the oracle's trajectory)."""
import sys, os
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from mycobotgym_b200 import mjcf
from mycobotgym_b200.vector_env import MyCobotVectorEnv
from oracle.oracle import OracleSim

wl = sys.argv[1] if len(sys.argv) > 1 else "push"
f = np.load(os.path.join(ROOT, "exp_build", f"rollout_parity_fail_{wl}.npz"))
flat = mjcf.load_compiled()
env = MyCobotVectorEnv(num_envs=1, has_object=True, block_gripper=(wl == "push"), target_in_the_air=False, reward_type="sparse", auto_reset=False, autotune=False, nefc_max=128)
for k, i in enumerate(f["idx"][:4]):
    sim = OracleSim(flat)
    sim.set_state(f["st_qpos"][k], f["st_qvel"][k], np.clip(f["acts"][k], -1, 1).astype(np.float64), f["st_qacc_warmstart"][k])
    for t in range(20):
        q, v, w = sim.qpos.copy(), sim.qvel.copy(), sim.qacc_warmstart.copy()
        env.set_state(qpos=q[None], qvel=v[None], ctrl=sim.ctrl.copy()[None], qacc_warmstart=w[None])
        sim.forward()
        d = env.debug_forward(0)
        oc = sim.contacts()
        msg = f"env {i} substep {t}: nefc gpu {d['nefc']} oracle {sim.nefc}; ncon {d['ncon']} / {sim.ncon}; iters {d['iters']} / {sim.solver_iter}"
        bad = d["nefc"] != sim.nefc or d["ncon"] != sim.ncon
        if not bad:
            eJ = np.abs(d["efc_J"] - sim.efc("J")).max(); eA = np.abs(d["efc_aref"] - sim.efc("aref")).max(); eD = np.abs(d["efc_D"] / sim.efc("D") - 1).max()
            eq = np.abs(d["qacc"] - sim.qacc).max()
            msg += f" |dJ| {eJ:.2e} |daref| {eA:.2e} |dD/D| {eD:.2e} |dqacc| {eq:.2e} (|qacc| {np.abs(sim.qacc).max():.2e})"
            bad = eJ > 1e-9 or eq > 1e-6 * max(1, np.abs(sim.qacc).max())
        if bad or t == 0:
            print(msg)
        if bad:
            print("  gpu contacts:", [(round(float(a), 9), np.round(p, 6).tolist(), np.round(nn, 6).tolist()) for a, p, nn in zip(d["contact_dist"], d["contact_pos"], d["contact_normal"])])
            print("  oracle contacts:", [(round(c["dist"], 9), np.round(c["pos"], 6).tolist(), np.round(c["frame"][0], 6).tolist(), c["geom1"], c["geom2"]) for c in oc])
            break
        sim.L.o_euler(__import__("ctypes").byref(sim.om), sim.d)

"""GPU-side diagnosis for the hull narrow phase: substep-by-substep stage comparison (mcb_debug_forward vs the oracle's forward)
along the oracle's trajectory, for push envs that start a step with hull contacts."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from mycobotgym_b200 import mjcf
from mycobotgym_b200.vector_env import MyCobotVectorEnv
from oracle.oracle import OracleSim

fm = mjcf.load_compiled()
ng = fm["ngeom"]
kw = dict(has_object=True, block_gripper=True, target_in_the_air=False, reward_type="sparse")
n = 1024
env = MyCobotVectorEnv(num_envs=n, seed=5, lockstep_warps=16, autotune=False, mesh_collision=True, **kw)
env.reset(); env.set_state(elapsed=torch.arange(n, dtype=torch.int32) % 50)
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
for _ in range(12): env.step(torch.rand(n, 7, device="cuda", generator=gen) * 2 - 1)
st = {k: v.cpu().numpy() for k, v in env.get_state().items()}
acts = (torch.rand(n, 7, device="cuda", generator=gen) * 2 - 1).cpu().numpy()
env.close()
e1 = MyCobotVectorEnv(num_envs=1, auto_reset=False, autotune=False, mesh_collision=True, nefc_max=128, **kw)
shown = 0
for i in range(n):
    sim = OracleSim(fm, mesh_collision=True)
    sim.set_state(st["qpos"][i], st["qvel"][i], np.clip(acts[i], -1, 1).astype(np.float64), st["qacc_warmstart"][i])
    sim.forward()
    if not any(c["geom2"] >= ng for c in sim.contacts()): continue
    sim.set_state(st["qpos"][i], st["qvel"][i], np.clip(acts[i], -1, 1).astype(np.float64), st["qacc_warmstart"][i])
    for t in range(20):
        q, v, w = sim.qpos.copy(), sim.qvel.copy(), sim.qacc_warmstart.copy()
        e1.set_state(qpos=q[None], qvel=v[None], ctrl=sim.ctrl.copy()[None], qacc_warmstart=w[None])
        sim.forward()
        d = e1.debug_forward(0)
        oc = sim.contacts()
        bad = d["nefc"] != sim.nefc or d["ncon"] != sim.ncon
        msg = f"env {i} substep {t}: nefc {d['nefc']}/{sim.nefc} ncon {d['ncon']}/{sim.ncon} iters {d['iters']}/{sim.solver_iter}"
        if not bad:
            ed = np.abs(d["contact_dist"] - np.array([c["dist"] for c in oc])).max() if oc else 0
            en = np.abs(d["contact_normal"] - np.array([c["frame"][0] for c in oc])).max() if oc else 0
            ep = np.abs(d["contact_pos"] - np.array([c["pos"] for c in oc])).max() if oc else 0
            eD = np.abs(d["efc_D"] / sim.efc("D") - 1).max(); eA = np.abs(d["efc_aref"] - sim.efc("aref")).max(); eq = np.abs(d["qacc"] - sim.qacc).max()
            msg += f" |ddist| {ed:.1e} |dnormal| {en:.1e} |dpos| {ep:.1e} |dD/D| {eD:.1e} |daref| {eA:.1e} |dqacc| {eq:.1e} (|qacc| {np.abs(sim.qacc).max():.1e})"
            bad = ed > 1e-7 or eq > 1e-5 * max(1, np.abs(sim.qacc).max())
        if bad or t == 0: print(msg)
        if bad or (t == 0 and ep > 1e-6):
            gn = list(fm["geom_names"]) + ["hull:" + x for x in fm["hull_names"]]
            for k, c in enumerate(oc): print("    oracle", gn[c["geom1"]], gn[c["geom2"]], round(c["dist"], 8), np.round(c["pos"], 5), np.round(c["frame"][0], 4))
            for k in range(d["ncon"]): print("    gpu   ", round(float(d["contact_dist"][k]), 8), np.round(d["contact_pos"][k], 5), np.round(d["contact_normal"][k], 4))
            if bad: break
        sim.L.o_euler(ctypes.byref(sim.om), sim.d)
    shown += 1
    if shown >= 6: break

"""GPU diagnostic: stage-level and step-level comparison of the CUDA engine against the CPU oracle,
plus a quick throughput probe.  Run on a B200 box:  python tools/gpu_check.py [--quick]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mycobotgym_b200 import mjcf  # noqa: E402
from mycobotgym_b200.vector_env import MyCobotVectorEnv  # noqa: E402
from oracle.oracle import OracleEnv, OracleSim  # noqa: E402

np.set_printoptions(precision=6, suppress=False, linewidth=200)
flat = mjcf.load_compiled()
JB = [b for b in range(flat["nbody"]) if flat["body_jntnum"][b] > 0]


def mk_states(n, rng, has_object=True, contact_free=False):
    qpos = np.tile(flat["qpos0"], (n, 1))
    qvel = np.zeros((n, 18))
    ctrl = np.zeros((n, 7))
    for i in range(1, n):
        qpos[i, :6] = rng.uniform(-1, 1, 6)
        qpos[i, 6] = qpos[i, 8] = rng.uniform(0.0, 0.5)
        qvel[i, :6] = rng.normal(size=6) * 0.5
        ctrl[i] = rng.uniform(-1, 1, 7)
        if has_object:
            qpos[i, 12:14] = rng.uniform(-0.1, 0.1, 2)
            qpos[i, 14] = 0.21 - 1e-5 if not contact_free else 0.3
            if i % 2 == 0:
                qvel[i, 12:18] = rng.normal(size=6) * 0.05
            if i % 3 == 0:
                q = np.array([1.0, 0, 0, 0]) + rng.normal(size=4) * 0.02
                qpos[i, 15:19] = q / np.linalg.norm(q)
    return qpos, qvel, ctrl


def cmp(name, a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        print(f"   {name:16s} SHAPE MISMATCH {a.shape} vs {b.shape}")
        return False
    err = np.abs(a - b).max() if a.size else 0.0
    scale = max(1.0, np.abs(b).max() if b.size else 1.0)
    ok = err <= tol * scale
    print(f"   {name:16s} max|diff| = {err:.3e} (scale {scale:.2e}) {'ok' if ok else 'FAIL'}")
    return ok


def stage_check(has_object):
    print(f"== stage check has_object={has_object}")
    n = 8
    rng = np.random.default_rng(3)
    qpos, qvel, ctrl = mk_states(n, rng, has_object)
    env = MyCobotVectorEnv(num_envs=n, has_object=has_object, reward_type="dense", auto_reset=False)
    env.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, qacc_warmstart=np.zeros((n, 18)))
    allok = True
    for i in range(n):
        sim = OracleSim(flat, disable_cube=not has_object)
        sim.set_state(qpos[i], qvel[i], ctrl[i], np.zeros(18))
        sim.forward()
        dbg = env.debug_forward(i)
        env.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, qacc_warmstart=np.zeros((n, 18)))
        nva = 18 if has_object else 12
        print(f" env {i}: nefc gpu {dbg['nefc']} oracle {sim.nefc}; ncon {dbg['ncon']} / {sim.ncon}; iters {dbg['iters']} / {sim.solver_iter}")
        ok = True
        ok &= cmp("xpos", dbg["xpos"][: (13 if has_object else 12)], sim.xpos[JB][: (13 if has_object else 12)], 1e-13)
        ok &= cmp("M", dbg["M"][:nva, :nva], sim.M[:nva, :nva], 1e-13)
        ok &= cmp("qfrc_bias", dbg["qfrc_bias"][:nva], sim.qfrc_bias[:nva], 1e-12)
        ok &= cmp("qfrc_smooth", dbg["qfrc_smooth"][:nva], sim.qfrc_smooth[:nva], 1e-11)
        ok &= cmp("qacc_smooth", dbg["qacc_smooth"][:nva], sim.qacc_smooth[:nva], 1e-9)
        if dbg["nefc"] == sim.nefc:
            ok &= cmp("efc_J", dbg["efc_J"][:, :nva], sim.efc("J")[:, :nva], 1e-12)
            ok &= cmp("efc_D", dbg["efc_D"], sim.efc("D"), 1e-12)
            ok &= cmp("efc_aref", dbg["efc_aref"], sim.efc("aref"), 1e-10)
        else:
            ok = False
        ok &= cmp("qacc", dbg["qacc"][:nva], sim.qacc[:nva], 1e-7)
        allok &= bool(ok)
    print("stage check", "PASS" if allok else "FAIL")
    env.close()
    return allok


def step_check(has_object, block_gripper=False, nsteps=3):
    print(f"== step check has_object={has_object} block_gripper={block_gripper}")
    n = 8
    rng = np.random.default_rng(5)
    qpos, qvel, ctrl = mk_states(n, rng, has_object)
    goals = rng.uniform(-0.1, 0.1, (n, 3)) + np.array([0, 0, 0.3])
    env = MyCobotVectorEnv(num_envs=n, has_object=has_object, block_gripper=block_gripper, reward_type="sparse" if has_object else "dense",
                           auto_reset=False)
    env.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, qacc_warmstart=np.zeros((n, 18)), goal=goals, elapsed=np.zeros(n, dtype=np.int32))
    oenvs = []
    for i in range(n):
        oe = OracleEnv(flat, has_object=has_object, block_gripper=block_gripper, reward_type="sparse" if has_object else "dense")
        oe.sim.set_state(qpos[i], qvel[i], ctrl[i], np.zeros(18))
        oe.goal = goals[i].copy()
        oenvs.append(oe)
    allok = True
    for t in range(nsteps):
        act = rng.uniform(-1, 1, (n, 7)).astype(np.float32)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act))
        st = env.get_state()
        torch.cuda.synchronize()
        gq, gv = st["qpos"].cpu().numpy(), st["qvel"].cpu().numpy()
        gobs = obs["observation"].cpu().numpy()
        grew = rew.cpu().numpy()
        nq = 19 if has_object else 12
        nv = 18 if has_object else 12
        worst = dict(qpos=0, qvel=0, obs=0, rew=0)
        for i in range(n):
            o, r, te, tr, inf = oenvs[i].step(act[i])
            worst["qpos"] = max(worst["qpos"], np.abs(gq[i, :nq] - oenvs[i].sim.qpos[:nq]).max())
            worst["qvel"] = max(worst["qvel"], np.abs(gv[i, :nv] - oenvs[i].sim.qvel[:nv]).max())
            worst["obs"] = max(worst["obs"], np.abs(gobs[i] - o["observation"]).max())
            worst["rew"] = max(worst["rew"], abs(float(grew[i]) - float(r)))
            if bool(term[i]) != te or bool(trunc[i]) != tr:
                print("   flag mismatch env", i, bool(term[i]), te, bool(trunc[i]), tr)
                allok = False
        print(f" step {t}: worst abs diff qpos {worst['qpos']:.3e} qvel {worst['qvel']:.3e} obs {worst['obs']:.3e} reward {worst['rew']:.3e}")
        tol = 1e-5 if has_object else 1e-9
        if t == 0 and (worst["qpos"] > tol or worst["qvel"] > tol * 100 or worst["obs"] > tol * 10):
            allok = False
    print("stats", env.stats().cpu().numpy())
    print("step check", "PASS" if allok else "FAIL")
    env.close()
    return allok


def timing(has_object, n, steps=10, nefc_max=0):
    env = MyCobotVectorEnv(num_envs=n, has_object=has_object, reward_type="sparse" if has_object else "dense", nefc_max=nefc_max)
    env.reset()
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    acts = torch.rand(steps + 3, n, 7, device="cuda", generator=g) * 2 - 1
    for t in range(3):
        env.step(acts[t])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        env.step(acts[3 + t])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = env.stats().cpu().numpy()
    print(f"timing has_object={has_object} N={n} nefc_max={nefc_max}: {ms:.3f} ms/step -> {n / ms * 1e3:.3e} env-steps/s; stats {st}")
    env.close()


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    print(torch.cuda.get_device_name(0))
    ok = True
    ok &= stage_check(True)
    ok &= stage_check(False)
    ok &= step_check(False)
    ok &= step_check(True)
    ok &= step_check(True, block_gripper=True)
    timing(False, 4096)
    timing(True, 16384)
    if not quick:
        timing(True, 16384, nefc_max=128)
        timing(False, 16384)
    print("ALL", "PASS" if ok else "FAIL")

"""Instrumented algorithmic FLOP count of one env-step ("to be replaced by an instrumented
count from the CPU oracle, then frozen").

Builds oracle/_count/liboracle_count.so = oracle/mjc_oracle.c compiled as C++ with `double` replaced by a counting wrapper
(oracle/flopcount.h), runs each bench workload's random-action protocol on it and prints add / mul / div / sqrt per
env-step (FMA = 2 FLOP falls out of counting the multiply and the add separately).  The counts are those of the oracle's
scalar restatement of MuJoCo 2.3.2's algorithms (sparse LDL on the 25-body tree, dense Newton on nefc x 18 rows); the
task layer's numpy arithmetic (observation, reward: < 1 kFLOP per step) is not included.

    python tools/count_flops.py            # prints a table + JSON; numbers are frozen in BASELINE.md / bench.py
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "oracle", "_count")
os.makedirs(OUT, exist_ok=True)
LIB = os.path.join(OUT, "liboracle_count.so")
subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fpermissive", "-w", "-shared", "-I", os.path.join(ROOT, "oracle"),
                       "-o", LIB, os.path.join(ROOT, "oracle", "flopcount_wrap.cc")])
import oracle.oracle as oo  # noqa: E402

oo.LIB_PATH = LIB
oo.build_lib = lambda force=False: LIB
from mycobotgym_b200 import mjcf  # noqa: E402
from oracle.oracle import OracleEnv  # noqa: E402

L = oo.lib()
fl = (C.c_longlong * 8).in_dll(L, "o_flops")
joint, mocap = mjcf.load_compiled(mjcf.COMPILED_JOINT), mjcf.load_compiled(mjcf.COMPILED_MOCAP)
WORK = {
    "reach": (joint, dict(has_object=False, reward_type="dense"), 7),
    "push": (joint, dict(has_object=True, block_gripper=True, target_in_the_air=False, reward_type="sparse"), 7),
    "pick": (joint, dict(has_object=True, reward_type="sparse"), 7),
    "ik": (joint, dict(has_object=True, reward_type="sparse", controller_type="IK"), 7),
    "mocap": (mocap, dict(has_object=True, reward_type="sparse", controller_type="mocap"), 8),
}


def measure(name, episodes=4):
    fm, kw, adim = WORK[name]
    env = OracleEnv(fm, **kw)
    rng = np.random.default_rng(0)
    tot, steps, iters, sub = np.zeros(6), 0, 0, 0
    for ep in range(episodes):
        env.reset(seed=ep)
        for t in range(50):
            for i in range(8):
                fl[i] = 0
            a = rng.uniform(-1, 1, adim).astype(np.float32)
            o, r, te, tr, info = env.step(a)
            tot += [fl[i] for i in range(6)]
            steps += 1
            if te:
                break
    return dict(add=tot[0] / steps, mul=tot[1] / steps, div=tot[2] / steps, sqrt=tot[3] / steps, transcendental=tot[4] / steps,
                flop_per_env_step=float(tot[:4].sum() / steps), env_steps=steps)


def measure_grasp():
    g = np.load(os.path.join(ROOT, "tests", "golden", "grasp_pick_sparse.npz"))
    env = OracleEnv(joint, has_object=True, reward_type="sparse")
    env.sim.set_state(g["qpos0"], g["qvel0"], g["ctrl0"], g["warm0"])
    env.goal = g["goal"].copy()
    tot, steps = np.zeros(6), 0
    for t in range(len(g["actions"])):
        for i in range(8):
            fl[i] = 0
        env.step(g["actions"][t])
        tot += [fl[i] for i in range(6)]
        steps += 1
    return dict(add=tot[0] / steps, mul=tot[1] / steps, div=tot[2] / steps, sqrt=tot[3] / steps, transcendental=tot[4] / steps,
                flop_per_env_step=float(tot[:4].sum() / steps), env_steps=steps)


if __name__ == "__main__":
    res = {k: measure(k) for k in WORK}
    res["grasp"] = measure_grasp()
    print(f"{'workload':8s} {'add':>10s} {'mul':>10s} {'div':>8s} {'sqrt':>7s} {'FLOP/env-step':>14s}")
    for k, v in res.items():
        print(f"{k:8s} {v['add']:10.0f} {v['mul']:10.0f} {v['div']:8.0f} {v['sqrt']:7.0f} {v['flop_per_env_step']:14.0f}")
    print(json.dumps(res))

import os, sys, random
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mycobotgym_b200 import mjcf
from mycobotgym_b200.vector_env import MyCobotVectorEnv
from oracle.oracle import OracleEnv
flat = mjcf.load_compiled()
n, adim = 6, 7
env = MyCobotVectorEnv(num_envs=n, has_object=False, reward_type="dense", controller_type="IK", auto_reset=False, goal_source="reference")
oes = [OracleEnv(flat, has_object=False, reward_type="dense", controller_type="IK") for _ in range(n)]
random.seed(11)
goals = []
for i, oe in enumerate(oes):
    oe.reset(seed=100 + i); goals.append(oe.goal.copy())
env.reset(goals=np.array(goals))
rng = np.random.default_rng(12)
for t in range(4):
    acts = rng.uniform(-1, 1, (n, adim)).astype(np.float32)
    obs, rew, term, trunc, info = env.step(torch.as_tensor(acts))
    st = env.get_state()
    dq, dc, dv = [], [], []
    for i, oe in enumerate(oes):
        oe.step(acts[i])
        dq.append(np.abs(st["qpos"][i, :12].cpu().numpy() - oe.sim.qpos[:12]).max())
        dv.append(np.abs(st["qvel"][i, :12].cpu().numpy() - oe.sim.qvel[:12]).max())
        dc.append(np.abs(st["ctrl"][i].cpu().numpy() - oe.sim.ctrl).max())
    print(t, "qpos", ["%.1e" % x for x in dq], "qvel", ["%.1e" % x for x in dv], "ctrl", ["%.1e" % x for x in dc])
    # re-sync GPU to the oracle state (single-step comparison from identical states, incl. the stale frames' qprev)

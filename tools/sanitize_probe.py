"""Small workload for compute-sanitizer (memcheck / racecheck): reset + 2 steps of 48 pick-and-place envs incl. a grasp
state that takes the fallback-layout path, and a reach batch."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mycobotgym_b200.vector_env import MyCobotVectorEnv  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "grasp_pick_sparse.npz"))
n = 48
env = MyCobotVectorEnv(num_envs=n, has_object=True, reward_type="sparse", seed=3)
env.reset()
st = env.get_state()
qpos, qvel, ctrl = st["qpos"].cpu().numpy(), st["qvel"].cpu().numpy(), st["ctrl"].cpu().numpy()
qpos[5], qvel[5], ctrl[5] = g["qpos0"], g["qvel0"], g["ctrl0"]
env.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, elapsed=np.full(n, 48, dtype=np.int32))
a = torch.rand(n, 7, device="cuda") * 2 - 1
for _ in range(3):
    env.step(a)
torch.cuda.synchronize()
print("pick stats", env.stats().cpu().numpy())
env2 = MyCobotVectorEnv(num_envs=20, has_object=False, reward_type="dense")
env2.reset()
env2.step(torch.zeros(20, 7, device="cuda"))
torch.cuda.synchronize()
print("reach ok")

"""Step time when EVERY env holds the cube between the finger layers (the contact-rich regime a trained pick-and-place
policy spends its episodes in; the benchmark's random actions never get there).  State: tests/golden/grasp_pick_sparse.npz."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from mycobotgym_b200.vector_env import MyCobotVectorEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = np.load(os.path.join(ROOT, "tests", "golden", "grasp_pick_sparse.npz"))
lw = int(sys.argv[2]) if len(sys.argv) > 2 else 0
env = MyCobotVectorEnv(num_envs=n, has_object=True, reward_type="sparse", auto_reset=False, seed=1, lockstep_warps=lw)
env.reset()
rep = lambda x: np.repeat(x[None], n, 0)
env.set_state(qpos=rep(g["qpos0"]), qvel=rep(g["qvel0"]), ctrl=rep(g["ctrl0"]), qacc_warmstart=rep(g["warm0"]), goal=rep(g["goal"]),
              elapsed=np.zeros(n, dtype=np.int32), qprev=rep(g["qpos0"][:6]))
act = torch.zeros(n, 7, device="cuda")
act[:, 6] = 0.8
for t in range(3):
    env.step(act)
print("lockstep", env.lockstep_warps, "fallback envs", env.last_fallback_envs())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for t in range(K):
    env.step(act)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"{n} envs all grasping: {ms:.2f} ms/step = {n / ms * 1e3 / 1e6:.3f} M env-steps/s; fallback envs {env.last_fallback_envs()}")

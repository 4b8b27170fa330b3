"""Soak run: every workload for many steps at the full batch with random actions and auto-reset; prints episode statistics, the
anomaly counter (constraint rows dropped in the last tier + bad-simulation resets) and state extremes.
usage: python tools/soak.py [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from mycobotgym_b200.vector_env import MyCobotVectorEnv  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
WORK = {
    "pick": (dict(has_object=True, reward_type="sparse"), 16384, 7),
    "push": (dict(has_object=True, block_gripper=True, target_in_the_air=False, reward_type="sparse"), 16384, 7),
    "reach": (dict(has_object=False, reward_type="dense"), 4096, 7),
    "ik": (dict(has_object=True, reward_type="sparse", controller_type="IK"), 16384, 7),
    "mocap": (dict(has_object=True, reward_type="sparse", controller_type="mocap", model_path="./assets/mycobot280_mocap.xml"), 16384, 8),
}
for name, (kw, n, adim) in WORK.items():
    env = MyCobotVectorEnv(num_envs=n, seed=1, **kw)
    env.reset()
    env.set_state(elapsed=torch.arange(n, dtype=torch.int32) % 50)      # staggered episode clocks: resets spread over the steps
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    k = steps if name != "ik" else max(50, steps // 5)
    t0 = time.perf_counter()
    for _ in range(k):
        env.step(torch.rand(n, adim, device="cuda", generator=gen) * 2 - 1)
    s = env.stats().cpu().numpy()
    st = env.get_state()
    dt = time.perf_counter() - t0
    finite = bool(torch.isfinite(st["qpos"]).all() and torch.isfinite(st["qvel"]).all())
    print(f"{name:6s} {k:5d} steps x {n} envs in {dt:6.1f} s: episodes {int(s[0])}, successes {int(s[1])}, mean return {s[2] / max(s[0], 1):8.3f}, "
          f"env-steps {int(s[4])}, anomalies {int(s[5])}, Newton iterations / substep {s[6] / max(s[7], 1):.3f}, finite {finite}, "
          f"max |qvel| {float(st['qvel'].abs().max()):.1f}, cube z range [{float(st['qpos'][:, 14].min()):.4f}, {float(st['qpos'][:, 14].max()):.4f}]", flush=True)
    env.close()

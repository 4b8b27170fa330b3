"""Step time and tier-0 overflow count over a long random rollout (is the bench's timed window representative?)."""
import sys, os
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import WORKLOADS
from mycobotgym_b200.vector_env import MyCobotVectorEnv

wl = sys.argv[1] if len(sys.argv) > 1 else "mocap"
kw, n, _ = WORKLOADS[wl]
env = MyCobotVectorEnv(num_envs=n, seed=1000, **kw)
env.reset()
env.set_state(elapsed=torch.arange(n, device="cuda", dtype=torch.int32) % 50)
gen = torch.Generator(device="cuda").manual_seed(1234)
for chunk in range(16):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(10):
        env.step(torch.rand(n, env.action_dim, device="cuda", generator=gen) * 2 - 1)
    e1.record()
    torch.cuda.synchronize()
    print(f"{wl} steps {chunk * 10:3d}-{chunk * 10 + 9:3d}: {e0.elapsed_time(e1) / 10:6.2f} ms/step, left tier 0 / tier 1: {env.last_fallback_envs()}, lockstep {env.lockstep_warps}")

# device entry point vs host-buffer entry point on the same steady-state batch
import time
import numpy as np
acts = torch.rand(20, n, env.action_dim, device="cuda", generator=gen) * 2 - 1
a_host = acts.cpu().pin_memory().numpy()
out = env.step_host(a_host[0])
def step_sync(t):
    env.step(acts[t])
    torch.cuda.synchronize()


for name, fn in (("step (device buffers)", lambda t: env.step(acts[t])), ("step + synchronize every step", step_sync),
                 ("step_host (pinned numpy)", lambda t: env.step_host(a_host[t], out))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(20):
        fn(t)
    torch.cuda.synchronize()
    print(f"{wl} {name}: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms/step wall")

# host-side cost of one step() call (returns before the GPU finishes unless something inside blocks)
torch.cuda.synchronize()
ts = []
for t in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.step(acts[t])
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"{wl} host time inside step(): " + " ".join(f"{x:.2f}" for x in ts) + " ms")

# GPU time per step (events) when the host synchronises after every step
evs = []
for t in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step(acts[t])
    e1.record()
    torch.cuda.current_stream().synchronize()
    evs.append(e0.elapsed_time(e1))
print(f"{wl} GPU event time per synced step: " + " ".join(f"{x:.2f}" for x in evs) + " ms")

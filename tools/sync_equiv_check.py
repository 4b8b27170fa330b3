"""Do back-to-back (unsynchronised) steps give the same bits as steps separated by a device synchronise?"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import WORKLOADS
from mycobotgym_b200.vector_env import MyCobotVectorEnv

wl = sys.argv[1] if len(sys.argv) > 1 else "mocap"
kw, n, _ = WORKLOADS[wl]
n = 4096
gen = torch.Generator(device="cuda").manual_seed(3)
acts = torch.rand(40, n, 8 if wl == "mocap" else 7, device="cuda", generator=gen) * 2 - 1
outs = []
for sync in (False, True):
    env = MyCobotVectorEnv(num_envs=n, seed=5, lockstep_warps=16, **kw)
    env.reset()
    fb = []
    for t in range(40):
        env.step(acts[t])
        if sync:
            torch.cuda.synchronize()
            fb.append(env.last_fallback_envs()[0])
    st = env.get_state()
    outs.append((st["qpos"].clone(), st["qvel"].clone(), env.stats(reset=False).clone()))
    if sync:
        print("fallback envs per step (synced run):", fb[::4])
    env.close()
print("qpos equal:", torch.equal(outs[0][0], outs[1][0]), "qvel equal:", torch.equal(outs[0][1], outs[1][1]), "stats equal:", torch.equal(outs[0][2], outs[1][2]))
print((outs[0][0] - outs[1][0]).abs().max().item(), outs[0][2].tolist(), outs[1][2].tolist())

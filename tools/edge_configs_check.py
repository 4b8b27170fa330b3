import sys, numpy as np, torch, time
sys.path.insert(0,'/root/repo')
from mycobotgym_b200.vector_env import MyCobotVectorEnv, make
def run(desc, **kw):
    t0=time.time()
    env = MyCobotVectorEnv(**kw)
    obs,_ = env.reset()
    a = torch.zeros(env.num_envs, env.action_dim, device='cuda')
    for t in range(3):
        obs, rew, term, trunc, info = env.step(a)
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(obs['observation']).all())
    print(f"{desc}: ok={ok} lockstep={env.lockstep_warps} fallback={env.last_fallback_envs()} {time.time()-t0:.1f}s")
    env.close()
run("1 env", num_envs=1)
run("17 envs frame_skip 5", num_envs=17, frame_skip=5)
run("131072 envs", num_envs=131072)
run("IK control_steps 2", num_envs=33, controller_type="IK", control_steps=2)
run("mocap fetch", num_envs=65, controller_type="mocap", model_path="./assets/mycobot280_mocap.xml", fetch_env=True)
run("reach dense block_gripper", num_envs=100, has_object=False, reward_type="dense", block_gripper=True)
run("push reward_shaping", num_envs=40, has_object=True, block_gripper=True, reward_type="reward_shaping")
for eid in ["MyCobotFetchPickAndPlace-Dense-IK-v0","MyCobotReach-Sparse-mocap-v0","MyCobotPickAndPlace-RewardShaping-IK-v0"]:
    e = make(eid, num_envs=8); e.reset(); e.step(torch.zeros(8, e.action_dim)); print(eid, "ok"); e.close()

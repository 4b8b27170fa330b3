"""Boundary behaviour of the C ABI added in round 2 (VERDICT r01 "boundary defects"): host-buffer reset,
reseeding of the device sampler, full checkpoints, counted launches, no hidden synchronisation inside mcb_step (it can be
captured in a CUDA graph), the env list behind the fallback count, loud refusals in the SB3-shaped adapter."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(**kw):
    from mycobotgym_b200.vector_env import MyCobotVectorEnv

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    kw.setdefault("autotune", False)
    return MyCobotVectorEnv(**kw)


def test_reset_host_equals_reset():
    # mycobot.py:506-514 through host buffers == through device buffers (same Philox key and counters => same draws)
    n = 64
    e1 = _env(num_envs=n, has_object=True, reward_type="sparse", seed=9)
    e2 = _env(num_envs=n, has_object=True, reward_type="sparse", seed=9)
    o1, _ = e1.reset(seed=21)
    o2, _ = e2.reset_host(seed=21)
    for k in ("observation", "achieved_goal", "desired_goal"):
        assert np.array_equal(o1[k].cpu().numpy(), o2[k]), k
    # masked reset with injected sampler outputs
    mask = np.zeros(n, dtype=np.uint8); mask[::3] = 1
    xy = np.tile([0.05, -0.02], (n, 1)); goals = np.tile([0.1, 0.03, 0.25], (n, 1))
    a = np.random.default_rng(0).uniform(-1, 1, (n, 7)).astype(np.float32)
    e1.step(torch.as_tensor(a)); e2.step(torch.as_tensor(a))
    o1, _ = e1.reset(mask=mask, object_xy=xy, goals=goals)
    o2, _ = e2.reset_host(mask=mask, object_xy=xy, goals=goals)
    s1, s2 = e1.get_state(), e2.get_state()
    for k in ("qpos", "qvel", "goal", "elapsed"):
        assert torch.equal(s1[k], s2[k]), k
    m = mask.astype(bool)
    assert np.array_equal(o1["observation"].cpu().numpy()[m], o2["observation"][m])
    assert np.array_equal(s1["goal"].cpu().numpy()[m], goals[m]) and not np.array_equal(s1["goal"].cpu().numpy()[~m], goals[~m])
    e1.close(); e2.close()


def test_reset_seed_reseeds_the_device_sampler():
    # mycobot.py:509-510: reset(seed=s) makes the following goal / cube draws a function of s (here: of (s, env index))
    n = 128
    e1 = _env(num_envs=n, has_object=True, reward_type="sparse", seed=1)
    e2 = _env(num_envs=n, has_object=True, reward_type="sparse", seed=2)
    g1 = e1.reset()[0]["desired_goal"].clone()
    g2 = e2.reset()[0]["desired_goal"].clone()
    assert not torch.equal(g1, g2)                                   # constructor seeds differ
    a1 = e1.reset(seed=5)[0]["desired_goal"].clone()
    a2 = e2.reset(seed=5)[0]["desired_goal"].clone()
    assert torch.equal(a1, a2) and torch.equal(e1.get_state()["qpos"], e2.get_state()["qpos"])
    b1 = e1.reset(seed=6)[0]["desired_goal"].clone()
    assert not torch.equal(a1, b1)
    assert torch.equal(e1.reset(seed=5)[0]["desired_goal"], a1)     # and it is repeatable
    # masked reseed: only the masked envs restart their stream
    mask = torch.zeros(n, dtype=torch.uint8); mask[:10] = 1
    e1.reset(); e2.reset(); e2.reset()                               # streams now at different positions
    c1 = e1.reset(seed=7, mask=mask)[0]["desired_goal"].clone()
    c2 = e2.reset(seed=7, mask=mask)[0]["desired_goal"].clone()
    assert torch.equal(c1[:10], c2[:10]) and not torch.equal(c1[10:], c2[10:])
    # SB3-style seed(): takes effect at the next reset
    e1.seed(11); e2.seed(11)
    assert torch.equal(e1.reset()[0]["desired_goal"], e2.reset()[0]["desired_goal"])
    e1.close(); e2.close()


def test_checkpoint_continues_goal_stream_and_episode_returns():
    n = 96
    rng = np.random.default_rng(1)
    acts = torch.as_tensor(rng.uniform(-1, 1, (70, n, 7)).astype(np.float32))
    a = _env(num_envs=n, has_object=True, reward_type="sparse", seed=4)
    a.reset()
    a.set_state(elapsed=torch.arange(n, dtype=torch.int32) % 50)
    for t in range(10):
        a.step(acts[t])
    st = {k: v.clone() for k, v in a.get_state().items()}
    assert set(st) >= {"env_seed", "rng_counter", "ep_return"} and int(st["rng_counter"].max()) > 0
    b = _env(num_envs=n, has_object=True, reward_type="sparse", seed=12345)     # a different batch, restored from the checkpoint
    b.reset()
    b.set_state(**st)
    a.stats(); b.stats()
    for t in range(10, 70):                                           # every env passes its TimeLimit at least once: goals resampled
        oa = a.step(acts[t]); ob = b.step(acts[t])
        assert torch.equal(oa[0]["observation"], ob[0]["observation"]) and torch.equal(oa[0]["desired_goal"], ob[0]["desired_goal"])
    sa, sb = a.stats().cpu().numpy(), b.stats().cpu().numpy()
    assert np.array_equal(sa, sb) and sa[0] >= n                      # incl. return_sum: running episode returns were restored
    a.close(); b.close()


def test_step_launch_count_and_cuda_graph_capture():
    # mcb_step only enqueues its kernels on the caller's stream: no autotune, no allocation, no synchronisation inside
    n = 256
    e = _env(num_envs=n, has_object=True, reward_type="sparse", seed=3)
    g = _env(num_envs=n, has_object=True, reward_type="sparse", seed=3)
    e.reset(); g.reset()
    acts = torch.rand(6, n, 7, device="cuda") * 2 - 1
    t0 = e.total_launches
    e.step(acts[0])
    assert e.last_step_launches == 3 and e.total_launches - t0 == 3  # one kernel per layout tier, counted at the launch sites
    g.step(acts[0])
    static_a = acts[1].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            out = g.step(static_a)
    torch.cuda.current_stream().wait_stream(side)
    for t in range(1, 6):
        static_a.copy_(acts[t])
        graph.replay()
        ref = e.step(acts[t])
        torch.cuda.synchronize()
        assert torch.equal(out[0]["observation"], ref[0]["observation"]) and torch.equal(out[1], ref[1])
    assert torch.equal(g.get_state()["qpos"], e.get_state()["qpos"])
    e.close(); g.close()


def test_fallback_list_names_the_contact_rich_envs():
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grasp_pick_sparse.npz"))
    n = 40
    env = _env(num_envs=n, has_object=True, reward_type="sparse", auto_reset=False)
    env.reset()
    st = env.get_state()
    qpos, qvel, ctrl = st["qpos"].cpu().numpy(), st["qvel"].cpu().numpy(), st["ctrl"].cpu().numpy()
    acts = np.zeros((n, 7), dtype=np.float32)
    for i in (7, 23):                                                # two envs hold the cube: coupled rows overflow the common layout
        qpos[i], qvel[i], ctrl[i] = g["qpos0"], g["qvel0"], g["ctrl0"]
        acts[i] = g["actions"][0]
    env.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl)
    env.step(torch.as_tensor(acts))
    assert env.last_fallback_envs()[0] == 2
    assert sorted(env.last_fallback_list().tolist()) == [7, 23]
    assert len(env.last_fallback_list(cap=1)) == 1
    env.close()


def test_adapter_and_replay_refuse_what_they_cannot_serve():
    from mycobotgym_b200.her import DeviceHerReplayBuffer
    from mycobotgym_b200.sb3_adapter import MyCobotSB3VecEnv

    with pytest.raises(ValueError):
        MyCobotSB3VecEnv(4, auto_reset=False)
    with pytest.raises(ValueError):
        MyCobotSB3VecEnv(4, goal_source="reference")
    v = MyCobotSB3VecEnv(8, has_object=True, reward_type="sparse", autotune=False)
    v.seed(3)
    o1 = v.reset()
    v.seed(3)
    o2 = v.reset()
    assert np.array_equal(o1["desired_goal"], o2["desired_goal"])      # VecEnv.seed reaches the device sampler
    # terminal_observation rows come from a buffer the library zero-fills: finite everywhere, real values where done
    for _ in range(50):
        obs, rew, dones, infos = v.step(np.zeros((8, 7), dtype=np.float32))
    assert dones.all() and all(np.isfinite(i["terminal_observation"]["observation"]).all() for i in infos)
    env = v.venv
    buf = DeviceHerReplayBuffer(64 * 8, env, seed=1)
    obs = env._obs_dict()                                              # NOT cloned: aliases the step's output buffers
    a = torch.zeros(8, 7, device="cuda")
    out = env.step(a)
    with pytest.raises(ValueError, match="aliases"):
        buf.add_step(obs, a, out)
    buf.close()
    v.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_env_on_another_gpu_leaves_the_current_device_alone():
    torch.cuda.set_device(0)
    e = _env(num_envs=4, has_object=False, reward_type="dense", device="cuda:1")
    assert torch.cuda.current_device() == 0
    e.reset(); e.step(torch.zeros(4, 7))
    assert torch.cuda.current_device() == 0 and e.get_state()["qpos"].device.index == 1
    e.close()

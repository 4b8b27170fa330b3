"""HER replay: the numpy restatement of SB3 2.0.0a0's HerReplayBuffer (oracle/her_oracle.py) against its
own invariants on CPU, and the device buffer (mcb_her_*) against it bit for bit on the GPU."""
import numpy as np
import pytest
import torch

from oracle.her_oracle import HerOracle


def _feed(rng, T, N, od, ad, steps, sinks, max_len=7, reward_f64=False):
    """Random transitions with random episode ends, pushed into every sink (callables taking the numpy arrays)."""
    age = np.zeros(N, int)
    for k in range(steps):
        obs, nobs = rng.normal(size=(N, od)), rng.normal(size=(N, od))
        ag, nag, dg = rng.normal(size=(N, 3)) * 0.02, rng.normal(size=(N, 3)) * 0.02, rng.normal(size=(N, 3)) * 0.02
        act = rng.uniform(-1, 1, (N, ad)).astype(np.float32)
        rew = rng.normal(size=N) if reward_f64 else -(rng.random(N) < 0.7).astype(np.float32)
        age += 1
        term = rng.random(N) < 0.1
        trunc = np.logical_or(age >= max_len, term & (rng.random(N) < 0.5))
        age[term | trunc] = 0
        for s in sinks:
            s(obs, ag, dg, nobs, nag, act, rew, term, trunc)


def test_her_oracle_episode_table_and_future_relabelling():
    rng = np.random.default_rng(0)
    T, N = 23, 5
    o = HerOracle(T, N, 10, 7, n_sampled_goal=4, reward_type="sparse", distance_threshold=0.01)
    assert abs(o.her_ratio - 0.8) < 1e-15
    _feed(rng, T, N, 10, 7, 90, [o.add])
    assert o.full
    # every tagged transition sits inside its episode, the episode is contiguous on the ring and ends with done
    for t, e in zip(*np.nonzero(o.ep_length)):
        es, el = o.ep_start[t, e], o.ep_length[t, e]
        assert 0 <= (t - es) % T < el
        last = (es + el - 1) % T
        assert o.dones[last, e] == 1 and np.all(o.ep_length[np.arange(es, es + el) % T, e] == el)
        assert np.all(o.dones[np.arange(es, es + el - 1) % T, e] == 0)
    # the running episodes (not yet done) are not sampleable
    for e in range(N):
        k = o.cur_start[e]
        while k != o.pos:
            assert o.ep_length[k, e] == 0
            k = (k + 1) % T
    valid = o.valid_indices()
    idx = rng.choice(valid, 64)
    s = o.sample(idx, lambda cur, el: rng.integers(cur, el))
    assert s["nb_virtual"] == 51
    t, e = np.unravel_index(idx, (T, N))
    # relabelling with the current transition itself makes the stored next_achieved_goal the goal: success, reward -0.0
    s0 = o.sample(idx, lambda cur, el: cur)
    assert np.all(s0["rewards"][:51] == 0) and np.all(np.signbit(s0["rewards"][:51]))
    np.testing.assert_array_equal(s0["dg"][:51], o.next_ag[t[:51], e[:51]])
    np.testing.assert_array_equal(s["dg"][51:], o.dg[t[51:], e[51:]])
    np.testing.assert_array_equal(s["rewards"][51:], o.rewards[t[51:], e[51:]])
    assert np.all(s["source"] % N == e[:51])                                 # goals come from the same env's episode
    assert set(np.unique(s["dones"])) <= {0.0, 1.0}


@pytest.mark.gpu
@pytest.mark.parametrize("reward_type,od", [("sparse", 25), ("dense", 10)])
def test_device_her_matches_oracle_bit_for_bit(reward_type, od):
    from mycobotgym_b200.her import DeviceHerReplayBuffer

    rng = np.random.default_rng(1)
    T, N, ad = 23, 37, 7
    o = HerOracle(T, N, od, ad, n_sampled_goal=4, reward_type=reward_type, distance_threshold=0.01)
    buf = DeviceHerReplayBuffer(T * N, n_envs=N, obs_dim=od, action_dim=ad, n_sampled_goal=4, reward_type=reward_type,
                                distance_threshold=0.01, seed=3)
    assert buf.buffer_steps == T
    dev = buf.device
    f64 = reward_type == "dense"

    def dev_add(obs, ag, dg, nobs, nag, act, rew, term, trunc):
        tt = lambda x, dt=torch.float64: torch.as_tensor(x, dtype=dt, device=dev)
        buf.add({"observation": tt(obs), "achieved_goal": tt(ag), "desired_goal": tt(dg)},
                {"observation": tt(nobs), "achieved_goal": tt(nag)}, tt(act, torch.float32),
                tt(rew, torch.float64 if f64 else torch.float32), tt(term, torch.uint8), tt(trunc, torch.uint8))

    for chunk in range(6):
        _feed(rng, T, N, od, ad, 17, [o.add, dev_add], reward_f64=f64)
        es, el, nvalid = buf.episode_table()
        np.testing.assert_array_equal(el.cpu().numpy(), o.ep_length)
        tagged = o.ep_length > 0
        np.testing.assert_array_equal(es.cpu().numpy()[tagged], o.ep_start[tagged])
        assert nvalid == int(tagged.sum()) and buf.size() == (T if o.full else o.pos) * N
        valid = o.valid_indices()
        B = 257
        idx = rng.choice(valid, B)
        draws = {}
        want = o.sample(idx, lambda cur, el_: draws.setdefault("f", rng.integers(cur, el_)))
        fut = np.zeros(B, np.int32); fut[:want["nb_virtual"]] = draws["f"]
        got = buf.sample(B, indices=idx, future=fut, return_indices=True)
        assert buf.failed_samples() == 0
        g = lambda x: x.cpu().numpy()
        np.testing.assert_array_equal(g(got["observations"]["observation"]), want["obs"])
        np.testing.assert_array_equal(g(got["observations"]["achieved_goal"]), want["ag"])
        np.testing.assert_array_equal(g(got["observations"]["desired_goal"]), want["dg"])
        np.testing.assert_array_equal(g(got["next_observations"]["observation"]), want["next_obs"])
        np.testing.assert_array_equal(g(got["next_observations"]["achieved_goal"]), want["next_ag"])
        np.testing.assert_array_equal(g(got["actions"]), want["actions"])
        r = g(got["rewards"])[:, 0]
        assert r.dtype == np.float32
        if reward_type == "sparse":
            np.testing.assert_array_equal(r, want["rewards"])
            np.testing.assert_array_equal(np.signbit(r), np.signbit(want["rewards"]))       # -0.0 for success
        else:
            np.testing.assert_allclose(r, want["rewards"], rtol=1.2e-7, atol=0)             # sqrt of a 3-term sum, then float32
        np.testing.assert_array_equal(g(got["dones"])[:, 0], want["dones"])
        ix = g(got["indices"])
        np.testing.assert_array_equal(ix[:, 0], idx)
        np.testing.assert_array_equal(ix[:want["nb_virtual"], 1], want["source"])
        assert np.all(ix[want["nb_virtual"]:, 1] == -1)
    # device RNG path: every sample is a transition of a complete episode; virtual ones take a goal from [current, end)
    B = 4096
    got = buf.sample(B, return_indices=True)
    assert buf.failed_samples() == 0
    ix = got["indices"].cpu().numpy()
    t, e = np.unravel_index(ix[:, 0], (T, N))
    assert np.all(o.ep_length[t, e] > 0)
    nbv = int(o.her_ratio * B)
    ts, es_ = np.unravel_index(ix[:nbv, 1], (T, N))
    assert np.all(es_ == e[:nbv]) and np.all(ix[nbv:, 1] == -1)
    start, ln = o.ep_start[t[:nbv], e[:nbv]], o.ep_length[t[:nbv], e[:nbv]]
    cur, fut = (t[:nbv] - start) % T, (ts - start) % T
    assert np.all(fut >= cur) and np.all(fut < ln)
    assert (fut > cur).mean() > 0.3                                            # really the future, not just the same step
    counts = np.bincount(ix[:, 0], minlength=T * N)[o.valid_indices()]
    assert counts.min() >= 0 and abs(counts.mean() - B / len(o.valid_indices())) < 1e-9 and counts.max() < 8 * max(1.0, counts.mean()) + 8
    np.testing.assert_array_equal(got["observations"]["desired_goal"].cpu().numpy()[:nbv], o.next_ag[ts, es_])
    buf.close()


@pytest.mark.gpu
def test_device_her_on_env_rollouts():
    # rollout -> replay -> relabel without leaving the device; relabelled rewards equal env.compute_reward
    from mycobotgym_b200.her import DeviceHerReplayBuffer
    from mycobotgym_b200.vector_env import MyCobotVectorEnv

    n = 64
    env = MyCobotVectorEnv(num_envs=n, has_object=True, reward_type="sparse", max_episode_steps=6, seed=4)
    buf = DeviceHerReplayBuffer(20 * n, env, seed=1)
    obs, _ = env.reset()
    gen = torch.Generator(device=env.device).manual_seed(0)
    for t in range(15):
        prev = {k: v.clone() for k, v in obs.items()}
        act = torch.rand(n, 7, device=env.device, generator=gen) * 2 - 1
        out = env.step(act)
        buf.add_step(prev, act, out)
        obs = out[0]
    es, el, nvalid = buf.episode_table()
    assert nvalid == 12 * n and buf.size() == 15 * n                 # two complete 6-step episodes per env, the third is running
    s = buf.sample(512, return_indices=True)
    assert buf.failed_samples() == 0
    nbv = int(0.8 * 512)
    r = env.compute_reward(s["next_observations"]["achieved_goal"], s["observations"]["desired_goal"], None)
    assert torch.equal(r[:nbv], s["rewards"][:nbv, 0])
    # the terminal transition of an episode stores the FINAL observation, not the post-reset one
    ix = s["indices"][:, 0]
    t_idx = (ix // n).cpu().numpy()
    last = np.isin(t_idx, [5, 11])
    assert last.any() and torch.all(s["dones"][torch.as_tensor(last, device=env.device), 0] == 0)      # TimeLimit: done * (1 - timeout) = 0
    nag = s["next_observations"]["achieved_goal"]
    assert torch.equal(nag, s["next_observations"]["observation"][:, 3:6])
    buf.close(); env.close()

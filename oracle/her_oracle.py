"""TEST INFRASTRUCTURE -- CPU restatement of stable_baselines3==2.0.0a0 `HerReplayBuffer` (her/her_replay_buffer.py;
third-party, pinned in the reference's requirements.txt:8, not vendored and not installable here) as
mycobotgym/scripts/train.py:89-97 configures it: n_sampled_goal=4, goal_selection_strategy="future".

Restated from the published algorithm: `add` (episode-start tag per slot, invalidation of the episode being
overwritten, `_compute_episode_length` when an env is done), `sample` (valid = transitions of complete episodes, the
first int(her_ratio * batch) samples are virtual), `_sample_goals` (future: randint(current index in episode, episode
length)), rewards of virtual samples from env.compute_reward(next_achieved_goal, new_goal) (mycobot.py:289-295), dones
returned as done * (1 - timeout).  The random draws are arguments so the CUDA path can be compared bit for bit.
Parity unpinned against SB3 itself (no wheel offline).  Only tests/ may import this module.
"""
import numpy as np


class HerOracle:
    def __init__(self, buffer_steps, n_envs, obs_dim, action_dim, n_sampled_goal=4, reward_type="sparse", distance_threshold=0.01):
        T, N = buffer_steps, n_envs
        self.T, self.N = T, N
        self.obs = np.zeros((T, N, obs_dim)); self.next_obs = np.zeros((T, N, obs_dim))
        self.ag = np.zeros((T, N, 3)); self.next_ag = np.zeros((T, N, 3)); self.dg = np.zeros((T, N, 3))
        self.actions = np.zeros((T, N, action_dim), np.float32)
        self.rewards = np.zeros((T, N), np.float32)
        self.dones = np.zeros((T, N), np.float32); self.timeouts = np.zeros((T, N), np.float32)
        self.ep_start = np.zeros((T, N), np.int64); self.ep_length = np.zeros((T, N), np.int64)
        self.cur_start = np.zeros(N, np.int64)
        self.pos, self.full = 0, False
        self.her_ratio = 1 - (1.0 / (n_sampled_goal + 1))
        self.reward_type, self.thr = reward_type, distance_threshold

    def add(self, obs, ag, dg, next_obs, next_ag, action, reward, terminated, truncated):
        T = self.T
        for e in range(self.N):
            es, el = self.ep_start[self.pos, e], self.ep_length[self.pos, e]
            if el > 0:
                idx = np.arange(self.pos, es + el) % T
                self.ep_length[idx, e] = 0
        self.ep_start[self.pos] = self.cur_start
        p = self.pos
        self.obs[p], self.next_obs[p], self.ag[p], self.next_ag[p], self.dg[p] = obs, next_obs, ag, next_ag, dg
        self.actions[p], self.rewards[p] = action, np.asarray(reward).astype(np.float32)
        done = np.logical_or(terminated, truncated)
        self.dones[p] = done
        self.timeouts[p] = np.logical_and(truncated, np.logical_not(terminated))      # info["TimeLimit.truncated"]
        self.pos += 1
        if self.pos == T:
            self.full, self.pos = True, 0
        for e in range(self.N):
            if done[e]:
                start, end = self.cur_start[e], self.pos
                if end < start:
                    end += T
                idx = np.arange(start, end) % T
                self.ep_length[idx, e] = end - start
                self.cur_start[e] = self.pos

    def valid_indices(self):
        return np.flatnonzero(self.ep_length > 0)

    def compute_reward(self, ag, g):
        d = np.linalg.norm(ag - g, axis=-1)
        if self.reward_type == "sparse":
            return -(d > self.thr).astype(np.float32)
        return -d

    def sample(self, sampled_indices, future_draw):
        """`sampled_indices`: flat indices (np.random.choice(valid_indices, batch)); `future_draw(cur, length)` ->
        np.random.randint(cur, length) replacement for the virtual part."""
        B = len(sampled_indices)
        t, e = np.unravel_index(np.asarray(sampled_indices), self.ep_length.shape)
        nb_virtual = int(self.her_ratio * B)
        dg = self.dg[t, e].copy()
        rew = self.rewards[t, e].copy()
        vt, ve = t[:nb_virtual], e[:nb_virtual]
        es, el = self.ep_start[vt, ve], self.ep_length[vt, ve]
        cur = (vt - es) % self.T
        fut = np.asarray(future_draw(cur, el))
        assert np.all(fut >= cur) and np.all(fut < el)
        tr = (fut + es) % self.T
        new_goal = self.next_ag[tr, ve]
        dg[:nb_virtual] = new_goal
        rew[:nb_virtual] = np.asarray(self.compute_reward(self.next_ag[vt, ve], new_goal)).astype(np.float32)
        return dict(obs=self.obs[t, e], ag=self.ag[t, e], dg=dg, next_obs=self.next_obs[t, e], next_ag=self.next_ag[t, e],
                    actions=self.actions[t, e], rewards=rew, dones=self.dones[t, e] * (1 - self.timeouts[t, e]),
                    future=fut, source=tr * self.N + ve, nb_virtual=nb_virtual)

"""Device-resident HER replay buffer -- host-side mirror of stable_baselines3.HerReplayBuffer as the reference
configures it (mycobotgym/scripts/train.py:89-97: n_sampled_goal=4, goal_selection_strategy="future").

The reference asserts `num_env == 1` for HER (train.py:92) because SB3's buffer of that vintage stores one env; the
device buffer keeps SB3 2.0.0a0's per-(step, env) episode table, so all N envs of a `MyCobotVectorEnv` feed it and
`sample()` relabels on the GPU (`mcb_her_sample` in include/mycobot_b200.h): the rollout -> replay -> relabel ->
compute_reward path never leaves the device.  torch owns the output tensors, the ring lives inside the library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class DeviceHerReplayBuffer:
    def __init__(self, buffer_size, env=None, *, n_envs=None, obs_dim=None, action_dim=None, n_sampled_goal=4,
                 goal_selection_strategy="future", reward_type=None, distance_threshold=None, device=None, seed=0):
        """`buffer_size` counts transitions over all envs like SB3's (the ring holds buffer_size // n_envs steps)."""
        if goal_selection_strategy != "future":
            raise NotImplementedError("only the 'future' strategy (the reference's, train.py:95) is built")
        if env is not None:
            n_envs, obs_dim, action_dim = env.num_envs, env.obs_dim, env.action_dim
            reward_type = env.reward_type if reward_type is None else reward_type
            distance_threshold = env.distance_threshold if distance_threshold is None else distance_threshold
            device = env.device if device is None else device
            self.has_object = env.has_object
        else:
            self.has_object = obs_dim == 25
        if reward_type not in ("sparse", "dense"):
            raise NotImplementedError("reward_shaping cannot be relabelled: it depends on the simulation state (mycobot.py:296-298)")
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceHerReplayBuffer needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:0")
        self.n_envs, self.obs_dim, self.action_dim = int(n_envs), int(obs_dim), int(action_dim)
        self.buffer_steps = max(int(buffer_size) // self.n_envs, 2)
        self.n_sampled_goal = int(n_sampled_goal)
        self.her_ratio = 1 - (1.0 / (self.n_sampled_goal + 1))
        self.reward_type = reward_type
        self._L = _lib.load()
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        with torch.cuda.device(self._dev_index):
            _lib.check(self._L.mcb_her_create(self.n_envs, self.buffer_steps, self.obs_dim, self.action_dim, self.n_sampled_goal,
                                              0 if reward_type == "sparse" else 1, float(distance_threshold), int(seed), C.byref(h)))
        self._h = h
        self._fail = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def size(self):
        return int(self._L.mcb_her_size(self._h))

    def add(self, obs, next_obs, action, reward, terminated, truncated):
        """One transition per env.  `obs` / `next_obs`: dicts with observation / achieved_goal / desired_goal ([N, .]
        float64 CUDA tensors); `next_obs` must already hold the TERMINAL observation where the episode ended
        (SB3's `_store_transition` does that from infos["terminal_observation"]; see `add_step`)."""
        f64 = torch.float64

        def c(t, dt):
            return t.to(device=self.device, dtype=dt).contiguous()

        o, ag, dg = c(obs["observation"], f64), c(obs["achieved_goal"], f64), c(obs["desired_goal"], f64)
        no, nag = c(next_obs["observation"], f64), c(next_obs["achieved_goal"], f64)
        a = c(action, torch.float32)
        r64 = reward.dtype == torch.float64
        r = c(reward, f64 if r64 else torch.float32)
        te, tr = c(terminated, torch.uint8), c(truncated, torch.uint8)
        assert o.shape == (self.n_envs, self.obs_dim) and a.shape == (self.n_envs, self.action_dim) and r.shape == (self.n_envs,)
        with torch.cuda.device(self._dev_index):
            _lib.check(self._L.mcb_her_add(self._h, _ptr(o), _ptr(ag), _ptr(dg), _ptr(no), _ptr(nag), _ptr(a), _ptr(r), int(r64),
                                           _ptr(te), _ptr(tr), self._stream()))

    def add_step(self, prev_obs, actions, step_out):
        """Feed one `MyCobotVectorEnv.step` result: the auto-reset envs' next observation is info["final_observation"]
        (the achieved goal is the object / gripper position inside it, mycobot.py:342-388)."""
        obs, rew, term, trunc, info = step_out
        nobs, nag = obs["observation"], obs["achieved_goal"]
        # MyCobotVectorEnv.step returns its persistent output buffers: a `prev_obs` that was not cloned before the step now
        # holds the NEW observation and the stored transition would have obs == next_obs
        for k in ("observation", "achieved_goal"):
            if torch.is_tensor(prev_obs[k]) and prev_obs[k].data_ptr() == obs[k].data_ptr():
                raise ValueError(f"add_step: prev_obs[{k!r}] aliases the step's output buffer -- clone the observation before calling env.step()")
        if "final_observation" in info:
            done = info["_final_observation"]
            fo = info["final_observation"]
            fag = fo[:, 3:6] if self.has_object else fo[:, 0:3]
            nobs = torch.where(done[:, None], fo, nobs)
            nag = torch.where(done[:, None], fag, nag)
        self.add(prev_obs, {"observation": nobs, "achieved_goal": nag}, actions, rew, term, trunc)

    def alloc_batch(self, batch_size):
        """Output tensors of one `sample()`; pass them back as `out=` to reuse them (no allocation per gradient step)."""
        B, dev, f64 = int(batch_size), self.device, torch.float64
        return dict(obs=torch.empty(B, self.obs_dim, dtype=f64, device=dev), ag=torch.empty(B, 3, dtype=f64, device=dev),
                    dg=torch.empty(B, 3, dtype=f64, device=dev), nobs=torch.empty(B, self.obs_dim, dtype=f64, device=dev),
                    nag=torch.empty(B, 3, dtype=f64, device=dev), act=torch.empty(B, self.action_dim, dtype=torch.float32, device=dev),
                    rew=torch.empty(B, dtype=torch.float32, device=dev), done=torch.empty(B, dtype=torch.float32, device=dev))

    def sample(self, batch_size, *, indices=None, future=None, return_indices=False, out=None):
        """HerReplayBuffer.sample: dict observations / next_observations, actions, rewards [B,1] float32, dones [B,1].
        `indices` (flat step * n_envs + env) and `future` (index inside the episode) inject the random draws."""
        B, dev = int(batch_size), self.device
        if out is None:
            out = self.alloc_batch(B)
        assert out["obs"].shape[0] == B
        idx_out = torch.empty(B, 2, dtype=torch.int64, device=dev) if return_indices else None
        ii = None if indices is None else torch.as_tensor(np.asarray(indices), dtype=torch.int64).to(dev).contiguous()
        ff = None if future is None else torch.as_tensor(np.asarray(future), dtype=torch.int32).to(dev).contiguous()
        with torch.cuda.device(self._dev_index):
            _lib.check(self._L.mcb_her_sample(self._h, B, _ptr(ii), _ptr(ff), _ptr(out["obs"]), _ptr(out["ag"]), _ptr(out["dg"]),
                                              _ptr(out["nobs"]), _ptr(out["nag"]), _ptr(out["act"]), _ptr(out["rew"]), _ptr(out["done"]),
                                              _ptr(idx_out), _ptr(self._fail), self._stream()))
        res = dict(
            observations={"observation": out["obs"], "achieved_goal": out["ag"], "desired_goal": out["dg"]},
            next_observations={"observation": out["nobs"], "achieved_goal": out["nag"], "desired_goal": out["dg"]},
            actions=out["act"], rewards=out["rew"].reshape(-1, 1), dones=out["done"].reshape(-1, 1))
        if return_indices:
            res["indices"] = idx_out
        return res

    def failed_samples(self):
        """Samples of the last `sample()` that found no complete episode (SB3 raises in that case); synchronises."""
        return int(self._fail.item())

    def episode_table(self):
        T, N = self.buffer_steps, self.n_envs
        es = torch.empty(T, N, dtype=torch.int32, device=self.device)
        el = torch.empty(T, N, dtype=torch.int32, device=self.device)
        nv = C.c_int64()
        with torch.cuda.device(self._dev_index):
            _lib.check(self._L.mcb_her_episode_table(self._h, _ptr(es), _ptr(el), C.byref(nv), self._stream()))
        return es, el, int(nv.value)

    def close(self):
        if self._h:
            self._L.mcb_her_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Stable-Baselines3-shaped front end over the device-resident vector env.

The reference trains through SB3 `DummyVecEnv` / `SubprocVecEnv` of single envs wrapped in `Monitor`
(mycobotgym/scripts/train.py:25-33,80-85) and evaluates with `info["is_success"]`
(scripts/eval_model.py:131).  This class offers the same call pattern -- `reset()`, `step_async` /
`step_wait` / `step`, numpy dict observations, `dones`, per-env `infos` with `terminal_observation`,
`TimeLimit.truncated`, `is_success` and Monitor's `episode` record, `env_method("compute_reward", ...)`
for `HerReplayBuffer` -- without importing stable_baselines3 (not installable in this image), so it can be
handed to SB3 algorithms by a maintainer as a `VecEnv` duck type or subclassed from `VecEnv` where SB3 exists.
"""
from __future__ import annotations

import numpy as np

from .vector_env import MyCobotVectorEnv


class MyCobotSB3VecEnv:
    def __init__(self, num_envs, **kwargs):
        # SB3 VecEnvs always reset finished sub-envs inside step_wait; here that happens inside the step kernel, which needs
        # the device sampler.  Configurations that would leave finished episodes running are refused, not half-served.
        if not kwargs.setdefault("auto_reset", True):
            raise ValueError("MyCobotSB3VecEnv: auto_reset=False is not a VecEnv (step_wait must return the first observation of the next episode)")
        if kwargs.get("goal_source", "device") != "device":
            raise ValueError("MyCobotSB3VecEnv: goal_source='reference' resets on the host between steps; use MyCobotVectorEnv.step() for it")
        self.venv = MyCobotVectorEnv(num_envs=num_envs, **kwargs)
        self.num_envs = self.venv.num_envs
        self.observation_space = self.venv.single_observation_space
        self.action_space = self.venv.single_action_space
        self._actions = None
        self._out = None
        self._ep_ret = np.zeros(self.num_envs)
        self._ep_len = np.zeros(self.num_envs, dtype=np.int64)
        self._goal = np.zeros((self.num_envs, 3))      # desired goal of the running episode (for terminal observations)

    # -- VecEnv API ------------------------------------------------------------------------------
    def reset(self):
        obs, _ = self.venv.reset()
        self._ep_ret[:] = 0
        self._ep_len[:] = 0
        out = {k: v.cpu().numpy().copy() for k, v in obs.items()}
        self._goal[:] = out["desired_goal"]
        return out

    def seed(self, seed=None):
        # VecEnv.seed: takes effect at the next reset of each env, like `env.reset(seed=...)` in the reference (mycobot.py:509-510)
        return self.venv.seed(seed)

    def step_async(self, actions):
        self._actions = np.ascontiguousarray(actions, dtype=np.float32)

    def step_wait(self):
        out = self._out = self.venv.step_host(self._actions, self._out, want_final_obs=True)
        term, trunc, succ = out["terminated"].astype(bool), out["truncated"].astype(bool), out["is_success"].astype(bool)
        dones = term | trunc
        rew = out["reward"].astype(np.float64)
        self._ep_ret += rew
        self._ep_len += 1
        infos = [{"is_success": bool(succ[i]), "TimeLimit.truncated": bool(trunc[i] and not term[i])} for i in range(self.num_envs)]
        for i in np.nonzero(dones)[0]:
            fo = out["final_observation"][i].copy()
            ag = fo[3:6] if self.venv.has_object else fo[0:3]          # mycobot.py:258-261
            infos[i]["terminal_observation"] = {"observation": fo, "achieved_goal": ag.copy(), "desired_goal": self._goal[i].copy()}
            infos[i]["episode"] = {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i])}   # Monitor (train.py:27-31)
            self._ep_ret[i] = 0
            self._ep_len[i] = 0
        obs = {k: out[k].copy() for k in ("observation", "achieved_goal", "desired_goal")}
        self._goal[:] = obs["desired_goal"]
        return obs, out["reward"].copy(), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def env_method(self, method_name, *args, indices=None, **kwargs):
        if method_name != "compute_reward":
            raise NotImplementedError(method_name)
        r = self.venv.compute_reward(*args, **kwargs)          # batched: HerReplayBuffer passes [N,3] arrays
        return [r]

    def get_attr(self, attr_name, indices=None):
        return [getattr(self.venv, attr_name)] * self.num_envs

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def close(self):
        self.venv.close()

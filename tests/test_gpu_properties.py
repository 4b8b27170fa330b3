"""Size-independent properties at BASELINE.json's full batch (16384 envs): what must hold for ANY batch, checked where the oracle
is too slow to be run on every env."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N = 16384


def _env(**kw):
    from mycobotgym_b200.vector_env import MyCobotVectorEnv

    kw.setdefault("autotune", False)
    return MyCobotVectorEnv(**kw)


def _rolled(seed=2, steps=8, **kw):
    env = _env(num_envs=N, seed=seed, **kw)
    env.reset()
    env.set_state(elapsed=torch.arange(N, dtype=torch.int32) % 50)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    for _ in range(steps):
        env.step(torch.rand(N, 7, device="cuda", generator=gen) * 2 - 1)
    return env, gen


def test_env_permutation_equivariance():
    """Envs are independent units: stepping a permuted batch gives the permuted result, bit for bit (no cross-env state, no
    dependence on the CTA / warp an env lands in, on its neighbours' tier or on the lockstep grouping)."""
    env, gen = _rolled(has_object=True, reward_type="sparse")
    st = {k: v.clone() for k, v in env.get_state().items()}
    acts = torch.rand(N, 7, device="cuda", generator=gen) * 2 - 1
    env.close()
    perm = torch.randperm(N, device="cuda", generator=gen)
    outs = []
    for p, lw in ((None, 1), (perm, 16)):
        e = _env(num_envs=N, has_object=True, reward_type="sparse", auto_reset=False, lockstep_warps=lw)
        sel = (lambda x: x) if p is None else (lambda x: x[p])
        e.set_state(**{k: sel(st[k]) for k in ("qpos", "qvel", "ctrl", "qacc_warmstart", "goal", "elapsed", "qprev")})
        obs, rew, term, trunc, info = e.step(sel(acts))
        outs.append((e.get_state()["qpos"].clone(), e.get_state()["qvel"].clone(), obs["observation"].clone(), rew.clone(), term.clone()))
        e.close()
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a[perm], b)


def test_forward_is_idempotent_and_masked_reset_touches_only_the_mask():
    env, gen = _rolled(has_object=True, reward_type="sparse")
    o1 = {k: v.clone() for k, v in env.forward().items()}
    s1 = {k: v.clone() for k, v in env.get_state().items()}
    o2 = env.forward()
    s2 = env.get_state()
    for k in o1:
        assert torch.equal(o1[k], o2[k]), k
    for k in ("qpos", "qvel", "ctrl", "goal", "elapsed"):
        assert torch.equal(s1[k], s2[k]), k
    mask = torch.zeros(N, dtype=torch.uint8, device="cuda")
    mask[::7] = 1
    env.reset(mask=mask)
    s3 = env.get_state()
    keep = ~mask.bool()
    for k in ("qpos", "qvel", "ctrl", "goal", "elapsed"):
        assert torch.equal(s1[k][keep], s3[k][keep]), k
    assert bool((s3["elapsed"][mask.bool()] == 0).all()) and bool((s3["qvel"][mask.bool()] == 0).all())
    env.close()


def test_airborne_cubes_follow_the_damped_ballistic_law():
    """Free-joint dynamics in closed form for every env: v' = (v + h g) / (1 + h b / m) per substep (implicit joint damping
    b = 0.01 on the 0.008 kg cube, mycobot280_main.xml:260-263), positions by semi-implicit Euler -- 16384 cubes thrown with
    different velocities, none touching anything for one env-step."""
    env = _env(num_envs=N, has_object=True, reward_type="sparse", auto_reset=False)
    env.reset()
    st = env.get_state()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    qpos, qvel = st["qpos"].clone(), torch.zeros_like(st["qvel"])
    qpos[:, 12:14] = (torch.rand(N, 2, device="cuda", dtype=torch.float64, generator=gen) - 0.5) * 0.2
    qpos[:, 12] += 0.5                               # half a metre beside the robot: nothing to touch
    qpos[:, 14] = 0.6
    qvel[:, 12:15] = (torch.rand(N, 3, device="cuda", dtype=torch.float64, generator=gen) - 0.5) * 0.5
    env.set_state(qpos=qpos, qvel=qvel)
    env.step(torch.zeros(N, 7))
    after = env.get_state()
    h, b_over_m = 0.002, 0.01 / 0.008
    v, p = qvel[:, 12:15].clone(), qpos[:, 12:15].clone()
    g = torch.tensor([0.0, 0.0, -9.81], device="cuda", dtype=torch.float64)
    for _ in range(20):
        v = (v + h * g) / (1 + h * b_over_m)
        p = p + h * v
    assert float((after["qvel"][:, 12:15] - v).abs().max()) < 1e-12
    assert float((after["qpos"][:, 12:15] - p).abs().max()) < 1e-13
    assert float((after["qpos"][:, 15:19].norm(dim=1) - 1).abs().max()) < 1e-14
    env.close()


def test_reward_and_flags_are_functions_of_the_returned_observation():
    """compute_reward(achieved_goal, desired_goal) on the step's own outputs reproduces its reward, success and termination
    for every env (mycobot.py:199-205, 285-298), sparse and dense."""
    for reward_type in ("sparse", "dense"):
        env, gen = _rolled(has_object=True, reward_type=reward_type, steps=4)
        obs, rew, term, trunc, info = env.step(torch.rand(N, 7, device="cuda", generator=gen) * 2 - 1)
        done = term | trunc
        ag, dg = obs["achieved_goal"], obs["desired_goal"]
        keep = ~done                                     # auto-reset envs already show the next episode's first observation
        d = (ag - dg).norm(dim=1)
        r = env.compute_reward(ag, dg, None)
        assert torch.equal(r[keep], rew[keep])
        edge = (d - 0.01).abs() < 1e-12
        assert torch.equal((d < 0.01)[keep & ~edge], info["is_success"][keep & ~edge])
        assert not bool(term[keep].any())
        env.close()

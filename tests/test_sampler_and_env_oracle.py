"""Goal / cube-xy sampler golden vectors and the env-level oracle (mycobot.py:132-400)."""
import os
import random

import numpy as np
import pytest

from mycobotgym_b200 import mjcf
from oracle.oracle import OracleEnv

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def flat():
    return mjcf.load_compiled()


def test_sampler_golden_has_object_seed0(flat):
    env = OracleEnv(flat, has_object=True)
    random.seed(0)
    env.reset(seed=0)
    assert env.sim.qpos[12].hex() == "-0x1.ea6585257338ep-4" and env.sim.qpos[13].hex() == "-0x1.940bb35ca5400p-11"
    assert [v.hex() for v in env.goal] == ["0x1.695e447adb0b8p-4", "-0x1.f77de0424967cp-6", "0x1.ae147ae147ae1p-3"]
    env.reset()
    np.testing.assert_allclose(env.sim.qpos[12:14], [-0.100693, -0.02159345], atol=5e-9)
    np.testing.assert_allclose(env.goal, [0.00190575, 0.05194006, 0.21], atol=5e-9)
    env.reset()
    np.testing.assert_allclose(env.sim.qpos[12:14], [-0.09382612, 0.00615207], atol=5e-9)
    assert env.goal[2].hex() == "0x1.11908d354f20ep-2"


def test_sampler_golden_seed4_and_reach(flat):
    env = OracleEnv(flat, has_object=True)
    random.seed(4)
    env.reset(seed=4)  # the seed used by scripts/test_human_gym.py:24
    np.testing.assert_allclose(env.sim.qpos[12:14], [-0.06334846, -0.04762008], atol=5e-9)
    assert env.goal[2].hex() == "0x1.153bb56ed247ep-2"
    env = OracleEnv(flat, has_object=False)
    random.seed(0)
    env.reset(seed=0)
    np.testing.assert_allclose(env.goal, [-0.11972572, -0.00077066, 0.21], atol=5e-9)
    env.reset()
    np.testing.assert_allclose(env.goal, [-0.100693, -0.02159345, 0.25858354], atol=5e-9)


def test_product_sampler_matches_oracle_protocol(flat):
    from mycobotgym_b200.vector_env import ReferenceGoalSampler

    env = OracleEnv(flat, has_object=True)
    random.seed(0)
    env.reset(seed=0)
    s = ReferenceGoalSampler(1, env.height_offset, env.initial_gripper_xpos[:2], True, True)
    random.seed(0)
    s.seed(0)
    xy, g = s.sample([0])
    assert np.array_equal(xy[0], env.sim.qpos[12:14]) and np.array_equal(g[0], env.goal)


def test_env_semantics(flat):
    env = OracleEnv(flat, has_object=True, reward_type="sparse")
    random.seed(1)
    obs, info = env.reset(seed=1)
    assert obs["observation"].shape == (25,) and obs["achieved_goal"].shape == (3,) and info == {}
    # fresh observation after reset: object_pos == cube qpos, velocities zero
    np.testing.assert_array_equal(obs["observation"][3:6], env.sim.qpos[12:15])
    np.testing.assert_array_equal(obs["observation"][14:25], 0)
    o, r, te, tr, inf = env.step(np.zeros(7, dtype=np.float32))
    assert r.dtype == np.float32 and float(r) == -1.0 and not te and not tr and inf["is_success"] is False
    for _ in range(49):
        o, r, te, tr, inf = env.step(np.zeros(7, dtype=np.float32))
    assert tr and not te                                                     # TimeLimit(50), __init__.py:34
    # success => terminated and truncated (mycobot.py:390-400); sparse reward -0.0 (mycobot.py:293)
    env.goal = o["achieved_goal"].copy()
    assert float(env.compute_reward(o["achieved_goal"], env.goal)) == 0.0
    assert np.signbit(env.compute_reward(o["achieved_goal"], env.goal))
    # batched compute_reward for HER (train.py:93-97)
    ag = np.zeros((5, 3)); g = np.zeros((5, 3)); g[2:, 0] = 0.02
    np.testing.assert_array_equal(env.compute_reward(ag, g), np.array([-0.0, -0.0, -1, -1, -1], dtype=np.float32))
    # at d == threshold: not success, reward -0.0
    g1 = np.array([0.01, 0, 0.0])
    assert float(env.compute_reward(np.zeros(3), g1)) == 0.0


def test_observation_is_one_substep_stale(flat):
    # site poses in the obs come from the last substep's pre-advance kinematics
    env = OracleEnv(flat, has_object=False, reward_type="dense")
    random.seed(0)
    env.reset(seed=0)
    o, *_ = env.step(np.full(7, 0.5, dtype=np.float32))
    stale = o["observation"][:3].copy()
    env.sim.forward()
    fresh = env.sim.site_xpos[1].copy()
    assert 1e-7 < np.abs(stale - fresh).max() < 1e-2
    # block_gripper re-runs forward => fresh (mycobot.py:300-306)
    envb = OracleEnv(flat, has_object=True, block_gripper=True)
    random.seed(0)
    envb.reset(seed=0)
    o, *_ = envb.step(np.full(7, 0.5, dtype=np.float32))
    np.testing.assert_array_equal(o["observation"][:3], envb.sim.site_xpos[1])
    assert envb.sim.qpos[7] == 0 and envb.sim.qpos[9] == 0


@pytest.mark.parametrize("name", ["reach_dense_seed0", "reach_dense_seed1_perturbed", "pick_sparse_seed0",
                                  "pick_sparse_seed4_perturbed", "push_sparse_seed2", "grasp_pick_sparse"])
def test_oracle_reproduces_committed_golden(flat, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    env = OracleEnv(flat, has_object=bool(g["has_object"]), block_gripper=bool(g["block_gripper"]), reward_type=str(g["reward_type"]))
    env.sim.set_state(g["qpos0"], g["qvel0"], g["ctrl0"], g["warm0"])
    env.goal = g["goal"].copy()
    for t in range(len(g["actions"])):
        o, r, te, tr, info = env.step(g["actions"][t])
        np.testing.assert_allclose(env.sim.qpos, g["qpos"][t], atol=1e-12)
        np.testing.assert_allclose(env.sim.qvel, g["qvel"][t], atol=1e-10)
        np.testing.assert_allclose(o["observation"], g["obs"][t], atol=1e-12)
        assert float(r) == pytest.approx(float(g["reward"][t]), abs=1e-12)
        assert te == bool(g["terminated"][t]) and tr == bool(g["truncated"][t])


CTRL_GOLDEN = ["ik_pick_dense_seed3", "ik_fetch_pick_dense_seed5", "mocap_pick_dense_seed6", "mocap_fetch_pick_dense_seed7"]


@pytest.mark.parametrize("name", CTRL_GOLDEN)
def test_oracle_reproduces_controller_golden(flat, name):
    # IK / mocap controller fixtures (tools/make_golden.py controller_rollout): each step restarts from its recorded state,
    # stale frames included (kinematics at qprev), so a regression in either controller path of the oracle shows up here
    from mycobotgym_b200 import mjcf

    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    controller, fetch = str(g["controller"]), bool(g["fetch"])
    fm = mjcf.load_compiled(mjcf.COMPILED_MOCAP) if controller == "mocap" else flat
    env = OracleEnv(fm, has_object=True, reward_type="dense", controller_type=controller, fetch_env=fetch)
    env.goal = g["goal"].copy()
    for t in range(len(g["actions"])):
        # the full previous-substep configuration: the Jacobian's rounding depends on subtree_com, i.e. on every joint, and the
        # bang-bang actuators amplify one ulp to 1e-6 within a 100-substep IK step (the GPU tests use only the arm's `qprev0`)
        q_stale = g["qstale0"][t].copy()
        env.sim.set_state(q_stale, g["qvel0"][t], g["ctrl0"][t], g["warm0"][t])
        env.sim.kinematics()
        env.sim.qpos[:] = g["qpos0"][t]
        env.sim.mocap_pos[:], env.sim.mocap_quat[:] = g["mocap0"][t][:3], g["mocap0"][t][3:]
        o, r, te, tr, info = env.step(g["actions"][t])
        np.testing.assert_allclose(env.sim.qpos, g["qpos"][t], atol=1e-12)
        np.testing.assert_allclose(env.sim.ctrl, g["ctrl"][t], atol=1e-12)
        np.testing.assert_allclose(np.concatenate((env.sim.mocap_pos, env.sim.mocap_quat)), g["mocap"][t], atol=1e-14)
        np.testing.assert_allclose(o["observation"], g["obs"][t], atol=1e-12)
        assert float(r) == pytest.approx(float(g["reward"][t]), abs=1e-12)


def test_staged_reward_on_the_oracle(flat):
    # mycobot.py:402-448: reach term only while the cube is not held; grasp / lift terms once both finger layers touch it
    env = OracleEnv(flat, has_object=True, reward_type="reward_shaping")
    random.seed(3)
    env.reset(seed=3)
    o, r, te, tr, info = env.step(np.zeros(7, dtype=np.float32))
    d = np.linalg.norm(o["observation"][0:3] - o["observation"][3:6])
    assert r == pytest.approx((1 - np.tanh(d)) * 0.2 * 100, abs=1e-12) and r < 20
    g = np.load(os.path.join(GOLDEN, "grasp_pick_sparse.npz"))
    env.sim.set_state(g["qpos0"], g["qvel0"], g["ctrl0"], g["warm0"])
    o, r, te, tr, info = env.step(g["actions"][0])
    tgt = np.array([-0.15, 0.0, 0.21])                       # target0 site, mycobot280_main.xml:83
    lift = 0.5 + (1 - np.tanh(np.linalg.norm(o["observation"][3:6] - tgt))) * 0.4
    assert r == pytest.approx(lift * 100, abs=1e-9) and r > 50


def test_ik_controller_on_the_oracle(flat):
    # mycobot.py:134-170 + utils.py:499-556: the EEF site is driven towards current_pos + 0.2 * action[:3]
    for fetch in (False, True):
        env = OracleEnv(flat, has_object=True, controller_type="IK", fetch_env=fetch)
        random.seed(0)
        o, _ = env.reset(seed=0)
        g0 = o["observation"][:3].copy()
        if fetch:
            assert abs(env.height_offset - 0.209981) < 1e-9                     # keyframe cube height (mycobot280.xml:6)
            np.testing.assert_allclose(env.initial_gripper_xpos, [-0.05154491, 0.01053502, 0.3448586], atol=5e-9)
        a = np.zeros(4 if fetch else 7, dtype=np.float32)
        a[0] = 0.5
        for _ in range(3):
            o, r, te, tr, info = env.step(a)
        moved = o["observation"][:3] - g0
        assert 0.03 < moved[0] < 0.3 and abs(moved[1]) < 0.05
        assert abs(env.sim.ctrl[6] - 0.5) < 1e-15                                # gripper: centre + 0 * range
    # the reference solves an 18 x 18 damped least-squares problem; only the 6 arm columns of the site Jacobian are non-zero
    env = OracleEnv(flat, has_object=True, controller_type="IK")
    env.reset(seed=1)
    jp, jr = env.sim.jac_site(env.site_eef)
    assert np.all(jp[:, 6:] == 0) and np.all(jr[:, 6:] == 0)
    J6 = np.concatenate((jp, jr))[:, :6]
    err = np.array([0.01, -0.02, 0.03, 0.001, 0.002, -0.001])
    full = np.linalg.lstsq(np.concatenate((jp, jr)).T @ np.concatenate((jp, jr)) + 0.3 * np.eye(18), np.concatenate((jp, jr)).T @ err, rcond=-1)[0]
    np.testing.assert_allclose(full[:6], np.linalg.solve(J6.T @ J6 + 0.3 * np.eye(6), J6.T @ err), atol=1e-14)
    assert np.abs(full[6:]).max() < 1e-16


def test_mocap_controller_on_the_oracle():
    # mycobot.py:172-189: the mocap body is placed at the (stale) tool pose + 0.1 * action and a weld drags the arm there
    from mycobotgym_b200 import mjcf

    fm = mjcf.load_compiled(mjcf.COMPILED_MOCAP)
    assert (fm.nbody, fm.nu, fm.neq, fm.nmocap) == (26, 1, 4, 1) and list(fm.eq_type) == [1, 0, 0, 2]
    np.testing.assert_allclose(fm.eq_data[0, 6:10], [np.sqrt(0.5), 0, 0, np.sqrt(0.5)], atol=1e-12)   # tool frame vs mocap frame at qpos0
    for fetch in (False, True):
        env = OracleEnv(fm, has_object=True, controller_type="mocap", fetch_env=fetch)
        random.seed(0)
        o, _ = env.reset(seed=0)
        assert env.sim.nefc >= 13                       # weld (6) + 2 connects (6) + joint coupling (1)
        g0 = o["observation"][:3].copy()
        a = np.zeros(4 if fetch else 8, dtype=np.float32)
        a[0] = 1.0
        if not fetch:
            a[3:7] = env.sim.xquat[fm.body_names.index("gripper_tcp")]
        for _ in range(5):
            o, r, te, tr, info = env.step(a)
        moved = o["observation"][:3] - g0
        if fetch:
            assert 0.15 < moved[0] < 0.6 and abs(moved[1]) < 0.1    # 5 x 0.1 m commanded along x from the bent keyframe pose
        else:
            # from qpos0 the arm is upright (singular): the tool can not translate along x without pitching, and the weld's
            # rotational rows carry the translational inverse weight (2.3.2, pinned by the mocap keyframe), so it sags instead
            assert np.linalg.norm(moved) > 0.1 and np.linalg.norm(env.sim.mocap_pos - o["observation"][:3]) < 0.15
        assert abs(np.linalg.norm(env.sim.mocap_quat) - 1) < 1e-12 and abs(env.sim.ctrl[0] - 0.5) < 1e-15
    if True:
        env = OracleEnv(fm, has_object=True, controller_type="mocap", fetch_env=True)
        np.testing.assert_allclose(env.sim.mocap_pos, [-0.05154491, 0.01053502, 0.3448586], atol=1e-12)   # mycobot280_mocap.xml:8
        assert abs(env.height_offset - 0.209981) < 1e-9


def test_reach_reward_shaping_simulates_the_hidden_cube(flat):
    # MyCobotReach-RewardShaping-* (mycobotgym/__init__.py:6-35): compute_reward -> stage_rewards reads the object0 site
    # (mycobot.py:402-448) of the cube that _env_setup only hid (zero geom / site size, mycobot.py:475-481)
    env = OracleEnv(flat, has_object=False, reward_type="reward_shaping")
    random.seed(0)
    o, _ = env.reset(seed=0)
    assert o["observation"].shape == (10,)
    z0 = env.sim.qpos[14]
    for _ in range(5):
        o, r, te, tr, info = env.step(np.zeros(7, dtype=np.float32))
    assert 0.1999 < env.sim.qpos[14] < 0.2 and z0 > 0.2099              # the zero-size box fell from 0.21 onto the table top (one contact: 4x the depth)
    d = np.linalg.norm(env.sim.site_xpos[env.site_eef] - env.sim.site_xpos[env.site_obj])
    assert abs(float(r) - (1 - np.tanh(d)) * 0.2 * 100) < 1e-12           # reach term only: nothing can grasp a point
    assert np.array_equal(o["achieved_goal"], o["observation"][:3])       # reach: achieved goal = gripper position

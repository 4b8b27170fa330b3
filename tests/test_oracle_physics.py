"""The CPU oracle pinned by the reference's known answers and by physics identities.

PARITY UNPINNED against MuJoCo 2.3.2 itself (not installable here): these tests are what anchors it."""
import numpy as np
import pytest

from mycobotgym_b200 import mjcf
from oracle.oracle import OracleSim, mat2euler


@pytest.fixture(scope="module")
def flat():
    return mjcf.load_compiled()


def test_fk_kat_qpos0_and_keyframe(flat):
    s = OracleSim(flat)
    s.forward()
    np.testing.assert_allclose(s.site_xpos[1], [0.0138673, 0.01864658, 0.61236], atol=5e-9)   # mocap.xml:3
    np.testing.assert_allclose(s.site_xpos[2], [-0.05, 0, 0.21], atol=1e-15)                   # mycobot280_main.xml:260
    s.qpos[:] = flat.key_qpos[0]
    s.forward()
    np.testing.assert_allclose(s.site_xpos[1], [-0.05154491, 0.01053502, 0.3448586], atol=5e-9)  # mycobot280_mocap.xml:8
    # link origins at qpos0
    s.qpos[:] = flat.qpos0
    s.forward()
    np.testing.assert_allclose(s.xpos[3], [0.0038673, -0.2, 0.2774], atol=1e-12)
    np.testing.assert_allclose(s.xpos[8][1:], [-0.11135342, 0.61336], atol=1e-8)


def test_mass_matrix_equals_brute_force_and_rne_identity(flat):
    rng = np.random.default_rng(0)
    s = OracleSim(flat)
    for _ in range(5):
        s.qpos[:12] = rng.uniform(-1, 1, 12)
        q = rng.normal(size=4)
        s.qpos[15:19] = q / np.linalg.norm(q)
        s.qpos[12:15] = rng.uniform(-0.1, 0.1, 3) + [0, 0, 0.5]
        s.qvel[:] = rng.normal(size=18)
        s.forward()
        M2 = mjcf.mass_matrix_numpy(flat, mjcf.fk_numpy(flat, s.qpos.copy()))
        np.testing.assert_allclose(s.M, M2, atol=1e-15, rtol=1e-11)
        assert np.all(np.linalg.eigvalsh(s.M) > 0)
        qacc = rng.normal(size=18)
        np.testing.assert_allclose(s.rne_acc(qacc), s.M @ qacc + s.qfrc_bias, atol=1e-12)


def test_free_cube_in_flight_is_ballistic(flat):
    s = OracleSim(flat)
    s.qpos[14] = 0.6
    s.qvel[12:15] = [0.1, -0.2, 0.3]
    s.qvel[15:18] = [1.0, 2.0, -1.5]
    h, n = flat.timestep, 50
    v0 = s.qvel[12:18].copy()
    s.step(n)
    # linear damping 0.01 on a 0.008 kg body, implicit: v' = (v + h*g) / (1 + h*b/m) per step
    v = v0[:3].copy()
    for _ in range(n):
        v = (v + h * np.array([0, 0, -9.81])) / (1 + h * 0.01 / 0.008)
    np.testing.assert_allclose(s.qvel[12:15], v, rtol=1e-9)
    assert abs(np.linalg.norm(s.qpos[15:19]) - 1) < 1e-12


def test_cube_rests_on_four_corner_contacts(flat):
    """Solver consistency only (Newton reaches the force balance of the rows it is given): the rest depth itself is checked
    against the reference's keyframes in tests/test_keyframe_equilibria.py, where it is an OPEN discrepancy."""
    s = OracleSim(flat)
    s.step(1500)
    assert s.ncon == 4 and s.nefc >= 31
    f = s.efc("force")[-24:]
    assert np.all(f > 0) and abs(f.sum() - 0.008 * 9.81) < 1e-9      # 24 edge forces, each with unit normal component
    assert abs(s.qvel[14]) < 1e-6


def test_newton_solution_satisfies_kkt(flat):
    rng = np.random.default_rng(1)
    s = OracleSim(flat)
    s.qpos[:6] = rng.uniform(-0.5, 0.5, 6)
    s.qpos[6] = s.qpos[8] = -0.01            # gear limits active
    s.qvel[:12] = rng.normal(size=12) * 0.2
    s.ctrl[:] = rng.uniform(-1, 1, 7)
    s.forward()
    J, D, aref, f = s.efc("J"), s.efc("D"), s.efc("aref"), s.efc("force").copy()
    typ = s.efc("type")
    jar = J @ s.qacc - aref
    act = (typ == 0) | (jar < 0)
    np.testing.assert_allclose(f, np.where(act, -D * jar, 0.0), atol=1e-9)
    assert np.all(f[typ != 0] >= 0)
    # stationarity: M qacc - qfrc_smooth - J' f = 0 (to solver tolerance)
    res = s.M @ s.qacc - s.qfrc_smooth - J.T @ f
    assert np.abs(res).max() < 1e-5 * max(1.0, np.abs(s.qfrc_smooth).max())


def test_equality_only_problem_is_solved_in_one_newton_step(flat):
    s = OracleSim(flat, disable_cube=True)
    s.qpos[:6] = [0.3, -0.2, 0.4, 0.1, -0.3, 0.2]
    s.qpos[6] = s.qpos[8] = 0.2
    s.forward()
    assert s.nefc == 7 and s.solver_iter <= 2
    J, D, aref = s.efc("J")[:, :12], s.efc("D"), s.efc("aref")
    M = s.M[:12, :12]
    H = M + J.T @ (D[:, None] * J)
    qacc = np.linalg.solve(H, s.qfrc_smooth[:12] + J.T @ (D * aref))
    np.testing.assert_allclose(s.qacc[:12], qacc, rtol=1e-9, atol=1e-9)


def test_four_bar_closure_stays_closed(flat):
    s = OracleSim(flat, disable_cube=True)
    s.ctrl[6] = 1.0
    s.step(400)
    s.forward()
    assert np.abs(s.efc("pos")[:7]).max() < 2e-3        # soft constraints, (0.005, 1) time constant
    assert 0.05 < s.qpos[6] <= 0.71 and abs(s.qpos[6] - s.qpos[8]) < 5e-3


def test_contact_frames_and_pyramid_rows(flat):
    s = OracleSim(flat)
    s.qpos[14] = 0.21 - 2e-5
    s.forward()
    cons = s.contacts()
    assert len(cons) == 4
    for c in cons:
        np.testing.assert_allclose(c["frame"][0], [0, 0, 1], atol=1e-12)      # table -> cube
        np.testing.assert_allclose(c["frame"] @ c["frame"].T, np.eye(3), atol=1e-12)
        assert abs(c["dist"] + 2e-5) < 1e-12 and c["dim"] == 4
        assert abs(c["pos"][2] - (0.2 - 1e-5)) < 1e-12                          # mid-surface point
    R = s.efc("R")[7:]
    assert np.allclose(R, R[0]) and s.nefc == 31


def test_mat2euler_roundtrip():
    rng = np.random.default_rng(2)
    for _ in range(20):
        e = rng.uniform(-1.2, 1.2, 3)
        cx, cy, cz = np.cos(e)
        sx, sy, sz = np.sin(e)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, cy * 0 + 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        np.testing.assert_allclose(mat2euler(Rx @ Ry @ Rz), e, atol=1e-12)  # rotations.euler2mat convention


@pytest.mark.parametrize("which", ["joint", "mocap"])
def test_invweight0_second_opinion_from_the_c_pipeline(which):
    """body_invweight0 / dof_invweight0 / meaninertia set every constraint's regulariser (the weld's stiffness is
    body_invweight0[gripper_tcp] directly).  mjcf.set_const computes them with its own numpy FK and mass matrix; here they are
    recomputed from the C oracle's pipeline at qpos0 -- mj_kinematics / mj_comPos / mj_crb restated in C, mj_jac for the body
    Jacobians at the inertial frame origins -- i.e. (1/3) tr(J M^-1 J') as mj_setConst defines them, with no code shared."""
    from mycobotgym_b200 import mjcf

    fm = mjcf.load_compiled(mjcf.COMPILED_MOCAP if which == "mocap" else mjcf.COMPILED_JOINT)
    s = OracleSim(fm)
    s.qpos[:] = fm["qpos0"]
    s.forward()                                     # kinematics, comPos, crb at qpos0
    M = s.M.copy()
    nv = fm["nv"]
    Minv = np.linalg.inv(M)
    for b in range(1, fm["nbody"]):
        if fm["body_weldid"][b] == 0:
            assert np.all(fm["body_invweight0"][b] == 0)
            continue
        jp, jr = s.jac(b, s.xipos[b])
        tran, rot = np.trace(jp @ Minv @ jp.T) / 3, np.trace(jr @ Minv @ jr.T) / 3
        np.testing.assert_allclose(fm["body_invweight0"][b], [tran, rot], rtol=1e-10, err_msg=fm["body_names"][b])
    d = np.diag(Minv)
    for j in range(fm["njnt"]):
        a = fm["jnt_dofadr"][j]
        if fm["jnt_type"][j] == 0:                  # free joint: translational and rotational averages
            np.testing.assert_allclose(fm["dof_invweight0"][a:a + 3], d[a:a + 3].mean(), rtol=1e-10)
            np.testing.assert_allclose(fm["dof_invweight0"][a + 3:a + 6], d[a + 3:a + 6].mean(), rtol=1e-10)
        else:
            np.testing.assert_allclose(fm["dof_invweight0"][a], d[a], rtol=1e-10)
    assert abs(fm["stat_meaninertia"] - np.trace(M) / nv) < 1e-12 * np.trace(M)
    if which == "mocap":
        tcp = list(fm["body_names"]).index("gripper_tcp")
        assert abs(fm["body_invweight0"][tcp][0] - 0.394148058) < 1e-8      # the value the weld rows are regularised with

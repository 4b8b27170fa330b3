"""The reference's two keyframes are the only recorded MuJoCo states it holds; they pin the oracle's constraint model.

* `mycobot280_mocap.xml:7-9` (mocap variant): qpos, mocap pose, qvel 0.  The arm has no actuators there, so the state is a
  true equilibrium of gravity vs the mocap weld -- at a near-singular elbow pose (q2 = -0.0057), where the weld's
  translational and rotational rows fight each other with ~20 N / ~5 N m of internal load.  That makes the equilibrium a
  sharp probe of the weld rows' relative regularisation: MuJoCo 2.3.2 evaluates ONE impedance per weld at the norm of the
  6-residual (getposdim) and gives all six rows the translational inverse weight.  With the per-row impedance / rotational
  inverse weight the oracle used in round 1 the arm's residual acceleration at the keyframe is 6.5 rad/s^2 and the weld
  offset settles at 0.42 mm; the recorded offset is 1.149 mm.
* `mycobot280.xml:6-8` (joint variant): a snapshot of the arm *inside* its actuator chatter (kv h / I >> 2) with
  qvel typed as zeros -- no equilibrium, but its `ctrl - qpos` of up to 0.052 rad can only persist in the bang-bang regime,
  and its gripper angles must lie inside the envelope the chatter shakes them through.
* both: cube z = 0.209981, i.e. 1.9e-5 m rest depth.  NOT reproduced (0.96e-5): see `test_cube_rest_depth_open_question`, and
  `test_keyframes_predate_the_current_cube_element` for why the datum is most likely stale.

The keyframe numbers below are typed from the reference XML (CPU test; /root/reference is not read at run time).
"""
import itertools

import numpy as np
import pytest

from mycobotgym_b200 import mjcf
from oracle.oracle import OracleSim

# mycobot280_mocap.xml:7-9
MOCAP_QPOS = np.array([-1.5385, -0.666859, -0.00566179, -0.90322, 1.5695, 1.59939, 5.27478e-05, 0.000325629, 5.27511e-05,
                       0.000330339, 6.72723e-05, -7.0035e-05, 0.0914324, -0.000575927, 0.209981, 1, 1.31323e-13, 1.08636e-12,
                       0.000191033])
MOCAP_MPOS = np.array([-0.05154491, 0.01053502, 0.3448586])
MOCAP_MQUAT = np.array([0.50235287, -0.499, -0.5, 0.49764296])
# mycobot280.xml:6-8
JOINT_QPOS = np.array([-1.53825, -0.641282, 0.0433963, -0.872323, 1.56575, 1.56731, 0.000252284, 0.00058776, 0.000252672,
                       -0.000188413, 0.000429246, -0.00048648, 0.1, -8.13492e-16, 0.209981, 1, -6.01422e-16, 6.04142e-17,
                       -2.59125e-14])
JOINT_CTRL = np.array([-1.55902942e+00, -6.00806595e-01, -8.66798778e-03, -8.63875032e-01, 1.57012168e+00, 1.56181935e+00, 0.0])


@pytest.fixture(scope="module")
def mocap_model():
    return mjcf.load_compiled(mjcf.COMPILED_MOCAP)


def _mocap_sim(fm):
    s = OracleSim(fm)
    s.qpos[:] = MOCAP_QPOS
    s.qvel[:] = 0
    s.mocap_pos[:] = MOCAP_MPOS
    s.mocap_quat[:] = MOCAP_MQUAT
    return s


def test_mocap_keyframe_is_an_equilibrium_of_the_oracle(mocap_model):
    """Residual acceleration AT the recorded state.  The keyframe prints qpos to 6 digits; 1e-6 rad on a 1.3e5 N/m weld is
    ~0.04 N ~ 0.1-0.2 rad/s^2 on the armature-dominated joints, so 0.25 is the print-precision floor (round 1: 6.5)."""
    s = _mocap_sim(mocap_model)
    s.forward()
    tcp = list(mocap_model["body_names"]).index("gripper_tcp")
    off = s.xpos[tcp] - MOCAP_MPOS
    assert abs(np.linalg.norm(off) - 1.1486e-3) < 1e-6          # FK of the keyframe: |tcp - mocap| = 1.149 mm
    assert np.abs(s.qacc[:6]).max() < 0.25, s.qacc[:6]
    assert np.abs(s.qacc[6:12]).max() < 3.0, s.qacc[6:12]        # gripper: armature 0.005 -> 20x the sensitivity


def test_mocap_keyframe_settles_onto_the_recorded_weld_offset(mocap_model):
    s = _mocap_sim(mocap_model)
    tcp = list(mocap_model["body_names"]).index("gripper_tcp")
    for _ in range(5000):
        s.step(1)
    assert np.abs(s.qvel).max() < 1e-9
    off = s.xpos[tcp] - MOCAP_MPOS
    ref = np.array([0.19058e-3, -0.67609e-3, -0.90874e-3])        # FK(keyframe qpos) - mpos
    np.testing.assert_allclose(off, ref, atol=8e-6)               # round 1: (0.070, -0.234, -0.347) mm
    assert abs(np.linalg.norm(off) - 1.1486e-3) < 3e-6
    # gripper four-bar (2 connects + joint coupling + tendon actuator + limits): 6 recorded angles to their printed precision
    np.testing.assert_allclose(s.qpos[6:12], MOCAP_QPOS[6:12], atol=5e-7)
    # arm: the elbow-singular direction is soft (a 2 % change of the gravity torque moves it by 3e-3 rad while the tcp moves
    # by 6 um); the mesh-derived masses of flange / gripper_base are the unverified part of that torque
    np.testing.assert_allclose(s.qpos[:6], MOCAP_QPOS[:6], atol=3.5e-3)   # round 1: 1.0e-2
    np.testing.assert_allclose(s.qpos[[0, 4, 5]], MOCAP_QPOS[[0, 4, 5]], atol=5e-5)


@pytest.mark.xfail(strict=True, reason="open question: reference keyframes rest the cube 1.9e-5 m deep, MuJoCo 2.3.2's documented "
                                      "formulas with 4 corner contacts x 6 pyramid edges give 0.96e-5 m (DESIGN.md section 5)")
def test_cube_rest_depth_open_question(mocap_model):
    """Strict xfail: flips loudly the day the contact model reproduces the recorded depth."""
    s = _mocap_sim(mocap_model)
    for _ in range(3000):
        s.step(1)
    assert abs(s.qpos[14] - 0.209981) < 5e-7


def test_cube_rest_depth_formula_family():
    """What the recorded depth can and cannot be: depth = g * 2 mu^2 (1 + mu^2) (1 - d) / (d^2 K n_edges) is independent of
    the cube's mass, so only the mixing rule, the cone and the number of active pyramid edges enter.  Enumerate the discrete
    alternatives of MuJoCo 2.3.2's contact pipeline (mj_contactParam mixing, condim, contact count, friction): the recorded
    1.9e-5 needs 12 active edges (two condim-4 contacts or three condim-3 contacts), which no face-face box manifold gives;
    four corner contacts x six edges -- what the oracle and the kernel build -- give 0.959e-5."""
    g = 9.81
    cube = dict(ref=(0.001, 1.0), imp=(0.999, 0.999), fr=0.95, dim=4)      # mycobot280_main.xml:262-263
    table = dict(ref=(0.02, 1.0), imp=(0.9, 0.95), fr=1.0, dim=3)          # defaults, mycobot280_main.xml:86-89
    mix = lambda a, b: tuple(0.5 * x + 0.5 * y for x, y in zip(a, b))
    refs = {"mix": mix(cube["ref"], table["ref"]), "cube": cube["ref"], "table": table["ref"]}
    imps = {"mix": mix(cube["imp"], table["imp"]), "cube": cube["imp"], "table": table["imp"]}
    hits, ours = [], None
    for (rn, ref), (im, simp), dim, ncon, mu in itertools.product(refs.items(), imps.items(), (3, 4), (1, 2, 3, 4, 8), (0.95, 1.0)):
        tc = max(ref[0], 2 * 0.002)
        K = 1.0 / (simp[1] ** 2 * tc ** 2 * ref[1] ** 2)
        d0 = simp[0]                                                        # depth << width: impedance = d0
        depth = g * 2 * mu * mu * (1 + mu * mu) * (1 - d0) / (d0 * d0 * K * ncon * 2 * (dim - 1))
        if rn == "mix" and im == "mix" and dim == 4 and ncon == 4 and mu == 1.0:
            ours = depth
        if 1.85e-5 < depth < 1.95e-5:
            hits.append((rn, im, dim, ncon, mu))
    assert abs(ours - 0.959e-5) < 2e-8
    assert hits and all(h[0] == "mix" and h[1] == "mix" and h[3] * (h[2] - 1) == 6 for h in hits), hits


def test_keyframes_predate_the_current_cube_element():
    """Why the recorded depth need not be reproducible at all: the keyframes were saved under an earlier edit of the cube's
    XML element.  (1) The joint keyframe holds an UNTOUCHED cube (y = -8e-16, quaternion 1 to 1e-14) at x = 0.1 exactly, and the
    mocap keyframe a nudged one at x = 0.0914 -- but `mycobot280_main.xml:260` now places the body at x = -0.05, and neither the
    viewer nor the env (which draws the cube's xy at random, mycobot.py:216-227) puts an untouched cube at exactly 0.1.  So the
    `<body name="object0">` element was edited after the keyframes were recorded; its geom's contact attributes sit on the next
    two lines.  (2) Under the very same formulas that give 0.959e-5 today, ordinary earlier attribute sets rest the cube at the
    recorded depth, e.g. condim 3 with solref 0.004 (= 2 timestep) and today's solimp 0.999.  This is a plausibility
    argument, not a fit: nothing in the oracle or the kernel was changed for it, and the strict xfail above stays."""
    fm = mjcf.load_compiled(mjcf.COMPILED_JOINT)
    cube = list(fm["body_names"]).index("object0")
    assert abs(fm["body_pos"][cube][0] - (-0.05)) < 1e-12 and abs(fm["body_pos"][cube][2] - 0.21) < 1e-12
    assert JOINT_QPOS[12] == 0.1 and abs(JOINT_QPOS[13]) < 1e-12 and np.abs(JOINT_QPOS[16:]).max() < 1e-12
    g, table_ref, table_imp = 9.81, 0.02, (0.9, 0.95)
    def depth(dim, tc_cube, d0_cube, dmax_cube, mu_cube):
        tc = max(0.5 * (tc_cube + table_ref), 2 * 0.002)
        d0, dm, mu = 0.5 * (d0_cube + table_imp[0]), 0.5 * (dmax_cube + table_imp[1]), max(1.0, mu_cube)
        return g * 2 * mu * mu * (1 + mu * mu) * (1 - d0) * dm * dm * tc * tc / (d0 * d0 * 4 * 2 * (dim - 1))
    assert abs(depth(4, 0.001, 0.999, 0.999, 0.95) - 0.959e-5) < 2e-8             # today's element
    recorded = lambda d: abs(round(0.21 - d, 6) - 0.209981) < 1e-9                # what the keyframe's six digits print
    assert recorded(depth(3, 0.004, 0.999, 0.999, 0.95)) and recorded(depth(3, 0.004, 0.998, 0.998, 0.95)) and recorded(depth(4, 0.004, 0.95, 0.99, 0.95))


def test_joint_keyframe_is_a_chatter_snapshot():
    fm = mjcf.load_compiled(mjcf.COMPILED_JOINT)
    err_key = JOINT_CTRL[:6] - JOINT_QPOS[:6]
    # (1) dead band of the bang-bang regime: force = clamp(kp e - kv v) flips with v every substep while kv |v| - fmax > kp |e|;
    #     |v| = h fmax / I with I >= armature.  The recorded errors are far above the static error fmax-free PD would leave
    #     (gravity torque / kp ~ 2e-4) and inside the band.
    kp = np.array([4500, 4500, 3500, 2000, 2000, 2000.0])
    kv = kp / 10
    fmax = np.array([87, 87, 87, 12, 12, 12.0])
    h, arm = fm["timestep"], 0.1
    band = (kv * h * fmax / arm - fmax) / kp
    assert np.all(np.abs(err_key[:3]) > 0.02) and np.all(np.abs(err_key) < 1.45 * band), (err_key, band)
    # (2) the oracle started from the snapshot stays in that regime (errors frozen, velocities banging) and shakes the
    #     gripper through an envelope that contains the six recorded gripper angles
    s = OracleSim(fm)
    s.qpos[:] = JOINT_QPOS
    s.qvel[:] = 0
    s.ctrl[:] = JOINT_CTRL
    errs, grip, vmax = [], [], 0.0
    for i in range(4000):
        s.step(1)
        if i >= 1000:
            errs.append(JOINT_CTRL[:6] - s.qpos[:6])
            grip.append(s.qpos[6:12].copy())
            vmax = max(vmax, np.abs(s.qvel[:6]).max())
    errs, grip = np.array(errs), np.array(grip)
    assert vmax > 0.5                                                      # still banging after 8 s
    assert np.all(np.abs(errs.mean(0)[:3]) > 0.02)                         # errors of the recorded size persist
    np.testing.assert_allclose(errs.mean(0)[1:4], err_key[1:4], rtol=0.2)  # joints 2-4 froze near their recorded errors
    assert np.all(grip.min(0) <= JOINT_QPOS[6:12]) and np.all(JOINT_QPOS[6:12] <= grip.max(0))

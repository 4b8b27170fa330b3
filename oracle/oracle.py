"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see oracle/mjc_oracle.c header).

`OracleSim` wraps one (model, data) pair of the C restatement; `OracleEnv` restates the
reference's task logic (mycobotgym/envs/mycobot.py:132-133,190-306,342-400,450-481,506-514 and
mycobotgym/utils.py:14-26, 469-556) on top of it (joint, IK and mocap controllers).  PARITY UNPINNED vs MuJoCo
2.3.2 (not installable here); pinned by the reference's FK known answers, sampler protocol and recorded keyframes
(tests/test_keyframe_equilibria.py).
"""
from __future__ import annotations

import ctypes as C
import os
import random
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmjc_oracle.so")

_INT_SCALARS = ["nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "neq", "nexclude", "nM",
                "iterations", "ls_iterations", "disable_cube", "nmocap"]
_DBL_SCALARS = ["timestep", "tolerance", "ls_tolerance", "impratio", "meaninertia"]
_PTRS = [
    ("i", ["body_parentid", "body_rootid", "body_weldid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr", "body_mocapid"]),
    ("d", ["body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_subtreemass", "body_invweight0"]),
    ("i", ["jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited"]),
    ("d", ["jnt_pos", "jnt_axis", "jnt_range", "jnt_margin", "jnt_solref", "jnt_solimp"]),
    ("i", ["dof_bodyid", "dof_jntid", "dof_parentid", "dof_Madr"]),
    ("d", ["dof_armature", "dof_damping", "dof_invweight0"]),
    ("i", ["geom_type", "geom_bodyid", "geom_condim", "geom_contype", "geom_conaffinity"]),
    ("d", ["geom_pos", "geom_quat", "geom_size", "geom_friction", "geom_solref", "geom_solimp", "geom_solmix", "geom_margin", "geom_gap", "geom_rbound"]),
    ("i", ["site_bodyid"]),
    ("d", ["site_pos", "site_quat"]),
    ("i", ["eq_type", "eq_obj1id", "eq_obj2id"]),
    ("d", ["eq_data", "eq_solref", "eq_solimp"]),
    ("i", ["exclude"]),
    ("d", ["actuator_moment", "actuator_gain", "actuator_biasprm", "actuator_ctrlrange", "actuator_forcerange"]),
    ("i", ["actuator_ctrllimited", "actuator_forcelimited"]),
    ("d", ["qpos0"]),
]
# appended to the struct after the pointer block above: the convex hulls of the mesh geoms
_HULL_INTS = ["nhull", "mesh_collision"]
_HULL_PTRS = [("i", ["hull_bodyid", "hull_mult", "hull_vertadr", "hull_vertnum", "hull_condim"]),
              ("d", ["hull_vert", "hull_center", "hull_rbound", "hull_friction", "hull_solref", "hull_solimp", "hull_solmix"])]


def _fields():
    f = [(n, C.c_int) for n in _INT_SCALARS] + [(n, C.c_double) for n in _DBL_SCALARS] + [("gravity", C.c_double * 3)]
    for kind, names in _PTRS:
        for n in names:
            f.append((n, C.POINTER(C.c_int if kind == "i" else C.c_double)))
    f += [(n, C.c_int) for n in _HULL_INTS]
    for kind, names in _HULL_PTRS:
        for n in names:
            f.append((n, C.POINTER(C.c_int if kind == "i" else C.c_double)))
    return f


class OModel(C.Structure):
    _fields_ = _fields()


def build_lib(force=False):
    src = os.path.join(HERE, "mjc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_lib()
        _lib = C.CDLL(LIB_PATH)
        _lib.o_sizeof_data.restype = C.c_int
        _lib.o_sizeof_model.restype = C.c_int
        assert _lib.o_sizeof_model() == C.sizeof(OModel), (_lib.o_sizeof_model(), C.sizeof(OModel))
        for name in ["mocap_pos", "mocap_quat", "qpos", "qvel", "ctrl", "qacc_warmstart", "qacc", "xpos", "xquat", "xmat", "xipos", "site_xpos", "site_xmat",
                     "geom_xpos", "geom_xmat", "subtree_com", "cdof", "cinert", "Mfull", "qfrc_bias", "qfrc_smooth", "qacc_smooth",
                     "qfrc_constraint", "qfrc_actuator", "qfrc_passive", "actuator_force", "efc_J", "efc_pos", "efc_D", "efc_R",
                     "efc_aref", "efc_force", "efc_diagApprox", "efc_KBIP"]:
            getattr(_lib, "o_" + name).restype = C.POINTER(C.c_double)
        _lib.o_efc_type.restype = C.POINTER(C.c_int)
    return _lib


class OracleSim:
    """One environment of the CPU restatement."""

    def __init__(self, flat, disable_cube=False, mesh_collision=None):
        """`mesh_collision`: collide the convex hulls of the mesh geoms too (mjc_Convex / mjc_PlaneConvex restated as MPR);
        default: the model's own "mesh_collision" entry (absent = off)."""
        self.flat = flat
        self.L = lib()
        self._keep = []
        om = OModel()
        for n in _INT_SCALARS:
            if n in ("nexclude", "disable_cube"):
                continue
            setattr(om, n, int(flat[n]))
        om.nexclude = int(flat["exclude"].shape[0])
        om.disable_cube = int(disable_cube)
        for n in ["timestep", "tolerance", "ls_tolerance", "impratio"]:
            setattr(om, n, float(flat[n]))
        om.meaninertia = float(flat["stat_meaninertia"])
        om.gravity[:] = list(flat["gravity"])
        for kind, names in _PTRS:
            for n in names:
                a = np.ascontiguousarray(flat[n], dtype=np.int32 if kind == "i" else np.float64)
                self._keep.append(a)
                setattr(om, n, a.ctypes.data_as(C.POINTER(C.c_int if kind == "i" else C.c_double)))
        om.nhull = int(flat.get("nhull", 0))
        om.mesh_collision = int(bool(flat.get("mesh_collision", False) if mesh_collision is None else mesh_collision)) if om.nhull else 0
        if om.nhull:
            for kind, names in _HULL_PTRS:
                for n in names:
                    a = np.ascontiguousarray(flat[n], dtype=np.int32 if kind == "i" else np.float64)
                    self._keep.append(a)
                    setattr(om, n, a.ctypes.data_as(C.POINTER(C.c_int if kind == "i" else C.c_double)))
        self.om = om
        self.nq, self.nv, self.nu = int(flat["nq"]), int(flat["nv"]), int(flat["nu"])
        self.buf = C.create_string_buffer(self.L.o_sizeof_data())
        self.d = C.cast(self.buf, C.c_void_p)
        self.qpos[:] = flat["qpos0"]
        if self.flat["nmocap"]:
            mb = list(flat["body_mocapid"]).index(0)
            self.mocap_pos[:] = flat["body_pos"][mb]
            self.mocap_quat[:] = flat["body_quat"][mb]
        else:
            self.mocap_quat[:] = [1, 0, 0, 0]

    def _arr(self, name, n, dtype=np.float64):
        p = getattr(self.L, "o_" + name)(self.d)
        return np.ctypeslib.as_array(p, shape=(n,))

    mocap_pos = property(lambda s: s._arr("mocap_pos", 3))
    mocap_quat = property(lambda s: s._arr("mocap_quat", 4))
    qpos = property(lambda s: s._arr("qpos", s.nq))
    qvel = property(lambda s: s._arr("qvel", s.nv))
    ctrl = property(lambda s: s._arr("ctrl", s.nu))
    qacc = property(lambda s: s._arr("qacc", s.nv))
    qacc_warmstart = property(lambda s: s._arr("qacc_warmstart", s.nv))
    qacc_smooth = property(lambda s: s._arr("qacc_smooth", s.nv))
    qfrc_bias = property(lambda s: s._arr("qfrc_bias", s.nv))
    qfrc_smooth = property(lambda s: s._arr("qfrc_smooth", s.nv))
    qfrc_constraint = property(lambda s: s._arr("qfrc_constraint", s.nv))
    qfrc_actuator = property(lambda s: s._arr("qfrc_actuator", s.nv))
    actuator_force = property(lambda s: s._arr("actuator_force", s.nu))
    xpos = property(lambda s: s._arr("xpos", 3 * s.flat["nbody"]).reshape(-1, 3))
    xquat = property(lambda s: s._arr("xquat", 4 * s.flat["nbody"]).reshape(-1, 4))
    xmat = property(lambda s: s._arr("xmat", 9 * s.flat["nbody"]).reshape(-1, 3, 3))
    xipos = property(lambda s: s._arr("xipos", 3 * s.flat["nbody"]).reshape(-1, 3))
    site_xpos = property(lambda s: s._arr("site_xpos", 3 * s.flat["nsite"]).reshape(-1, 3))
    site_xmat = property(lambda s: s._arr("site_xmat", 9 * s.flat["nsite"]).reshape(-1, 3, 3))
    geom_xpos = property(lambda s: s._arr("geom_xpos", 3 * s.flat["ngeom"]).reshape(-1, 3))
    subtree_com = property(lambda s: s._arr("subtree_com", 3 * s.flat["nbody"]).reshape(-1, 3))
    cdof = property(lambda s: s._arr("cdof", 6 * s.nv).reshape(-1, 6))

    @property
    def M(self):
        return self._arr("Mfull", self.nv * self.nv).reshape(self.nv, self.nv)

    @property
    def nefc(self):
        return self.L.o_nefc(self.d)

    @property
    def ncon(self):
        return self.L.o_ncon(self.d)

    @property
    def solver_iter(self):
        return self.L.o_solver_iter(self.d)

    def efc(self, name):
        n = self.nefc
        if name == "J":
            return self._arr("efc_J", n * self.nv).reshape(n, self.nv)
        if name == "type":
            return np.ctypeslib.as_array(self.L.o_efc_type(self.d), shape=(n,))
        if name == "KBIP":
            return self._arr("efc_KBIP", 4 * n).reshape(n, 4)
        return self._arr("efc_" + name, n)

    def contacts(self):
        out = []
        buf = (C.c_double * 16)()
        for i in range(self.ncon):
            self.L.o_contact(self.d, i, buf)
            a = np.array(buf[:])
            out.append(dict(dist=a[0], pos=a[1:4].copy(), frame=a[4:13].reshape(3, 3).copy(), geom1=int(a[13]), geom2=int(a[14]), dim=int(a[15])))
        return out

    def forward(self):
        self.L.o_forward(C.byref(self.om), self.d)

    def step(self, nstep=1):
        self.L.o_step(C.byref(self.om), self.d, int(nstep))

    def kinematics(self):
        self.L.o_kinematics(C.byref(self.om), self.d)
        self.L.o_compos(C.byref(self.om), self.d)

    def jac_site(self, site):
        jp = np.zeros((3, self.nv))
        jr = np.zeros((3, self.nv))
        self.L.o_jac_site(C.byref(self.om), self.d, jp.ctypes.data_as(C.POINTER(C.c_double)), jr.ctypes.data_as(C.POINTER(C.c_double)), int(site))
        return jp, jr

    def jac(self, body, point):
        jp = np.zeros((3, self.nv))
        jr = np.zeros((3, self.nv))
        pt = np.ascontiguousarray(point, dtype=np.float64)
        self.L.o_jac(C.byref(self.om), self.d, jp.ctypes.data_as(C.POINTER(C.c_double)), jr.ctypes.data_as(C.POINTER(C.c_double)),
                     pt.ctypes.data_as(C.POINTER(C.c_double)), int(body))
        return jp, jr

    def rne_acc(self, qacc):
        out = np.zeros(self.nv)
        qa = np.ascontiguousarray(qacc, dtype=np.float64)
        self.L.o_rne_acc(C.byref(self.om), self.d, qa.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def get_state(self):
        return dict(qpos=self.qpos.copy(), qvel=self.qvel.copy(), ctrl=self.ctrl.copy(), qacc_warmstart=self.qacc_warmstart.copy())

    def set_state(self, qpos=None, qvel=None, ctrl=None, qacc_warmstart=None):
        if qpos is not None:
            self.qpos[:] = qpos
        if qvel is not None:
            self.qvel[:] = qvel
        if ctrl is not None:
            self.ctrl[:] = ctrl
        if qacc_warmstart is not None:
            self.qacc_warmstart[:] = qacc_warmstart


# ----------------------------------------------------------------------------------------
# gymnasium_robotics.utils.rotations.mat2euler (1.2.0), restated


def mat2euler(mat):
    mat = np.asarray(mat, dtype=np.float64)
    cy = np.sqrt(mat[..., 2, 2] * mat[..., 2, 2] + mat[..., 1, 2] * mat[..., 1, 2])
    condition = cy > np.finfo(np.float64).eps * 4.0
    euler = np.empty(mat.shape[:-1], dtype=np.float64)
    euler[..., 2] = np.where(condition, -np.arctan2(mat[..., 0, 1], mat[..., 0, 0]), -np.arctan2(-mat[..., 1, 0], mat[..., 1, 1]))
    euler[..., 1] = np.where(condition, -np.arctan2(-mat[..., 0, 2], cy), -np.arctan2(-mat[..., 0, 2], cy))
    euler[..., 0] = np.where(condition, -np.arctan2(mat[..., 1, 2], mat[..., 2, 2]), 0.0)
    return euler


def goal_distance(a, b):
    """mycobotgym/utils.py:24-26"""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    return np.linalg.norm(a - b, axis=-1)


class ReferenceSampler:
    """The reference's goal / cube-xy sampling protocol (mycobot.py:207-243, utils.py:14-21):
    x,y from the GLOBAL stdlib `random`, the in-the-air coin and offset from the env's numpy
    Generator (gymnasium seeding.np_random == Generator(PCG64(SeedSequence(seed))))."""

    def __init__(self, height_offset, target_in_the_air=True, rng=None):
        self.height_offset = height_offset
        self.target_in_the_air = target_in_the_air
        self.np_random = rng if rng is not None else np.random.Generator(np.random.PCG64(np.random.SeedSequence()))

    def seed(self, seed):
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    def sample_goal(self):
        x = random.uniform(-0.12, 0.12)
        y = random.uniform(-0.06, 0.06)
        goal = [x, y, self.height_offset]
        if self.target_in_the_air and self.np_random.uniform() < 0.5:
            goal[2] += self.np_random.uniform(0, 0.1)
        return np.array(goal)


# ----------------------------------------------------------------------------------------
# helpers of the IK branch: mujoco.mju_mat2Quat / mju_mulQuat / mju_negQuat / mju_quat2Vel (engine_util_spatial.c)
# and gymnasium_robotics.utils.rotations.euler2quat, restated


def mju_mat2quat(mat):
    m = np.asarray(mat, dtype=np.float64).ravel()
    q = np.zeros(4)
    if m[0] + m[4] + m[8] > 0:
        q[0] = 0.5 * np.sqrt(1 + m[0] + m[4] + m[8])
        q[1] = 0.25 * (m[7] - m[5]) / q[0]
        q[2] = 0.25 * (m[2] - m[6]) / q[0]
        q[3] = 0.25 * (m[3] - m[1]) / q[0]
    elif m[0] > m[4] and m[0] > m[8]:
        q[1] = 0.5 * np.sqrt(1 + m[0] - m[4] - m[8])
        q[0] = 0.25 * (m[7] - m[5]) / q[1]
        q[2] = 0.25 * (m[1] + m[3]) / q[1]
        q[3] = 0.25 * (m[2] + m[6]) / q[1]
    elif m[4] > m[8]:
        q[2] = 0.5 * np.sqrt(1 - m[0] + m[4] - m[8])
        q[0] = 0.25 * (m[2] - m[6]) / q[2]
        q[1] = 0.25 * (m[1] + m[3]) / q[2]
        q[3] = 0.25 * (m[5] + m[7]) / q[2]
    else:
        q[3] = 0.5 * np.sqrt(1 - m[0] - m[4] + m[8])
        q[0] = 0.25 * (m[3] - m[1]) / q[3]
        q[1] = 0.25 * (m[2] + m[6]) / q[3]
        q[2] = 0.25 * (m[5] + m[7]) / q[3]
    return q / np.linalg.norm(q)


def mju_mulquat(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def mju_quat2vel(quat, dt):
    axis = np.array(quat[1:4], dtype=np.float64)
    sin_a_2 = np.linalg.norm(axis)
    axis = axis / sin_a_2 if sin_a_2 >= 1e-15 else np.array([1.0, 0, 0])
    speed = 2 * np.arctan2(sin_a_2, quat[0])
    if speed > np.pi:
        speed -= 2 * np.pi
    return axis * (speed / dt)


def euler2quat(euler):
    euler = np.asarray(euler, dtype=np.float64)
    ai, aj, ak = euler[2] / 2, -euler[1] / 2, euler[0] / 2
    si, sj, sk = np.sin(ai), np.sin(aj), np.sin(ak)
    ci, cj, ck = np.cos(ai), np.cos(aj), np.cos(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([cj * cc + sj * ss, cj * cs - sj * sc, -(cj * ss + sj * cc), cj * sc - sj * cs])


class OracleEnv:
    """Restatement of MyCobotEnv (joint and IK controllers, incl. the Fetch IK variant) + TimeLimit(50) on the CPU oracle."""

    def __init__(self, flat, has_object=True, block_gripper=False, target_in_the_air=True, distance_threshold=0.01,
                 reward_type="sparse", frame_skip=20, max_episode_steps=50, controller_type="joint", fetch_env=False,
                 control_steps=5):
        self.flat = flat
        self.controller_type, self.fetch_env, self.control_steps = controller_type, fetch_env, control_steps
        assert controller_type in ("joint", "IK", "mocap") and not (fetch_env and controller_type == "joint")
        assert (controller_type == "mocap") == bool(flat["nmocap"]), "the mocap controller needs the mocap model variant (mycobot280_mocap.xml)"
        self.has_object, self.block_gripper = has_object, block_gripper
        self.target_in_the_air, self.distance_threshold = target_in_the_air, distance_threshold
        self.reward_type, self.frame_skip, self.max_episode_steps = reward_type, frame_skip, max_episode_steps
        if not has_object and reward_type == "reward_shaping":
            # "Hide object in Reach env" (mycobot.py:475-481): the cube stays in the simulation with a zero-size box (geom_rbound
            # keeps its compiled value); only this reward reads it (mycobot.py:402-448), so only here it is not frozen
            hidden = type(flat)(flat)
            gs = np.array(flat["geom_size"], dtype=np.float64).copy()
            gs[list(flat["geom_names"]).index("object0")] = 0.0
            hidden["geom_size"] = gs
            self.sim = OracleSim(hidden, disable_cube=False)
        else:
            self.sim = OracleSim(flat, disable_cube=not has_object)
        self.site_eef = flat["site_names"].index("EEF")
        self.site_obj = flat["site_names"].index("object0")
        jn = flat["jnt_names"]
        self.robot_jnts = [i for i, n in enumerate(jn) if n.startswith("robot")]
        self.j_rf, self.j_lf = jn.index("right_finger_joint"), jn.index("left_finger_joint")
        self.j_obj = jn.index("object0:joint")
        self.goal = np.zeros(3)
        # _env_setup (mycobot.py:450-481): fetch envs start from keyframe 0 (mycobot280.xml:4-9)
        if fetch_env:
            self.sim.qpos[:] = flat["key_qpos"][0]
            self.sim.qvel[:] = flat["key_qvel"][0]
            self.sim.ctrl[:] = flat["key_ctrl"][0]
            if flat["nmocap"]:
                self.sim.mocap_pos[:] = flat["key_mpos"][0]
                self.sim.mocap_quat[:] = flat["key_mquat"][0]
        self.sim.forward()
        self.initial_gripper_xpos = self.sim.site_xpos[self.site_eef].copy()
        self.height_offset = float(self.sim.site_xpos[self.site_obj][2])
        self.init_qpos = self.sim.qpos.copy()
        self.init_qvel = self.sim.qvel.copy()
        self.init_ctrl = self.sim.ctrl.copy()
        self.sampler = ReferenceSampler(self.height_offset, target_in_the_air)
        self.elapsed = 0
        self.dt = frame_skip * flat["timestep"]

    # mycobot.py:342-388 + 245-283
    def _get_obs(self):
        s = self.sim
        grip_pos = s.site_xpos[self.site_eef].copy()
        jp_e, _ = s.jac_site(self.site_eef)
        grip_velp = jp_e @ s.qvel * self.dt
        qa = [self.flat["jnt_qposadr"][j] for j in self.robot_jnts]
        da = [self.flat["jnt_dofadr"][j] for j in self.robot_jnts]
        robot_qpos, robot_qvel = s.qpos[qa].copy(), s.qvel[da].copy()
        if self.has_object:
            object_pos = s.site_xpos[self.site_obj].copy()
            object_rot = mat2euler(s.site_xmat[self.site_obj])
            jp_o, jr_o = s.jac_site(self.site_obj)
            object_velp = jp_o @ s.qvel * self.dt
            object_velr = jr_o @ s.qvel * self.dt
            object_rel_pos = object_pos - grip_pos
            object_velp = object_velp - grip_velp
        else:
            object_pos = object_rot = object_velp = object_velr = object_rel_pos = np.zeros(0)
        gripper_state = robot_qpos[-2:]
        gripper_vel = robot_qvel[-2:] * self.dt
        achieved = grip_pos.copy() if not self.has_object else object_pos.copy()
        obs = np.concatenate([grip_pos, object_pos.ravel(), object_rel_pos.ravel(), gripper_state, object_rot.ravel(),
                              object_velp.ravel(), object_velr.ravel(), grip_velp, gripper_vel])
        self.achieved_goal = achieved.copy()
        return {"observation": obs, "achieved_goal": achieved, "desired_goal": self.goal.copy()}

    def compute_reward(self, achieved_goal, goal, info=None):
        d = goal_distance(achieved_goal, goal)
        if self.reward_type == "sparse":
            return -(d > self.distance_threshold).astype(np.float32)
        if self.reward_type == "dense":
            return -d
        if self.reward_type == "reward_shaping":
            return max(self.stage_rewards()) * 100

    def stage_rewards(self):
        """mycobot.py:402-448; target0 keeps its XML position because only render() moves it (mycobot.py:309-311)."""
        s, gn = self.sim, self.flat["geom_names"]
        grip_pos, object_pos = s.site_xpos[self.site_eef], s.site_xpos[self.site_obj]
        target_pos = s.site_xpos[self.flat["site_names"].index("target0")]
        r_reach = (1 - np.tanh(goal_distance(grip_pos, object_pos))) * 0.2
        rl, ll, ob = gn.index("right_finger_layer"), gn.index("left_finger_layer"), gn.index("object0")
        pairs = {(c["geom1"], c["geom2"]) for c in s.contacts()}
        touch = lambda g: (g, ob) in pairs or (ob, g) in pairs
        r_grasp = int(touch(rl) and touch(ll)) * 0.5
        r_lift = 0.0
        if r_grasp > 0.0:
            r_lift = 0.5 + (1 - np.tanh(goal_distance(object_pos, target_pos))) * (0.9 - 0.5)
        return r_reach, r_grasp, r_lift

    def reset(self, seed=None, object_xy=None, goal=None):
        """mycobot.py:506-514 + 207-236.  `object_xy`/`goal` inject sampler outputs (parity harness)."""
        if seed is not None:
            self.sampler.seed(seed)
        s = self.sim
        s.qpos[:] = self.init_qpos
        s.qvel[:] = self.init_qvel
        s.ctrl[:] = self.init_ctrl
        s.forward()
        object_xpos = self.initial_gripper_xpos[:2]
        if self.has_object:
            if object_xy is None:
                while np.linalg.norm(object_xpos - self.initial_gripper_xpos[:2]) < 0.1:
                    object_xpos = self.sampler.sample_goal()[:2]
            else:
                object_xpos = np.asarray(object_xy, dtype=np.float64)
            qa = self.flat["jnt_qposadr"][self.j_obj]
            s.qpos[qa:qa + 2] = object_xpos
        s.forward()
        if goal is None:
            self.goal = self.sampler.sample_goal()
            while np.linalg.norm(self.goal[:2] - object_xpos) < 0.1:
                self.goal = self.sampler.sample_goal()
        else:
            self.goal = np.asarray(goal, dtype=np.float64).copy()
        self.elapsed = 0
        return self._get_obs(), {}

    # mycobotgym/utils.py:499-556 (IKController.compute_qpos_delta / solve_DLS)
    def _ik_delta(self, target_pos, target_quat):
        s = self.sim
        err = np.empty(6)
        err[:3] = target_pos - s.site_xpos[self.site_eef]
        eef_q = mju_mat2quat(s.site_xmat[self.site_eef])
        neg = np.array([eef_q[0], -eef_q[1], -eef_q[2], -eef_q[3]])
        err[3:] = mju_quat2vel(mju_mulquat(target_quat, neg), 50)
        jp, jr = s.jac_site(self.site_eef)
        jac = np.concatenate((jp, jr), axis=0)
        hess = jac.T.dot(jac) + np.eye(jac.shape[1]) * 0.3
        return np.linalg.lstsq(hess, jac.T.dot(err), rcond=-1)[0]

    def step(self, action):
        action = np.clip(np.asarray(action, dtype=np.float32), np.float32(-1.0), np.float32(1.0))
        s = self.sim
        if self.controller_type == "IK":      # mycobot.py:134-170
            target_pos = s.site_xpos[self.site_eef] + action[:3] * np.float32(0.2)     # float32 product, float64 sum
            if self.fetch_env:
                target_quat = np.array([0, -0.707, 0, 0.707])
            else:
                quat_rot = euler2quat(action[3:6] * np.float32(0.5))
                target_quat = mju_mulquat(quat_rot, mju_mat2quat(s.site_xmat[self.site_eef]))
            ctrl_action = np.zeros(7)
            ctrl_action[-1] = 0.5 + np.float64(action[-1]) * 0.5                         # actuation_center / range of ctrlrange [0, 1]
            for _ in range(self.control_steps):
                delta = self._ik_delta(target_pos, target_quat)
                ctrl_action[:6] = s.ctrl[:6] + delta[:6]
                s.ctrl[:] = ctrl_action
                s.step(self.frame_skip)
        elif self.controller_type == "mocap":   # mycobot.py:172-189 + gymnasium_robotics mocap_set_action / reset_mocap2body_xpos
            tcp = self.flat["body_names"].index("gripper_tcp")
            mocap_action = np.zeros(7)
            mocap_action[:3] = action[:3] * np.float32(0.1)                       # float32 product
            grip_tcp_quat = s.xquat[tcp].copy()                                   # stale frame, like every site / body pose here
            mocap_action[3:7] = np.array([0.5, -0.5, -0.5, 0.5]) if self.fetch_env else action[3:7]
            mocap_action[3:7] -= grip_tcp_quat
            s.mocap_pos[:] = s.xpos[tcp]                                          # reset_mocap2body_xpos
            s.mocap_quat[:] = s.xquat[tcp]
            s.mocap_pos[:] = s.mocap_pos + mocap_action[:3]
            s.mocap_quat[:] = s.mocap_quat + mocap_action[3:7]
            s.ctrl[-1] = 0.5 + np.float64(action[-1]) * 0.5
            s.step(self.frame_skip)
        else:
            s.ctrl[:] = action.astype(np.float64)  # do_simulation: ctrl[:] = action (absolute; mycobot.py:192-193)
            s.step(self.frame_skip)
        if self.block_gripper:  # mycobot.py:300-306
            s.qpos[self.flat["jnt_qposadr"][self.j_rf]] = 0.0
            s.qpos[self.flat["jnt_qposadr"][self.j_lf]] = 0.0
            s.forward()
        obs = self._get_obs()
        d = goal_distance(self.achieved_goal, self.goal)
        is_success = bool(d < self.distance_threshold)
        reward = self.compute_reward(self.achieved_goal, self.goal, {})
        terminated = is_success
        self.elapsed += 1
        truncated = is_success or self.elapsed >= self.max_episode_steps
        return obs, reward, terminated, truncated, {"is_success": is_success}

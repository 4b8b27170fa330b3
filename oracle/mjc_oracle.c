/*
 * oracle/mjc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Single-environment, scalar fp64 CPU restatement of the physics the reference
 * obtains from `mujoco.mj_step` / `mujoco.mj_forward` (call sites
 * mycobotgym/envs/mycobot.py:170,189,193,213,229,306,453,468).  The arithmetic
 * itself lives in the un-vendored dependency mujoco==2.3.2 (requirements.txt:4), which
 * is not installable in this image, so every stage below restates MuJoCo 2.3.2's
 * published algorithm from its C sources (file names cited per function) rather than
 * following in-tree reference lines.
 *
 * PARITY UNPINNED against MuJoCo itself: no mujoco wheel is available offline.  The
 * oracle is pinned by (a) the reference's own known-answer values (FK of the EEF site at
 * qpos0 = mocap.xml:3, at the keyframe = mycobot280_mocap.xml:8), (b) the reference's two
 * recorded keyframes as equilibria (tests/test_keyframe_equilibria.py: the mocap keyframe fixes
 * the weld rows -- one impedance per weld at the residual norm, translational inverse weight on
 * all six rows; the cube's recorded rest depth, 1.9e-5 m, is NOT reproduced: 0.96e-5),
 * (c) physics identities (tests/test_oracle_physics.py), see SURVEY.md Appendix B.9/C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The CUDA product never links or calls it.
 *
 * Model = the unmerged mjModel-like table produced by mycobotgym_b200/mjcf.py
 * (25 bodies, 13 joints, nv=18), passed as `omodel` (pointers into numpy arrays).
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999

#define MAXBODY 32
#define MAXJNT 24
#define MAXV 24
#define MAXQ 28
#define MAXU 8
#define MAXGEOM 8
#define MAXSITE 4
#define MAXCON 64
#define MAXEFC 448

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2 };
enum { GEOM_PLANE = 0, GEOM_BOX = 6 };
enum { EFC_EQUALITY = 0, EFC_LIMIT = 1, EFC_CONTACT = 2 };

typedef struct {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, neq, nexclude, nM;
  int iterations, ls_iterations, disable_cube, nmocap;
  double timestep, tolerance, ls_tolerance, impratio, meaninertia;
  double gravity[3];
  const int *body_parentid, *body_rootid, *body_weldid, *body_jntnum, *body_jntadr, *body_dofnum, *body_dofadr, *body_mocapid;
  const double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_subtreemass, *body_invweight0;
  const int *jnt_type, *jnt_qposadr, *jnt_dofadr, *jnt_bodyid, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_margin, *jnt_solref, *jnt_solimp;
  const int *dof_bodyid, *dof_jntid, *dof_parentid, *dof_Madr;
  const double *dof_armature, *dof_damping, *dof_invweight0;
  const int *geom_type, *geom_bodyid, *geom_condim, *geom_contype, *geom_conaffinity;
  const double *geom_pos, *geom_quat, *geom_size, *geom_friction, *geom_solref, *geom_solimp, *geom_solmix, *geom_margin, *geom_gap, *geom_rbound;
  const int *site_bodyid;
  const double *site_pos, *site_quat;
  const int *eq_type, *eq_obj1id, *eq_obj2id;
  const double *eq_data, *eq_solref, *eq_solimp;
  const int *exclude;
  const double *actuator_moment, *actuator_gain, *actuator_biasprm, *actuator_ctrlrange, *actuator_forcerange;
  const int *actuator_ctrllimited, *actuator_forcelimited;
  const double *qpos0;
  /* convex hulls of the mesh geoms (mjcf.py: body frame, identical copies folded into hull_mult) */
  int nhull, mesh_collision;
  const int *hull_bodyid, *hull_mult, *hull_vertadr, *hull_vertnum, *hull_condim;
  const double *hull_vert, *hull_center, *hull_rbound, *hull_friction, *hull_solref, *hull_solimp, *hull_solmix;
} omodel;

typedef struct {
  double dist, pos[3], frame[9], friction[5], solref[2], solimp[5], includemargin;
  int dim, geom1, geom2, efc_address;   /* geom ids >= ngeom are hulls (ngeom + hull index) */
  int mult;                             /* identical contacts MuJoCo would generate here (twin mesh geoms): scales D */
} ocontact;

typedef struct {
  /* state */
  double qpos[MAXQ], qvel[MAXV], ctrl[MAXU], qacc_warmstart[MAXV], time;
  double mocap_pos[3], mocap_quat[4];   /* one mocap body at most (robot0:mocap, mocap.xml:3) */
  /* kinematics */
  double xpos[MAXBODY * 3], xquat[MAXBODY * 4], xmat[MAXBODY * 9], xipos[MAXBODY * 3], ximat[MAXBODY * 9];
  double xanchor[MAXJNT * 3], xaxis[MAXJNT * 3];
  double geom_xpos[MAXGEOM * 3], geom_xmat[MAXGEOM * 9], site_xpos[MAXSITE * 3], site_xmat[MAXSITE * 9];
  double subtree_com[MAXBODY * 3], cinert[MAXBODY * 10], crb[MAXBODY * 10], cdof[MAXV * 6], cdof_dot[MAXV * 6];
  double cvel[MAXBODY * 6];
  double qM[MAXV * MAXV], qLD[MAXV * MAXV], qLDiagInv[MAXV], Mfull[MAXV * MAXV];
  double actuator_length[MAXU], actuator_velocity[MAXU], actuator_force[MAXU];
  double qfrc_bias[MAXV], qfrc_passive[MAXV], qfrc_actuator[MAXV], qfrc_smooth[MAXV], qacc_smooth[MAXV];
  double qfrc_constraint[MAXV], qacc[MAXV];
  /* contacts + constraints */
  int ncon, nefc, ne, nl;
  ocontact contact[MAXCON];
  int efc_type[MAXEFC], efc_id[MAXEFC];
  double efc_J[MAXEFC * MAXV], efc_pos[MAXEFC], efc_margin[MAXEFC], efc_diagApprox[MAXEFC];
  double efc_R[MAXEFC], efc_D[MAXEFC], efc_KBIP[MAXEFC * 4], efc_vel[MAXEFC], efc_aref[MAXEFC], efc_force[MAXEFC];
  /* diagnostics */
  int solver_iter, ls_total; double solver_cost;
  long flop_count;
} odata;

/* ------------------------------------------------------------------ small math (engine_util_blas / _spatial) */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double* r, const double* a, const double* b) {
  r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}
static double normalize3(double* v) {
  double n = sqrt(dot3(v, v));
  if (n < MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; } else { v[0] /= n; v[1] /= n; v[2] /= n; }
  return n;
}
static double normalize4(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n; }
  return n;
}
static void mulquat(double* r, const double* a, const double* b) {
  double t[4];
  t[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  t[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  t[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  t[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  memcpy(r, t, sizeof t);
}
static void quat2mat(double* m, const double* q) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03);
  m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
static void mulmatvec3(double* r, const double* m, const double* v) {
  double t0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double t1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double t2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
static void rotvecquat(double* r, const double* v, const double* q) {
  double m[9];
  quat2mat(m, q);
  mulmatvec3(r, m, v);
}
static void axisangle2quat(double* r, const double* axis, double angle) {
  if (angle == 0) { r[0] = 1; r[1] = r[2] = r[3] = 0; return; }
  double s = sin(angle * 0.5);
  r[0] = cos(angle * 0.5); r[1] = axis[0] * s; r[2] = axis[1] * s; r[3] = axis[2] * s;
}
/* spatial vectors are [rot(3); lin(3)] (engine_util_spatial.c) */
static void cross_motion(double* r, const double* vel, const double* v) {
  r[0] = -vel[2] * v[1] + vel[1] * v[2];
  r[1] = vel[2] * v[0] - vel[0] * v[2];
  r[2] = -vel[1] * v[0] + vel[0] * v[1];
  r[3] = -vel[2] * v[4] + vel[1] * v[5];
  r[4] = vel[2] * v[3] - vel[0] * v[5];
  r[5] = -vel[1] * v[3] + vel[0] * v[4];
  r[3] += -vel[5] * v[1] + vel[4] * v[2];
  r[4] += vel[5] * v[0] - vel[3] * v[2];
  r[5] += -vel[4] * v[0] + vel[3] * v[1];
}
static void cross_force(double* r, const double* vel, const double* f) {
  r[0] = -vel[2] * f[1] + vel[1] * f[2];
  r[1] = vel[2] * f[0] - vel[0] * f[2];
  r[2] = -vel[1] * f[0] + vel[0] * f[1];
  r[3] = -vel[2] * f[4] + vel[1] * f[5];
  r[4] = vel[2] * f[3] - vel[0] * f[5];
  r[5] = -vel[1] * f[3] + vel[0] * f[4];
  r[0] += -vel[5] * f[4] + vel[4] * f[5];
  r[1] += vel[5] * f[3] - vel[3] * f[5];
  r[2] += -vel[4] * f[3] + vel[3] * f[4];
}
static void mul_inert_vec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void inert_com(double* res, const double* inert, const double* mat, const double* dif, double mass) {
  double tmp[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) tmp[3 * r + c] = mat[3 * r + c] * inert[c];
  /* res = tmp * mat' */
  double R[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[3 * r + c] = tmp[3 * r] * mat[3 * c] + tmp[3 * r + 1] * mat[3 * c + 1] + tmp[3 * r + 2] * mat[3 * c + 2];
  res[0] = R[0]; res[1] = R[4]; res[2] = R[8]; res[3] = R[1]; res[4] = R[2]; res[5] = R[5];
  res[0] += mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] += mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] += mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] -= mass * dif[0] * dif[1];
  res[4] -= mass * dif[0] * dif[2];
  res[5] -= mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2]; res[9] = mass;
}

/* ------------------------------------------------------------------ engine_core_smooth.c: mj_kinematics */
static void local2global(const odata* d, double* xpos, double* xmat, const double* pos, const double* quat, int body) {
  double q[4];
  mulmatvec3(xpos, d->xmat + 9 * body, pos);
  for (int k = 0; k < 3; k++) xpos[k] += d->xpos[3 * body + k];
  mulquat(q, d->xquat + 4 * body, quat);
  quat2mat(xmat, q);
}
void o_kinematics(const omodel* m, odata* d) {
  memset(d->xpos, 0, 3 * sizeof(double));
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == JNT_FREE) normalize4(d->qpos + m->jnt_qposadr[j] + 3);
  for (int i = 1; i < m->nbody; i++) {
    int pid = m->body_parentid[i], ja = m->body_jntadr[i], jn = m->body_jntnum[i];
    double *xp = d->xpos + 3 * i, *xq = d->xquat + 4 * i;
    if (jn == 1 && m->jnt_type[ja] == JNT_FREE) {
      int qa = m->jnt_qposadr[ja];
      memcpy(xp, d->qpos + qa, 3 * sizeof(double));
      memcpy(xq, d->qpos + qa + 3, 4 * sizeof(double));
      memcpy(d->xanchor + 3 * ja, xp, 3 * sizeof(double));
      rotvecquat(d->xaxis + 3 * ja, m->jnt_axis + 3 * ja, xq);
    } else {
      const double* bpos = m->body_pos + 3 * i;
      const double* bquat = m->body_quat + 4 * i;
      if (m->body_mocapid[i] >= 0) { normalize4(d->mocap_quat); bpos = d->mocap_pos; bquat = d->mocap_quat; }   /* mj_kinematics: mocap pose */
      mulmatvec3(xp, d->xmat + 9 * pid, bpos);
      for (int k = 0; k < 3; k++) xp[k] += d->xpos[3 * pid + k];
      mulquat(xq, d->xquat + 4 * pid, bquat);
      for (int j = ja; j < ja + jn; j++) {
        double vec[3], qloc[4];
        int qa = m->jnt_qposadr[j];
        rotvecquat(d->xaxis + 3 * j, m->jnt_axis + 3 * j, xq);
        rotvecquat(vec, m->jnt_pos + 3 * j, xq);
        for (int k = 0; k < 3; k++) d->xanchor[3 * j + k] = vec[k] + xp[k];
        /* hinge only in this model */
        axisangle2quat(qloc, m->jnt_axis + 3 * j, d->qpos[qa] - m->qpos0[qa]);
        mulquat(xq, xq, qloc);
        rotvecquat(vec, m->jnt_pos + 3 * j, xq);
        for (int k = 0; k < 3; k++) xp[k] = d->xanchor[3 * j + k] - vec[k];
      }
    }
    normalize4(xq);
    quat2mat(d->xmat + 9 * i, xq);
  }
  for (int i = 0; i < m->nbody; i++) local2global(d, d->xipos + 3 * i, d->ximat + 9 * i, m->body_ipos + 3 * i, m->body_iquat + 4 * i, i);
  for (int g = 0; g < m->ngeom; g++) local2global(d, d->geom_xpos + 3 * g, d->geom_xmat + 9 * g, m->geom_pos + 3 * g, m->geom_quat + 4 * g, m->geom_bodyid[g]);
  for (int s = 0; s < m->nsite; s++) local2global(d, d->site_xpos + 3 * s, d->site_xmat + 9 * s, m->site_pos + 3 * s, m->site_quat + 4 * s, m->site_bodyid[s]);
}

/* engine_core_smooth.c: mj_comPos */
void o_compos(const omodel* m, odata* d) {
  int nb = m->nbody;
  memset(d->subtree_com, 0, sizeof(double) * 3 * nb);
  for (int i = nb - 1; i >= 0; i--) {
    for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] += m->body_mass[i] * d->xipos[3 * i + k];
    if (i) for (int k = 0; k < 3; k++) d->subtree_com[3 * m->body_parentid[i] + k] += d->subtree_com[3 * i + k];
  }
  for (int i = 0; i < nb; i++) {
    if (m->body_subtreemass[i] < MINVAL) memcpy(d->subtree_com + 3 * i, d->xipos + 3 * i, 3 * sizeof(double));
    else for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] /= m->body_subtreemass[i];
  }
  memset(d->cinert, 0, 10 * sizeof(double));
  for (int i = 1; i < nb; i++) {
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = d->xipos[3 * i + k] - d->subtree_com[3 * m->body_rootid[i] + k];
    inert_com(d->cinert + 10 * i, m->body_inertia + 3 * i, d->ximat + 9 * i, off, m->body_mass[i]);
  }
  for (int j = 0; j < m->njnt; j++) {
    int bi = m->jnt_bodyid[j], da = 6 * m->jnt_dofadr[j];
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = d->subtree_com[3 * m->body_rootid[bi] + k] - d->xanchor[3 * j + k];
    if (m->jnt_type[j] == JNT_FREE) {
      memset(d->cdof + da, 0, 18 * sizeof(double));
      for (int k = 0; k < 3; k++) d->cdof[da + 3 + 7 * k] = 1;
      for (int k = 0; k < 3; k++) {
        double ax[3] = {d->xmat[9 * bi + k], d->xmat[9 * bi + k + 3], d->xmat[9 * bi + k + 6]};
        double* r = d->cdof + da + 18 + 6 * k;
        r[0] = ax[0]; r[1] = ax[1]; r[2] = ax[2];
        cross3(r + 3, ax, off);
      }
    } else { /* hinge */
      double* r = d->cdof + da;
      const double* ax = d->xaxis + 3 * j;
      r[0] = ax[0]; r[1] = ax[1]; r[2] = ax[2];
      cross3(r + 3, ax, off);
    }
  }
}

/* engine_core_smooth.c: mj_crb, mj_factorM, mj_solveLD */
void o_crb(const omodel* m, odata* d) {
  int nb = m->nbody, nv = m->nv;
  memcpy(d->crb, d->cinert, sizeof(double) * 10 * nb);
  for (int i = nb - 1; i > 0; i--)
    if (m->body_parentid[i] > 0)
      for (int k = 0; k < 10; k++) d->crb[10 * m->body_parentid[i] + k] += d->crb[10 * i + k];
  memset(d->qM, 0, sizeof(double) * m->nM);
  memset(d->Mfull, 0, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    double buf[6];
    d->qM[adr] = m->dof_armature[i];
    mul_inert_vec(buf, d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[6 * j + k] * buf[k];
      d->qM[adr++] += s;
    }
    adr = m->dof_Madr[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      d->Mfull[i * nv + j] = d->Mfull[j * nv + i] = d->qM[adr++];
    }
  }
}
static void factor_ld(const omodel* m, double* qLD, double* diaginv) {
  int nv = m->nv;
  for (int k = nv - 1; k >= 0; k--) {
    int Mkk = m->dof_Madr[k], i = m->dof_parentid[k], Mki = Mkk + 1;
    while (i >= 0) {
      double tmp = qLD[Mki] / qLD[Mkk];
      int Mii = m->dof_Madr[i], cnt = 0;
      for (int j = i; j >= 0; j = m->dof_parentid[j]) cnt++;
      for (int c = 0; c < cnt; c++) qLD[Mii + c] -= tmp * qLD[Mki + c];
      qLD[Mki] = tmp;
      i = m->dof_parentid[i];
      Mki++;
    }
    diaginv[k] = 1.0 / qLD[Mkk];
  }
}
static void solve_ld(const omodel* m, double* x, const double* qLD, const double* diaginv) {
  int nv = m->nv;
  for (int i = nv - 1; i >= 0; i--) {
    if (x[i] != 0) {
      int adr = m->dof_Madr[i] + 1;
      for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= qLD[adr++] * x[i];
    }
  }
  for (int i = 0; i < nv; i++) x[i] *= diaginv[i];
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[i] -= qLD[adr++] * x[j];
  }
}
void o_factorM(const omodel* m, odata* d) {
  memcpy(d->qLD, d->qM, sizeof(double) * m->nM);
  factor_ld(m, d->qLD, d->qLDiagInv);
}
static void mulM(const omodel* m, const odata* d, double* res, const double* v) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) s += d->Mfull[i * nv + j] * v[j];
    res[i] = s;
  }
}

/* engine_support.c: mj_jac -- Jacobian of a world point attached to a body */
void o_jac(const omodel* m, const odata* d, double* jacp, double* jacr, const double* point, int body) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, sizeof(double) * 3 * nv);
  if (jacr) memset(jacr, 0, sizeof(double) * 3 * nv);
  double off[3];
  for (int k = 0; k < 3; k++) off[k] = point[k] - d->subtree_com[3 * m->body_rootid[body] + k];
  while (body && m->body_dofnum[body] == 0) body = m->body_parentid[body];
  if (!body) return;
  int i = m->body_dofadr[body] + m->body_dofnum[body] - 1;
  while (i >= 0) {
    const double* c = d->cdof + 6 * i;
    if (jacr) { jacr[i] = c[0]; jacr[nv + i] = c[1]; jacr[2 * nv + i] = c[2]; }
    if (jacp) {
      double t[3];
      cross3(t, c, off);
      jacp[i] = c[3] + t[0]; jacp[nv + i] = c[4] + t[1]; jacp[2 * nv + i] = c[5] + t[2];
    }
    i = m->dof_parentid[i];
  }
}
void o_jac_site(const omodel* m, const odata* d, double* jacp, double* jacr, int site) {
  o_jac(m, d, jacp, jacr, d->site_xpos + 3 * site, m->site_bodyid[site]);
}

/* ------------------------------------------------------------------ collision (engine_collision_driver.c / _primitive.c / _box.c) */
static void make_frame(double* f) { /* engine_util_misc? mju_makeFrame */
  normalize3(f);
  f[3] = f[4] = f[5] = 0;
  if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  double t = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= t * f[k];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

typedef struct { double dist, pos[3], normal[3]; } rawcon;

/* mjc_PlaneBox: corners below the plane, at most 4 */
static int plane_box(rawcon* out, const double* ppos, const double* pmat, const double* bpos, const double* bmat, const double* bsize, double margin) {
  double norm[3] = {pmat[2], pmat[5], pmat[8]}, dif[3];
  for (int k = 0; k < 3; k++) dif[k] = bpos[k] - ppos[k];
  double dist = dot3(dif, norm);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    double vec[3], corner[3];
    vec[0] = (i & 1 ? bsize[0] : -bsize[0]);
    vec[1] = (i & 2 ? bsize[1] : -bsize[1]);
    vec[2] = (i & 4 ? bsize[2] : -bsize[2]);
    mulmatvec3(corner, bmat, vec);
    double ldist = dot3(norm, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    double cd = dist + ldist;
    out[cnt].dist = cd;
    for (int k = 0; k < 3; k++) { out[cnt].pos[k] = corner[k] + bpos[k] - norm[k] * cd * 0.5; out[cnt].normal[k] = norm[k]; }
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

/* Box-box: separating-axis test over the 15 candidate axes, then either a face manifold
 * (incident face clipped against the reference face, up to 8 points) or one edge-edge point.
 * MuJoCo's mjc_BoxBox (engine_collision_box.c) is the algorithm restated; the degenerate
 * tie-breaking here (first face axis wins; an edge axis must beat faces by 5%) is ours and
 * is mirrored exactly by the CUDA kernel. Normal points from box 1 to box 2. */
#define CLIP_EPS 1e-12
static int clip_poly(double* px, double* py, int n, double hx, double hy) {
  /* Sutherland-Hodgman against |x|<=hx, |y|<=hy */
  double qx[16], qy[16];
  for (int side = 0; side < 4; side++) {
    int cnt = 0;
    for (int i = 0; i < n; i++) {
      int j = (i + 1) % n;
      double ax = px[i], ay = py[i], bx = px[j], by = py[j];
      double da, db;
      if (side == 0) { da = hx - ax; db = hx - bx; }
      else if (side == 1) { da = hx + ax; db = hx + bx; }
      else if (side == 2) { da = hy - ay; db = hy - by; }
      else { da = hy + ay; db = hy + by; }
      /* a vertex within CLIP_EPS of the clip line counts as inside: when an edge of the incident face coincides with the
         border of the reference face (the two finger pads share their y extent) its end points sit +-1 ulp around the
         line and a strict test would cut it at a point chosen by rounding noise -- a contact of arbitrary depth */
      if (da >= -CLIP_EPS) { qx[cnt] = ax; qy[cnt] = ay; cnt++; }
      if ((da >= -CLIP_EPS) != (db >= -CLIP_EPS)) {
        double t = da / (da - db);
        qx[cnt] = ax + t * (bx - ax); qy[cnt] = ay + t * (by - ay); cnt++;
      }
    }
    n = cnt;
    for (int i = 0; i < n; i++) { px[i] = qx[i]; py[i] = qy[i]; }
    if (n == 0) return 0;
  }
  return n;
}
static int box_box(rawcon* out, const double* p1, const double* R1, const double* s1, const double* p2, const double* R2, const double* s2, double margin) {
  double d[3], A[3][3], B[3][3];
  for (int k = 0; k < 3; k++) d[k] = p2[k] - p1[k];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) { A[i][k] = R1[3 * k + i]; B[i][k] = R2[3 * k + i]; }
  double C[3][3], Q[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { C[i][j] = dot3(A[i], B[j]); Q[i][j] = fabs(C[i][j]); }
  double best = -1e300; int code = -1; double bn[3] = {0, 0, 0};
  /* face axes of box 1, then box 2 */
  for (int i = 0; i < 3; i++) {
    double t = dot3(d, A[i]);
    double s = fabs(t) - (s1[i] + s2[0] * Q[i][0] + s2[1] * Q[i][1] + s2[2] * Q[i][2]);
    if (s > margin) return 0;
    if (s > best + (code >= 0 ? 1e-10 : 0.0)) { best = s; code = i; for (int k = 0; k < 3; k++) bn[k] = (t < 0 ? -A[i][k] : A[i][k]); }
  }
  for (int j = 0; j < 3; j++) {
    double t = dot3(d, B[j]);
    double s = fabs(t) - (s2[j] + s1[0] * Q[0][j] + s1[1] * Q[1][j] + s1[2] * Q[2][j]);
    if (s > margin) return 0;
    if (s > best + 1e-10) { best = s; code = 3 + j; for (int k = 0; k < 3; k++) bn[k] = (t < 0 ? -B[j][k] : B[j][k]); }
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double L[3];
      cross3(L, A[i], B[j]);
      double l = sqrt(dot3(L, L));
      if (l < 1e-6) continue;
      for (int k = 0; k < 3; k++) L[k] /= l;
      double t = dot3(d, L);
      double ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) { ra += s1[k] * fabs(dot3(A[k], L)); rb += s2[k] * fabs(dot3(B[k], L)); }
      double s = fabs(t) - (ra + rb);
      if (s > margin) return 0;
      if (s * 1.05 > best + 1e-10 && s > best) { best = s; code = 6 + 3 * i + j; for (int k = 0; k < 3; k++) bn[k] = (t < 0 ? -L[k] : L[k]); }
    }
  if (code < 0) return 0;
  if (code >= 6) {
    /* edge-edge: support edges */
    int i = (code - 6) / 3, j = (code - 6) % 3;
    double pa[3], pb[3];
    for (int k = 0; k < 3; k++) { pa[k] = p1[k]; pb[k] = p2[k]; }
    for (int a = 0; a < 3; a++) {
      if (a == i) continue;
      double sg = dot3(A[a], bn) > 0 ? 1.0 : -1.0;
      for (int k = 0; k < 3; k++) pa[k] += sg * s1[a] * A[a][k];
    }
    for (int b = 0; b < 3; b++) {
      if (b == j) continue;
      double sg = dot3(B[b], bn) > 0 ? -1.0 : 1.0;
      for (int k = 0; k < 3; k++) pb[k] += sg * s2[b] * B[b][k];
    }
    /* closest points of lines pa + u*A[i], pb + v*B[j] */
    double w[3];
    for (int k = 0; k < 3; k++) w[k] = pa[k] - pb[k];
    double b_ = C[i][j], dd = dot3(A[i], w), e = dot3(B[j], w);
    double den = 1 - b_ * b_;
    double u = (b_ * e - dd) / den, v = (e - b_ * dd) / den;
    for (int k = 0; k < 3; k++) {
      double ca = pa[k] + u * A[i][k], cb = pb[k] + v * B[j][k];
      out[0].pos[k] = 0.5 * (ca + cb);
      out[0].normal[k] = bn[k];
    }
    out[0].dist = best;
    return 1;
  }
  /* face contact: reference box owns the axis */
  const double *pr, *pi_, *sr, *si; double (*Ar)[3], (*Ai)[3]; double nref[3]; int ax;
  if (code < 3) { pr = p1; pi_ = p2; sr = s1; si = s2; Ar = A; Ai = B; ax = code; for (int k = 0; k < 3; k++) nref[k] = bn[k]; }
  else { pr = p2; pi_ = p1; sr = s2; si = s1; Ar = B; Ai = A; ax = code - 3; for (int k = 0; k < 3; k++) nref[k] = -bn[k]; }
  /* incident face: axis of incident box most anti-parallel to nref */
  int ia = 0; double bestd = -1;
  for (int a = 0; a < 3; a++) { double v = fabs(dot3(Ai[a], nref)); if (v > bestd) { bestd = v; ia = a; } }
  double isg = dot3(Ai[ia], nref) > 0 ? -1.0 : 1.0;
  int i1 = (ia + 1) % 3, i2 = (ia + 2) % 3;
  int r1 = (ax + 1) % 3, r2 = (ax + 2) % 3;
  double fc[3];
  for (int k = 0; k < 3; k++) fc[k] = pi_[k] + isg * si[ia] * Ai[ia][k] - pr[k];
  double px[16], py[16], vz[4][3];
  const double sgn[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
  for (int c = 0; c < 4; c++) {
    for (int k = 0; k < 3; k++) vz[c][k] = fc[k] + sgn[c][0] * si[i1] * Ai[i1][k] + sgn[c][1] * si[i2] * Ai[i2][k];
    px[c] = dot3(vz[c], Ar[r1]); py[c] = dot3(vz[c], Ar[r2]);
  }
  /* plane of the incident face in reference coordinates: depth is affine in (x,y) */
  double o_n = dot3(fc, nref);
  double u1 = dot3(Ai[i1], nref), u2 = dot3(Ai[i2], nref);
  double a11 = dot3(Ai[i1], Ar[r1]), a12 = dot3(Ai[i1], Ar[r2]), a21 = dot3(Ai[i2], Ar[r1]), a22 = dot3(Ai[i2], Ar[r2]);
  double det = a11 * a22 - a12 * a21;
  double fx = dot3(fc, Ar[r1]), fy = dot3(fc, Ar[r2]);
  int n = clip_poly(px, py, 4, sr[r1], sr[r2]);
  int cnt = 0, ncand = 0;
  double cand[16][3];
  for (int c = 0; c < n && cnt < 8; c++) {
    /* recover height above reference center along nref */
    double dx = px[c] - fx, dy = py[c] - fy, h;
    if (fabs(det) > 1e-12) {
      double al = (dx * a22 - dy * a21) / det, be = (dy * a11 - dx * a12) / det;
      h = o_n + al * u1 + be * u2;
    } else h = o_n;
    double depth = sr[ax] - h;
    if (-depth >= margin) continue;
    /* skip duplicates: a point within 1e-10 of an EARLIER PENETRATING point (accepted or not) is dropped */
    double pt[3];
    for (int k = 0; k < 3; k++) pt[k] = pr[k] + px[c] * Ar[r1][k] + py[c] * Ar[r2][k] + (h + 0.5 * depth) * nref[k];
    int dup = 0;
    for (int e = 0; e < ncand; e++) {
      double q[3] = {pt[0] - cand[e][0], pt[1] - cand[e][1], pt[2] - cand[e][2]};
      if (dot3(q, q) < 1e-20) dup = 1;
    }
    for (int k = 0; k < 3; k++) cand[ncand][k] = pt[k];
    ncand++;
    if (dup) continue;
    for (int k = 0; k < 3; k++) {
      out[cnt].pos[k] = pr[k] + px[c] * Ar[r1][k] + py[c] * Ar[r2][k] + (h + 0.5 * depth) * nref[k];
      out[cnt].normal[k] = bn[k];
    }
    out[cnt].dist = -depth;
    cnt++;
  }
  return cnt;
}

/* ------------------------------------------------------------------ convex hulls: engine_collision_convex.c (mjc_Convex, mjc_PlaneConvex)
 * MuJoCo hands hull pairs to libccd's ccdMPRPenetration (Minkowski Portal Refinement, mpr_tolerance 1e-6, 50 iterations) with
 * support mappings over the mesh vertices / the box corners and the geom centres as interior points; one contact per pair.
 * libccd is absent from /root/reference (a MuJoCo dependency): the published algorithm (mpr.c: discoverPortal, refinePortal,
 * findPenetr, findPos) is restated here.  PARITY UNPINNED like the rest of the physics. */
typedef struct { int is_box; const double *pos, *mat, *verts, *size; int n; double center[3]; } cvx;   /* pos / mat: world pose of the vertex frame */
typedef struct { double v[3], v1[3], v2[3]; } mpt;
#define CCD_EPS 2.220446049250313e-16
#define SUPPORT_TIE 1e-11
static int ccd_zero(double x) { return fabs(x) < CCD_EPS; }
static int ccd_eq(double a, double b) {
  double ab = fabs(a - b);
  if (ab < CCD_EPS) return 1;
  a = fabs(a); b = fabs(b);
  return b > a ? ab < CCD_EPS * b : ab < CCD_EPS * a;
}
static void cvx_support(const cvx* o, const double* dir, double* out) {
  double ld[3], best[3] = {0, 0, 0};
  for (int k = 0; k < 3; k++) ld[k] = o->mat[k] * dir[0] + o->mat[3 + k] * dir[1] + o->mat[6 + k] * dir[2];   /* mat' dir */
  /* box: the corner by the signs of the local direction; a component within SUPPORT_TIE of zero (a direction along a face normal:
     all four corners of the face tie) counts as positive, so that rounding noise does not pick the corner */
  if (o->is_box) { for (int k = 0; k < 3; k++) best[k] = ld[k] >= -SUPPORT_TIE ? o->size[k] : -o->size[k]; }
  else {
    /* the FIRST vertex within SUPPORT_TIE of the maximum: hull faces carry many coplanar vertices (cylinder caps, flat sides), whose
       support values tie up to rounding; with a plain arg-max the winner -- and with it the portal MPR ends on -- would be decided
       by the last bit, differently on every platform */
    double bd = -1e300;
    for (int i = 0; i < o->n; i++) {
      const double* v = o->verts + 3 * i;
      double t = v[0] * ld[0] + v[1] * ld[1] + v[2] * ld[2];
      if (t > bd) bd = t;
    }
    for (int i = 0; i < o->n; i++) {
      const double* v = o->verts + 3 * i;
      double t = v[0] * ld[0] + v[1] * ld[1] + v[2] * ld[2];
      if (t >= bd - SUPPORT_TIE) { best[0] = v[0]; best[1] = v[1]; best[2] = v[2]; break; }
    }
  }
  mulmatvec3(out, o->mat, best);
  for (int k = 0; k < 3; k++) out[k] += o->pos[k];
}
static void mpr_support(const cvx* a, const cvx* b, const double* dir, mpt* p) {   /* Minkowski difference A - B */
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  cvx_support(a, dir, p->v1);
  cvx_support(b, nd, p->v2);
  for (int k = 0; k < 3; k++) p->v[k] = p->v1[k] - p->v2[k];
}
static void portal_dir(const mpt* P, double* dir) {
  double a[3], b[3];
  for (int k = 0; k < 3; k++) { a[k] = P[2].v[k] - P[1].v[k]; b[k] = P[3].v[k] - P[1].v[k]; }
  cross3(dir, a, b);
  normalize3(dir);
}
static int portal_reach_tol(const mpt* P, const mpt* v4, const double* dir, double tol) {
  double dv4 = dot3(v4->v, dir);
  double d1 = dv4 - dot3(P[1].v, dir), d2 = dv4 - dot3(P[2].v, dir), d3 = dv4 - dot3(P[3].v, dir);
  d1 = fmin(d1, fmin(d2, d3));
  return ccd_eq(d1, tol) || d1 < tol;
}
static void expand_portal(mpt* P, const mpt* v4) {
  double v4v0[3];
  cross3(v4v0, v4->v, P[0].v);
  if (dot3(P[1].v, v4v0) > 0) { if (dot3(P[2].v, v4v0) > 0) P[1] = *v4; else P[3] = *v4; }
  else { if (dot3(P[3].v, v4v0) > 0) P[2] = *v4; else P[1] = *v4; }
}
/* squared distance of the origin to triangle (a, b, c); closest point in w (Ericson, Real-Time Collision Detection 5.1.5) */
static double origin_tri_dist2(const double* a, const double* b, const double* c, double* w) {
  double ab[3], ac[3], ap[3], bp[3], cp[3];
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = -a[k]; bp[k] = -b[k]; cp[k] = -c[k]; }
  double d1 = dot3(ab, ap), d2 = dot3(ac, ap), d3 = dot3(ab, bp), d4 = dot3(ac, bp), d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  double u, v;
  if (d1 <= 0 && d2 <= 0) { u = 0; v = 0; }
  else if (d3 >= 0 && d4 <= d3) { u = 1; v = 0; }
  else if (d1 * d4 - d3 * d2 <= 0 && d1 >= 0 && d3 <= 0) { u = d1 / (d1 - d3); v = 0; }
  else if (d6 >= 0 && d5 <= d6) { u = 0; v = 1; }
  else if (d5 * d2 - d1 * d6 <= 0 && d2 >= 0 && d6 <= 0) { u = 0; v = d2 / (d2 - d6); }
  else if (d3 * d6 - d5 * d4 <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { v = (d4 - d3) / ((d4 - d3) + (d5 - d6)); u = 1 - v; }
  else { double va = d3 * d6 - d5 * d4, vb = d5 * d2 - d1 * d6, vc = d1 * d4 - d3 * d2, den = 1 / (va + vb + vc); u = vb * den; v = vc * den; }
  for (int k = 0; k < 3; k++) w[k] = a[k] + u * ab[k] + v * ac[k];
  return dot3(w, w);
}
/* returns 1 and (depth, dir from A to B, pos) if the hulls intersect */
static int mpr_penetration(const cvx* A, const cvx* B, double* depth, double* dir, double* pos) {
  const double tol = 1e-6; const int maxiter = 50;
  mpt P[4], v4;
  double d[3], va[3], vb[3], dot;
  /* discoverPortal */
  for (int k = 0; k < 3; k++) { P[0].v1[k] = A->center[k]; P[0].v2[k] = B->center[k]; P[0].v[k] = P[0].v1[k] - P[0].v2[k]; }
  if (ccd_zero(P[0].v[0]) && ccd_zero(P[0].v[1]) && ccd_zero(P[0].v[2])) P[0].v[0] += CCD_EPS * 10;
  for (int k = 0; k < 3; k++) d[k] = -P[0].v[k];
  normalize3(d);
  mpr_support(A, B, d, &P[1]);
  dot = dot3(P[1].v, d);
  if (ccd_zero(dot) || dot < 0) return 0;
  cross3(d, P[0].v, P[1].v);
  if (ccd_zero(dot3(d, d))) {
    /* origin on the segment v0-v1 (or on v1): findPenetrSegment / findPenetrTouch */
    *depth = sqrt(dot3(P[1].v, P[1].v));
    for (int k = 0; k < 3; k++) { dir[k] = P[1].v[k]; pos[k] = 0.5 * (P[1].v1[k] + P[1].v2[k]); }
    normalize3(dir);
    return 1;
  }
  normalize3(d);
  mpr_support(A, B, d, &P[2]);
  dot = dot3(P[2].v, d);
  if (ccd_zero(dot) || dot < 0) return 0;
  for (int k = 0; k < 3; k++) { va[k] = P[1].v[k] - P[0].v[k]; vb[k] = P[2].v[k] - P[0].v[k]; }
  cross3(d, va, vb);
  normalize3(d);
  if (dot3(d, P[0].v) > 0) { mpt t = P[1]; P[1] = P[2]; P[2] = t; for (int k = 0; k < 3; k++) d[k] = -d[k]; }
  for (int guard = 0; guard < 100; guard++) {
    mpr_support(A, B, d, &P[3]);
    dot = dot3(P[3].v, d);
    if (ccd_zero(dot) || dot < 0) return 0;
    int cont = 0;
    cross3(va, P[1].v, P[3].v);
    dot = dot3(va, P[0].v);
    if (dot < 0 && !ccd_zero(dot)) { P[2] = P[3]; cont = 1; }
    if (!cont) {
      cross3(va, P[3].v, P[2].v);
      dot = dot3(va, P[0].v);
      if (dot < 0 && !ccd_zero(dot)) { P[1] = P[3]; cont = 1; }
    }
    if (!cont) break;
    for (int k = 0; k < 3; k++) { va[k] = P[1].v[k] - P[0].v[k]; vb[k] = P[2].v[k] - P[0].v[k]; }
    cross3(d, va, vb);
    normalize3(d);
    if (guard == 99) return 0;
  }
  /* refinePortal */
  for (int guard = 0;; guard++) {
    portal_dir(P, d);
    dot = dot3(d, P[1].v);
    if (ccd_zero(dot) || dot > 0) break;                      /* portalEncapsulesOrigin */
    mpr_support(A, B, d, &v4);
    dot = dot3(v4.v, d);
    if (!(ccd_zero(dot) || dot > 0) || portal_reach_tol(P, &v4, d, tol)) return 0;
    expand_portal(P, &v4);
    if (guard > 1000) return 0;
  }
  /* findPenetr */
  for (int it = 0;; it++) {
    portal_dir(P, d);
    mpr_support(A, B, d, &v4);
    if (portal_reach_tol(P, &v4, d, tol) || it > maxiter) {
      double w[3];
      *depth = sqrt(origin_tri_dist2(P[1].v, P[2].v, P[3].v, w));
      if (ccd_zero(w[0]) && ccd_zero(w[1]) && ccd_zero(w[2])) { for (int k = 0; k < 3; k++) dir[k] = d[k]; *depth = 0; }
      else { for (int k = 0; k < 3; k++) dir[k] = w[k]; normalize3(dir); }
      /* findPos: barycentric coordinates of the origin in the portal tetrahedron */
      double b[4], t[3], sum;
      cross3(t, P[1].v, P[2].v); b[0] = dot3(t, P[3].v);
      cross3(t, P[3].v, P[2].v); b[1] = dot3(t, P[0].v);
      cross3(t, P[0].v, P[1].v); b[2] = dot3(t, P[3].v);
      cross3(t, P[2].v, P[1].v); b[3] = dot3(t, P[0].v);
      sum = b[0] + b[1] + b[2] + b[3];
      if (ccd_zero(sum) || sum < 0) {
        b[0] = 0;
        cross3(t, P[2].v, P[3].v); b[1] = dot3(t, d);
        cross3(t, P[3].v, P[1].v); b[2] = dot3(t, d);
        cross3(t, P[1].v, P[2].v); b[3] = dot3(t, d);
        sum = b[1] + b[2] + b[3];
      }
      double inv = 1.0 / sum;
      for (int k = 0; k < 3; k++) {
        double p1 = b[0] * P[0].v1[k] + b[1] * P[1].v1[k] + b[2] * P[2].v1[k] + b[3] * P[3].v1[k];
        double p2 = b[0] * P[0].v2[k] + b[1] * P[1].v2[k] + b[2] * P[2].v2[k] + b[3] * P[3].v2[k];
        pos[k] = 0.5 * (p1 + p2) * inv;
      }
      return 1;
    }
    expand_portal(P, &v4);
  }
}

static int body_filter(const omodel* m, int b1, int b2) {
  int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
  if (w1 == w2) return 1;
  int lo = b1 < b2 ? b1 : b2, hi = b1 < b2 ? b2 : b1;
  for (int e = 0; e < m->nexclude; e++)
    if (m->exclude[2 * e] == lo && m->exclude[2 * e + 1] == hi) return 1;
  if (w1 && w2 && (m->body_weldid[m->body_parentid[w1]] == w2 || m->body_weldid[m->body_parentid[w2]] == w1)) return 1;
  return 0;
}

static int con_body(const omodel* m, int g) { return g < m->ngeom ? m->geom_bodyid[g] : m->hull_bodyid[g - m->ngeom]; }
/* mj_contactParam (same priority): condim max, friction max, solref mixed (or min for direct stiffness), solimp mixed */
static void mix_params(ocontact* con, int cd1, const double* fr1, const double* ra, const double* ia, double sa,
                       int cd2, const double* fr2, const double* rb, const double* ib, double sb) {
  con->dim = cd1 > cd2 ? cd1 : cd2;
  double fr[3];
  for (int k = 0; k < 3; k++) fr[k] = fmax(fr1[k], fr2[k]);
  con->friction[0] = fr[0]; con->friction[1] = fr[0]; con->friction[2] = fr[1]; con->friction[3] = fr[2]; con->friction[4] = fr[2];
  double mix;
  if (sa >= MINVAL && sb >= MINVAL) mix = sa / (sa + sb);
  else if (sa < MINVAL && sb < MINVAL) mix = 0.5;
  else if (sa < MINVAL) mix = 0.0; else mix = 1.0;
  if (ra[0] > 0 && rb[0] > 0) for (int k = 0; k < 2; k++) con->solref[k] = mix * ra[k] + (1 - mix) * rb[k];
  else for (int k = 0; k < 2; k++) con->solref[k] = fmin(ra[k], rb[k]);
  for (int k = 0; k < 5; k++) con->solimp[k] = mix * ia[k] + (1 - mix) * ib[k];
}
static void hull_cvx(const omodel* m, const odata* d, int h, cvx* o) {
  int b = m->hull_bodyid[h];
  o->is_box = 0; o->pos = d->xpos + 3 * b; o->mat = d->xmat + 9 * b; o->verts = m->hull_vert + 3 * m->hull_vertadr[h]; o->n = m->hull_vertnum[h]; o->size = NULL;
  mulmatvec3(o->center, o->mat, m->hull_center + 3 * h);
  for (int k = 0; k < 3; k++) o->center[k] += o->pos[k];
}
static void add_convex_contact(const omodel* m, odata* d, int g1, int g2, double dist, const double* pos, const double* nrm, int mult) {
  if (d->ncon >= MAXCON) return;
  ocontact* con = d->contact + d->ncon++;
  con->dist = dist;
  memcpy(con->pos, pos, sizeof con->pos);
  memcpy(con->frame, nrm, 3 * sizeof(double));
  make_frame(con->frame);
  con->geom1 = g1; con->geom2 = g2; con->mult = mult; con->includemargin = 0; con->efc_address = -1;
  const int h1 = g1 - m->ngeom, h2 = g2 - m->ngeom;
  const int cd1 = h1 >= 0 ? m->hull_condim[h1] : m->geom_condim[g1], cd2 = h2 >= 0 ? m->hull_condim[h2] : m->geom_condim[g2];
  const double* f1 = h1 >= 0 ? m->hull_friction + 3 * h1 : m->geom_friction + 3 * g1; const double* f2 = h2 >= 0 ? m->hull_friction + 3 * h2 : m->geom_friction + 3 * g2;
  const double* r1 = h1 >= 0 ? m->hull_solref + 2 * h1 : m->geom_solref + 2 * g1; const double* r2 = h2 >= 0 ? m->hull_solref + 2 * h2 : m->geom_solref + 2 * g2;
  const double* i1 = h1 >= 0 ? m->hull_solimp + 5 * h1 : m->geom_solimp + 5 * g1; const double* i2 = h2 >= 0 ? m->hull_solimp + 5 * h2 : m->geom_solimp + 5 * g2;
  mix_params(con, cd1, f1, r1, i1, h1 >= 0 ? m->hull_solmix[h1] : m->geom_solmix[g1], cd2, f2, r2, i2, h2 >= 0 ? m->hull_solmix[h2] : m->geom_solmix[g2]);
}
/* hull x {plane, box, hull}: mjc_PlaneConvex / mjc_Convex, one contact per pair, after the primitive pairs */
static void collide_hulls(const omodel* m, odata* d) {
  for (int h = 0; h < m->nhull; h++) {
    cvx A; hull_cvx(m, d, h, &A);
    const int bh = m->hull_bodyid[h];
    const double rh = m->hull_rbound[h];
    for (int g = 0; g < m->ngeom; g++) {                      /* primitive first: geom type order plane < box < mesh */
      int bg = m->geom_bodyid[g];
      if (body_filter(m, bg, bh)) continue;
      if (m->disable_cube && m->body_dofnum[bg] == 6) continue;
      const double *pg = d->geom_xpos + 3 * g, *Rg = d->geom_xmat + 9 * g;
      if (m->geom_type[g] == GEOM_PLANE) {
        double nrm[3] = {Rg[2], Rg[5], Rg[8]}, nn[3] = {-Rg[2], -Rg[5], -Rg[8]}, dif[3], p[3];
        for (int k = 0; k < 3; k++) dif[k] = A.center[k] - pg[k];
        if (dot3(dif, nrm) > rh) continue;
        cvx_support(&A, nn, p);
        for (int k = 0; k < 3; k++) dif[k] = p[k] - pg[k];
        double dist = dot3(dif, nrm);
        if (dist > 0) continue;
        for (int k = 0; k < 3; k++) p[k] -= 0.5 * dist * nrm[k];
        add_convex_contact(m, d, g, m->ngeom + h, dist, p, nrm, m->hull_mult[h]);
      } else {
        double dif[3], bound = rh + m->geom_rbound[g];
        for (int k = 0; k < 3; k++) dif[k] = A.center[k] - pg[k];
        if (dot3(dif, dif) > bound * bound) continue;
        cvx B; B.is_box = 1; B.pos = pg; B.mat = Rg; B.size = m->geom_size + 3 * g; B.verts = NULL; B.n = 0;
        for (int k = 0; k < 3; k++) B.center[k] = pg[k];
        double depth, dir[3], pos[3];
        if (mpr_penetration(&B, &A, &depth, dir, pos)) add_convex_contact(m, d, g, m->ngeom + h, -depth, pos, dir, m->hull_mult[h]);
      }
    }
    for (int h2 = h + 1; h2 < m->nhull; h2++) {
      if (body_filter(m, bh, m->hull_bodyid[h2])) continue;
      cvx B; hull_cvx(m, d, h2, &B);
      double dif[3], bound = rh + m->hull_rbound[h2];
      for (int k = 0; k < 3; k++) dif[k] = A.center[k] - B.center[k];
      if (dot3(dif, dif) > bound * bound) continue;
      double depth, dir[3], pos[3];
      if (mpr_penetration(&A, &B, &depth, dir, pos)) add_convex_contact(m, d, m->ngeom + h, m->ngeom + h2, -depth, pos, dir, m->hull_mult[h] * m->hull_mult[h2]);
    }
  }
}

void o_collision(const omodel* m, odata* d) {
  d->ncon = 0;
  for (int g1 = 0; g1 < m->ngeom; g1++)
    for (int g2 = g1 + 1; g2 < m->ngeom; g2++) {
      int a = g1, b = g2;
      if (m->geom_type[a] > m->geom_type[b]) { a = g2; b = g1; }
      int b1 = m->geom_bodyid[a], b2 = m->geom_bodyid[b];
      if (body_filter(m, b1, b2)) continue;
      if (!((m->geom_contype[a] & m->geom_conaffinity[b]) || (m->geom_contype[b] & m->geom_conaffinity[a]))) continue;
      if (m->disable_cube && (m->body_dofnum[b1] == 6 || m->body_dofnum[b2] == 6)) continue;
      double margin = fmax(m->geom_margin[a], m->geom_margin[b]);
      double gap = fmax(m->geom_gap[a], m->geom_gap[b]);
      const double *pa = d->geom_xpos + 3 * a, *pb = d->geom_xpos + 3 * b, *Ra = d->geom_xmat + 9 * a, *Rb = d->geom_xmat + 9 * b;
      rawcon rc[8]; int n = 0;
      if (m->geom_type[a] == GEOM_PLANE && m->geom_type[b] == GEOM_BOX) {
        double nrm[3] = {Ra[2], Ra[5], Ra[8]}, dif[3];
        for (int k = 0; k < 3; k++) dif[k] = pb[k] - pa[k];
        if (dot3(dif, nrm) > margin + m->geom_rbound[b]) continue;
        n = plane_box(rc, pa, Ra, pb, Rb, m->geom_size + 3 * b, margin);
      } else if (m->geom_type[a] == GEOM_BOX && m->geom_type[b] == GEOM_BOX) {
        double dif[3];
        for (int k = 0; k < 3; k++) dif[k] = pb[k] - pa[k];
        double bound = margin + m->geom_rbound[a] + m->geom_rbound[b];
        if (dot3(dif, dif) > bound * bound) continue;
        n = box_box(rc, pa, Ra, m->geom_size + 3 * a, pb, Rb, m->geom_size + 3 * b, margin);
      } else continue;
      for (int c = 0; c < n && d->ncon < MAXCON; c++) {
        ocontact* con = d->contact + d->ncon++;
        con->dist = rc[c].dist;
        memcpy(con->pos, rc[c].pos, sizeof con->pos);
        memcpy(con->frame, rc[c].normal, 3 * sizeof(double));
        make_frame(con->frame);
        con->geom1 = a; con->geom2 = b;
        con->mult = 1;
        mix_params(con, m->geom_condim[a], m->geom_friction + 3 * a, m->geom_solref + 2 * a, m->geom_solimp + 5 * a, m->geom_solmix[a],
                   m->geom_condim[b], m->geom_friction + 3 * b, m->geom_solref + 2 * b, m->geom_solimp + 5 * b, m->geom_solmix[b]);
        con->includemargin = margin - gap;
        con->efc_address = -1;
      }
    }
  if (m->mesh_collision && m->nhull > 0) collide_hulls(m, d);
}

/* ------------------------------------------------------------------ engine_core_constraint.c */
static int add_row(odata* d, int nv, const double* jac, double pos, double margin, int type, int id) {
  int r = d->nefc;
  if (r >= MAXEFC) return -1;
  memcpy(d->efc_J + r * nv, jac, sizeof(double) * nv);
  d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_type[r] = type; d->efc_id[r] = id;
  d->nefc++;
  return r;
}
static void get_impedance(const double* solimp_in, double pos, double margin, double* imp) {
  double s[5];
  s[0] = fmin(MAXIMP, fmax(MINIMP, solimp_in[0]));
  s[1] = fmin(MAXIMP, fmax(MINIMP, solimp_in[1]));
  s[2] = fmax(0, solimp_in[2]);
  s[3] = fmin(MAXIMP, fmax(MINIMP, solimp_in[3]));
  s[4] = fmax(1, solimp_in[4]);
  if (s[0] == s[1] || s[2] <= MINVAL) { *imp = 0.5 * (s[0] + s[1]); return; }
  double x = (pos - margin) / s[2];
  if (x < 0) x = -x;
  if (x >= 1 || x <= 0) { *imp = (x >= 1 ? s[1] : s[0]); return; }
  double y;
  if (s[4] == 1) y = x;
  else if (x <= s[3]) { double a = 1 / pow(s[3], s[4] - 1); y = a * pow(x, s[4]); }
  else { double b = 1 / pow(1 - s[3], s[4] - 1); y = 1 - b * pow(1 - x, s[4]); }
  *imp = s[0] + y * (s[1] - s[0]);
}

void o_make_constraint(const omodel* m, odata* d) {
  int nv = m->nv;
  double jac[MAXV], jp1[3 * MAXV], jp2[3 * MAXV], jr1[3 * MAXV], jr2[3 * MAXV];
  d->nefc = 0;
  /* equality (mj_instantiateEquality) */
  for (int e = 0; e < m->neq; e++) {
    const double* data = m->eq_data + 11 * e;
    int o1 = m->eq_obj1id[e], o2 = m->eq_obj2id[e];
    if (m->eq_type[e] == EQ_CONNECT) {
      double p1[3], p2[3];
      mulmatvec3(p1, d->xmat + 9 * o1, data);
      mulmatvec3(p2, d->xmat + 9 * o2, data + 3);
      for (int k = 0; k < 3; k++) { p1[k] += d->xpos[3 * o1 + k]; p2[k] += d->xpos[3 * o2 + k]; }
      o_jac(m, d, jp1, NULL, p1, o1);
      o_jac(m, d, jp2, NULL, p2, o2);
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < nv; c++) jac[c] = jp1[r * nv + c] - jp2[r * nv + c];
        int row = add_row(d, nv, jac, p1[r] - p2[r], 0, EFC_EQUALITY, e);
        d->efc_diagApprox[row] = m->body_invweight0[2 * o1] + m->body_invweight0[2 * o2];
      }
    } else if (m->eq_type[e] == EQ_WELD) {
      /* mj_instantiateEquality, mjEQ_WELD (2.3.x): data = anchor in body2 (3), anchor in body1 (3), relpose quat (4), torquescale */
      double p1[3], p2[3], quat[4], quat1[4], quat2[4], cpos[6];
      const double torquescale = data[10];
      mulmatvec3(p1, d->xmat + 9 * o1, data + 3);
      mulmatvec3(p2, d->xmat + 9 * o2, data);
      for (int k = 0; k < 3; k++) { p1[k] += d->xpos[3 * o1 + k]; p2[k] += d->xpos[3 * o2 + k]; cpos[k] = p1[k] - p2[k]; }
      o_jac(m, d, jp1, jr1, p1, o1);
      o_jac(m, d, jp2, jr2, p2, o2);
      mulquat(quat, d->xquat + 4 * o1, data + 6);                        /* q1 * relpose */
      quat1[0] = d->xquat[4 * o2]; quat1[1] = -d->xquat[4 * o2 + 1]; quat1[2] = -d->xquat[4 * o2 + 2]; quat1[3] = -d->xquat[4 * o2 + 3];
      mulquat(quat2, quat1, quat);                                       /* neg(q2) * q1 * relpose */
      for (int k = 0; k < 3; k++) cpos[3 + k] = torquescale * quat2[1 + k];
      double tran = m->body_invweight0[2 * o1] + m->body_invweight0[2 * o2];
      double rot = m->body_invweight0[2 * o1 + 1] + m->body_invweight0[2 * o2 + 1];
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < nv; c++) jac[c] = jp1[r * nv + c] - jp2[r * nv + c];
        int row = add_row(d, nv, jac, cpos[r], 0, EFC_EQUALITY, e);
        d->efc_diagApprox[row] = tran;
      }
      /* rotational rows: 0.5 * neg(q2) * (jacr1 - jacr2)_col * q1 * relpose, axis part, times torquescale */
      double Jrot[3 * MAXV];
      for (int c = 0; c < nv; c++) {
        double ax[3] = {jr1[c] - jr2[c], jr1[nv + c] - jr2[nv + c], jr1[2 * nv + c] - jr2[2 * nv + c]};
        double qa[4] = {-quat1[1] * ax[0] - quat1[2] * ax[1] - quat1[3] * ax[2],
                        quat1[0] * ax[0] + quat1[2] * ax[2] - quat1[3] * ax[1],
                        quat1[0] * ax[1] + quat1[3] * ax[0] - quat1[1] * ax[2],
                        quat1[0] * ax[2] + quat1[1] * ax[1] - quat1[2] * ax[0]};   /* mju_mulQuatAxis(neg(q2), axis) */
        double q3[4];
        mulquat(q3, qa, quat);
        for (int r = 0; r < 3; r++) Jrot[r * nv + c] = 0.5 * q3[1 + r] * torquescale;
      }
      /* mj_diagApprox: in 2.3.2 connect and weld share one branch (body translation for every row); the
         `weldcnt > 2` split into translational / rotational inverse weight is a later fix.  The choice is PINNED by the
         reference's mocap keyframe (mycobot280_mocap.xml:7-9, a recorded equilibrium at a near-singular elbow pose):
         with `rot` here the arm's residual acceleration at that state is 6.5 rad/s^2 and the weld offset settles at
         0.42 mm instead of the recorded 1.149 mm; with `tran` (and getposdim's norm below) 0.18 rad/s^2 / 1.147 mm
         (tests/test_keyframe_equilibria.py). */
      (void)rot;
      for (int r = 0; r < 3; r++) {
        int row = add_row(d, nv, Jrot + r * nv, cpos[3 + r], 0, EFC_EQUALITY, e);
        d->efc_diagApprox[row] = tran;
      }
    } else if (m->eq_type[e] == EQ_JOINT) {
      int q1 = m->jnt_qposadr[o1], q2 = m->jnt_qposadr[o2];
      double pos1 = d->qpos[q1] - m->qpos0[q1], dif = d->qpos[q2] - m->qpos0[q2];
      double cpos = pos1 - data[0] - data[1] * dif - data[2] * dif * dif - data[3] * dif * dif * dif - data[4] * dif * dif * dif * dif;
      double deriv = data[1] + 2 * data[2] * dif + 3 * data[3] * dif * dif + 4 * data[4] * dif * dif * dif;
      memset(jac, 0, sizeof(double) * nv);
      jac[m->jnt_dofadr[o1]] = 1;
      jac[m->jnt_dofadr[o2]] = -deriv;
      int row = add_row(d, nv, jac, cpos, 0, EFC_EQUALITY, e);
      d->efc_diagApprox[row] = m->dof_invweight0[m->jnt_dofadr[o1]] + m->dof_invweight0[m->jnt_dofadr[o2]];
    }
  }
  d->ne = d->nefc;
  /* limits (mj_instantiateLimit) */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] != JNT_HINGE) continue;
    double value = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - value);
      if (dist < margin) {
        memset(jac, 0, sizeof(double) * nv);
        jac[m->jnt_dofadr[j]] = -side;
        int row = add_row(d, nv, jac, dist, margin, EFC_LIMIT, j);
        d->efc_diagApprox[row] = m->dof_invweight0[m->jnt_dofadr[j]];
      }
    }
  }
  d->nl = d->nefc - d->ne;
  /* contacts, pyramidal (mj_instantiateContact) */
  for (int c = 0; c < d->ncon; c++) {
    ocontact* con = d->contact + c;
    if (con->dist >= con->includemargin) continue;
    int b1 = con_body(m, con->geom1), b2 = con_body(m, con->geom2);
    int dim = con->dim;
    o_jac(m, d, jp1, jr1, con->pos, b1);
    o_jac(m, d, jp2, jr2, con->pos, b2);
    double J[6 * MAXV];
    for (int r = 0; r < 3; r++)
      for (int col = 0; col < nv; col++) {
        double s = 0, sr = 0;
        for (int k = 0; k < 3; k++) {
          s += con->frame[3 * r + k] * (jp2[k * nv + col] - jp1[k * nv + col]);
          sr += con->frame[3 * r + k] * (jr2[k * nv + col] - jr1[k * nv + col]);
        }
        J[r * nv + col] = s; J[(3 + r) * nv + col] = sr;
      }
    double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    double rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    con->efc_address = d->nefc;
    if (dim == 1) {
      int row = add_row(d, nv, J, con->dist, con->includemargin, EFC_CONTACT, c);
      d->efc_diagApprox[row] = tran;
      continue;
    }
    for (int k = 1; k < dim; k++) {
      double mu = con->friction[k - 1];
      for (int sgn = 0; sgn < 2; sgn++) {
        for (int col = 0; col < nv; col++) jac[col] = J[col] + (sgn ? -mu : mu) * J[k * nv + col];
        int row = add_row(d, nv, jac, con->dist, con->includemargin, EFC_CONTACT, c);
        if (row < 0) return;
        int jrow = 2 * (k - 1) + sgn;
        double f = con->friction[jrow / 2];
        d->efc_diagApprox[row] = tran + f * f * (jrow < 4 ? tran : rot);
      }
    }
  }
  /* mj_makeImpedance */
  for (int i = 0; i < d->nefc; i++) {
    double solref[2]; const double* solimp;
    int id = d->efc_id[i];
    if (d->efc_type[i] == EFC_EQUALITY) { solref[0] = m->eq_solref[2 * id]; solref[1] = m->eq_solref[2 * id + 1]; solimp = m->eq_solimp + 5 * id; }
    else if (d->efc_type[i] == EFC_LIMIT) { solref[0] = m->jnt_solref[2 * id]; solref[1] = m->jnt_solref[2 * id + 1]; solimp = m->jnt_solimp + 5 * id; }
    else { solref[0] = d->contact[id].solref[0]; solref[1] = d->contact[id].solref[1]; solimp = d->contact[id].solimp; }
    if (solref[0] > 0) solref[0] = fmax(solref[0], 2 * m->timestep); /* refsafe */
    double imp, ipos = d->efc_pos[i];
    /* getposdim(): connect (3 rows) and weld (6 rows) share ONE impedance, evaluated at the norm of the whole residual
       (that is why torquescale is "notionally in units of length") */
    if (d->efc_type[i] == EFC_EQUALITY && m->eq_type[id] != EQ_JOINT) {
      int n = m->eq_type[id] == EQ_WELD ? 6 : 3, i0 = i;
      while (i0 > 0 && d->efc_type[i0 - 1] == EFC_EQUALITY && d->efc_id[i0 - 1] == id) i0--;
      double s2 = 0;
      for (int k = i0; k < i0 + n; k++) s2 += d->efc_pos[k] * d->efc_pos[k];
      ipos = sqrt(s2);
    }
    get_impedance(solimp, ipos, d->efc_margin[i], &imp);
    d->efc_R[i] = fmax(MINVAL, (1 - imp) * d->efc_diagApprox[i] / imp);
    double dmax = fmin(MAXIMP, fmax(MINIMP, solimp[1]));
    double K, B;
    if (solref[0] > 0) {
      K = 1 / fmax(MINVAL, dmax * dmax * solref[0] * solref[0] * solref[1] * solref[1]);
      B = 2 / fmax(MINVAL, dmax * solref[0]);
    } else {
      K = -solref[0] / fmax(MINVAL, dmax * dmax);
      B = -solref[1] / fmax(MINVAL, dmax);
    }
    d->efc_KBIP[4 * i] = K; d->efc_KBIP[4 * i + 1] = B; d->efc_KBIP[4 * i + 2] = imp; d->efc_KBIP[4 * i + 3] = 0;
  }
  /* pyramidal regulariser scaling */
  for (int i = 0; i < d->nefc; i++) {
    if (d->efc_type[i] != EFC_CONTACT) continue;
    ocontact* con = d->contact + d->efc_id[i];
    if (con->dim > 1) {
      double mu = con->friction[0] / sqrt(m->impratio);
      double Rpy = 2 * mu * mu * d->efc_R[i];
      int n = 2 * (con->dim - 1);
      Rpy /= (double)con->mult;      /* `mult` identical contacts (twin mesh geoms) == one contact with mult x the D of each */
      for (int j = 0; j < n; j++) d->efc_R[i + j] = Rpy;
      i += n - 1;
    }
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1 / d->efc_R[i];
}

/* ------------------------------------------------------------------ velocity stage (engine_core_smooth.c mj_comVel, mj_rne; engine_passive.c) */
void o_fwd_velocity(const omodel* m, odata* d) {
  int nb = m->nbody, nv = m->nv;
  memset(d->cvel, 0, 6 * sizeof(double));
  for (int i = 1; i < nb; i++) {
    double cv[6];
    memcpy(cv, d->cvel + 6 * m->body_parentid[i], sizeof cv);
    int bda = m->body_dofadr[i], dn = m->body_dofnum[i];
    int j = 0;
    while (j < dn) {
      int jt = m->jnt_type[m->dof_jntid[bda + j]];
      if (jt == JNT_FREE) {
        memset(d->cdof_dot + 6 * bda, 0, 18 * sizeof(double));
        for (int k = 0; k < 3; k++)
          for (int c = 0; c < 6; c++) cv[c] += d->cdof[6 * (bda + k) + c] * d->qvel[bda + k];
        for (int k = 3; k < 6; k++) cross_motion(d->cdof_dot + 6 * (bda + k), cv, d->cdof + 6 * (bda + k));
        for (int k = 3; k < 6; k++)
          for (int c = 0; c < 6; c++) cv[c] += d->cdof[6 * (bda + k) + c] * d->qvel[bda + k];
        j += 6;
      } else {
        cross_motion(d->cdof_dot + 6 * (bda + j), cv, d->cdof + 6 * (bda + j));
        for (int c = 0; c < 6; c++) cv[c] += d->cdof[6 * (bda + j) + c] * d->qvel[bda + j];
        j++;
      }
    }
    memcpy(d->cvel + 6 * i, cv, sizeof cv);
  }
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  for (int r = 0; r < d->nefc; r++) {
    double s = 0;
    for (int c = 0; c < nv; c++) s += d->efc_J[r * nv + c] * d->qvel[c];
    d->efc_vel[r] = s;
    d->efc_aref[r] = -d->efc_KBIP[4 * r + 1] * s - d->efc_KBIP[4 * r] * d->efc_KBIP[4 * r + 2] * (d->efc_pos[r] - d->efc_margin[r]);
  }
  /* mj_rne with flg_acc = 0 */
  double cacc[MAXBODY * 6], cfrc[MAXBODY * 6];
  memset(cacc, 0, sizeof cacc);
  for (int k = 0; k < 3; k++) cacc[3 + k] = -m->gravity[k];
  memset(cfrc, 0, 6 * sizeof(double));
  for (int i = 1; i < nb; i++) {
    int bda = m->body_dofadr[i], dn = m->body_dofnum[i];
    double tmp[6], tmp1[6];
    memcpy(cacc + 6 * i, cacc + 6 * m->body_parentid[i], 6 * sizeof(double));
    for (int j = 0; j < dn; j++)
      for (int c = 0; c < 6; c++) cacc[6 * i + c] += d->cdof_dot[6 * (bda + j) + c] * d->qvel[bda + j];
    mul_inert_vec(cfrc + 6 * i, d->cinert + 10 * i, cacc + 6 * i);
    mul_inert_vec(tmp, d->cinert + 10 * i, d->cvel + 6 * i);
    cross_force(tmp1, d->cvel + 6 * i, tmp);
    for (int c = 0; c < 6; c++) cfrc[6 * i + c] += tmp1[c];
  }
  for (int i = nb - 1; i > 0; i--)
    if (m->body_parentid[i])
      for (int c = 0; c < 6; c++) cfrc[6 * m->body_parentid[i] + c] += cfrc[6 * i + c];
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int c = 0; c < 6; c++) s += d->cdof[6 * i + c] * cfrc[6 * m->dof_bodyid[i] + c];
    d->qfrc_bias[i] = s;
  }
}

/* engine_forward.c mj_fwdActuation: general actuators, dyntype none, fixed gain, affine bias */
void o_fwd_actuation(const omodel* m, odata* d) {
  int nv = m->nv;
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  for (int a = 0; a < m->nu; a++) {
    const double* mom = m->actuator_moment + a * nv;
    double len = 0, vel = 0;
    for (int i = 0; i < nv; i++) {
      if (mom[i] == 0) continue;
      int q = m->jnt_qposadr[m->dof_jntid[i]];
      len += mom[i] * d->qpos[q];
      vel += mom[i] * d->qvel[i];
    }
    d->actuator_length[a] = len; d->actuator_velocity[a] = vel;
    double ctrl = d->ctrl[a];
    if (m->actuator_ctrllimited[a]) ctrl = fmax(m->actuator_ctrlrange[2 * a], fmin(m->actuator_ctrlrange[2 * a + 1], ctrl));
    const double* bp = m->actuator_biasprm + 3 * a;
    double f = m->actuator_gain[a] * ctrl + bp[0] + bp[1] * len + bp[2] * vel;
    if (m->actuator_forcelimited[a]) f = fmax(m->actuator_forcerange[2 * a], fmin(m->actuator_forcerange[2 * a + 1], f));
    d->actuator_force[a] = f;
    for (int i = 0; i < nv; i++) d->qfrc_actuator[i] += mom[i] * f;
  }
}

void o_fwd_acceleration(const omodel* m, odata* d) {
  for (int i = 0; i < m->nv; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  solve_ld(m, d->qacc_smooth, d->qLD, d->qLDiagInv);
}

/* ------------------------------------------------------------------ engine_solver.c: Newton, pyramidal cones */
typedef struct {
  const omodel* m; odata* d;
  int nv, nefc;
  double Ma[MAXV], Jaref[MAXEFC], grad[MAXV], Mgrad[MAXV], search[MAXV], Mv[MAXV], Jv[MAXEFC];
  double quad[MAXEFC * 3], quadGauss[3];
  int active[MAXEFC];
  double cost, gauss, scale;
  double H[MAXV * MAXV];
} nctx;

static double constraint_update(nctx* c, const double* jar, int store) {
  odata* d = c->d; int nv = c->nv;
  double cost = 0;
  for (int i = 0; i < c->nefc; i++) {
    int act = (d->efc_type[i] == EFC_EQUALITY) || jar[i] < 0;
    if (store) { c->active[i] = act; d->efc_force[i] = act ? -d->efc_D[i] * jar[i] : 0; }
    if (act) cost += 0.5 * d->efc_D[i] * jar[i] * jar[i];
  }
  if (store) {
    for (int k = 0; k < nv; k++) d->qfrc_constraint[k] = 0;
    for (int i = 0; i < c->nefc; i++) {
      double f = d->efc_force[i];
      if (f == 0) continue;
      for (int k = 0; k < nv; k++) d->qfrc_constraint[k] += d->efc_J[i * nv + k] * f;
    }
  }
  return cost;
}
static double gauss_cost(nctx* c, const double* Ma, const double* qacc) {
  double g = 0;
  for (int i = 0; i < c->nv; i++) g += 0.5 * (Ma[i] - c->d->qfrc_smooth[i]) * (qacc[i] - c->d->qacc_smooth[i]);
  return g;
}
static void chol_factor(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double t = A[j * n + j];
    for (int k = 0; k < j; k++) t -= A[j * n + k] * A[j * n + k];
    if (t < MINVAL) t = MINVAL;
    A[j * n + j] = sqrt(t);
    double inv = 1 / A[j * n + j];
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s * inv;
    }
  }
}
static void chol_solve(const double* L, double* x, const double* b, int n) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}
static void update_all(nctx* c) {
  odata* d = c->d; int nv = c->nv;
  double cc = constraint_update(c, c->Jaref, 1);
  c->gauss = gauss_cost(c, c->Ma, d->qacc);
  c->cost = c->gauss + cc;
  for (int i = 0; i < nv; i++) c->grad[i] = c->Ma[i] - d->qfrc_smooth[i] - d->qfrc_constraint[i];
  for (int i = 0; i < nv * nv; i++) c->H[i] = d->Mfull[i];
  for (int r = 0; r < c->nefc; r++) {
    if (!c->active[r]) continue;
    const double* J = d->efc_J + r * nv; double D = d->efc_D[r];
    for (int i = 0; i < nv; i++) {
      if (J[i] == 0) continue;
      double t = D * J[i];
      for (int j = 0; j <= i; j++) c->H[i * nv + j] += t * J[j];
    }
  }
  chol_factor(c->H, nv);
  chol_solve(c->H, c->Mgrad, c->grad, nv);
}
typedef struct { double alpha, cost, d0, d1; } lspt;
static void ls_eval(nctx* c, lspt* p) {
  double a = p->alpha, q0 = c->quadGauss[0], q1 = c->quadGauss[1], q2 = c->quadGauss[2];
  odata* d = c->d;
  for (int i = 0; i < c->nefc; i++) {
    if (d->efc_type[i] == EFC_EQUALITY || c->Jaref[i] + a * c->Jv[i] < 0) {
      q0 += c->quad[3 * i]; q1 += c->quad[3 * i + 1]; q2 += c->quad[3 * i + 2];
    }
  }
  p->cost = a * a * q2 + a * q1 + q0;
  p->d0 = 2 * a * q2 + q1;
  p->d1 = 2 * q2;
  if (p->d1 <= 0) p->d1 = MINVAL;
  d->ls_total++;
}
static double line_search(nctx* c) {
  const omodel* m = c->m; odata* d = c->d; int nv = c->nv;
  double snorm = 0;
  for (int i = 0; i < nv; i++) snorm += c->search[i] * c->search[i];
  snorm = sqrt(snorm);
  if (snorm < MINVAL) return 0;
  double gtol = m->tolerance * m->ls_tolerance * snorm / c->scale;
  mulM(m, d, c->Mv, c->search);
  for (int r = 0; r < c->nefc; r++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[r * nv + k] * c->search[k];
    c->Jv[r] = s;
  }
  double g1 = 0, g2 = 0;
  for (int i = 0; i < nv; i++) { g1 += c->search[i] * (c->Ma[i] - d->qfrc_smooth[i]); g2 += c->search[i] * c->Mv[i]; }
  c->quadGauss[0] = c->gauss; c->quadGauss[1] = g1; c->quadGauss[2] = 0.5 * g2;
  for (int i = 0; i < c->nefc; i++) {
    double D = d->efc_D[i];
    c->quad[3 * i] = 0.5 * D * c->Jaref[i] * c->Jaref[i];
    c->quad[3 * i + 1] = D * c->Jaref[i] * c->Jv[i];
    c->quad[3 * i + 2] = 0.5 * D * c->Jv[i] * c->Jv[i];
  }
  lspt p0, p1, p2, pmid, p1next, p2next;
  p0.alpha = 0; ls_eval(c, &p0);
  p1.alpha = p0.alpha - p0.d0 / p0.d1; ls_eval(c, &p1);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.d0) < gtol) return p1.alpha;
  int dir = p1.d0 < 0 ? 1 : -1, iter = 0, p2update = 0;
  p2 = p1;
  while (p1.d0 * dir <= -gtol && iter < m->ls_iterations) {
    p2 = p1; p2update = 1;
    p1.alpha = p1.alpha - p1.d0 / p1.d1; ls_eval(c, &p1); iter++;
    if (fabs(p1.d0) < gtol) return p1.alpha;
  }
  if (iter >= m->ls_iterations || !p2update) return p1.alpha;
  p2next = p1;
  p1next.alpha = p1.alpha - p1.d0 / p1.d1; ls_eval(c, &p1next);
  /* now p2 (deriv*dir<0 side) and p1 (deriv*dir>0 side) bracket the minimum */
  while (iter < m->ls_iterations) {
    pmid.alpha = 0.5 * (p1.alpha + p2.alpha); ls_eval(c, &pmid); iter++;
    lspt* cand[3] = {&p1next, &p2next, &pmid};
    int besti = -1;
    for (int i = 0; i < 3; i++)
      if (fabs(cand[i]->d0) < gtol && (besti < 0 || cand[i]->cost < cand[besti]->cost)) besti = i;
    if (besti >= 0) return cand[besti]->alpha;
    int b1 = 0, b2 = 0;
    for (int i = 0; i < 3; i++) {
      /* p2 holds the side with deriv*dir<0, p1 the side with deriv*dir>0 */
      if (cand[i]->d0 * dir < 0 && (cand[i]->alpha - p2.alpha) * dir > 0 && (p1.alpha - cand[i]->alpha) * dir > 0) { p2 = *cand[i]; b2 = 1; }
      else if (cand[i]->d0 * dir > 0 && (p1.alpha - cand[i]->alpha) * dir > 0 && (cand[i]->alpha - p2.alpha) * dir > 0) { p1 = *cand[i]; b1 = 1; }
    }
    if (!b1 && !b2) break;
    if (b1) { p1next.alpha = p1.alpha - p1.d0 / p1.d1; ls_eval(c, &p1next); }
    if (b2) { p2next.alpha = p2.alpha - p2.d0 / p2.d1; ls_eval(c, &p2next); }
  }
  return (p1.cost < p2.cost ? p1.alpha : p2.alpha);
}

void o_fwd_constraint(const omodel* m, odata* d) {
  int nv = m->nv, nefc = d->nefc;
  d->solver_iter = 0;
  if (!nefc) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * nv);
    memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
    return;
  }
  static __thread nctx cs;
  nctx* c = &cs;
  c->m = m; c->d = d; c->nv = nv; c->nefc = nefc;
  c->scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  /* warmstart(): better of qacc_warmstart and qacc_smooth */
  {
    double jar[MAXEFC], Ma[MAXV];
    for (int r = 0; r < nefc; r++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += d->efc_J[r * nv + k] * d->qacc_warmstart[k];
      jar[r] = s - d->efc_aref[r];
    }
    double cw = constraint_update(c, jar, 0);
    mulM(m, d, Ma, d->qacc_warmstart);
    cw += gauss_cost(c, Ma, d->qacc_warmstart);
    for (int r = 0; r < nefc; r++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += d->efc_J[r * nv + k] * d->qacc_smooth[k];
      jar[r] = s - d->efc_aref[r];
    }
    double csm = constraint_update(c, jar, 0);
    if (cw > csm) memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    else memcpy(d->qacc, d->qacc_warmstart, sizeof(double) * nv);
  }
  mulM(m, d, c->Ma, d->qacc);
  for (int r = 0; r < nefc; r++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[r * nv + k] * d->qacc[k];
    c->Jaref[r] = s - d->efc_aref[r];
  }
  update_all(c);
  for (int i = 0; i < nv; i++) c->search[i] = -c->Mgrad[i];
  int iter = 0;
  while (iter < m->iterations) {
    double alpha = line_search(c);
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * c->search[i]; c->Ma[i] += alpha * c->Mv[i]; }
    for (int r = 0; r < nefc; r++) c->Jaref[r] += alpha * c->Jv[r];
    double oldcost = c->cost;
    update_all(c);
    iter++;
    double gn = 0;
    for (int i = 0; i < nv; i++) gn += c->grad[i] * c->grad[i];
    double improvement = c->scale * (oldcost - c->cost), gradient = c->scale * sqrt(gn);
    if (improvement < m->tolerance || gradient < m->tolerance) break;
    for (int i = 0; i < nv; i++) c->search[i] = -c->Mgrad[i];
  }
  d->solver_iter = iter;
  d->solver_cost = c->cost;
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
}

/* ------------------------------------------------------------------ engine_forward.c */
void o_forward(const omodel* m, odata* d) {
  o_kinematics(m, d);
  o_compos(m, d);
  o_crb(m, d);
  o_factorM(m, d);
  o_collision(m, d);
  o_make_constraint(m, d);
  o_fwd_velocity(m, d);
  o_fwd_actuation(m, d);
  o_fwd_acceleration(m, d);
  o_fwd_constraint(m, d);
}

/* mj_Euler (implicit in joint damping) + mj_advance / mj_integratePos */
void o_euler(const omodel* m, odata* d) {
  int nv = m->nv;
  double qacc[MAXV], h = m->timestep;
  int damp = 0;
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damp = 1;
  if (!damp) memcpy(qacc, d->qacc, sizeof(double) * nv);
  else {
    double MhB[MAXV * MAXV], dinv[MAXV];
    memcpy(MhB, d->qM, sizeof(double) * m->nM);
    for (int i = 0; i < nv; i++) MhB[m->dof_Madr[i]] += h * m->dof_damping[i];
    factor_ld(m, MhB, dinv);
    for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    solve_ld(m, qacc, MhB, dinv);
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      if (m->disable_cube) { for (int k = 0; k < 6; k++) d->qvel[da + k] = 0; continue; }
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      double ax[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]}, qr[4];
      double ang = h * normalize3(ax);
      axisangle2quat(qr, ax, ang);
      normalize4(d->qpos + qa + 3);
      mulquat(d->qpos + qa + 3, d->qpos + qa + 3, qr);
    } else d->qpos[qa] += h * d->qvel[da];
  }
  d->time += h;
}

void o_step(const omodel* m, odata* d, int nstep) {
  for (int s = 0; s < nstep; s++) {
    o_forward(m, d);
    o_euler(m, d);
  }
}

/* inverse-dynamics style helpers for identity tests: full RNE with qacc (M*qacc + bias) */
void o_rne_acc(const omodel* m, odata* d, const double* qacc, double* result) {
  int nb = m->nbody, nv = m->nv;
  double cacc[MAXBODY * 6], cfrc[MAXBODY * 6];
  memset(cacc, 0, sizeof cacc);
  for (int k = 0; k < 3; k++) cacc[3 + k] = -m->gravity[k];
  memset(cfrc, 0, 6 * sizeof(double));
  for (int i = 1; i < nb; i++) {
    int bda = m->body_dofadr[i], dn = m->body_dofnum[i];
    double tmp[6], tmp1[6];
    memcpy(cacc + 6 * i, cacc + 6 * m->body_parentid[i], 6 * sizeof(double));
    for (int j = 0; j < dn; j++)
      for (int c = 0; c < 6; c++) cacc[6 * i + c] += d->cdof_dot[6 * (bda + j) + c] * d->qvel[bda + j] + d->cdof[6 * (bda + j) + c] * qacc[bda + j];
    mul_inert_vec(cfrc + 6 * i, d->cinert + 10 * i, cacc + 6 * i);
    mul_inert_vec(tmp, d->cinert + 10 * i, d->cvel + 6 * i);
    cross_force(tmp1, d->cvel + 6 * i, tmp);
    for (int c = 0; c < 6; c++) cfrc[6 * i + c] += tmp1[c];
  }
  for (int i = nb - 1; i > 0; i--)
    if (m->body_parentid[i])
      for (int c = 0; c < 6; c++) cfrc[6 * m->body_parentid[i] + c] += cfrc[6 * i + c];
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int c = 0; c < 6; c++) s += d->cdof[6 * i + c] * cfrc[6 * m->dof_bodyid[i] + c];
    result[i] = s + m->dof_armature[i] * qacc[i];
  }
}

int o_sizeof_data(void) { return (int)sizeof(odata); }
int o_sizeof_model(void) { return (int)sizeof(omodel); }
int o_maxv(void) { return MAXV; }
int o_maxefc(void) { return MAXEFC; }

/* field accessors so the Python side never has to mirror the odata layout */
#define ACC(name, type) type* o_##name(odata* d) { return d->name; }
ACC(mocap_pos, double) ACC(mocap_quat, double) ACC(qpos, double) ACC(qvel, double) ACC(ctrl, double) ACC(qacc_warmstart, double) ACC(qacc, double)
ACC(xpos, double) ACC(xquat, double) ACC(xmat, double) ACC(xipos, double) ACC(site_xpos, double) ACC(site_xmat, double)
ACC(geom_xpos, double) ACC(geom_xmat, double)
ACC(subtree_com, double) ACC(cdof, double) ACC(cinert, double) ACC(Mfull, double) ACC(qfrc_bias, double) ACC(qfrc_smooth, double)
ACC(qacc_smooth, double) ACC(qfrc_constraint, double) ACC(qfrc_actuator, double) ACC(qfrc_passive, double) ACC(actuator_force, double)
ACC(efc_J, double) ACC(efc_pos, double) ACC(efc_D, double) ACC(efc_R, double) ACC(efc_aref, double) ACC(efc_force, double) ACC(efc_type, int)
ACC(efc_diagApprox, double) ACC(efc_KBIP, double)
int o_nefc(odata* d) { return d->nefc; }
int o_ncon(odata* d) { return d->ncon; }
int o_solver_iter(odata* d) { return d->solver_iter; }
void o_contact(odata* d, int i, double* out /* dist, pos3, frame9, geom1, geom2, dim */) {
  ocontact* c = d->contact + i;
  out[0] = c->dist; memcpy(out + 1, c->pos, 3 * sizeof(double)); memcpy(out + 4, c->frame, 9 * sizeof(double));
  out[13] = c->geom1; out[14] = c->geom2; out[15] = c->dim;
}

/* test hook: MPR on two vertex sets given in world coordinates (interior points = vertex means); out = depth, dir[3], pos[3] */
int o_test_mpr(const double* va, int na, const double* vb, int nb, double* out) {
  static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z3[3] = {0, 0, 0};
  cvx A = {0, Z3, I3, va, NULL, na, {0, 0, 0}}, B = {0, Z3, I3, vb, NULL, nb, {0, 0, 0}};
  for (int i = 0; i < na; i++) for (int k = 0; k < 3; k++) A.center[k] += va[3 * i + k] / na;
  for (int i = 0; i < nb; i++) for (int k = 0; k < 3; k++) B.center[k] += vb[3 * i + k] / nb;
  return mpr_penetration(&A, &B, out, out + 1, out + 4);
}

/* ------------------------------------------------------------------ batched CPU rollout used as the bench cpu_baseline ("port") */
/* Each call steps `n` independent envs `nstep` substeps with the given ctrl (one thread; the
 * Python side fans out across threads with the GIL released by ctypes). */
void o_step_batch(const omodel* m, odata* d, int n, const double* ctrl, int nstep) {
  for (int e = 0; e < n; e++) {
    memcpy(d[e].ctrl, ctrl + e * m->nu, sizeof(double) * m->nu);
    o_step(m, d + e, nstep);
  }
}

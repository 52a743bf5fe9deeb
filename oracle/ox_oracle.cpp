/*
 * ox_oracle.cpp — CPU restatement (fp64, scalar, array-of-structs, one env at a time) of the
 * arithmetic behind oxide_control's hot path  Physics::step -> rusty_mujoco::mj_step
 * (reference src/physics.rs:44-46), plus Physics::forward (src/physics.rs:48-50) and
 * Physics::reset (src/physics.rs:52-54).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it. The product path
 * (oxide_control_b200/csrc) never links, includes or calls anything in this directory.
 *
 * PARITY UNPINNED. The arithmetic of the reference path lives in third-party code that is absent
 * from /root/reference: crate rusty_mujoco "0.1.0" (Cargo.toml:16) -> MuJoCo 3.3.2
 * (.github/workflows/CI.yml:28). The reference holds no tests, fixtures or golden vectors
 *  and MuJoCo cannot be built or imported here. This file restates
 * MuJoCo's published algorithm (stage list and formulas: SURVEY.md Appendix A; decisions where the
 * documentation leaves room: ORACLE_DECISIONS.md). It is pinned instead by closed-form and
 * algorithm-independent checks in tests/  and by the golden hook
 * tools/dump_mujoco_golden.py (output goes to tests/golden/).
 *
 * Input model = the ox_model_tables struct of include/ox_b200.h (mirror of the mjModel fields the
 * path reads). Stage functions are exported individually so the GPU path can be compared stage by stage.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/ox_b200.h"

typedef ox_model_tables Model;
static const double OX_PI_D = 3.14159265358979323846;

struct oxo_data {
  // state (mjData fields of the same names)
  std::vector<double> qpos, qvel, ctrl, qfrc_applied, xfrc_applied, qacc_warmstart, act, act_dot, mocap_pos, mocap_quat, eq_active, ten_length, ten_J;
  double time = 0;
  // position stage
  std::vector<double> xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, geom_xpos, geom_xmat, site_xpos, site_xmat;
  std::vector<double> subtree_com, cinert, cdof, crb, qM, qLD, qLDiagInv;
  // velocity / actuation / acceleration
  std::vector<double> cvel, cdof_dot, qfrc_bias, qfrc_passive, actuator_force, qfrc_actuator, qfrc_smooth, qacc_smooth;
  // contacts
  int ncon = 0;
  std::vector<double> con_dist, con_pos, con_frame;
  std::vector<int> con_pair;
  // constraints
  int nefc = 0, ne = 0, nf = 0;  // ne equality rows (first, two-sided quadratic), then nf dry-friction rows (Huber cost, |force| <= frictionloss)
  std::vector<double> efc_J, efc_pos, efc_margin, efc_D, efc_R, efc_aref, efc_vel, efc_force, efc_diagApprox, efc_floss;
  std::vector<int> efc_type, efc_id;
  // solution
  std::vector<double> qacc, qfrc_constraint;
  int solver_niter = 0;
  std::vector<double> sensordata;
  int diverged = 0;  // number of auto-resets (mj_checkPos/Vel/Acc warnings)
};
typedef oxo_data Data;

namespace {

// ---------------------------------------------------------------- small math (mju_* equivalents)
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
inline double normalize3(double* v) {
  double n = std::sqrt(dot3(v, v));
  if (n < OX_MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; } else { double s = 1 / n; v[0] *= s; v[1] *= s; v[2] *= s; }
  return n;
}
inline void normalize4(double* q) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < OX_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { double s = 1 / n; for (int i = 0; i < 4; i++) q[i] *= s; }
}
inline void mulQuat(double* r, const double* a, const double* b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  std::memcpy(r, t, sizeof t);
}
inline void quat2Mat(double* m, const double* q) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03);
  m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
inline void mulMatVec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
inline void rotVecQuat(double* r, const double* v, const double* q) {
  double m[9];
  quat2Mat(m, q);
  mulMatVec3(r, m, v);
}
inline void axisAngle2Quat(double* q, const double* axis, double angle) {
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double s = std::sin(angle * 0.5);
  q[0] = std::cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
// quaternion difference as a 3D velocity: qb * quat(res) = qa  (mju_subQuat)
inline void subQuat(double* res, const double* qa, const double* qb) {
  double qneg[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, qdif[4];
  mulQuat(qdif, qneg, qa);
  double axis[3] = {qdif[1], qdif[2], qdif[3]};
  double sin_a_2 = normalize3(axis);
  double speed = 2 * std::atan2(sin_a_2, qdif[0]);
  if (speed > M_PI) speed -= 2 * M_PI;
  res[0] = axis[0] * speed; res[1] = axis[1] * speed; res[2] = axis[2] * speed;
}
inline void quatIntegrate(double* quat, const double* vel, double scale) {
  double tmp[3] = {vel[0], vel[1], vel[2]}, qrot[4];
  double angle = scale * normalize3(tmp);
  axisAngle2Quat(qrot, tmp, angle);
  normalize4(quat);
  mulQuat(quat, quat, qrot);
}
// spatial algebra, 6-vectors are [angular; linear]
inline void crossMotion(double* r, const double* vel, const double* v) {
  r[0] = -vel[2] * v[1] + vel[1] * v[2];
  r[1] = vel[2] * v[0] - vel[0] * v[2];
  r[2] = -vel[1] * v[0] + vel[0] * v[1];
  r[3] = -vel[2] * v[4] + vel[1] * v[5] - vel[5] * v[1] + vel[4] * v[2];
  r[4] = vel[2] * v[3] - vel[0] * v[5] + vel[5] * v[0] - vel[3] * v[2];
  r[5] = -vel[1] * v[3] + vel[0] * v[4] - vel[4] * v[0] + vel[3] * v[1];
}
inline void crossForce(double* r, const double* vel, const double* f) {
  r[0] = -vel[2] * f[1] + vel[1] * f[2] - vel[5] * f[4] + vel[4] * f[5];
  r[1] = vel[2] * f[0] - vel[0] * f[2] + vel[5] * f[3] - vel[3] * f[5];
  r[2] = -vel[1] * f[0] + vel[0] * f[1] - vel[4] * f[3] + vel[3] * f[4];
  r[3] = -vel[2] * f[4] + vel[1] * f[5];
  r[4] = vel[2] * f[3] - vel[0] * f[5];
  r[5] = -vel[1] * f[3] + vel[0] * f[4];
}
// 10-number com-frame inertia times motion vector
inline void mulInertVec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
inline double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
inline bool disabled(const Model* m, int bit) { return (m->disableflags & bit) != 0; }

// ---------------------------------------------------------------- A.1 kinematics
void kinematics(const Model* m, Data* d) {
  double* xpos = d->xpos.data();
  double* xquat = d->xquat.data();
  double* xmat = d->xmat.data();
  xpos[0] = xpos[1] = xpos[2] = 0;
  xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
  for (int k = 0; k < 9; k++) { xmat[k] = (k % 4 == 0); d->ximat[k] = (k % 4 == 0); }
  d->xipos[0] = d->xipos[1] = d->xipos[2] = 0;
  for (int i = 1; i < m->nbody; i++) {
    double pos[3], quat[4];
    int jntadr = m->body_jntadr[i], jntnum = m->body_jntnum[i];
    if (m->nmocap > 0 && m->body_mocapid[i] >= 0) {
      // mj_kinematics: a mocap body takes its pose from mjData.mocap_pos / mocap_quat (normalised copy)
      int id = m->body_mocapid[i];
      std::memcpy(pos, &d->mocap_pos[3 * id], 3 * sizeof(double));
      std::memcpy(quat, &d->mocap_quat[4 * id], 4 * sizeof(double));
      normalize4(quat);
    } else if (jntnum == 1 && m->jnt_type[jntadr] == OX_JNT_FREE) {
      int qadr = m->jnt_qposadr[jntadr];
      std::memcpy(pos, &d->qpos[qadr], 3 * sizeof(double));
      std::memcpy(quat, &d->qpos[qadr + 3], 4 * sizeof(double));
      normalize4(quat);
      std::memcpy(&d->xanchor[3 * jntadr], pos, 3 * sizeof(double));
      std::memcpy(&d->xaxis[3 * jntadr], &m->jnt_axis[3 * jntadr], 3 * sizeof(double));
    } else {
      int pid = m->body_parentid[i];
      mulMatVec3(pos, xmat + 9 * pid, m->body_pos + 3 * i);
      for (int k = 0; k < 3; k++) pos[k] += xpos[3 * pid + k];
      mulQuat(quat, xquat + 4 * pid, m->body_quat + 4 * i);
      for (int j = 0; j < jntnum; j++) {
        int jid = jntadr + j, qadr = m->jnt_qposadr[jid], jt = m->jnt_type[jid];
        double anchor[3], axis[3];
        rotVecQuat(axis, m->jnt_axis + 3 * jid, quat);
        rotVecQuat(anchor, m->jnt_pos + 3 * jid, quat);
        for (int k = 0; k < 3; k++) anchor[k] += pos[k];
        if (jt == OX_JNT_SLIDE) {
          double dq = d->qpos[qadr] - m->qpos0[qadr];
          for (int k = 0; k < 3; k++) pos[k] += axis[k] * dq;
        } else if (jt == OX_JNT_BALL || jt == OX_JNT_HINGE) {
          double qloc[4];
          if (jt == OX_JNT_BALL) {
            std::memcpy(qloc, &d->qpos[qadr], 4 * sizeof(double));
            normalize4(qloc);
          } else {
            axisAngle2Quat(qloc, m->jnt_axis + 3 * jid, d->qpos[qadr] - m->qpos0[qadr]);
          }
          mulQuat(quat, quat, qloc);
          double vec[3];
          rotVecQuat(vec, m->jnt_pos + 3 * jid, quat);
          for (int k = 0; k < 3; k++) pos[k] = anchor[k] - vec[k];
        }
        std::memcpy(&d->xanchor[3 * jid], anchor, sizeof anchor);
        std::memcpy(&d->xaxis[3 * jid], axis, sizeof axis);
      }
    }
    normalize4(quat);
    std::memcpy(xquat + 4 * i, quat, sizeof quat);
    std::memcpy(xpos + 3 * i, pos, sizeof pos);
    quat2Mat(xmat + 9 * i, quat);
  }
  for (int i = 1; i < m->nbody; i++) {
    double v[3], q[4];
    mulMatVec3(v, xmat + 9 * i, m->body_ipos + 3 * i);
    for (int k = 0; k < 3; k++) d->xipos[3 * i + k] = xpos[3 * i + k] + v[k];
    mulQuat(q, xquat + 4 * i, m->body_iquat + 4 * i);
    quat2Mat(&d->ximat[9 * i], q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double v[3], q[4];
    mulMatVec3(v, xmat + 9 * b, m->geom_pos + 3 * g);
    for (int k = 0; k < 3; k++) d->geom_xpos[3 * g + k] = xpos[3 * b + k] + v[k];
    mulQuat(q, xquat + 4 * b, m->geom_quat + 4 * g);
    quat2Mat(&d->geom_xmat[9 * g], q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    double v[3], q[4];
    mulMatVec3(v, xmat + 9 * b, m->site_pos + 3 * s);
    for (int k = 0; k < 3; k++) d->site_xpos[3 * s + k] = xpos[3 * b + k] + v[k];
    mulQuat(q, xquat + 4 * b, m->site_quat + 4 * s);
    quat2Mat(&d->site_xmat[9 * s], q);
  }
}

// ---------------------------------------------------------------- A.2 com-frame quantities
void comPos(const Model* m, Data* d) {
  double* sc = d->subtree_com.data();
  std::fill(d->subtree_com.begin(), d->subtree_com.end(), 0.0);
  for (int i = m->nbody - 1; i >= 0; i--) {
    for (int k = 0; k < 3; k++) sc[3 * i + k] += d->xipos[3 * i + k] * m->body_mass[i];
    if (i) {
      int j = m->body_parentid[i];
      for (int k = 0; k < 3; k++) sc[3 * j + k] += sc[3 * i + k];
    }
    if (m->body_subtreemass[i] < OX_MINVAL) std::memcpy(sc + 3 * i, &d->xipos[3 * i], 3 * sizeof(double));
    else for (int k = 0; k < 3; k++) sc[3 * i + k] /= m->body_subtreemass[i];
  }
  std::fill(d->cinert.begin(), d->cinert.begin() + 10, 0.0);
  for (int i = 1; i < m->nbody; i++) {
    const double* mat = &d->ximat[9 * i];
    const double* inert = m->body_inertia + 3 * i;
    double mass = m->body_mass[i], dif[3];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[3 * i + k] - sc[3 * m->body_rootid[i] + k];
    double* res = &d->cinert[10 * i];
    double tmp[9];  // diag(inert) * mat'
    tmp[0] = mat[0] * inert[0]; tmp[3] = mat[1] * inert[1]; tmp[6] = mat[2] * inert[2];
    tmp[1] = mat[3] * inert[0]; tmp[4] = mat[4] * inert[1]; tmp[7] = mat[5] * inert[2];
    tmp[2] = mat[6] * inert[0]; tmp[5] = mat[7] * inert[1]; tmp[8] = mat[8] * inert[2];
    res[0] = mat[0] * tmp[0] + mat[1] * tmp[3] + mat[2] * tmp[6];
    res[1] = mat[3] * tmp[1] + mat[4] * tmp[4] + mat[5] * tmp[7];
    res[2] = mat[6] * tmp[2] + mat[7] * tmp[5] + mat[8] * tmp[8];
    res[3] = mat[0] * tmp[1] + mat[1] * tmp[4] + mat[2] * tmp[7];
    res[4] = mat[0] * tmp[2] + mat[1] * tmp[5] + mat[2] * tmp[8];
    res[5] = mat[3] * tmp[2] + mat[4] * tmp[5] + mat[5] * tmp[8];
    res[0] += mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    res[1] += mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    res[2] += mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    res[3] -= mass * dif[0] * dif[1];
    res[4] -= mass * dif[0] * dif[2];
    res[5] -= mass * dif[1] * dif[2];
    res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2];
    res[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) {
    int bi = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    double offset[3];
    for (int k = 0; k < 3; k++) offset[k] = sc[3 * m->body_rootid[bi] + k] - d->xanchor[3 * j + k];
    double* cdof = d->cdof.data();
    int skip = 0;
    switch (m->jnt_type[j]) {
      case OX_JNT_FREE:
        std::fill(cdof + 6 * da, cdof + 6 * da + 18, 0.0);
        for (int i = 0; i < 3; i++) cdof[6 * (da + i) + 3 + i] = 1;
        skip = 3;
        // fallthrough
      case OX_JNT_BALL:
        for (int i = 0; i < 3; i++) {
          double axis[3] = {d->xmat[9 * bi + i], d->xmat[9 * bi + i + 3], d->xmat[9 * bi + i + 6]};
          double* r = cdof + 6 * (da + i + skip);
          r[0] = axis[0]; r[1] = axis[1]; r[2] = axis[2];
          cross3(r + 3, axis, offset);
        }
        break;
      case OX_JNT_SLIDE: {
        double* r = cdof + 6 * da;
        r[0] = r[1] = r[2] = 0;
        std::memcpy(r + 3, &d->xaxis[3 * j], 3 * sizeof(double));
        break;
      }
      case OX_JNT_HINGE: {
        double* r = cdof + 6 * da;
        std::memcpy(r, &d->xaxis[3 * j], 3 * sizeof(double));
        cross3(r + 3, &d->xaxis[3 * j], offset);
        break;
      }
    }
  }
}

// ---------------------------------------------------------------- A.3 composite rigid body
void crb(const Model* m, Data* d) {
  d->crb = d->cinert;
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * i + k];
  }
  std::fill(d->qM.begin(), d->qM.end(), 0.0);
  for (int i = 0; i < m->nv; i++) {
    int adr = m->dof_Madr[i];
    d->qM[adr] = m->dof_armature[i];
    double buf[6];
    mulInertVec(buf, &d->crb[10 * m->dof_bodyid[i]], &d->cdof[6 * i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) d->qM[adr++] += dot6(&d->cdof[6 * j], buf);
  }
}

// ---------------------------------------------------------------- A.4 sparse L'DL
void factorLD(const Model* m, double* qLD, double* qLDiagInv) {
  for (int k = m->nv - 1; k >= 0; k--) {
    int Madr_kk = m->dof_Madr[k], Madr_ki = Madr_kk + 1, i = m->dof_parentid[k];
    while (i >= 0) {
      double tmp = qLD[Madr_ki] / qLD[Madr_kk];
      // row i holds M(i,i), M(i,parent(i)), ...; row k from Madr_ki holds M(k,i), M(k,parent(i)), ...
      int n = 1;
      for (int a = m->dof_parentid[i]; a >= 0; a = m->dof_parentid[a]) n++;
      for (int c = 0; c < n; c++) qLD[m->dof_Madr[i] + c] -= tmp * qLD[Madr_ki + c];
      qLD[Madr_ki] = tmp;
      i = m->dof_parentid[i];
      Madr_ki++;
    }
    qLDiagInv[k] = 1.0 / qLD[Madr_kk];
  }
}
void factorM(const Model* m, Data* d) {
  d->qLD = d->qM;
  factorLD(m, d->qLD.data(), d->qLDiagInv.data());
}
void solveLD(const Model* m, const double* qLD, const double* qLDiagInv, double* x) {
  for (int i = m->nv - 1; i >= 0; i--) {
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= qLD[adr++] * x[i];
  }
  for (int i = 0; i < m->nv; i++) x[i] *= qLDiagInv[i];
  for (int i = 0; i < m->nv; i++) {
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[i] -= qLD[adr++] * x[j];
  }
}
void mulM(const Model* m, const double* qM, double* res, const double* v) {
  for (int i = 0; i < m->nv; i++) res[i] = 0;
  for (int i = 0; i < m->nv; i++) {
    int adr = m->dof_Madr[i];
    res[i] += qM[adr] * v[i];
    int k = 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j], k++) {
      res[i] += qM[adr + k] * v[j];
      res[j] += qM[adr + k] * v[i];
    }
  }
}

// ---------------------------------------------------------------- A.5 collision
struct RawContact { double dist, pos[3], frame[9]; };

int planeSphere(RawContact* con, double margin, const double* pos1, const double* mat1, const double* pos2, double radius) {
  double normal[3] = {mat1[2], mat1[5], mat1[8]};
  double tmp[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double cdist = dot3(tmp, normal);
  if (cdist > margin + radius) return 0;
  con->dist = cdist - radius;
  std::memcpy(con->frame, normal, sizeof normal);
  con->frame[3] = con->frame[4] = con->frame[5] = 0;
  for (int k = 0; k < 3; k++) con->pos[k] = pos2[k] + normal[k] * (-con->dist / 2 - radius);
  return 1;
}
int sphereSphere(RawContact* con, double margin, const double* pos1, double r1, const double* pos2, double r2) {
  double dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double cdist2 = dot3(dif, dif), mind = margin + r1 + r2;
  if (cdist2 > mind * mind) return 0;
  con->dist = std::sqrt(cdist2) - r1 - r2;
  std::memcpy(con->frame, dif, sizeof dif);
  normalize3(con->frame);
  con->frame[3] = con->frame[4] = con->frame[5] = 0;
  for (int k = 0; k < 3; k++) con->pos[k] = pos1[k] + con->frame[k] * (r1 + con->dist / 2);
  return 1;
}
inline double clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// sphere (centre, radius) against a box (pos2, mat2, half sizes): mjc_SphereBox. The closest point of the box to the centre is
// the centre clamped to the box in the box frame; a centre inside the box is pushed out through the nearest face.
// The normal points from the sphere (geom1) to the box (geom2).
int sphereBox(RawContact* con, double margin, const double* centre, double radius, const double* pos2, const double* mat2, const double* size2) {
  double tmp[3] = {centre[0] - pos2[0], centre[1] - pos2[1], centre[2] - pos2[2]}, c[3], clamped[3], dir[3], lpos[3], lnorm[3];
  for (int k = 0; k < 3; k++) c[k] = mat2[k] * tmp[0] + mat2[3 + k] * tmp[1] + mat2[6 + k] * tmp[2];   // mat2' * tmp
  for (int k = 0; k < 3; k++) { clamped[k] = clip(c[k], -size2[k], size2[k]); dir[k] = clamped[k] - c[k]; }
  double dist = std::sqrt(dot3(dir, dir));
  if (dist - radius > margin) return 0;
  if (dist <= OX_MINVAL) {   // centre inside the box
    double closest = 2 * std::max(size2[0], std::max(size2[1], size2[2]));
    int face = 0;
    for (int i = 0; i < 6; i++) {
      double fd = std::fabs((i % 2 ? 1.0 : -1.0) * size2[i / 2] - c[i / 2]);
      if (fd < closest) { closest = fd; face = i; }
    }
    lnorm[0] = lnorm[1] = lnorm[2] = 0;
    lnorm[face / 2] = face % 2 ? -1.0 : 1.0;
    for (int k = 0; k < 3; k++) lpos[k] = c[k] + lnorm[k] * (radius - closest) / 2;
    con->dist = -closest - radius;
  } else {
    for (int k = 0; k < 3; k++) lnorm[k] = dir[k] / dist;
    for (int k = 0; k < 3; k++) lpos[k] = 0.5 * (clamped[k] + c[k] + lnorm[k] * radius);   // midway between the box point and the deepest sphere point
    con->dist = dist - radius;
  }
  mulMatVec3(con->frame, mat2, lnorm);
  con->frame[3] = con->frame[4] = con->frame[5] = 0;
  mulMatVec3(tmp, mat2, lpos);
  for (int k = 0; k < 3; k++) con->pos[k] = tmp[k] + pos2[k];
  return 1;
}

// capsule against a box. NOT a restatement of mjc_CapsuleBox (its case analysis is not reproducible from the documentation; see
// ORACLE_DECISIONS.md #12): the capsule is a swept sphere, so the first contact is sphereBox at the point of the axis segment
// closest to the box - found EXACTLY: the derivative g(t) of half the squared distance along the segment is piecewise linear and
// non-decreasing, with breakpoints where a coordinate crosses a face; its zero crossing lies between the last candidate with
// g < 0 and the first with g >= 0. The second contact is the far end cap, if it is within the margin (a capsule lying along a
// face or an edge gets two contacts, as with the plane).
int capsuleBox(RawContact* con, double margin, const double* pos1, const double* mat1, const double* size1, const double* pos2, const double* mat2,
               const double* size2) {
  const double hl = size1[1], radius = size1[0];
  double tmp[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]}, ax[3] = {mat1[2] * hl, mat1[5] * hl, mat1[8] * hl}, c[3], a[3];
  for (int k = 0; k < 3; k++) {
    c[k] = mat2[k] * tmp[0] + mat2[3 + k] * tmp[1] + mat2[6 + k] * tmp[2];
    a[k] = mat2[k] * ax[0] + mat2[3 + k] * ax[1] + mat2[6 + k] * ax[2];
  }
  auto g = [&](double t) {
    double s = 0;
    for (int k = 0; k < 3; k++) { double pk = c[k] + t * a[k]; s += a[k] * (pk - clip(pk, -size2[k], size2[k])); }
    return s;
  };
  double cand[8] = {-1, 1, 0, 0, 0, 0, 0, 0};
  int nc = 2;
  for (int k = 0; k < 3; k++)
    if (std::fabs(a[k]) > OX_MINVAL)
      for (double sgn : {-1.0, 1.0}) {
        double t = (sgn * size2[k] - c[k]) / a[k];
        if (t > -1 && t < 1) cand[nc++] = t;
      }
  // the axis itself cuts through the box (deep penetration): every point of the cut is at distance zero, take its middle - the
  // nearest-face rule of sphereBox needs a point well inside, not one on the surface
  double tin = -1, tout = 1;
  bool cut = true;
  for (int k = 0; k < 3; k++) {
    if (std::fabs(a[k]) > OX_MINVAL) {
      double t1 = (-size2[k] - c[k]) / a[k], t2 = (size2[k] - c[k]) / a[k];
      if (t1 > t2) std::swap(t1, t2);
      tin = std::max(tin, t1); tout = std::min(tout, t2);
    } else if (std::fabs(c[k]) > size2[k]) cut = false;
  }
  double tstar;
  if (cut && tin <= tout) tstar = 0.5 * (tin + tout);
  else if (g(-1) >= 0) tstar = -1;
  else if (g(1) <= 0) tstar = 1;
  else {
    double ta = -1, ga = g(-1), tb = 1, gb = g(1);
    for (int i = 2; i < nc; i++) {
      double gi = g(cand[i]);
      if (gi < 0) { if (cand[i] > ta) { ta = cand[i]; ga = gi; } }
      else if (cand[i] < tb) { tb = cand[i]; gb = gi; }
    }
    tstar = gb - ga > OX_MINVAL ? ta - ga * (tb - ta) / (gb - ga) : ta;
  }
  double pt[3];
  for (int k = 0; k < 3; k++) pt[k] = pos1[k] + tstar * ax[k];
  int n = sphereBox(con, margin, pt, radius, pos2, mat2, size2);
  const double tfar = tstar <= 0 ? 1.0 : -1.0;
  for (int k = 0; k < 3; k++) pt[k] = pos1[k] + tfar * ax[k];
  n += sphereBox(con + n, margin, pt, radius, pos2, mat2, size2);
  for (int i = 0; i < n; i++) { con[i].frame[3] = mat1[2]; con[i].frame[4] = mat1[5]; con[i].frame[5] = mat1[8]; }   // tangent hint: the capsule axis
  return n;
}

// box against box. NOT a restatement of mjc_BoxBox (ORACLE_DECISIONS.md #32): separating-axis test over the 6 face normals and
// the 9 edge-edge cross products; the axis of least penetration decides the contact kind. Face axis: the face of the other box
// most opposed to the normal is clipped (Sutherland-Hodgman) against the four side planes of the reference face and every vertex
// of the clipped polygon within the margin of the reference face becomes a contact (up to 8). Edge axis: one contact between the
// closest points of the two supporting edges. Face axes win ties (an edge axis must be better by a relative 1e-3 of the box size).
int boxBox(RawContact* con, double margin, const double* pos1, const double* mat1, const double* size1, const double* pos2, const double* mat2,
           const double* size2) {
  double A[3][3], B[3][3], dvec[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) { A[i][k] = mat1[3 * k + i]; B[i][k] = mat2[3 * k + i]; }   // axis i = column i
  auto radius = [&](const double (*ax)[3], const double* sz, const double* L) {
    return sz[0] * std::fabs(dot3(ax[0], L)) + sz[1] * std::fabs(dot3(ax[1], L)) + sz[2] * std::fabs(dot3(ax[2], L));
  };
  double bestFace = -1e300, bestEdge = -1e300;
  int faceIdx = -1, edgeI = -1, edgeJ = -1;
  for (int f = 0; f < 6; f++) {
    const double* L = f < 3 ? A[f] : B[f - 3];
    const double sep = std::fabs(dot3(dvec, L)) - radius(A, size1, L) - radius(B, size2, L);
    if (sep > margin) return 0;
    if (sep > bestFace) { bestFace = sep; faceIdx = f; }
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double L[3];
      cross3(L, A[i], B[j]);
      const double n = std::sqrt(dot3(L, L));
      if (n < 1e-6) continue;   // parallel edges: a face axis covers it
      for (int k = 0; k < 3; k++) L[k] /= n;
      const double sep = std::fabs(dot3(dvec, L)) - radius(A, size1, L) - radius(B, size2, L);
      if (sep > margin) return 0;
      if (sep > bestEdge) { bestEdge = sep; edgeI = i; edgeJ = j; }
    }
  const double scale = std::max(std::max(size1[0], size1[1]), std::max(size1[2], std::max(size2[0], std::max(size2[1], size2[2]))));
  if (edgeI >= 0 && bestEdge > bestFace + 1e-3 * scale) {
    double L[3];
    cross3(L, A[edgeI], B[edgeJ]);
    normalize3(L);
    if (dot3(L, dvec) < 0) for (double& v : L) v = -v;   // from box 1 to box 2
    double c1[3] = {pos1[0], pos1[1], pos1[2]}, c2[3] = {pos2[0], pos2[1], pos2[2]};
    for (int k = 0; k < 3; k++) {
      if (k != edgeI) { const double sg = dot3(A[k], L) > 0 ? 1.0 : -1.0; for (int c = 0; c < 3; c++) c1[c] += sg * size1[k] * A[k][c]; }
      if (k != edgeJ) { const double sg = dot3(B[k], L) > 0 ? -1.0 : 1.0; for (int c = 0; c < 3; c++) c2[c] += sg * size2[k] * B[k][c]; }
    }
    // closest points of the lines c1 + s A_i, c2 + t B_j, clamped to the edges
    const double* u = A[edgeI]; const double* v = B[edgeJ];
    double w[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    const double b = dot3(u, v), dd = dot3(u, w), e = dot3(v, w), den = 1 - b * b;
    double sA = den > 1e-12 ? (b * e - dd) / den : 0.0, tB = den > 1e-12 ? (e - b * dd) / den : 0.0;
    sA = clip(sA, -size1[edgeI], size1[edgeI]); tB = clip(tB, -size2[edgeJ], size2[edgeJ]);
    double pA[3], pB[3];
    for (int c = 0; c < 3; c++) { pA[c] = c1[c] + sA * u[c]; pB[c] = c2[c] + tB * v[c]; }
    const double diff[3] = {pB[0] - pA[0], pB[1] - pA[1], pB[2] - pA[2]};
    con->dist = dot3(diff, L);
    if (con->dist > margin) return 0;
    for (int c = 0; c < 3; c++) { con->pos[c] = 0.5 * (pA[c] + pB[c]); con->frame[c] = L[c]; con->frame[3 + c] = 0; }
    return 1;
  }
  // face contact: reference box R (the owner of the axis), incident box I
  const bool refA = faceIdx < 3;
  const int ra = refA ? faceIdx : faceIdx - 3;
  const double (*Rax)[3] = refA ? A : B; const double (*Iax)[3] = refA ? B : A;
  const double* Rpos = refA ? pos1 : pos2; const double* Ipos = refA ? pos2 : pos1;
  const double* Rsz = refA ? size1 : size2; const double* Isz = refA ? size2 : size1;
  double n[3], toI[3] = {Ipos[0] - Rpos[0], Ipos[1] - Rpos[1], Ipos[2] - Rpos[2]};
  const double sgn = dot3(toI, Rax[ra]) >= 0 ? 1.0 : -1.0;
  for (int c = 0; c < 3; c++) n[c] = sgn * Rax[ra][c];   // outward normal of the reference face, towards the incident box
  int ia = 0;
  double best = -1;
  for (int k = 0; k < 3; k++) { const double a = std::fabs(dot3(Iax[k], n)); if (a > best) { best = a; ia = k; } }
  const double isg = dot3(Iax[ia], n) > 0 ? -1.0 : 1.0;   // the incident face looks against n
  const int iu = (ia + 1) % 3, iv = (ia + 2) % 3;
  double poly[16][3], tmp[16][3];
  int np = 4;
  static const double su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
  for (int q = 0; q < 4; q++)
    for (int c = 0; c < 3; c++)
      poly[q][c] = Ipos[c] + isg * Isz[ia] * Iax[ia][c] + su[q] * Isz[iu] * Iax[iu][c] + sv[q] * Isz[iv] * Iax[iv][c];
  for (int side = 0; side < 4 && np > 0; side++) {   // clip against the four side planes of the reference face
    const int ta = (ra + 1 + side / 2) % 3;
    const double ps = side % 2 ? -1.0 : 1.0;
    auto inside = [&](const double* pt) {
      const double rel[3] = {pt[0] - Rpos[0], pt[1] - Rpos[1], pt[2] - Rpos[2]};
      return Rsz[ta] - ps * dot3(rel, Rax[ta]);   // >= 0 inside
    };
    int no = 0;
    for (int q = 0; q < np; q++) {
      const double* P = poly[q]; const double* Q = poly[(q + 1) % np];
      const double dp = inside(P), dq = inside(Q);
      if (dp >= 0) { for (int c = 0; c < 3; c++) tmp[no][c] = P[c]; no++; }
      if ((dp >= 0) != (dq >= 0)) { const double t = dp / (dp - dq); for (int c = 0; c < 3; c++) tmp[no][c] = P[c] + t * (Q[c] - P[c]); no++; }
    }
    np = no;
    for (int q = 0; q < np; q++) for (int c = 0; c < 3; c++) poly[q][c] = tmp[q][c];
  }
  int cnt = 0;
  for (int q = 0; q < np && cnt < 8; q++) {
    const double rel[3] = {poly[q][0] - Rpos[0], poly[q][1] - Rpos[1], poly[q][2] - Rpos[2]};
    const double depth = dot3(rel, n) - Rsz[ra];
    if (depth > margin) continue;
    con[cnt].dist = depth;
    for (int c = 0; c < 3; c++) {
      con[cnt].pos[c] = poly[q][c] - 0.5 * depth * n[c];
      con[cnt].frame[c] = refA ? n[c] : -n[c];   // from box 1 to box 2
      con[cnt].frame[3 + c] = 0;
    }
    cnt++;
  }
  return cnt;
}

int collidePair(const Model* m, const Data* d, int p, RawContact* con) {
  int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
  int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  const double *pos1 = &d->geom_xpos[3 * g1], *mat1 = &d->geom_xmat[9 * g1], *size1 = m->geom_size + 3 * g1;
  const double *pos2 = &d->geom_xpos[3 * g2], *mat2 = &d->geom_xmat[9 * g2], *size2 = m->geom_size + 3 * g2;
  double margin = m->pair_margin[p];
  if (t1 == OX_GEOM_PLANE && t2 == OX_GEOM_SPHERE) return planeSphere(con, margin, pos1, mat1, pos2, size2[0]);
  if (t1 == OX_GEOM_PLANE && t2 == OX_GEOM_CAPSULE) {
    double axis[3] = {mat2[2], mat2[5], mat2[8]}, seg[3], pt[3];
    for (int k = 0; k < 3; k++) seg[k] = axis[k] * size2[1];
    for (int k = 0; k < 3; k++) pt[k] = pos2[k] + seg[k];
    int n1 = planeSphere(con, margin, pos1, mat1, pt, size2[0]);
    for (int k = 0; k < 3; k++) pt[k] = pos2[k] - seg[k];
    int n2 = planeSphere(con + n1, margin, pos1, mat1, pt, size2[0]);
    for (int c = 0; c < n1 + n2; c++) std::memcpy(con[c].frame + 3, axis, sizeof axis);
    return n1 + n2;
  }
  if (t1 == OX_GEOM_PLANE && t2 == OX_GEOM_BOX) {
    double norm[3] = {mat1[2], mat1[5], mat1[8]};
    double dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
    double dist = dot3(dif, norm);
    int cnt = 0;
    for (int i = 0; i < 8; i++) {
      double vec[3] = {(i & 1 ? size2[0] : -size2[0]), (i & 2 ? size2[1] : -size2[1]), (i & 4 ? size2[2] : -size2[2])};
      double corner[3];
      mulMatVec3(corner, mat2, vec);
      double ldist = dot3(norm, corner);
      if (dist + ldist > margin || ldist > 0) continue;
      con[cnt].dist = dist + ldist;
      std::memcpy(con[cnt].frame, norm, sizeof norm);
      con[cnt].frame[3] = con[cnt].frame[4] = con[cnt].frame[5] = 0;
      for (int k = 0; k < 3; k++) con[cnt].pos[k] = corner[k] + pos2[k] + norm[k] * (-con[cnt].dist / 2);
      if (++cnt >= 4) return 4;
    }
    return cnt;
  }
  if (t1 == OX_GEOM_SPHERE && t2 == OX_GEOM_SPHERE) return sphereSphere(con, margin, pos1, size1[0], pos2, size2[0]);
  if (t1 == OX_GEOM_SPHERE && t2 == OX_GEOM_CAPSULE) {
    double axis[3] = {mat2[2], mat2[5], mat2[8]};
    double vec[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
    double x = clip(dot3(axis, vec), -size2[1], size2[1]);
    for (int k = 0; k < 3; k++) vec[k] = pos2[k] + axis[k] * x;
    return sphereSphere(con, margin, pos1, size1[0], vec, size2[0]);
  }
  if (t1 == OX_GEOM_BOX && t2 == OX_GEOM_BOX) return boxBox(con, margin, pos1, mat1, size1, pos2, mat2, size2);
  if (t1 == OX_GEOM_SPHERE && t2 == OX_GEOM_BOX) return sphereBox(con, margin, pos1, size1[0], pos2, mat2, size2);
  if (t1 == OX_GEOM_CAPSULE && t2 == OX_GEOM_BOX) return capsuleBox(con, margin, pos1, mat1, size1, pos2, mat2, size2);
  if (t1 == OX_GEOM_CAPSULE && t2 == OX_GEOM_CAPSULE) {
    double axis1[3] = {mat1[2] * size1[1], mat1[5] * size1[1], mat1[8] * size1[1]};
    double axis2[3] = {mat2[2] * size2[1], mat2[5] * size2[1], mat2[8] * size2[1]};
    double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
    double ma = dot3(axis1, axis1), mb = -dot3(axis1, axis2), mc = dot3(axis2, axis2);
    double u = -dot3(axis1, dif), v = dot3(axis2, dif), det = ma * mc - mb * mb;
    double vec1[3], vec2[3];
    if (std::fabs(det) >= OX_MINVAL) {
      double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
      if (x1 > 1) { x1 = 1; x2 = (v - mb) / mc; } else if (x1 < -1) { x1 = -1; x2 = (v + mb) / mc; }
      if (x2 > 1) { x2 = 1; x1 = (u - mb) / ma; } else if (x2 < -1) { x2 = -1; x1 = (u + mb) / ma; }
      x1 = clip(x1, -1, 1);
      for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k] * x1; vec2[k] = pos2[k] + axis2[k] * x2; }
      return sphereSphere(con, margin, vec1, size1[0], vec2, size2[0]);
    }
    // parallel axes: up to two contacts
    int n = 0;
    double x2 = clip((v - mb) / mc, -1, 1);
    for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k]; vec2[k] = pos2[k] + axis2[k] * x2; }
    n += sphereSphere(con + n, margin, vec1, size1[0], vec2, size2[0]);
    x2 = clip((v + mb) / mc, -1, 1);
    for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] - axis1[k]; vec2[k] = pos2[k] + axis2[k] * x2; }
    n += sphereSphere(con + n, margin, vec1, size1[0], vec2, size2[0]);
    if (n >= 2) return n;
    double x1 = clip((u - mb) / ma, -1, 1);
    for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k] * x1; vec2[k] = pos2[k] + axis2[k]; }
    n += sphereSphere(con + n, margin, vec1, size1[0], vec2, size2[0]);
    if (n >= 2) return n;
    x1 = clip((u + mb) / ma, -1, 1);
    for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k] * x1; vec2[k] = pos2[k] - axis2[k]; }
    n += sphereSphere(con + n, margin, vec1, size1[0], vec2, size2[0]);
    return n;
  }
  return 0;
}

void makeFrame(double* frame) {
  normalize3(frame);
  if (std::sqrt(dot3(frame + 3, frame + 3)) < 0.5) {
    frame[3] = frame[4] = frame[5] = 0;
    if (frame[1] < 0.5 && frame[1] > -0.5) frame[4] = 1; else frame[5] = 1;
  }
  double t = dot3(frame, frame + 3);
  for (int k = 0; k < 3; k++) frame[3 + k] -= t * frame[k];
  normalize3(frame + 3);
  cross3(frame + 6, frame, frame + 3);
}

void collision(const Model* m, Data* d) {
  d->ncon = 0;
  if (disabled(m, OX_DSBL_CONTACT) || disabled(m, OX_DSBL_CONSTRAINT)) return;
  for (int p = 0; p < m->npair; p++) {
    RawContact rc[8];
    int n = collidePair(m, d, p, rc);
    for (int c = 0; c < n; c++) {
      makeFrame(rc[c].frame);
      int k = d->ncon++;
      d->con_dist[k] = rc[c].dist;
      std::memcpy(&d->con_pos[3 * k], rc[c].pos, 3 * sizeof(double));
      std::memcpy(&d->con_frame[9 * k], rc[c].frame, 9 * sizeof(double));
      d->con_pair[k] = p;
    }
  }
}

// ---------------------------------------------------------------- A.6 constraint assembly
// weld orientation rows: residual[3] = ts * imag(conj(q2) q1 qrel) and the 3x3 map G with d residual / dt = G (w1 - w2):
// column c of G = ts/2 * imag(conj(q2) [0, e_c] q1 qrel)
void weldRotMap(const double* q1, const double* q2, const double* qrel, double ts, double* G, double* residual) {
  double quat[4], q2c[4] = {q2[0], -q2[1], -q2[2], -q2[3]}, e[4], t1[4], t2[4];
  mulQuat(quat, q1, qrel);
  mulQuat(e, q2c, quat);
  for (int k = 0; k < 3; k++) residual[k] = ts * e[1 + k];
  for (int c = 0; c < 3; c++) {
    double ax[4] = {0, 0, 0, 0};
    ax[1 + c] = 1;
    mulQuat(t1, q2c, ax);
    mulQuat(t2, t1, quat);
    for (int k = 0; k < 3; k++) G[3 * k + c] = 0.5 * ts * t2[1 + k];
  }
}
// translational Jacobian of a world point attached to `body` (mj_jac, jacp rows only), accumulated with sign
void addJacP(const Model* m, const Data* d, int body, const double* point, double sign, double* jacp /*3 x nv*/) {
  int nv = m->nv;
  double offset[3];
  for (int k = 0; k < 3; k++) offset[k] = point[k] - d->subtree_com[3 * m->body_rootid[body] + k];
  while (body && m->body_dofnum[body] == 0) body = m->body_parentid[body];
  if (!body) return;
  for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i]) {
    const double* cd = &d->cdof[6 * i];
    double tmp[3];
    cross3(tmp, cd, offset);
    for (int k = 0; k < 3; k++) jacp[k * nv + i] += sign * (cd[3 + k] + tmp[k]);
  }
}

// rotational Jacobian of `body` (mj_jac, jacr rows): the angular part of cdof along the body's dof chain
void addJacR(const Model* m, const Data* d, int body, double sign, double* jacr /*3 x nv*/) {
  int nv = m->nv;
  while (body && m->body_dofnum[body] == 0) body = m->body_parentid[body];
  if (!body) return;
  for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i])
    for (int k = 0; k < 3; k++) jacr[k * nv + i] += sign * d->cdof[6 * i + k];
}

void getImpedance(const double* solimp, double pos, double margin, double* imp) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (dmin == dmax || width <= OX_MINVAL) { *imp = 0.5 * (dmin + dmax); return; }
  double x = std::fabs(pos - margin) / width;
  if (x >= 1) { *imp = dmax; return; }
  if (x <= 0) { *imp = dmin; return; }
  double y;
  if (power == 1) y = x;
  else if (power == 2) y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
  else if (x <= mid) y = std::pow(x, power) / std::pow(mid, power - 1);
  else y = 1 - std::pow(1 - x, power) / std::pow(1 - mid, power - 1);
  *imp = dmin + y * (dmax - dmin);
}

void addRow(const Model* m, Data* d, const double* jrow, double pos, double margin, double diagApprox, const double* solref,
            const double* solimp, int type, int id) {
  int r = d->nefc++, nv = m->nv;
  std::memcpy(&d->efc_J[(size_t)r * nv], jrow, nv * sizeof(double));
  d->efc_pos[r] = pos;
  d->efc_margin[r] = margin;
  d->efc_type[r] = type;
  d->efc_id[r] = id;
  d->efc_floss[r] = 0;
  d->efc_diagApprox[r] = diagApprox;
  double imp;
  getImpedance(solimp, pos, margin, &imp);
  imp = clip(imp, OX_MINIMP, OX_MAXIMP);
  d->efc_R[r] = std::max(OX_MINVAL, (1 - imp) * diagApprox / imp);
  // stiffness / damping of the reference acceleration
  double dmax = solimp[1], K, Bd;
  if (solref[0] > 0) {
    double tc = solref[0], dr = solref[1];
    if (!disabled(m, OX_DSBL_REFSAFE)) tc = std::max(tc, 2 * m->timestep);
    K = 1 / std::max(OX_MINVAL, dmax * dmax * tc * tc * dr * dr);
    Bd = 2 / std::max(OX_MINVAL, dmax * tc);
  } else {
    K = -solref[0] / std::max(OX_MINVAL, dmax * dmax);
    Bd = -solref[1] / std::max(OX_MINVAL, dmax);
  }
  double vel = 0;
  for (int i = 0; i < nv; i++) vel += jrow[i] * d->qvel[i];
  d->efc_vel[r] = vel;
  d->efc_aref[r] = -Bd * vel - K * imp * (pos - margin);
}

// mj_tendon: length and Jacobian row of every tendon. Fixed: length = sum of coef * joint coordinate, J = the coefficients.
// Spatial (sites only): length = sum of segment lengths, J = sum dir' (Jp(site k+1) - Jp(site k)).
void addJacP(const Model* m, const Data* d, int body, const double* point, double sign, double* jacp);
void tendonLength(const Model* m, Data* d) {
  const int nv = m->nv;
  std::fill(d->ten_J.begin(), d->ten_J.end(), 0.0);
  std::vector<double> jacp(3 * nv);
  for (int i = 0; i < m->ntendon; i++) {
    double L = 0;
    double* J = &d->ten_J[(size_t)i * nv];
    const int adr = m->tendon_adr[i], num = m->tendon_num[i];
    if (m->tendon_type[i] == OX_TEN_FIXED) {
      for (int w = adr; w < adr + num; w++) {
        L += m->wrap_prm[w] * d->qpos[m->jnt_qposadr[m->wrap_objid[w]]];
        J[m->jnt_dofadr[m->wrap_objid[w]]] += m->wrap_prm[w];
      }
    } else {
      for (int w = adr; w + 1 < adr + num; w++) {
        const int sa = m->wrap_objid[w], sb = m->wrap_objid[w + 1];
        const double *pa = &d->site_xpos[3 * sa], *pb = &d->site_xpos[3 * sb];
        double dir[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
        const double len = std::sqrt(dot3(dir, dir));
        L += len;
        if (len < OX_MINVAL) continue;
        for (double& v : dir) v /= len;
        std::fill(jacp.begin(), jacp.end(), 0.0);
        addJacP(m, d, m->site_bodyid[sb], pb, +1, jacp.data());
        addJacP(m, d, m->site_bodyid[sa], pa, -1, jacp.data());
        for (int k = 0; k < nv; k++) J[k] += dir[0] * jacp[k] + dir[1] * jacp[nv + k] + dir[2] * jacp[2 * nv + k];
      }
    }
    d->ten_length[i] = L;
  }
}
double tendonVelocity(const Model* m, const Data* d, int i) {
  double v = 0;
  for (int k = 0; k < m->nv; k++) v += d->ten_J[(size_t)i * m->nv + k] * d->qvel[k];
  return v;
}

void makeConstraint(const Model* m, Data* d) {
  int nv = m->nv;
  tendonLength(m, d);   // position-stage quantity (mj_tendon); idempotent, so passive() may have computed it already
  d->nefc = 0;
  d->ne = 0;
  d->nf = 0;
  if (disabled(m, OX_DSBL_CONSTRAINT)) return;
  std::vector<double> jrow(nv), jacp(3 * nv), jac(3 * nv);
  // equality constraints first (mj_instantiateEquality): residual as pos, margin 0, always active in the solver
  if (!disabled(m, OX_DSBL_EQUALITY))
    for (int i = 0; i < m->neq; i++) {
      if (d->eq_active[i] == 0) continue;
      const double* data = m->eq_data + 11 * i;
      if (m->eq_type[i] == OX_EQ_CONNECT || m->eq_type[i] == OX_EQ_WELD) {
        // anchors of the two bodies in world coordinates must coincide: residual p1 - p2, Jacobian Jp(b1, p1) - Jp(b2, p2)
        int b1 = m->eq_obj1id[i], b2 = m->eq_obj2id[i];
        double p1[3], p2[3], w[3];
        mulMatVec3(w, &d->xmat[9 * b1], data);
        for (int k = 0; k < 3; k++) p1[k] = d->xpos[3 * b1 + k] + w[k];
        mulMatVec3(w, &d->xmat[9 * b2], data + 3);
        for (int k = 0; k < 3; k++) p2[k] = d->xpos[3 * b2 + k] + w[k];
        std::fill(jacp.begin(), jacp.end(), 0.0);
        addJacP(m, d, b1, p1, +1, jacp.data());
        addJacP(m, d, b2, p2, -1, jacp.data());
        double diag = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
        for (int k = 0; k < 3; k++)
          addRow(m, d, &jacp[(size_t)k * nv], p1[k] - p2[k], 0.0, diag, m->eq_solref + 2 * i, m->eq_solimp + 5 * i, 3, i);
        if (m->eq_type[i] == OX_EQ_WELD) {
          // orientation: e = conj(q2) (q1 q_rel) is the identity when body2 sits at its welded orientation; residual =
          // torquescale * imag(e); d/dt imag(e) = 1/2 imag(conj(q2) [0, w1 - w2] q1 q_rel), linear in the relative angular velocity
          double G[9];
          weldRotMap(&d->xquat[4 * b1], &d->xquat[4 * b2], data + 6, data[10], G, jrow.data() /*scratch: residual in [0..2]*/);
          const double res[3] = {jrow[0], jrow[1], jrow[2]};
          std::vector<double> jacr(3 * nv, 0.0);
          addJacR(m, d, b1, +1, jacr.data());
          addJacR(m, d, b2, -1, jacr.data());
          const double rdiag = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
          for (int k = 0; k < 3; k++) {
            for (int c = 0; c < nv; c++) jrow[c] = G[3 * k] * jacr[c] + G[3 * k + 1] * jacr[nv + c] + G[3 * k + 2] * jacr[2 * nv + c];
            addRow(m, d, jrow.data(), res[k], 0.0, rdiag, m->eq_solref + 2 * i, m->eq_solimp + 5 * i, 3, i);
          }
        }
      } else {
        // joint coupling: q1 - q1_0 = c0 + c1 dq2 + ... + c4 dq2^4 with dq2 = q2 - q2_0 (only c0 without a second joint)
        int j1 = m->eq_obj1id[i], j2 = m->eq_obj2id[i];
        std::fill(jrow.begin(), jrow.end(), 0.0);
        double pos = d->qpos[m->jnt_qposadr[j1]] - m->qpos0[m->jnt_qposadr[j1]];
        double diag = m->dof_invweight0[m->jnt_dofadr[j1]];
        jrow[m->jnt_dofadr[j1]] = 1;
        if (j2 >= 0) {
          double dq = d->qpos[m->jnt_qposadr[j2]] - m->qpos0[m->jnt_qposadr[j2]];
          pos -= data[0] + data[1] * dq + data[2] * dq * dq + data[3] * dq * dq * dq + data[4] * dq * dq * dq * dq;
          jrow[m->jnt_dofadr[j2]] = -(data[1] + 2 * data[2] * dq + 3 * data[3] * dq * dq + 4 * data[4] * dq * dq * dq);
          diag += m->dof_invweight0[m->jnt_dofadr[j2]];
        } else {
          pos -= data[0];
        }
        addRow(m, d, jrow.data(), pos, 0.0, diag, m->eq_solref + 2 * i, m->eq_solimp + 5 * i, 3, i);
      }
    }
  d->ne = d->nefc;
  // dry joint friction (mj_instantiateFriction): one row per dof with frictionloss > 0, J = e_dof, pos = margin = 0
  if (!disabled(m, OX_DSBL_FRICTIONLOSS))
    for (int i = 0; i < nv; i++) {
      if (!(m->dof_frictionloss[i] > 0)) continue;
      std::fill(jrow.begin(), jrow.end(), 0.0);
      jrow[i] = 1;
      addRow(m, d, jrow.data(), 0.0, 0.0, m->dof_invweight0[i], m->dof_solref_fri + 2 * i, m->dof_solimp_fri + 5 * i, 5, i);
      d->efc_floss[d->nefc - 1] = m->dof_frictionloss[i];
    }
  d->nf = d->nefc - d->ne;
  // joint limits
  if (!disabled(m, OX_DSBL_LIMIT))
    for (int j = 0; j < m->njnt; j++) {
      if (!m->jnt_limited[j]) continue;
      int jt = m->jnt_type[j];
      if (jt == OX_JNT_BALL) {
        // mj_instantiateLimit, ball: the rotation vector of the joint quaternion (mju_quat2Vel, dt = 1) gives angle and axis;
        // dist = max(range) - angle, one row with J = -axis on the joint's dofs
        double q[4], axis[3];
        std::memcpy(q, &d->qpos[m->jnt_qposadr[j]], sizeof q);
        normalize4(q);
        double sn = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), ang = 2 * std::atan2(sn, q[0]);
        if (ang > OX_PI_D) ang -= 2 * OX_PI_D;
        for (int k = 0; k < 3; k++) axis[k] = sn > OX_MINVAL ? (ang < 0 ? -1.0 : 1.0) * q[1 + k] / sn : 0.0;
        double dist = std::max(m->jnt_range[2 * j], m->jnt_range[2 * j + 1]) - std::fabs(ang), margin = m->jnt_margin[j];
        if (dist < margin) {
          std::fill(jrow.begin(), jrow.end(), 0.0);
          for (int k = 0; k < 3; k++) jrow[m->jnt_dofadr[j] + k] = -axis[k];
          addRow(m, d, jrow.data(), dist, margin, m->dof_invweight0[m->jnt_dofadr[j]], m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j, 0, j);
        }
        continue;
      }
      if (jt != OX_JNT_SLIDE && jt != OX_JNT_HINGE) continue;
      double value = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
      for (int side = -1; side <= 1; side += 2) {
        double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - value);
        if (dist < margin) {
          std::fill(jrow.begin(), jrow.end(), 0.0);
          jrow[m->jnt_dofadr[j]] = -side;
          addRow(m, d, jrow.data(), dist, margin, m->dof_invweight0[m->jnt_dofadr[j]], m->jnt_solref + 2 * j,
                 m->jnt_solimp + 5 * j, 0, j);
        }
      }
    }
  // tendon limits (after the joint limits, mj_instantiateLimit order)
  if (!disabled(m, OX_DSBL_LIMIT))
    for (int i = 0; i < m->ntendon; i++) {
      if (!m->tendon_limited[i]) continue;
      const double value = d->ten_length[i], margin = m->tendon_margin[i];
      for (int side = -1; side <= 1; side += 2) {
        double dist = side * (m->tendon_range[2 * i + (side + 1) / 2] - value);
        if (dist < margin) {
          std::fill(jrow.begin(), jrow.end(), 0.0);
          for (int k = 0; k < nv; k++) jrow[k] = -side * d->ten_J[(size_t)i * nv + k];
          addRow(m, d, jrow.data(), dist, margin, m->tendon_invweight0[i], m->tendon_solref_lim + 2 * i, m->tendon_solimp_lim + 5 * i, 4, i);
        }
      }
    }
  // contacts (pyramidal cone)
  for (int c = 0; c < d->ncon; c++) {
    int p = d->con_pair[c];
    double includemargin = m->pair_margin[p] - m->pair_gap[p];
    if (d->con_dist[c] >= includemargin) continue;
    int b1 = m->geom_bodyid[m->pair_geom1[p]], b2 = m->geom_bodyid[m->pair_geom2[p]];
    std::fill(jacp.begin(), jacp.end(), 0.0);
    addJacP(m, d, b2, &d->con_pos[3 * c], +1, jacp.data());
    addJacP(m, d, b1, &d->con_pos[3 * c], -1, jacp.data());
    const double* frame = &d->con_frame[9 * c];
    for (int r = 0; r < 3; r++)
      for (int i = 0; i < nv; i++)
        jac[r * nv + i] = frame[3 * r] * jacp[i] + frame[3 * r + 1] * jacp[nv + i] + frame[3 * r + 2] * jacp[2 * nv + i];
    double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2], rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    if (m->pair_dim[p] > 3) {   // condim 4 / 6: rows 3..5 = relative angular velocity about the normal (torsion) and the tangents (rolling)
      jac.resize(6 * nv);
      std::vector<double> jacr(3 * nv, 0.0);
      addJacR(m, d, b2, +1, jacr.data());
      addJacR(m, d, b1, -1, jacr.data());
      for (int r = 0; r < 3; r++)
        for (int i = 0; i < nv; i++)
          jac[(3 + r) * nv + i] = frame[3 * r] * jacr[i] + frame[3 * r + 1] * jacr[nv + i] + frame[3 * r + 2] * jacr[2 * nv + i];
    }
    const double* fri = m->pair_friction + 5 * p;
    int dim = m->pair_dim[p];
    if (dim == 1) {
      addRow(m, d, jac.data(), d->con_dist[c], includemargin, tran, m->pair_solref + 2 * p, m->pair_solimp + 5 * p, 1, c);
    } else if (m->cone == OX_CONE_ELLIPTIC) {
      // elliptic cone: rows = normal, then one row per friction direction (tangents; torsion / rolling for condim 4 / 6). Only the
      // normal row has a position term; the friction rows get R_j = R_n mu^2 / friction_j^2 with the regularised friction
      // mu = friction_1 / sqrt(impratio), which makes the cone circular in the scaled coordinates of the cost (ellipticCost).
      const int first = d->nefc;
      addRow(m, d, jac.data(), d->con_dist[c], includemargin, tran, m->pair_solref + 2 * p, m->pair_solimp + 5 * p, 6, c);
      const double mu = fri[0] * std::sqrt(1 / m->impratio);
      for (int k = 1; k < dim; k++) {
        addRow(m, d, &jac[(size_t)k * nv], 0.0, 0.0, tran, m->pair_solref + 2 * p, m->pair_solimp + 5 * p, 6, c);
        d->efc_R[d->nefc - 1] = d->efc_R[first] * mu * mu / (fri[k - 1] * fri[k - 1]);
      }
    } else {
      int first = d->nefc;
      for (int k = 1; k < dim; k++)
        for (int s = 0; s < 2; s++) {
          double sg = s ? -1.0 : 1.0;
          for (int i = 0; i < nv; i++) jrow[i] = jac[i] + sg * fri[k - 1] * jac[k * nv + i];
          addRow(m, d, jrow.data(), d->con_dist[c], includemargin, tran + fri[k - 1] * fri[k - 1] * (k < 3 ? tran : rot), m->pair_solref + 2 * p,
                 m->pair_solimp + 5 * p, 2, c);
        }
      // pyramidal regularisation: every edge gets R = 2 mu^2 R(first edge), mu = friction[0]/sqrt(impratio)
      double mu = fri[0] * std::sqrt(1 / m->impratio);
      double Rpy = 2 * mu * mu * d->efc_R[first];
      for (int r = first; r < d->nefc; r++) d->efc_R[r] = Rpy;
    }
  }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1 / d->efc_R[r];
}

// ---------------------------------------------------------------- A.7 velocity
void comVel(const Model* m, Data* d) {
  std::fill(d->cvel.begin(), d->cvel.begin() + 6, 0.0);
  for (int i = 1; i < m->nbody; i++) {
    int bda = m->body_dofadr[i];
    double cvel[6], cdofdot[36], tmp[6];
    std::memcpy(cvel, &d->cvel[6 * m->body_parentid[i]], sizeof cvel);
    int dofnum = m->body_dofnum[i];
    for (int j = 0; j < dofnum; j++) {
      int jt = m->jnt_type[m->dof_jntid[bda + j]];
      if (jt == OX_JNT_FREE) {
        std::fill(cdofdot, cdofdot + 18, 0.0);
        for (int k = 0; k < 3; k++)
          for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * (bda + k) + c] * d->qvel[bda + k];
        j += 3;
      }
      if (jt == OX_JNT_FREE || jt == OX_JNT_BALL) {
        for (int k = 0; k < 3; k++) crossMotion(cdofdot + 6 * (j + k), cvel, &d->cdof[6 * (bda + j + k)]);
        for (int k = 0; k < 3; k++)
          for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * (bda + j + k) + c] * d->qvel[bda + j + k];
        j += 2;
      } else {
        crossMotion(cdofdot + 6 * j, cvel, &d->cdof[6 * (bda + j)]);
        for (int c = 0; c < 6; c++) tmp[c] = d->cdof[6 * (bda + j) + c] * d->qvel[bda + j];
        for (int c = 0; c < 6; c++) cvel[c] += tmp[c];
      }
    }
    std::memcpy(&d->cvel[6 * i], cvel, sizeof cvel);
    if (dofnum) std::memcpy(&d->cdof_dot[6 * bda], cdofdot, 6 * dofnum * sizeof(double));
  }
}

void passive(const Model* m, Data* d) {
  std::fill(d->qfrc_passive.begin(), d->qfrc_passive.end(), 0.0);
  tendonLength(m, d);
  if (disabled(m, OX_DSBL_PASSIVE)) return;
  for (int j = 0; j < m->njnt; j++) {
    double k = m->jnt_stiffness[j];
    if (k == 0) continue;
    int pa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case OX_JNT_FREE:
        for (int c = 0; c < 3; c++) d->qfrc_passive[da + c] -= k * (d->qpos[pa + c] - m->qpos_spring[pa + c]);
        pa += 3; da += 3;
        // fallthrough
      case OX_JNT_BALL: {
        double q[4], dif[3];
        std::memcpy(q, &d->qpos[pa], sizeof q);
        normalize4(q);
        subQuat(dif, q, m->qpos_spring + pa);
        for (int c = 0; c < 3; c++) d->qfrc_passive[da + c] -= k * dif[c];
        break;
      }
      default: d->qfrc_passive[da] -= k * (d->qpos[pa] - m->qpos_spring[pa]);
    }
  }
  for (int i = 0; i < m->nv; i++) d->qfrc_passive[i] -= m->dof_damping[i] * d->qvel[i];
  // tendon spring (dead band [lengthspring0, lengthspring1]) and damper, mapped through J' = the coefficients
  for (int i = 0; i < m->ntendon; i++) {
    const double L = d->ten_length[i], lo = m->tendon_lengthspring[2 * i], hi = m->tendon_lengthspring[2 * i + 1];
    double f = 0;
    if (L > hi) f = m->tendon_stiffness[i] * (hi - L);
    else if (L < lo) f = m->tendon_stiffness[i] * (lo - L);
    f -= m->tendon_damping[i] * tendonVelocity(m, d, i);
    if (f != 0)
      for (int k = 0; k < m->nv; k++) d->qfrc_passive[k] += d->ten_J[(size_t)i * m->nv + k] * f;
  }
  // body gravcomp (mj_passive, qfrc_gravcomp): -gravity * mass * gravcomp applied at xipos
  if (m->ngravcomp && !disabled(m, OX_DSBL_GRAVITY))
    for (int b = 1; b < m->nbody; b++) {
      if (m->body_gravcomp[b] == 0) continue;
      double offset[3], frc[3];
      for (int k = 0; k < 3; k++) offset[k] = d->xipos[3 * b + k] - d->subtree_com[3 * m->body_rootid[b] + k];
      for (int k = 0; k < 3; k++) frc[k] = -m->gravity[k] * m->body_mass[b] * m->body_gravcomp[b];
      int body = b;
      while (body && m->body_dofnum[body] == 0) body = m->body_parentid[body];
      if (!body) continue;
      for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i]) {
        const double* cd = &d->cdof[6 * i];
        double jp[3];
        cross3(jp, cd, offset);
        for (int k = 0; k < 3; k++) jp[k] += cd[3 + k];
        d->qfrc_passive[i] += dot3(jp, frc);
      }
    }
  // mj_inertiaBoxFluidModel (engine_passive.c): viscous and quadratic drag on the equivalent inertia box of every body. Done from
  // mjOption density / viscosity / wind and body_inertia as MuJoCo does at run time (the product uses a table compiled from them).
  if (m->viscosity > 0 || m->density > 0)
    for (int b = 1; b < m->nbody; b++) {
      const double mass = m->body_mass[b], *I = &m->body_inertia[3 * b];
      if (mass < OX_MINVAL) continue;
      const double* R = &d->ximat[9 * b];
      double offset[3], vel[3], lvel[6], lfrc[6] = {0, 0, 0, 0, 0, 0}, box[3];
      for (int k = 0; k < 3; k++) offset[k] = d->xipos[3 * b + k] - d->subtree_com[3 * m->body_rootid[b] + k];
      // mj_objectVelocity(mjOBJ_BODY, flg_local = 1): velocity at xipos, rotated into the inertial frame; minus the wind
      cross3(vel, &d->cvel[6 * b], offset);
      for (int k = 0; k < 3; k++) vel[k] += d->cvel[6 * b + 3 + k] - m->wind[k];
      for (int k = 0; k < 3; k++) {
        lvel[k] = R[k] * d->cvel[6 * b] + R[3 + k] * d->cvel[6 * b + 1] + R[6 + k] * d->cvel[6 * b + 2];
        lvel[3 + k] = R[k] * vel[0] + R[3 + k] * vel[1] + R[6 + k] * vel[2];
      }
      box[0] = std::sqrt(std::max(OX_MINVAL, I[1] + I[2] - I[0]) / mass * 6.0);
      box[1] = std::sqrt(std::max(OX_MINVAL, I[0] + I[2] - I[1]) / mass * 6.0);
      box[2] = std::sqrt(std::max(OX_MINVAL, I[0] + I[1] - I[2]) / mass * 6.0);
      if (m->viscosity > 0) {
        const double diam = (box[0] + box[1] + box[2]) / 3.0;
        for (int k = 0; k < 3; k++) lfrc[k] = -OX_PI_D * diam * diam * diam * m->viscosity * lvel[k];
        for (int k = 0; k < 3; k++) lfrc[3 + k] = -3.0 * OX_PI_D * diam * m->viscosity * lvel[3 + k];
      }
      if (m->density > 0) {
        lfrc[3] -= 0.5 * m->density * box[1] * box[2] * std::fabs(lvel[3]) * lvel[3];
        lfrc[4] -= 0.5 * m->density * box[0] * box[2] * std::fabs(lvel[4]) * lvel[4];
        lfrc[5] -= 0.5 * m->density * box[0] * box[1] * std::fabs(lvel[5]) * lvel[5];
        lfrc[0] -= m->density * box[0] * (std::pow(box[1], 4) + std::pow(box[2], 4)) * std::fabs(lvel[0]) * lvel[0] / 64.0;
        lfrc[1] -= m->density * box[1] * (std::pow(box[0], 4) + std::pow(box[2], 4)) * std::fabs(lvel[1]) * lvel[1] / 64.0;
        lfrc[2] -= m->density * box[2] * (std::pow(box[0], 4) + std::pow(box[1], 4)) * std::fabs(lvel[2]) * lvel[2] / 64.0;
      }
      double trq[3], frc[3];
      for (int k = 0; k < 3; k++) {
        trq[k] = R[3 * k] * lfrc[0] + R[3 * k + 1] * lfrc[1] + R[3 * k + 2] * lfrc[2];
        frc[k] = R[3 * k] * lfrc[3] + R[3 * k + 1] * lfrc[4] + R[3 * k + 2] * lfrc[5];
      }
      // mj_applyFT at xipos
      int body = b;
      while (body && m->body_dofnum[body] == 0) body = m->body_parentid[body];
      if (!body) continue;
      for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i]) {
        const double* cd = &d->cdof[6 * i];
        double jp[3];
        cross3(jp, cd, offset);
        for (int k = 0; k < 3; k++) jp[k] += cd[3 + k];
        d->qfrc_passive[i] += dot3(jp, frc) + dot3(cd, trq);
      }
    }
}

// ---------------------------------------------------------------- A.8 bias forces (RNE, flg_acc = 0)
void rne(const Model* m, Data* d) {
  int nb = m->nbody;
  std::vector<double> cacc(6 * nb, 0.0), cfrc(6 * nb, 0.0);
  if (!disabled(m, OX_DSBL_GRAVITY)) for (int k = 0; k < 3; k++) cacc[3 + k] = -m->gravity[k];
  for (int i = 1; i < nb; i++) {
    int bda = m->body_dofadr[i];
    double tmp[6] = {0, 0, 0, 0, 0, 0}, tmp1[6];
    for (int j = 0; j < m->body_dofnum[i]; j++)
      for (int c = 0; c < 6; c++) tmp[c] += d->cdof_dot[6 * (bda + j) + c] * d->qvel[bda + j];
    for (int c = 0; c < 6; c++) cacc[6 * i + c] = cacc[6 * m->body_parentid[i] + c] + tmp[c];
    mulInertVec(&cfrc[6 * i], &d->cinert[10 * i], &cacc[6 * i]);
    mulInertVec(tmp, &d->cinert[10 * i], &d->cvel[6 * i]);
    crossForce(tmp1, &d->cvel[6 * i], tmp);
    for (int c = 0; c < 6; c++) cfrc[6 * i + c] += tmp1[c];
  }
  for (int i = nb - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p) for (int c = 0; c < 6; c++) cfrc[6 * p + c] += cfrc[6 * i + c];
  }
  for (int i = 0; i < m->nv; i++) d->qfrc_bias[i] = dot6(&d->cdof[6 * i], &cfrc[6 * m->dof_bodyid[i]]);
}

// ---------------------------------------------------------------- A.9 actuation
void actuation(const Model* m, Data* d) {
  std::fill(d->qfrc_actuator.begin(), d->qfrc_actuator.end(), 0.0);
  std::fill(d->actuator_force.begin(), d->actuator_force.end(), 0.0);
  if (disabled(m, OX_DSBL_ACTUATION)) return;
  for (int i = 0; i < m->nu; i++) {
    // mj_transmission: joint coordinate, or the length of a fixed tendon (moment = gear * tendon Jacobian)
    const bool ten = m->actuator_trntype[i] == OX_TRN_TENDON;
    const int j = m->actuator_trnid[i];
    double gear = m->actuator_gear[i];
    double length = gear * (ten ? d->ten_length[j] : d->qpos[m->jnt_qposadr[j]]);
    double velocity = gear * (ten ? tendonVelocity(m, d, j) : d->qvel[m->jnt_dofadr[j]]);
    double ctrl = d->ctrl[i];
    if (m->actuator_ctrllimited[i] && !disabled(m, OX_DSBL_CLAMPCTRL))
      ctrl = clip(ctrl, m->actuator_ctrlrange[2 * i], m->actuator_ctrlrange[2 * i + 1]);
    // mj_fwdActuation, stateful actuators: act_dot from the control, force from the activation
    //   integrator: act_dot = ctrl;  filter / filterexact: act_dot = (ctrl - act) / max(mjMINVAL, dynprm[0])
    if (m->actuator_dyntype[i] != OX_DYN_NONE) {
      int aa = m->actuator_actadr[i];
      d->act_dot[aa] = m->actuator_dyntype[i] == OX_DYN_INTEGRATOR ? ctrl : (ctrl - d->act[aa]) / std::max(OX_MINVAL, m->actuator_dynprm[3 * i]);
      ctrl = d->act[aa];
    }
    const double* gp = m->actuator_gainprm + 3 * i;
    const double* bp = m->actuator_biasprm + 3 * i;
    double gain = gp[0];
    if (m->actuator_gaintype[i] == OX_GAIN_AFFINE) gain += gp[1] * length + gp[2] * velocity;
    double bias = 0;
    if (m->actuator_biastype[i] == OX_BIAS_AFFINE) bias = bp[0] + bp[1] * length + bp[2] * velocity;
    double force = gain * ctrl + bias;
    if (m->actuator_forcelimited[i]) force = clip(force, m->actuator_forcerange[2 * i], m->actuator_forcerange[2 * i + 1]);
    d->actuator_force[i] = force;
    if (ten)
      for (int k = 0; k < m->nv; k++) d->qfrc_actuator[k] += gear * force * d->ten_J[(size_t)j * m->nv + k];
    else d->qfrc_actuator[m->jnt_dofadr[j]] += gear * force;
  }
}

// ---------------------------------------------------------------- A.10 smooth acceleration
void fwdAcceleration(const Model* m, Data* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++)
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
  // Cartesian forces: force at xipos, torque
  for (int b = 1; b < m->nbody; b++) {
    const double* f = &d->xfrc_applied[6 * b];
    if (f[0] == 0 && f[1] == 0 && f[2] == 0 && f[3] == 0 && f[4] == 0 && f[5] == 0) continue;
    double offset[3];
    for (int k = 0; k < 3; k++) offset[k] = d->xipos[3 * b + k] - d->subtree_com[3 * m->body_rootid[b] + k];
    int body = b;
    while (body && m->body_dofnum[body] == 0) body = m->body_parentid[body];
    if (!body) continue;
    for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i]) {
      const double* cd = &d->cdof[6 * i];
      double jp[3];
      cross3(jp, cd, offset);
      for (int k = 0; k < 3; k++) jp[k] += cd[3 + k];
      d->qfrc_smooth[i] += dot3(jp, f) + dot3(cd, f + 3);
    }
  }
  d->qacc_smooth = d->qfrc_smooth;
  solveLD(m, d->qLD.data(), d->qLDiagInv.data(), d->qacc_smooth.data());
}

// ---------------------------------------------------------------- elliptic cone (mj_constraintUpdate, mjCNSTR_CONTACT_ELLIPTIC)
// Cost of one elliptic contact as a function of x = J a - aref of its rows (normal first). In the scaled coordinates
// N = mu x_0, T = |(friction_j x_j)| the regularised cone is circular and the cost has three zones:
//   top    (N >= mu T)     : the constraint is satisfied, cost 0
//   bottom (mu N + T <= 0) : every row is an ordinary quadratic row, cost 1/2 sum D_r x_r^2
//   middle (the cone)      : cost 1/2 Dm (N - mu T)^2 with Dm = D_n / (mu^2 (1 + mu^2))
// force = -d cost / d x;  Hc = d^2 cost / d x^2 (dim x dim, only in the middle zone; bottom = diag(D), top = 0).
// Returns the zone (0 top, 1 bottom, 2 middle).
struct EllipticContact { int row, dim; double mu, fri[5]; };
int ellipticCost(const EllipticContact& c, const double* D, const double* x, double* cost, double* force, double* Hc) {
  const int dim = c.dim;
  double U[6] = {0, 0, 0, 0, 0, 0}, scl[6];
  scl[0] = c.mu;
  for (int j = 1; j < dim; j++) scl[j] = c.fri[j - 1];
  for (int j = 0; j < dim; j++) U[j] = x[j] * scl[j];
  double T2 = 0;
  for (int j = 1; j < dim; j++) T2 += U[j] * U[j];
  const double N = U[0], T = std::sqrt(T2), mu = c.mu;
  if (force) std::fill(force, force + dim, 0.0);
  if (Hc) std::fill(Hc, Hc + dim * dim, 0.0);
  *cost = 0;
  if (N >= mu * T || (T <= 0 && N >= 0)) return 0;
  if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
    for (int j = 0; j < dim; j++) {
      *cost += 0.5 * D[j] * x[j] * x[j];
      if (force) force[j] = -D[j] * x[j];
      if (Hc) Hc[j * dim + j] = D[j];
    }
    return 1;
  }
  const double Dm = D[0] / (mu * mu * (1 + mu * mu)), NT = N - mu * T;
  *cost = 0.5 * Dm * NT * NT;
  // gradient in U: g_0 = Dm NT, g_j = -Dm NT mu U_j / T
  if (force) {
    force[0] = -Dm * NT * scl[0];
    for (int j = 1; j < dim; j++) force[j] = Dm * NT * mu * U[j] / T * scl[j];
  }
  if (Hc) {
    double HU[36];
    HU[0] = Dm;
    for (int j = 1; j < dim; j++) HU[j] = HU[j * dim] = -Dm * mu * U[j] / T;
    for (int j = 1; j < dim; j++)
      for (int k = 1; k < dim; k++)
        HU[j * dim + k] = Dm * mu * mu * U[j] * U[k] / T2 - Dm * NT * mu * ((j == k ? 1.0 : 0.0) / T - U[j] * U[k] / (T2 * T));
    for (int j = 0; j < dim; j++)
      for (int k = 0; k < dim; k++) Hc[j * dim + k] = HU[j * dim + k] * scl[j] * scl[k];
  }
  return 2;
}

// ---------------------------------------------------------------- A.11 constraint solver (primal: Newton / CG)
struct Solver {
  const Model* m;
  Data* d;
  int nv, nefc;
  std::vector<double> Ma, Jaref, grad, Mgrad, search, Mv, Jv, H, gradold, Mgradold;
  std::vector<char> active;
  std::vector<EllipticContact> ell;   // elliptic contacts (their rows are skipped by the row loops and handled per contact)
  std::vector<char> isell;
  std::vector<double> ellH;           // per contact: 36 entries of the cone Hessian at the current point
  std::vector<int> ellZone;
  double cost = 0, gauss = 0;
  double quadGauss[3];

  Solver(const Model* m_, Data* d_) : m(m_), d(d_), nv(m_->nv), nefc(d_->nefc) {
    Ma.resize(nv); Jaref.resize(nefc); grad.resize(nv); Mgrad.resize(nv); search.resize(nv); Mv.resize(nv); Jv.resize(nefc);
    H.resize((size_t)nv * nv); gradold.resize(nv); Mgradold.resize(nv); active.resize(nefc);
    isell.assign(nefc, 0);
    for (int r = 0; r < nefc; r++)
      if (d->efc_type[r] == 6 && (r == 0 || d->efc_type[r - 1] != 6 || d->efc_id[r - 1] != d->efc_id[r])) {
        const int p = d->con_pair[d->efc_id[r]];
        EllipticContact c;
        c.row = r; c.dim = m->pair_dim[p]; c.mu = m->pair_friction[5 * p] * std::sqrt(1 / m->impratio);
        for (int k = 0; k < 5; k++) c.fri[k] = m->pair_friction[5 * p + k];
        ell.push_back(c);
        for (int k = 0; k < c.dim; k++) isell[r + k] = 1;
      }
    ellH.assign(36 * ell.size(), 0.0); ellZone.assign(ell.size(), 0);
  }
  // efc_force, active set, cost, qfrc_constraint at the current Jaref / Ma / qacc
  void updateConstraint() {
    double c = 0;
    for (size_t k = 0; k < ell.size(); k++) {
      double ck;
      ellZone[k] = ellipticCost(ell[k], &d->efc_D[ell[k].row], &Jaref[ell[k].row], &ck, &d->efc_force[ell[k].row], &ellH[36 * k]);
      c += ck;
      for (int j = 0; j < ell[k].dim; j++) active[ell[k].row + j] = 0;   // their curvature enters through ellH
    }
    for (int r = 0; r < nefc; r++) {
      if (isell[r]) continue;
      const double fl = d->efc_floss[r];
      if (fl > 0 && std::fabs(Jaref[r]) >= fl / d->efc_D[r]) {   // dry friction, linear zone: |force| pinned at frictionloss
        const double rf = fl / d->efc_D[r], sgn = Jaref[r] > 0 ? 1.0 : -1.0;
        active[r] = 0;
        d->efc_force[r] = -sgn * fl;
        c += fl * (sgn * Jaref[r] - 0.5 * rf);
      } else if (Jaref[r] < 0 || r < d->ne || fl > 0) {   // equality rows and the quadratic zone of friction rows: two-sided
        active[r] = 1;
        d->efc_force[r] = -d->efc_D[r] * Jaref[r];
        c += 0.5 * d->efc_D[r] * Jaref[r] * Jaref[r];
      } else {
        active[r] = 0;
        d->efc_force[r] = 0;
      }
    }
    for (int i = 0; i < nv; i++) {
      double f = 0;
      for (int r = 0; r < nefc; r++) f += d->efc_J[(size_t)r * nv + i] * d->efc_force[r];
      d->qfrc_constraint[i] = f;
    }
    gauss = 0;
    for (int i = 0; i < nv; i++) gauss += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
    cost = c + gauss;
  }
  void updateGradient(bool newton) {
    for (int i = 0; i < nv; i++) grad[i] = Ma[i] - d->qfrc_smooth[i] - d->qfrc_constraint[i];
    if (!newton) {
      Mgrad = grad;
      solveLD(m, d->qLD.data(), d->qLDiagInv.data(), Mgrad.data());
      return;
    }
    // H = M + J' diag(D active) J, dense, lower triangle; Cholesky; Mgrad = H^-1 grad
    std::fill(H.begin(), H.end(), 0.0);
    for (int i = 0; i < nv; i++) {
      int adr = m->dof_Madr[i];
      for (int j = i; j >= 0; j = m->dof_parentid[j]) H[(size_t)i * nv + j] = d->qM[adr++];
    }
    for (int r = 0; r < nefc; r++) {
      if (!active[r]) continue;
      const double* J = &d->efc_J[(size_t)r * nv];
      double D = d->efc_D[r];
      for (int i = 0; i < nv; i++) {
        if (J[i] == 0) continue;
        double s = D * J[i];
        for (int j = 0; j <= i; j++) H[(size_t)i * nv + j] += s * J[j];
      }
    }
    for (size_t k = 0; k < ell.size(); k++) {   // J_c' Hc J_c of every elliptic contact that is not in the top zone
      if (ellZone[k] == 0) continue;
      const int dim = ell[k].dim, r0 = ell[k].row;
      for (int a = 0; a < dim; a++)
        for (int b = 0; b < dim; b++) {
          const double h = ellH[36 * k + a * dim + b];
          if (h == 0) continue;
          const double *Ja = &d->efc_J[(size_t)(r0 + a) * nv], *Jb = &d->efc_J[(size_t)(r0 + b) * nv];
          for (int i = 0; i < nv; i++) {
            if (Ja[i] == 0) continue;
            const double sa = h * Ja[i];
            for (int j = 0; j <= i; j++) H[(size_t)i * nv + j] += sa * Jb[j];
          }
        }
    }
    for (int j = 0; j < nv; j++) {
      double s = H[(size_t)j * nv + j];
      for (int k = 0; k < j; k++) s -= H[(size_t)j * nv + k] * H[(size_t)j * nv + k];
      s = std::sqrt(std::max(s, OX_MINVAL));
      H[(size_t)j * nv + j] = s;
      for (int i = j + 1; i < nv; i++) {
        double v = H[(size_t)i * nv + j];
        for (int k = 0; k < j; k++) v -= H[(size_t)i * nv + k] * H[(size_t)j * nv + k];
        H[(size_t)i * nv + j] = v / s;
      }
    }
    for (int i = 0; i < nv; i++) {
      double v = grad[i];
      for (int k = 0; k < i; k++) v -= H[(size_t)i * nv + k] * Mgrad[k];
      Mgrad[i] = v / H[(size_t)i * nv + i];
    }
    for (int i = nv - 1; i >= 0; i--) {
      double v = Mgrad[i];
      for (int k = i + 1; k < nv; k++) v -= H[(size_t)k * nv + i] * Mgrad[k];
      Mgrad[i] = v / H[(size_t)i * nv + i];
    }
  }
  struct Pt { double alpha, cost, d0, d1, s0; };
  Pt eval(double a) const {
    Pt p;
    p.alpha = a;
    p.cost = a * a * quadGauss[2] + a * quadGauss[1] + quadGauss[0];
    p.d0 = 2 * a * quadGauss[2] + quadGauss[1];
    p.d1 = 2 * quadGauss[2];
    p.s0 = std::fabs(2 * a * quadGauss[2]) + std::fabs(quadGauss[1]);
    for (size_t k = 0; k < ell.size(); k++) {   // elliptic contacts: exact cost, slope and curvature along the search direction
      const int dim = ell[k].dim, r0 = ell[k].row;
      double x[6], f[6], Hc[36], ck;
      for (int j = 0; j < dim; j++) x[j] = Jaref[r0 + j] + a * Jv[r0 + j];
      ellipticCost(ell[k], &d->efc_D[r0], x, &ck, f, Hc);
      p.cost += ck;
      for (int j = 0; j < dim; j++) {
        p.d0 -= f[j] * Jv[r0 + j];
        p.s0 += std::fabs(f[j] * Jv[r0 + j]);
        for (int l = 0; l < dim; l++) p.d1 += Jv[r0 + j] * Hc[j * dim + l] * Jv[r0 + l];
      }
    }
    for (int r = 0; r < nefc; r++) {
      if (isell[r]) continue;
      double x = Jaref[r] + a * Jv[r];
      const double fl = d->efc_floss[r];
      if (fl > 0 && std::fabs(x) >= fl / d->efc_D[r]) {  // friction row in its linear zone: floss (|x| - R floss / 2)
        const double sgn = x > 0 ? 1.0 : -1.0;
        p.cost += fl * (sgn * x - 0.5 * fl / d->efc_D[r]);
        p.d0 += sgn * fl * Jv[r];
        p.s0 += std::fabs(fl * Jv[r]);
      } else if (x < 0 || r < d->ne || fl > 0) {  // active at alpha: 1/2 D x^2 and its derivatives in alpha
        const double Dx = d->efc_D[r] * x, Dj = d->efc_D[r] * Jv[r];
        p.cost += 0.5 * Dx * x;
        p.d0 += Dx * Jv[r];
        p.d1 += Dj * Jv[r];
        p.s0 += std::fabs(Dx * Jv[r]);
      }
    }
    if (p.d1 < OX_MINVAL) p.d1 = OX_MINVAL;
    return p;
  }
  // exact line search on the convex piecewise-quadratic phi(alpha): safeguarded Newton on phi'
  double lineSearch(int ls_iterations) {
    double snorm = 0;
    for (int i = 0; i < nv; i++) snorm += search[i] * search[i];
    snorm = std::sqrt(snorm);
    if (snorm < OX_MINVAL) return 0;
    double gtol = m->tolerance * m->ls_tolerance * snorm * m->meaninertia * std::max(1, nv);
    mulM(m, d->qM.data(), Mv.data(), search.data());
    for (int r = 0; r < nefc; r++) {
      double v = 0;
      for (int i = 0; i < nv; i++) v += d->efc_J[(size_t)r * nv + i] * search[i];
      Jv[r] = v;
    }
    quadGauss[0] = gauss;
    quadGauss[1] = 0;
    quadGauss[2] = 0;
    for (int i = 0; i < nv; i++) {
      quadGauss[1] += search[i] * (Ma[i] - d->qfrc_smooth[i]);
      quadGauss[2] += 0.5 * search[i] * Mv[i];
    }
    const double eps = 4 * 2.220446049250313e-16;
    Pt p0 = eval(0);
    if (!(p0.d0 < 0)) return 0;
    Pt lo = p0, hi = p0, cur = p0;
    bool have_hi = false;
    for (int it = 0; it < ls_iterations; it++) {
      double a = cur.alpha - cur.d0 / cur.d1;
      if (have_hi && !(a > lo.alpha && a < hi.alpha)) a = 0.5 * (lo.alpha + hi.alpha);
      if (std::fabs(a - cur.alpha) <= eps * std::fabs(a)) break;
      cur = eval(a);
      if (std::fabs(cur.d0) < gtol || std::fabs(cur.d0) <= 8 * eps * cur.s0) break;  // converged, or phi' below its own round-off
      if (cur.d0 < 0) lo = cur; else { hi = cur; have_hi = true; }
    }
    return cur.cost <= p0.cost ? cur.alpha : 0;
  }
  void solve(bool newton, int maxiter, int ls_iterations) {
    mulM(m, d->qM.data(), Ma.data(), d->qacc.data());
    for (int r = 0; r < nefc; r++) {
      double v = -d->efc_aref[r];
      for (int i = 0; i < nv; i++) v += d->efc_J[(size_t)r * nv + i] * d->qacc[i];
      Jaref[r] = v;
    }
    updateConstraint();
    updateGradient(newton);
    for (int i = 0; i < nv; i++) search[i] = -Mgrad[i];
    double scale = 1 / (m->meaninertia * std::max(1, nv));
    int iter = 0;
    double gn = 0;
    for (int i = 0; i < nv; i++) gn += grad[i] * grad[i];
    if (scale * std::sqrt(gn) < m->tolerance) maxiter = 0;
    while (iter < maxiter) {
      double alpha = lineSearch(ls_iterations);
      if (alpha == 0) break;
      for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; }
      for (int r = 0; r < nefc; r++) Jaref[r] += alpha * Jv[r];
      if (!newton) { gradold = grad; Mgradold = Mgrad; }
      double oldcost = cost;
      updateConstraint();
      updateGradient(newton);
      iter++;
      double improvement = scale * (oldcost - cost);
      gn = 0;
      for (int i = 0; i < nv; i++) gn += grad[i] * grad[i];
      double gradient = scale * std::sqrt(gn);
      if (improvement < m->tolerance || gradient < m->tolerance) break;
      // floating-point floor (ORACLE_DECISIONS.md #7): inert in fp64, kept so oracle and GPU run the same rule
      if (oldcost - cost <= 8 * 4 * 2.220446049250313e-16 * (std::fabs(oldcost) + std::fabs(cost))) break;
      if (newton) for (int i = 0; i < nv; i++) search[i] = -Mgrad[i];
      else {
        double num = 0, den = 0;
        for (int i = 0; i < nv; i++) { num += grad[i] * (Mgrad[i] - Mgradold[i]); den += gradold[i] * Mgradold[i]; }
        double beta = num / std::max(OX_MINVAL, den);
        if (beta < 0) beta = 0;
        for (int i = 0; i < nv; i++) search[i] = -Mgrad[i] + beta * search[i];
      }
    }
    d->solver_niter = iter;
  }
  // total cost of a candidate acceleration (warm-start selection)
  double costAt(const double* qacc) {
    std::vector<double> ma(nv);
    mulM(m, d->qM.data(), ma.data(), qacc);
    double c = 0;
    for (int i = 0; i < nv; i++) c += 0.5 * (ma[i] - d->qfrc_smooth[i]) * (qacc[i] - d->qacc_smooth[i]);
    std::vector<double> jar(nefc);
    for (int r = 0; r < nefc; r++) {
      double v = -d->efc_aref[r];
      for (int i = 0; i < nv; i++) v += d->efc_J[(size_t)r * nv + i] * qacc[i];
      jar[r] = v;
    }
    for (size_t k = 0; k < ell.size(); k++) {
      double ck;
      ellipticCost(ell[k], &d->efc_D[ell[k].row], &jar[ell[k].row], &ck, nullptr, nullptr);
      c += ck;
    }
    for (int r = 0; r < nefc; r++) {
      if (isell[r]) continue;
      double v = jar[r];
      const double fl = d->efc_floss[r];
      if (fl > 0 && std::fabs(v) >= fl / d->efc_D[r]) c += fl * (std::fabs(v) - 0.5 * fl / d->efc_D[r]);
      else if (v < 0 || r < d->ne || fl > 0) c += 0.5 * d->efc_D[r] * v * v;
    }
    return c;
  }
};

// ---------------------------------------------------------------- dual solvers: PGS (mj_solPGS) and the noslip post-pass (mj_solNoSlip)
// Both work on the constraint forces f with  A = J M^-1 J' (+ diag R for PGS),  b = J qacc_smooth - aref.  The matrix is never
// formed: the solver carries w = qacc_smooth + M^-1 J' f, so that row r's residual is J_r w - aref_r (+ R_r f_r), and a change
// of f_r by delta moves w by delta * M^-1 J_r' (one sparse L'DL solve) - O(nefc nv) memory instead of O(nefc^2).
struct DualSolver {
  const Model* m;
  Data* d;
  int nv, nefc;
  std::vector<double> w, u;
  DualSolver(const Model* m_, Data* d_) : m(m_), d(d_), nv(m_->nv), nefc(d_->nefc), w(m_->nv), u(m_->nv) {}
  void minvJt(int r, double scale_b = 0, int rb = -1) {   // u = M^-1 (J_r - scale_b * J_rb)'
    for (int i = 0; i < nv; i++) u[i] = d->efc_J[(size_t)r * nv + i] - (rb >= 0 ? scale_b * d->efc_J[(size_t)rb * nv + i] : 0.0);
    solveLD(m, d->qLD.data(), d->qLDiagInv.data(), u.data());
  }
  double rowDot(int r, const std::vector<double>& v) const {
    double s = 0;
    for (int i = 0; i < nv; i++) s += d->efc_J[(size_t)r * nv + i] * v[i];
    return s;
  }
  // w = qacc_smooth + M^-1 J' f  (also the final primal acceleration: mj dual2Primal); qfrc_constraint = J' f
  void primalFromForces() {
    for (int i = 0; i < nv; i++) {
      double f = 0;
      for (int r = 0; r < nefc; r++) f += d->efc_J[(size_t)r * nv + i] * d->efc_force[r];
      d->qfrc_constraint[i] = f;
    }
    w = d->qfrc_constraint;
    solveLD(m, d->qLD.data(), d->qLDiagInv.data(), w.data());
    for (int i = 0; i < nv; i++) w[i] += d->qacc_smooth[i];
  }
  void pgs(int maxiter) {
    const double scale = 1 / (m->meaninertia * std::max(1, nv));
    // warm start: forces implied by qacc_warmstart (constraint update), kept only if their dual cost beats f = 0 (cost 0)
    bool warm = false;
    if (!disabled(m, OX_DSBL_WARMSTART)) {
      for (int r = 0; r < nefc; r++) {
        double jar = rowDot(r, d->qacc_warmstart) - d->efc_aref[r];
        d->efc_force[r] = (jar < 0 || r < d->ne || d->efc_floss[r] > 0) ? -d->efc_D[r] * jar : 0.0;
        if (d->efc_floss[r] > 0) d->efc_force[r] = clip(d->efc_force[r], -d->efc_floss[r], d->efc_floss[r]);
      }
      primalFromForces();
      double cost = 0;   // 1/2 f'(A + R) f + f'b  with  A f = J (w - qacc_smooth),  b = J qacc_smooth - aref
      for (int r = 0; r < nefc; r++) {
        const double f = d->efc_force[r], jw = rowDot(r, w), js = rowDot(r, d->qacc_smooth);
        cost += 0.5 * f * (jw - js + f / d->efc_D[r]) + f * (js - d->efc_aref[r]);
      }
      warm = cost < 0;
    }
    if (!warm) {
      std::fill(d->efc_force.begin(), d->efc_force.begin() + nefc, 0.0);
      w = d->qacc_smooth;
    }
    int iter = 0;
    while (iter < maxiter) {
      double improvement = 0;
      for (int r = 0; r < nefc; r++) {
        minvJt(r);
        const double R = 1 / d->efc_D[r], ARrr = rowDot(r, u) + R;
        const double old = d->efc_force[r];
        const double res = rowDot(r, w) - d->efc_aref[r] + R * old;
        double f = old - res / ARrr;
        if (d->efc_floss[r] > 0) f = clip(f, -d->efc_floss[r], d->efc_floss[r]);   // dry friction: a box
        else if (r >= d->ne && f < 0) f = 0;        // limits and pyramidal contact edges push only; equalities pull both ways
        const double delta = f - old;
        if (delta != 0) {
          d->efc_force[r] = f;
          for (int i = 0; i < nv; i++) w[i] += delta * u[i];
          improvement -= 0.5 * delta * delta * ARrr + delta * res;
        }
      }
      iter++;
      if (improvement * scale < m->tolerance) break;
    }
    d->solver_niter = iter;
  }
  // noslip: re-solves the friction dimensions WITHOUT regularisation (R = 0), normal forces fixed. For a pyramidal contact the
  // opposing edges (f+, f-) of each friction direction keep their sum (= the normal load they carry) and only their
  // difference y moves: 1-D quadratic in y with curvature K = (J+ - J-) M^-1 (J+ - J-)', clamped to |y| <= (f+ + f-)/2.
  void noslip(int maxiter) {
    const double scale = 1 / (m->meaninertia * std::max(1, nv));
    primalFromForces();
    for (int iter = 0; iter < maxiter; iter++) {
      double improvement = 0;
      for (int r = d->ne; r < d->ne + d->nf; r++) {   // dry-friction rows: the same update without the regulariser
        minvJt(r);
        const double Arr = rowDot(r, u);
        if (Arr < OX_MINVAL) continue;
        const double old = d->efc_force[r], res = rowDot(r, w) - d->efc_aref[r];
        const double f = clip(old - res / Arr, -d->efc_floss[r], d->efc_floss[r]), delta = f - old;
        if (delta != 0) {
          d->efc_force[r] = f;
          for (int i = 0; i < nv; i++) w[i] += delta * u[i];
          improvement -= 0.5 * delta * delta * Arr + delta * res;
        }
      }
      for (int r = d->ne; r < nefc; r++) {
        if (d->efc_type[r] != 2) continue;
        const int c = d->efc_id[r];
        int first = r;
        while (first > 0 && d->efc_type[first - 1] == 2 && d->efc_id[first - 1] == c) first--;
        if ((r - first) % 2) continue;               // r is the '+' edge of a pair
        const int ra = r, rb = r + 1;
        minvJt(ra, 1.0, rb);
        const double K = rowDot(ra, u) - rowDot(rb, u);
        if (K < OX_MINVAL) continue;
        const double fa = d->efc_force[ra], fb = d->efc_force[rb], mid = 0.5 * (fa + fb), yold = 0.5 * (fa - fb);
        const double dres = (rowDot(ra, w) - d->efc_aref[ra]) - (rowDot(rb, w) - d->efc_aref[rb]);
        double y = yold - dres / K;
        y = std::max(-mid, std::min(mid, y));
        const double delta = y - yold;
        if (delta != 0) {
          d->efc_force[ra] = mid + y; d->efc_force[rb] = mid - y;
          for (int i = 0; i < nv; i++) w[i] += delta * u[i];
          improvement -= 0.5 * delta * delta * K + delta * dres;
        }
      }
      if (improvement * scale < m->noslip_tolerance) break;
    }
  }
  void finish() {
    primalFromForces();
    d->qacc = w;
  }
};

void fwdConstraint(const Model* m, Data* d, int iterations, int ls_iterations) {
  int nv = m->nv;
  if (d->nefc == 0) {
    d->qacc = d->qacc_smooth;
    d->qacc_warmstart = d->qacc_smooth;
    std::fill(d->qfrc_constraint.begin(), d->qfrc_constraint.end(), 0.0);
    d->solver_niter = 0;
    return;
  }
  if (m->solver == OX_SOL_PGS) {
    DualSolver ds(m, d);
    ds.pgs(iterations);
    if (m->noslip_iterations > 0) ds.noslip(m->noslip_iterations);
    ds.finish();
  } else {
    Solver s(m, d);
    if (!disabled(m, OX_DSBL_WARMSTART)) {
      double cw = s.costAt(d->qacc_warmstart.data()), cs = s.costAt(d->qacc_smooth.data());
      d->qacc = cw > cs ? d->qacc_smooth : d->qacc_warmstart;
    } else d->qacc = d->qacc_smooth;
    s.solve(m->solver == OX_SOL_NEWTON, iterations, ls_iterations);
    if (m->noslip_iterations > 0) {
      DualSolver ds(m, d);
      ds.noslip(m->noslip_iterations);
      ds.finish();
    }
  }
  for (int i = 0; i < nv; i++) d->qacc_warmstart[i] = d->qacc[i];
}

// ---------------------------------------------------------------- sensors (N2 subset)
void subtreeLinvel(const Model* m, const Data* d, std::vector<double>& out) {
  int nb = m->nbody;
  out.assign(3 * nb, 0.0);
  // linear momentum of each body: m * (v_lin at body com); cvel is at subtree_com[root] in world axes
  for (int b = nb - 1; b > 0; b--) {
    double dif[3], v[3];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[3 * b + k] - d->subtree_com[3 * m->body_rootid[b] + k];
    cross3(v, &d->cvel[6 * b], dif);
    for (int k = 0; k < 3; k++) out[3 * b + k] += m->body_mass[b] * (d->cvel[6 * b + 3 + k] + v[k]);
    int p = m->body_parentid[b];
    for (int k = 0; k < 3; k++) out[3 * p + k] += out[3 * b + k];
  }
  for (int b = 0; b < nb; b++)
    for (int k = 0; k < 3; k++) out[3 * b + k] /= std::max(OX_MINVAL, m->body_subtreemass[b]);
}

void objFrame(const Model* m, const Data* d, int objtype, int id, const double** pos, const double** mat, int* body) {
  switch (objtype) {
    case OX_OBJ_BODY: *pos = &d->xipos[3 * id]; *mat = &d->ximat[9 * id]; *body = id; break;
    case OX_OBJ_XBODY: *pos = &d->xpos[3 * id]; *mat = &d->xmat[9 * id]; *body = id; break;
    case OX_OBJ_GEOM: *pos = &d->geom_xpos[3 * id]; *mat = &d->geom_xmat[9 * id]; *body = m->geom_bodyid[id]; break;
    default: *pos = &d->site_xpos[3 * id]; *mat = &d->site_xmat[9 * id]; *body = m->site_bodyid[id]; break;
  }
}
void mat2Quat(double* q, const double* m) {
  double tr = m[0] + m[4] + m[8];
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0) * 2;
    q[0] = 0.25 * s; q[1] = (m[7] - m[5]) / s; q[2] = (m[2] - m[6]) / s; q[3] = (m[3] - m[1]) / s;
  } else if (m[0] > m[4] && m[0] > m[8]) {
    double s = std::sqrt(1.0 + m[0] - m[4] - m[8]) * 2;
    q[0] = (m[7] - m[5]) / s; q[1] = 0.25 * s; q[2] = (m[1] + m[3]) / s; q[3] = (m[2] + m[6]) / s;
  } else if (m[4] > m[8]) {
    double s = std::sqrt(1.0 + m[4] - m[0] - m[8]) * 2;
    q[0] = (m[2] - m[6]) / s; q[1] = (m[1] + m[3]) / s; q[2] = 0.25 * s; q[3] = (m[5] + m[7]) / s;
  } else {
    double s = std::sqrt(1.0 + m[8] - m[0] - m[4]) * 2;
    q[0] = (m[3] - m[1]) / s; q[1] = (m[2] + m[6]) / s; q[2] = (m[5] + m[7]) / s; q[3] = 0.25 * s;
  }
  normalize4(q);
}

// body accelerations in the com frame (mj_rnePostConstraint's forward pass):
// cacc[0] = (0, -gravity); cacc[b] = cacc[parent] + sum over the body's dofs of cdof_dot*qvel + cdof*qacc
void bodyAcc(const Model* m, const Data* d, std::vector<double>& cacc) {
  cacc.assign(6 * m->nbody, 0.0);
  if (!(m->disableflags & OX_DSBL_GRAVITY))
    for (int k = 0; k < 3; k++) cacc[3 + k] = -m->gravity[k];
  for (int b = 1; b < m->nbody; b++) {
    const int p = m->body_parentid[b];
    for (int k = 0; k < 6; k++) cacc[6 * b + k] = cacc[6 * p + k];
    for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++)
      for (int k = 0; k < 6; k++) cacc[6 * b + k] += d->cdof_dot[6 * i + k] * d->qvel[i] + d->cdof[6 * i + k] * d->qacc[i];
  }
}

// mju_rayGeom restated for the site shapes a touch sensor may have (sphere, capsule, box): distance along `vec` from `pnt` to
// the nearest intersection with the shape, -1 if the ray misses. Written with explicit candidate lists rather than MuJoCo's
// in-place minimum so that it can be read against the documentation.
double rayGeom(const double* pos, const double* mat, const double* size, const double* pnt, const double* vec, int type) {
  double lp[3], lv[3];
  for (int k = 0; k < 3; k++) {
    lp[k] = mat[k] * (pnt[0] - pos[0]) + mat[3 + k] * (pnt[1] - pos[1]) + mat[6 + k] * (pnt[2] - pos[2]);
    lv[k] = mat[k] * vec[0] + mat[3 + k] * vec[1] + mat[6 + k] * vec[2];
  }
  std::vector<double> hits;
  auto roots = [](double a, double b, double c, double* x) {  // a x^2 + 2 b x + c = 0
    double det = b * b - a * c;
    if (det < OX_MINVAL) return false;
    x[0] = (-b - std::sqrt(det)) / a; x[1] = (-b + std::sqrt(det)) / a;
    return true;
  };
  double x[2];
  if (type == OX_GEOM_SPHERE) {
    if (roots(dot3(lv, lv), dot3(lv, lp), dot3(lp, lp) - size[0] * size[0], x)) { if (x[0] >= 0) hits.push_back(x[0]); else if (x[1] >= 0) hits.push_back(x[1]); }
  } else if (type == OX_GEOM_BOX) {
    for (int i = 0; i < 3; i++) {
      if (std::fabs(lv[i]) <= OX_MINVAL) continue;
      for (double side : {-1.0, 1.0}) {
        double t = (side * size[i] - lp[i]) / lv[i];
        int j = (i + 1) % 3, k = (i + 2) % 3;
        if (t >= 0 && std::fabs(lp[j] + t * lv[j]) <= size[j] && std::fabs(lp[k] + t * lv[k]) <= size[k]) hits.push_back(t);
      }
    }
  } else {  // capsule along z: radius size[0], half-length size[1]
    if (roots(lv[0] * lv[0] + lv[1] * lv[1], lv[0] * lp[0] + lv[1] * lp[1], lp[0] * lp[0] + lp[1] * lp[1] - size[0] * size[0], x)) {
      double t = x[0] >= 0 ? x[0] : (x[1] >= 0 ? x[1] : -1);
      if (t >= 0 && std::fabs(lp[2] + t * lv[2]) <= size[1]) hits.push_back(t);
    }
    for (double side : {-1.0, 1.0}) {
      double c[3] = {lp[0], lp[1], lp[2] - side * size[1]};
      if (!roots(dot3(lv, lv), dot3(lv, c), dot3(c, c) - size[0] * size[0], x)) continue;
      for (int i = 0; i < 2; i++) {
        double z = lp[2] + x[i] * lv[2];
        if (x[i] >= 0 && (side > 0 ? z >= size[1] : z <= -size[1])) hits.push_back(x[i]);
      }
    }
  }
  return hits.empty() ? -1.0 : *std::min_element(hits.begin(), hits.end());
}

// mj_rnePostConstraint: com-based interaction force of every body with its parent, cfrc_int = [torque; force] about
// subtree_com[root] in world axes.  cfrc_ext collects the Cartesian forces applied to each body (xfrc_applied, contact forces,
// connect-equality forces); cfrc_int[b] = cinert cacc + cvel x* (cinert cvel) - cfrc_ext, accumulated from the leaves to the root.
void addExtForce(const Model* m, const Data* d, std::vector<double>& ext, int body, const double* point, const double* force, const double* torque,
                 double sign) {
  if (body == 0) return;
  const double* sc = &d->subtree_com[3 * m->body_rootid[body]];
  double dif[3] = {point[0] - sc[0], point[1] - sc[1], point[2] - sc[2]}, t[3];
  cross3(t, dif, force);                       // moving the force to the com-frame origin adds (point - origin) x f
  for (int k = 0; k < 3; k++) {
    ext[6 * body + k] += sign * (t[k] + (torque ? torque[k] : 0.0));
    ext[6 * body + 3 + k] += sign * force[k];
  }
}
void rnePostConstraint(const Model* m, const Data* d, const std::vector<double>& cacc, std::vector<double>& cfrc_int) {
  const int nb = m->nbody, nv = m->nv;
  (void)nv;
  std::vector<double> ext(6 * nb, 0.0);
  for (int b = 1; b < nb; b++) {
    const double* f = &d->xfrc_applied[6 * b];
    bool any = false;
    for (int k = 0; k < 6; k++) any |= f[k] != 0;
    if (any) addExtForce(m, d, ext, b, &d->xipos[3 * b], f, f + 3, 1.0);
  }
  // contacts: mj_contactForce in the contact frame (pyramid: normal = sum of the edge forces, tangent k = (f+ - f-) mu_k), rotated
  // to the world; it pushes geom2's body along +normal and geom1's body the other way
  for (int c = 0; c < d->ncon; c++) {
    int first = -1, n = 0;
    for (int r = 0; r < d->nefc; r++)
      if ((d->efc_type[r] == 1 || d->efc_type[r] == 2 || d->efc_type[r] == 6) && d->efc_id[r] == c) { if (first < 0) first = r; n++; }
    if (first < 0) continue;
    const int p = d->con_pair[c];
    const double* fri = m->pair_friction + 5 * p;
    double lf[6] = {0, 0, 0, 0, 0, 0};   // contact-frame force (normal, tangents) and torque (torsion, rolling)
    const bool elliptic = d->efc_type[first] == 6;
    if (n == 1) lf[0] = d->efc_force[first];
    else if (elliptic)
      for (int k = 0; k < n; k++) lf[k] = d->efc_force[first + k];   // rows are the contact-frame components themselves
    else
      for (int k = 0; k < n / 2; k++) {
        lf[0] += d->efc_force[first + 2 * k] + d->efc_force[first + 2 * k + 1];
        lf[1 + k] = (d->efc_force[first + 2 * k] - d->efc_force[first + 2 * k + 1]) * fri[k];
      }
    const double* fr = &d->con_frame[9 * c];
    double wf[3], wt[3];
    for (int k = 0; k < 3; k++) {
      wf[k] = fr[k] * lf[0] + fr[3 + k] * lf[1] + fr[6 + k] * lf[2];
      wt[k] = fr[k] * lf[3] + fr[3 + k] * lf[4] + fr[6 + k] * lf[5];
    }
    const bool has_torque = elliptic ? n > 3 : n > 4;
    addExtForce(m, d, ext, m->geom_bodyid[m->pair_geom1[p]], &d->con_pos[3 * c], wf, has_torque ? wt : nullptr, -1.0);
    addExtForce(m, d, ext, m->geom_bodyid[m->pair_geom2[p]], &d->con_pos[3 * c], wf, has_torque ? wt : nullptr, +1.0);
  }
  // connect equalities: the three row forces are a world-frame force on body1 at its anchor and the opposite on body2 at its own
  for (int r = 0; r + 2 < d->ne; r++) {
    if (d->efc_type[r] != 3 || m->eq_type[d->efc_id[r]] == OX_EQ_JOINT) continue;
    const int i = d->efc_id[r], b1 = m->eq_obj1id[i], b2 = m->eq_obj2id[i];
    const double* data = m->eq_data + 11 * i;
    double p1[3], p2[3], w[3];
    mulMatVec3(w, &d->xmat[9 * b1], data);
    for (int k = 0; k < 3; k++) p1[k] = d->xpos[3 * b1 + k] + w[k];
    mulMatVec3(w, &d->xmat[9 * b2], data + 3);
    for (int k = 0; k < 3; k++) p2[k] = d->xpos[3 * b2 + k] + w[k];
    addExtForce(m, d, ext, b1, p1, &d->efc_force[r], nullptr, +1.0);
    addExtForce(m, d, ext, b2, p2, &d->efc_force[r], nullptr, -1.0);
    r += 2;
    if (m->eq_type[i] == OX_EQ_WELD) {   // the three orientation rows are a torque pair: tau = G' f on body1, -tau on body2
      double G[9], res[3], tau[3], zero[3] = {0, 0, 0};
      weldRotMap(&d->xquat[4 * b1], &d->xquat[4 * b2], data + 6, data[10], G, res);
      for (int c = 0; c < 3; c++) tau[c] = G[c] * d->efc_force[r + 1] + G[3 + c] * d->efc_force[r + 2] + G[6 + c] * d->efc_force[r + 3];
      addExtForce(m, d, ext, b1, p1, zero, tau, +1.0);
      addExtForce(m, d, ext, b2, p2, zero, tau, -1.0);
      r += 3;
    }
  }
  cfrc_int.assign(6 * nb, 0.0);
  for (int b = 1; b < nb; b++) {
    double ia[6], iv[6], cf[6];
    mulInertVec(ia, &d->cinert[10 * b], &cacc[6 * b]);
    mulInertVec(iv, &d->cinert[10 * b], &d->cvel[6 * b]);
    crossForce(cf, &d->cvel[6 * b], iv);
    for (int k = 0; k < 6; k++) cfrc_int[6 * b + k] = ia[k] + cf[k] - ext[6 * b + k];
  }
  for (int b = nb - 1; b > 0; b--)
    for (int k = 0; k < 6; k++) cfrc_int[6 * m->body_parentid[b] + k] += cfrc_int[6 * b + k];
}

void sensors(const Model* m, Data* d) {
  std::vector<double> slv, cacc, cfrc_int;
  for (int s = 0; s < m->nsensor; s++) {
    double* out = &d->sensordata[m->sensor_adr[s]];
    int id = m->sensor_objid[s], ot = m->sensor_objtype[s];
    const double *pos, *mat;
    int body;
    switch (m->sensor_type[s]) {
      case OX_SENS_JOINTPOS: out[0] = d->qpos[m->jnt_qposadr[id]]; break;
      case OX_SENS_JOINTVEL: out[0] = d->qvel[m->jnt_dofadr[id]]; break;
      case OX_SENS_ACTUATORPOS:
        out[0] = m->actuator_gear[id] * (m->actuator_trntype[id] == OX_TRN_TENDON ? d->ten_length[m->actuator_trnid[id]] : d->qpos[m->jnt_qposadr[m->actuator_trnid[id]]]);
        break;
      case OX_SENS_ACTUATORVEL:
        out[0] = m->actuator_gear[id] * (m->actuator_trntype[id] == OX_TRN_TENDON ? tendonVelocity(m, d, m->actuator_trnid[id]) : d->qvel[m->jnt_dofadr[m->actuator_trnid[id]]]);
        break;
      case OX_SENS_ACTUATORFRC: out[0] = d->actuator_force[id]; break;
      case OX_SENS_JOINTACTFRC: out[0] = d->qfrc_actuator[m->jnt_dofadr[id]]; break;   // mjSENS_JOINTACTFRC: net actuator force on the joint
      case OX_SENS_BALLQUAT: {                                                          // mjSENS_BALLQUAT: copy, then mju_normalize4
        const double* q = &d->qpos[m->jnt_qposadr[id]];
        const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        if (n < OX_MINVAL) { out[0] = 1; out[1] = out[2] = out[3] = 0; }
        else for (int k = 0; k < 4; k++) out[k] = q[k] / n;
        break;
      }
      case OX_SENS_BALLANGVEL: std::memcpy(out, &d->qvel[m->jnt_dofadr[id]], 3 * sizeof(double)); break;
      case OX_SENS_TENDONPOS: out[0] = d->ten_length[id]; break;
      case OX_SENS_TENDONVEL: out[0] = tendonVelocity(m, d, id); break;
      case OX_SENS_TOUCH: {
        // mj_sensorAcc / mjSENS_TOUCH: normal forces (mj_contactForce: sum of the pyramid edge forces, or the single row of a
        // frictionless contact) of contacts on the site's body whose point the site volume contains (ray along the normal)
        out[0] = 0;
        int sbody = m->site_bodyid[id];
        for (int c = 0; c < d->ncon; c++) {
          int p = d->con_pair[c], b1 = m->geom_bodyid[m->pair_geom1[p]], b2 = m->geom_bodyid[m->pair_geom2[p]];
          if (sbody != b1 && sbody != b2) continue;
          double fn = 0;
          bool has_rows = false;
          for (int r = 0; r < d->nefc; r++)
            if ((d->efc_type[r] == 1 || d->efc_type[r] == 2) && d->efc_id[r] == c) { fn += d->efc_force[r]; has_rows = true; }
            else if (d->efc_type[r] == 6 && d->efc_id[r] == c) { if (!has_rows) fn = d->efc_force[r]; has_rows = true; }   // elliptic: the first row IS the normal force
          if (!has_rows || fn <= 0) continue;
          double ray[3] = {d->con_frame[9 * c], d->con_frame[9 * c + 1], d->con_frame[9 * c + 2]};
          if (sbody == b2) for (double& v : ray) v = -v;
          if (rayGeom(&d->site_xpos[3 * id], &d->site_xmat[9 * id], m->site_size + 3 * id, &d->con_pos[3 * c], ray, m->site_type[id]) >= 0) out[0] += fn;
        }
        break;
      }
      case OX_SENS_SUBTREECOM: std::memcpy(out, &d->subtree_com[3 * id], 3 * sizeof(double)); break;
      case OX_SENS_SUBTREELINVEL:
        if (slv.empty()) subtreeLinvel(m, d, slv);
        std::memcpy(out, &slv[3 * id], 3 * sizeof(double));
        break;
      case OX_SENS_FRAMEPOS: objFrame(m, d, ot, id, &pos, &mat, &body); std::memcpy(out, pos, 3 * sizeof(double)); break;
      case OX_SENS_FRAMEQUAT: objFrame(m, d, ot, id, &pos, &mat, &body); mat2Quat(out, mat); break;
      case OX_SENS_FRAMEXAXIS: case OX_SENS_FRAMEYAXIS: case OX_SENS_FRAMEZAXIS: {     // column of xmat (mjSENS_FRAME[XYZ]AXIS)
        objFrame(m, d, ot, id, &pos, &mat, &body);
        const int c = m->sensor_type[s] - OX_SENS_FRAMEXAXIS;
        for (int k = 0; k < 3; k++) out[k] = mat[3 * k + c];
        break;
      }
      case OX_SENS_FRAMELINVEL: case OX_SENS_FRAMEANGVEL: case OX_SENS_VELOCIMETER: case OX_SENS_GYRO: {
        // object velocity (mj_objectVelocity): cvel transported from the com frame origin to the object position
        objFrame(m, d, ot, id, &pos, &mat, &body);
        const double* cv = &d->cvel[6 * body];
        double dif[3], lin[3], tmp[3];
        for (int k = 0; k < 3; k++) dif[k] = pos[k] - d->subtree_com[3 * m->body_rootid[body] + k];
        cross3(tmp, cv, dif);
        for (int k = 0; k < 3; k++) lin[k] = cv[3 + k] + tmp[k];
        int ty = m->sensor_type[s];
        const double* src = (ty == OX_SENS_FRAMELINVEL || ty == OX_SENS_VELOCIMETER) ? lin : cv;
        if (ty == OX_SENS_VELOCIMETER || ty == OX_SENS_GYRO) {  // local frame: mat' * v
          for (int k = 0; k < 3; k++) out[k] = mat[k] * src[0] + mat[3 + k] * src[1] + mat[6 + k] * src[2];
        } else std::memcpy(out, src, 3 * sizeof(double));
        break;
      }
      case OX_SENS_FRAMELINACC: case OX_SENS_FRAMEANGACC: case OX_SENS_ACCELEROMETER: {
        // mj_objectAcceleration(local): cacc and cvel transported to the site, plus omega x v, in the site frame.
        // At rest this reads -gravity (an accelerometer measures proper acceleration).
        // mjSENS_FRAMELINACC / FRAMEANGACC: mj_objectAcceleration(flg_local = 0) of any frame object - world axes.
        if (cacc.empty()) bodyAcc(m, d, cacc);
        objFrame(m, d, m->sensor_type[s] == OX_SENS_ACCELEROMETER ? (int)OX_OBJ_SITE : ot, id, &pos, &mat, &body);
        const double *cv = &d->cvel[6 * body], *ca = &cacc[6 * body];
        double dif[3], lv[3], la[3], t1[3], t2[3], t3[3];
        for (int k = 0; k < 3; k++) dif[k] = pos[k] - d->subtree_com[3 * m->body_rootid[body] + k];
        cross3(t1, cv, dif);
        cross3(t2, ca, dif);
        for (int k = 0; k < 3; k++) { lv[k] = cv[3 + k] + t1[k]; la[k] = ca[3 + k] + t2[k]; }
        cross3(t3, cv, lv);
        for (int k = 0; k < 3; k++) la[k] += t3[k];
        if (m->sensor_type[s] == OX_SENS_FRAMELINACC) { std::memcpy(out, la, sizeof la); break; }
        if (m->sensor_type[s] == OX_SENS_FRAMEANGACC) { std::memcpy(out, ca, 3 * sizeof(double)); break; }
        for (int k = 0; k < 3; k++) out[k] = mat[k] * la[0] + mat[3 + k] * la[1] + mat[6 + k] * la[2];
        break;
      }
      case OX_SENS_FORCE: case OX_SENS_TORQUE: {
        // interaction force / torque between the site's body and its parent, at the site, in the site frame (mj_sensorAcc)
        if (cacc.empty()) bodyAcc(m, d, cacc);
        if (cfrc_int.empty()) rnePostConstraint(m, d, cacc, cfrc_int);
        objFrame(m, d, OX_OBJ_SITE, id, &pos, &mat, &body);
        const double* ci = &cfrc_int[6 * body];
        double dif[3], t[3], v[3];
        for (int k = 0; k < 3; k++) dif[k] = pos[k] - d->subtree_com[3 * m->body_rootid[body] + k];
        cross3(t, dif, ci + 3);
        for (int k = 0; k < 3; k++) v[k] = m->sensor_type[s] == OX_SENS_FORCE ? ci[3 + k] : ci[k] - t[k];   // torque moved to the site: tau - (site - origin) x f
        for (int k = 0; k < 3; k++) out[k] = mat[k] * v[0] + mat[3 + k] * v[1] + mat[6 + k] * v[2];
        break;
      }
      case OX_SENS_CLOCK: out[0] = d->time; break;
      default: break;
    }
  }
}

// ---------------------------------------------------------------- forward / integrators / step
void forwardSkip(const Model* m, Data* d, bool skipsensor) {
  kinematics(m, d);
  comPos(m, d);
  crb(m, d);
  factorM(m, d);
  collision(m, d);
  comVel(m, d);  // (order within the position/velocity stages does not matter for independent outputs)
  makeConstraint(m, d);
  passive(m, d);
  rne(m, d);
  actuation(m, d);
  fwdAcceleration(m, d);
  fwdConstraint(m, d, m->iterations, m->ls_iterations);
  if (!skipsensor) sensors(m, d);
}

void integratePos(const Model* m, double* qpos, const double* qvel, double dt) {
  for (int j = 0; j < m->njnt; j++) {
    int pa = m->jnt_qposadr[j], va = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case OX_JNT_FREE:
        for (int i = 0; i < 3; i++) qpos[pa + i] += dt * qvel[va + i];
        pa += 3; va += 3;
        // fallthrough
      case OX_JNT_BALL: quatIntegrate(qpos + pa, qvel + va, dt); break;
      default: qpos[pa] += dt * qvel[va];
    }
  }
}
// mj_nextActivation: act + h act_dot, or the exact first-order-filter update act + act_dot tau (1 - exp(-h/tau)); then actrange
double nextActivation(const Model* m, int i, double act, double act_dot) {
  if (m->actuator_dyntype[i] == OX_DYN_FILTEREXACT) {
    double tau = std::max(OX_MINVAL, m->actuator_dynprm[3 * i]);
    act += act_dot * tau * (1 - std::exp(-m->timestep / tau));
  } else {
    act += act_dot * m->timestep;
  }
  if (m->actuator_actlimited[i]) act = clip(act, m->actuator_actrange[2 * i], m->actuator_actrange[2 * i + 1]);
  return act;
}
void advance(const Model* m, Data* d, const double* qacc, const double* qvel_override, const double* act_dot) {
  for (int i = 0; i < m->nu; i++)
    if (m->actuator_dyntype[i] != OX_DYN_NONE) {
      int aa = m->actuator_actadr[i];
      d->act[aa] = nextActivation(m, i, d->act[aa], act_dot[aa]);
    }
  for (int i = 0; i < m->nv; i++) d->qvel[i] += m->timestep * qacc[i];
  integratePos(m, d->qpos.data(), qvel_override ? qvel_override : d->qvel.data(), m->timestep);
  d->time += m->timestep;
}
void euler(const Model* m, Data* d) {
  int nv = m->nv;
  bool damping = false;
  const bool fast = m->integrator == OX_INT_IMPLICITFAST;
  if (!disabled(m, OX_DSBL_EULERDAMP))
    for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damping = true;
  if (!damping && !fast) { advance(m, d, d->qacc.data(), nullptr, d->act_dot.data()); return; }
  // implicit-in-velocity joint damping: (M + h B) qacc' = qfrc_smooth + qfrc_constraint.
  // MuJoCo factors into d->qLD, overwriting the factor of M; so do we.
  d->qLD = d->qM;
  for (int i = 0; i < nv; i++) d->qLD[m->dof_Madr[i]] += m->timestep * m->dof_damping[i];
  if (fast && !disabled(m, OX_DSBL_ACTUATION)) {
    // mj_implicit, implicitfast flavour: M - h d(qfrc_smooth)/d(qvel) with the RNE terms dropped. In this
    // subset the derivative is diagonal: -damping (above) + moment' (d force / d velocity) moment for affine gain / bias
    // terms (mjd_actuator_vel), skipped while the force is clamped at its forcerange.
    for (int i = 0; i < m->nu; i++) {
      double dfdv = 0;
      double in = d->ctrl[i];
      if (m->actuator_ctrllimited[i] && !disabled(m, OX_DSBL_CLAMPCTRL)) in = clip(in, m->actuator_ctrlrange[2 * i], m->actuator_ctrlrange[2 * i + 1]);
      if (m->actuator_dyntype[i] != OX_DYN_NONE) in = d->act[m->actuator_actadr[i]];
      if (m->actuator_gaintype[i] == OX_GAIN_AFFINE) dfdv += m->actuator_gainprm[3 * i + 2] * in;
      if (m->actuator_biastype[i] == OX_BIAS_AFFINE) dfdv += m->actuator_biasprm[3 * i + 2];
      if (dfdv == 0 && m->actuator_gaintype[i] != OX_GAIN_AFFINE && m->actuator_biastype[i] != OX_BIAS_AFFINE) continue;
      if (m->actuator_forcelimited[i] &&
          (d->actuator_force[i] <= m->actuator_forcerange[2 * i] || d->actuator_force[i] >= m->actuator_forcerange[2 * i + 1])) continue;
      int da = m->jnt_dofadr[m->actuator_trnid[i]];
      d->qLD[m->dof_Madr[da]] -= m->timestep * m->actuator_gear[i] * m->actuator_gear[i] * dfdv;
    }
  }
  factorLD(m, d->qLD.data(), d->qLDiagInv.data());
  std::vector<double> qacc(nv);
  for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
  solveLD(m, d->qLD.data(), d->qLDiagInv.data(), qacc.data());
  advance(m, d, qacc.data(), nullptr, d->act_dot.data());
}
void rk4(const Model* m, Data* d) {
  static const double A[9] = {0.5, 0, 0, 0, 0.5, 0, 0, 0, 1}, Bw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
  int nv = m->nv, nq = m->nq;
  double h = m->timestep, time = d->time;
  double C[3], T[3];
  for (int i = 0; i < 3; i++) { C[i] = A[3 * i] + A[3 * i + 1] + A[3 * i + 2]; T[i] = time + C[i] * h; }
  std::vector<double> X0q = d->qpos, X0v = d->qvel, X0act = d->act;
  std::vector<std::vector<double>> Fv(4), Fa(4), Fact(4);
  Fv[0] = d->qvel; Fa[0] = d->qacc; Fact[0] = d->act_dot;
  const int na = (int)d->act.size();
  std::vector<double> dXv(nv), dXa(nv), dXact(na);
  for (int i = 1; i < 4; i++) {
    std::fill(dXv.begin(), dXv.end(), 0.0);
    std::fill(dXa.begin(), dXa.end(), 0.0);
    std::fill(dXact.begin(), dXact.end(), 0.0);
    for (int j = 0; j < i; j++) {
      double a = A[(i - 1) * 3 + j];
      for (int k = 0; k < nv; k++) { dXv[k] += a * Fv[j][k]; dXa[k] += a * Fa[j][k]; }
      for (int k = 0; k < na; k++) dXact[k] += a * Fact[j][k];
    }
    for (int k = 0; k < na; k++) d->act[k] = X0act[k] + h * dXact[k];
    d->qpos = X0q;
    integratePos(m, d->qpos.data(), dXv.data(), h);
    for (int k = 0; k < nv; k++) d->qvel[k] = X0v[k] + h * dXa[k];
    d->time = T[i - 1];
    forwardSkip(m, d, true);
    Fv[i] = d->qvel; Fa[i] = d->qacc; Fact[i] = d->act_dot;
  }
  std::fill(dXv.begin(), dXv.end(), 0.0);
  std::fill(dXa.begin(), dXa.end(), 0.0);
  std::fill(dXact.begin(), dXact.end(), 0.0);
  for (int j = 0; j < 4; j++) {
    for (int k = 0; k < nv; k++) { dXv[k] += Bw[j] * Fv[j][k]; dXa[k] += Bw[j] * Fa[j][k]; }
    for (int k = 0; k < na; k++) dXact[k] += Bw[j] * Fact[j][k];
  }
  d->time = time;
  d->qpos = X0q; d->qvel = X0v; d->act = X0act;
  (void)nq;
  advance(m, d, dXa.data(), dXv.data(), dXact.data());
}

bool isBad(double x) { return std::isnan(x) || x > OX_MAXVAL || x < -OX_MAXVAL; }

void resetData(const Model* m, Data* d) {
  std::memcpy(d->qpos.data(), m->qpos0, m->nq * sizeof(double));
  auto z = [](std::vector<double>& v) { std::fill(v.begin(), v.end(), 0.0); };
  z(d->qvel); z(d->ctrl); z(d->qfrc_applied); z(d->xfrc_applied); z(d->qacc_warmstart); z(d->qacc); z(d->act); z(d->act_dot);
  for (int i = 1; i < m->nbody; i++)   // mj_resetData: mocap poses back to the model's body poses, equality constraints to their defaults
    if (m->nmocap > 0 && m->body_mocapid[i] >= 0) {
      std::memcpy(&d->mocap_pos[3 * m->body_mocapid[i]], m->body_pos + 3 * i, 3 * sizeof(double));
      std::memcpy(&d->mocap_quat[4 * m->body_mocapid[i]], m->body_quat + 4 * i, 4 * sizeof(double));
    }
  for (int i = 0; i < m->neq; i++) d->eq_active[i] = m->eq_active0[i];
  z(d->xpos); z(d->xquat); z(d->xmat); z(d->xipos); z(d->ximat); z(d->xanchor); z(d->xaxis); z(d->geom_xpos); z(d->geom_xmat);
  z(d->site_xpos); z(d->site_xmat); z(d->subtree_com); z(d->cinert); z(d->cdof); z(d->qM); z(d->qLD); z(d->qLDiagInv);
  z(d->cvel); z(d->cdof_dot); z(d->qfrc_bias); z(d->qfrc_passive); z(d->actuator_force); z(d->qfrc_actuator); z(d->qfrc_smooth);
  z(d->qacc_smooth); z(d->qfrc_constraint); z(d->sensordata); z(d->efc_force);
  d->time = 0; d->ncon = 0; d->nefc = 0; d->ne = 0; d->nf = 0; d->solver_niter = 0;
}

void step(const Model* m, Data* d) {
  bool bad = false;
  for (double x : d->qpos) bad |= isBad(x);
  for (double x : d->qvel) bad |= isBad(x);
  for (double x : d->act) bad |= isBad(x);
  if (bad) { resetData(m, d); d->diverged++; }
  forwardSkip(m, d, false);
  bad = false;
  for (double x : d->qacc) bad |= isBad(x);
  if (bad) { resetData(m, d); d->diverged++; forwardSkip(m, d, false); }
  if (m->integrator == OX_INT_RK4) rk4(m, d); else euler(m, d);
}

// ---------------------------------------------------------------- Philox4x32-10 control stream
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
inline double ctrl_from_bits(uint32_t x) {  // U(-1,1) on a 2^-23 lattice: exact in fp32 and fp64
  return (double)(int64_t)(((uint64_t)(x >> 9) * 2 + 1)) * (1.0 / 8388608.0) - 1.0;
}
double g_ctrl_scale = 1.0;  // amplitude of the benchmark control stream (oxo_set_ctrl_scale); a power of two keeps it exact in fp32
void fillCtrlPhilox(const Model* m, Data* d, uint64_t seed, int64_t genv, int64_t stepno) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int g = 0; g * 4 < m->nu; g++) {
    uint32_t ctr[4] = {(uint32_t)genv, (uint32_t)((uint64_t)genv >> 32), (uint32_t)stepno, (uint32_t)g}, out[4];
    philox4x32_10(ctr, key, out);
    for (int k = 0; k < 4 && g * 4 + k < m->nu; k++) d->ctrl[g * 4 + k] = g_ctrl_scale * ctrl_from_bits(out[k]);
  }
}

}  // namespace

// ================================================================== exported C interface
extern "C" {
#define OXO_API __attribute__((visibility("default")))

OXO_API oxo_data* oxo_make_data(const Model* m) {
  Data* d = new Data();
  int nb = m->nbody, nv = m->nv;
  d->qpos.resize(m->nq); d->qvel.resize(nv); d->ctrl.resize(m->nu); d->qfrc_applied.resize(nv); d->xfrc_applied.resize(6 * nb);
  d->qacc_warmstart.resize(nv); d->act.resize(m->na); d->act_dot.resize(m->na);
  d->mocap_pos.resize(3 * m->nmocap); d->mocap_quat.resize(4 * m->nmocap); d->eq_active.resize(m->neq); d->ten_length.resize(m->ntendon); d->ten_J.resize((size_t)m->ntendon * m->nv);
  d->xpos.resize(3 * nb); d->xquat.resize(4 * nb); d->xmat.resize(9 * nb); d->xipos.resize(3 * nb); d->ximat.resize(9 * nb);
  d->xanchor.resize(3 * m->njnt); d->xaxis.resize(3 * m->njnt); d->geom_xpos.resize(3 * m->ngeom); d->geom_xmat.resize(9 * m->ngeom);
  d->site_xpos.resize(3 * m->nsite); d->site_xmat.resize(9 * m->nsite);
  d->subtree_com.resize(3 * nb); d->cinert.resize(10 * nb); d->cdof.resize(6 * nv); d->crb.resize(10 * nb);
  d->qM.resize(m->nM); d->qLD.resize(m->nM); d->qLDiagInv.resize(nv);
  d->cvel.resize(6 * nb); d->cdof_dot.resize(6 * nv); d->qfrc_bias.resize(nv); d->qfrc_passive.resize(nv);
  d->actuator_force.resize(m->nu); d->qfrc_actuator.resize(nv); d->qfrc_smooth.resize(nv); d->qacc_smooth.resize(nv);
  int nc = std::max(1, m->nconmax), ne = std::max(1, m->nefcmax);
  d->con_dist.resize(nc); d->con_pos.resize(3 * nc); d->con_frame.resize(9 * nc); d->con_pair.resize(nc);
  d->efc_J.resize((size_t)ne * std::max(1, nv)); d->efc_pos.resize(ne); d->efc_margin.resize(ne); d->efc_D.resize(ne);
  d->efc_R.resize(ne); d->efc_aref.resize(ne); d->efc_vel.resize(ne); d->efc_force.resize(ne); d->efc_diagApprox.resize(ne);
  d->efc_type.resize(ne); d->efc_id.resize(ne); d->efc_floss.resize(ne);
  d->qacc.resize(nv); d->qfrc_constraint.resize(nv); d->sensordata.resize(m->nsensordata);
  resetData(m, d);
  return d;
}
OXO_API void oxo_free_data(oxo_data* d) { delete d; }
OXO_API void oxo_reset(const Model* m, oxo_data* d) { resetData(m, d); }
OXO_API void oxo_forward(const Model* m, oxo_data* d) { forwardSkip(m, d, false); }
OXO_API void oxo_step(const Model* m, oxo_data* d) { step(m, d); }
OXO_API void oxo_fill_ctrl_philox(const Model* m, oxo_data* d, uint64_t seed, int64_t genv, int64_t stepno) {
  fillCtrlPhilox(m, d, seed, genv, stepno);
}
OXO_API void oxo_set_ctrl_scale(double s) { g_ctrl_scale = s; }
OXO_API void oxo_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

// individual stages, in pipeline order
OXO_API void oxo_stage(const Model* m, oxo_data* d, const char* name) {
  std::string s(name);
  if (s == "kinematics") kinematics(m, d);
  else if (s == "comPos") comPos(m, d);
  else if (s == "crb") crb(m, d);
  else if (s == "factorM") factorM(m, d);
  else if (s == "collision") collision(m, d);
  else if (s == "makeConstraint") makeConstraint(m, d);
  else if (s == "comVel") comVel(m, d);
  else if (s == "passive") passive(m, d);
  else if (s == "rne") rne(m, d);
  else if (s == "actuation") actuation(m, d);
  else if (s == "fwdAcceleration") fwdAcceleration(m, d);
  else if (s == "fwdConstraint") fwdConstraint(m, d, m->iterations, m->ls_iterations);
  else if (s == "sensors") sensors(m, d);
}

// field access: pointer into the AoS arrays (valid until oxo_free_data)
OXO_API double* oxo_field(oxo_data* d, const char* name, int32_t* count) {
  std::string s(name);
#define F(f) if (s == #f) { *count = (int32_t)d->f.size(); return d->f.data(); }
  F(ten_length) F(ten_J) F(mocap_pos) F(mocap_quat) F(eq_active) F(act) F(act_dot) F(qpos) F(qvel) F(ctrl) F(qfrc_applied) F(xfrc_applied) F(qacc_warmstart) F(xpos) F(xquat) F(xmat) F(xipos) F(ximat)
  F(xanchor) F(xaxis) F(geom_xpos) F(geom_xmat) F(site_xpos) F(site_xmat) F(subtree_com) F(cinert) F(cdof) F(qM) F(qLD)
  F(qLDiagInv) F(cvel) F(cdof_dot) F(qfrc_bias) F(qfrc_passive) F(actuator_force) F(qfrc_actuator) F(qfrc_smooth) F(qacc_smooth)
  F(con_dist) F(con_pos) F(con_frame) F(efc_J) F(efc_pos) F(efc_margin) F(efc_D) F(efc_floss) F(efc_R) F(efc_aref) F(efc_vel) F(efc_force)
  F(efc_diagApprox) F(qacc) F(qfrc_constraint) F(sensordata)
#undef F
  if (s == "time") { *count = 1; return &d->time; }
  *count = -1;
  return nullptr;
}
OXO_API int32_t oxo_int(oxo_data* d, const char* name) {
  std::string s(name);
  if (s == "ncon") return d->ncon;
  if (s == "nefc") return d->nefc;
  if (s == "ne") return d->ne;
  if (s == "nf") return d->nf;
  if (s == "solver_niter") return d->solver_niter;
  if (s == "diverged") return d->diverged;
  return -1;
}
OXO_API const int32_t* oxo_int_field(oxo_data* d, const char* name, int32_t* count) {
  std::string s(name);
  if (s == "con_pair") { *count = d->ncon; return d->con_pair.data(); }
  if (s == "efc_type") { *count = d->nefc; return d->efc_type.data(); }
  if (s == "efc_id") { *count = d->nefc; return d->efc_id.data(); }
  *count = -1;
  return nullptr;
}

/* CPU baseline: step `nenv` envs for `nsteps` steps on `nthreads` threads with the same seeded initial
 * states (given as [nenv][nq], [nenv][nv]) and Philox controls as the GPU run. Returns wall seconds.
 * Final qpos/qvel are written back so the caller can cross-check against the GPU. */
OXO_API double oxo_bench(const Model* m, int32_t nenv, int32_t nsteps, int32_t nthreads, uint64_t seed, int64_t env_id_offset,
                         int64_t step0, double* qpos_io, double* qvel_io, double* stats_out /*4: ncon nefc niter diverged*/) {
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> th;
  std::vector<double> st(4 * (size_t)nthreads, 0.0);
  auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t]() {
      int lo = (int)((int64_t)nenv * t / nthreads), hi = (int)((int64_t)nenv * (t + 1) / nthreads);
      Data* d = oxo_make_data(m);
      for (int e = lo; e < hi; e++) {
        resetData(m, d);
        std::memcpy(d->qpos.data(), qpos_io + (size_t)e * m->nq, m->nq * sizeof(double));
        std::memcpy(d->qvel.data(), qvel_io + (size_t)e * m->nv, m->nv * sizeof(double));
        int div0 = d->diverged;
        for (int s = 0; s < nsteps; s++) {
          fillCtrlPhilox(m, d, seed, env_id_offset + e, step0 + s);
          step(m, d);
          st[4 * t] += d->ncon; st[4 * t + 1] += d->nefc; st[4 * t + 2] += d->solver_niter;
        }
        st[4 * t + 3] += d->diverged - div0;
        std::memcpy(qpos_io + (size_t)e * m->nq, d->qpos.data(), m->nq * sizeof(double));
        std::memcpy(qvel_io + (size_t)e * m->nv, d->qvel.data(), m->nv * sizeof(double));
      }
      delete d;
    });
  for (auto& x : th) x.join();
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats_out) {
    for (int k = 0; k < 4; k++) stats_out[k] = 0;
    for (int t = 0; t < nthreads; t++) for (int k = 0; k < 4; k++) stats_out[k] += st[4 * t + k];
    double n = (double)nenv * nsteps;
    for (int k = 0; k < 3; k++) stats_out[k] /= std::max(1.0, n);
  }
  return sec;
}
OXO_API int32_t oxo_hardware_threads(void) { return (int32_t)std::thread::hardware_concurrency(); }

}  // extern "C"

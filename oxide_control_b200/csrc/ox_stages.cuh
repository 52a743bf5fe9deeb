// mj_step stage arithmetic for ONE environment of the SoA batch, templated on the real type.
// One CUDA thread owns one environment (lane = env, so every batch access below is coalesced
// across the warp); the model tables are uniform across the warp and come from shared memory.
// Stage list / formulas: SURVEY.md Appendix A (A.1 .. A.12); reference entry Physics::step ->
// mj_step, src/physics.rs:44-46.
//
// The functions are __host__ __device__ only so that tests/native can run the very same code on the
// CPU against the oracle while no GPU is attached. libox_b200.so never instantiates them on the host.
#pragma once
#include <cmath>

#include "ox_blob.h"

namespace ox {

// ---------------------------------------------------------------- scalar helpers
OX_HD float ox_sqrt(float x) { return sqrtf(x); }
OX_HD double ox_sqrt(double x) { return sqrt(x); }
// 1/sqrt(x): one MUFU + a Newton step on the device instead of an IEEE sqrt followed by an IEEE division on the critical path
OX_HD float ox_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
OX_HD double ox_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
OX_HD float ox_abs(float x) { return fabsf(x); }
OX_HD double ox_abs(double x) { return fabs(x); }
OX_HD float ox_pow(float x, float y) { return powf(x, y); }
OX_HD double ox_pow(double x, double y) { return pow(x, y); }
OX_HD float ox_exp(float x) { return expf(x); }
OX_HD double ox_exp(double x) { return exp(x); }
OX_HD float ox_atan2(float y, float x) { return atan2f(y, x); }
OX_HD double ox_atan2(double y, double x) { return atan2(y, x); }
// sincos: ONE out-of-line copy per kernel. Inlined, CUDA's sincosf (fast path + Payne-Hanek slow path) is ~1.3 k
// instructions per call site; with one call per hinge it was 30 % of the specialised cheetah kernel's code
// (profiles/r1_notes.md), which is instruction-fetch bound.
#if defined(__CUDA_ARCH__)
__device__ __noinline__ void ox_sincos_dev(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __noinline__ void ox_sincos_dev(double a, double* s, double* c) { sincos(a, s, c); }
#endif
OX_HD void ox_sincos(float a, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  ox_sincos_dev(a, s, c);
#else
  *s = sinf(a); *c = cosf(a);
#endif
}
OX_HD void ox_sincos(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  ox_sincos_dev(a, s, c);
#else
  *s = sin(a); *c = cos(a);
#endif
}
template <typename T> OX_HD T ox_max(T a, T b) { return a > b ? a : b; }
template <typename T> OX_HD T ox_min(T a, T b) { return a < b ? a : b; }
template <typename T> OX_HD T ox_clip(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
template <typename T> OX_HD bool ox_bad(T x) { return !(x <= (T)OX_MAXVAL && x >= (T)-OX_MAXVAL); }  // NaN or too large
// solver floating-point floor (ORACLE_DECISIONS #8): stop when the cost decrease is below FLOOR_MULT * 4 eps * (|old| + |new|)
#ifndef OX_FLOOR_MULT
#define OX_FLOOR_MULT 8
#endif
template <typename T> struct Eps;
template <> struct Eps<float> { static OX_HD float v() { return 4 * 1.1920929e-07f; } };
template <> struct Eps<double> { static OX_HD double v() { return 4 * 2.220446049250313e-16; } };

// ---------------------------------------------------------------- 3-vectors / quaternions in registers
template <typename T> OX_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> OX_HD void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> OX_HD T normalize3(T* v) {
  T n = ox_sqrt(dot3(v, v));
  if (n < (T)OX_MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; }
  else { T s = (T)1 / n; v[0] *= s; v[1] *= s; v[2] *= s; }
  return n;
}
template <typename T> OX_HD void normalize4(T* q) {
  T n = ox_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < (T)OX_MINVAL) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; }
  else { T s = (T)1 / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
}
template <typename T> OX_HD void mul_quat(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <typename T> OX_HD void quat2mat(T* m, const T* q) {
  T q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03);
  m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
template <typename T> OX_HD void mat_vec3(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
    z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> OX_HD void rot_vec_quat(T* r, const T* v, const T* q) {
  T m[9];
  quat2mat(m, q);
  mat_vec3(r, m, v);
}
template <typename T> OX_HD void axis_angle2quat(T* q, const T* axis, T angle) {
  if (angle == 0) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; return; }
  T s, c;
  ox_sincos(angle * (T)0.5, &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
template <typename T> OX_HD void sub_quat(T* res, const T* qa, const T* qb) {
  T qneg[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, qdif[4];
  mul_quat(qdif, qneg, qa);
  T axis[3] = {qdif[1], qdif[2], qdif[3]};
  T s = normalize3(axis);
  T speed = 2 * ox_atan2(s, qdif[0]);
  if (speed > (T)3.14159265358979323846) speed -= (T)(2 * 3.14159265358979323846);
  res[0] = axis[0] * speed; res[1] = axis[1] * speed; res[2] = axis[2] * speed;
}
template <typename T> OX_HD void quat_integrate(T* quat, const T* vel, T scale) {
  T tmp[3] = {vel[0], vel[1], vel[2]}, qrot[4], out[4];
  T angle = scale * normalize3(tmp);
  axis_angle2quat(qrot, tmp, angle);
  normalize4(quat);
  mul_quat(out, quat, qrot);
  quat[0] = out[0]; quat[1] = out[1]; quat[2] = out[2]; quat[3] = out[3];
}
template <typename T> OX_HD void mat2quat(T* q, const T* m) {
  T tr = m[0] + m[4] + m[8];
  if (tr > 0) {
    T s = ox_sqrt(tr + (T)1) * 2;
    q[0] = (T)0.25 * s; q[1] = (m[7] - m[5]) / s; q[2] = (m[2] - m[6]) / s; q[3] = (m[3] - m[1]) / s;
  } else if (m[0] > m[4] && m[0] > m[8]) {
    T s = ox_sqrt((T)1 + m[0] - m[4] - m[8]) * 2;
    q[0] = (m[7] - m[5]) / s; q[1] = (T)0.25 * s; q[2] = (m[1] + m[3]) / s; q[3] = (m[2] + m[6]) / s;
  } else if (m[4] > m[8]) {
    T s = ox_sqrt((T)1 + m[4] - m[0] - m[8]) * 2;
    q[0] = (m[2] - m[6]) / s; q[1] = (m[1] + m[3]) / s; q[2] = (T)0.25 * s; q[3] = (m[5] + m[7]) / s;
  } else {
    T s = ox_sqrt((T)1 + m[8] - m[0] - m[4]) * 2;
    q[0] = (m[3] - m[1]) / s; q[1] = (m[2] + m[6]) / s; q[2] = (m[5] + m[7]) / s; q[3] = (T)0.25 * s;
  }
  normalize4(q);
}
// spatial vectors [angular; linear]
template <typename T> OX_HD void cross_motion(T* r, const T* vel, const T* v) {
  r[0] = -vel[2] * v[1] + vel[1] * v[2];
  r[1] = vel[2] * v[0] - vel[0] * v[2];
  r[2] = -vel[1] * v[0] + vel[0] * v[1];
  r[3] = -vel[2] * v[4] + vel[1] * v[5] - vel[5] * v[1] + vel[4] * v[2];
  r[4] = vel[2] * v[3] - vel[0] * v[5] + vel[5] * v[0] - vel[3] * v[2];
  r[5] = -vel[1] * v[3] + vel[0] * v[4] - vel[4] * v[0] + vel[3] * v[1];
}
template <typename T> OX_HD void cross_force(T* r, const T* vel, const T* f) {
  r[0] = -vel[2] * f[1] + vel[1] * f[2] - vel[5] * f[4] + vel[4] * f[5];
  r[1] = vel[2] * f[0] - vel[0] * f[2] + vel[5] * f[3] - vel[3] * f[5];
  r[2] = -vel[1] * f[0] + vel[0] * f[1] - vel[4] * f[3] + vel[3] * f[4];
  r[3] = -vel[2] * f[4] + vel[1] * f[5];
  r[4] = vel[2] * f[3] - vel[0] * f[5];
  r[5] = -vel[1] * f[3] + vel[0] * f[4];
}
template <typename T> OX_HD void mul_inert_vec(T* r, const T* i, const T* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
template <typename T> OX_HD T dot6(const T* a, const T* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

// Philox4x32-10 (control stream). KATs in tests/test_philox.py.
OX_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// model-table load into registers: dst[k] = m.table(first + k)
#define OX_LDM(N, dst, table, first)                         \
  _Pragma("unroll") for (int k_ = 0; k_ < (N); k_++) (dst)[k_] = m.table((first) + k_)

// Loops over the model's structure (bodies, joints, dofs, geoms, pairs, chains): fully unrolled when the model policy
// makes their bounds compile-time constants (LOCAL specialisation), left as loops in the generic kernels.
#define OX_MLOOP _Pragma("unroll")
// loops of the constraint solver over dofs: unrolled only for small specialised models (see Env::UNROLL_NV)
// loops over constraint rows in the solver: left to the compiler's own unrolling (x4). Forcing them rolled shrinks the Newton
// iteration body and speeds up the mean warp (K steps in one launch: 104 -> 126 M env-steps/s) but slows the slowest warp of
// every launch, which is what a one-step launch waits for (0.076 -> 0.087 ms) - measured, profiles/r1_notes.md
#define OX_ROWLOOP
#define OX_NVLOOP _Pragma("unroll (UNROLL_NV ? 64 : 1)")  // no count: full unroll iff the trip count is a compile-time constant, else none

// ---------------------------------------------------------------- one environment
// M     : model policy. DevModel<T> reads the runtime tables staged in shared memory; a generated Spec_* type
//         (ox_specgen) answers the same calls with compile-time constants so every model loop unrolls.
// LOCAL : false = fields live in the SoA batch arena ([element][env], coalesced); true = fields are per-thread
//         arrays (stride 1), which after unrolling become registers.
template <typename T, typename M = DevModel<T>, bool LOCAL = false>
struct Env {
  // Specialisation policy by model size: small models (cheetah-class) unroll the solver's O(nv^2)..O(nv^3) loops and use
  // static contact slots so that H, the vectors and the contact list are registers; large ones (humanoid-class) keep
  // those as loops over per-thread arrays - fully unrolled they are megabytes of straight-line code that the
  // instruction cache cannot hold (profiles/r1_notes.md).
  static constexpr bool UNROLL_NV = LOCAL && (M::Hdr::nv <= 12);
  static constexpr bool STATIC_CON = LOCAL && (M::Hdr::nconmax <= 24);
  M m;
  const DevBatch<T>& b;  // kernel parameter (arena), per-thread struct (specialised kernels) or per-group struct in shared memory (ox_coop.cu)
  int e;
  uint32_t S;  // env stride; every field index fits 32 bits (checked at batch creation)
  bool slots = false;  // contacts go to static slots pair_conadr(p) + k at run time too (cooperative kernel: lanes own pairs)

  OX_HD Env(const M& m_, const DevBatch<T>& b_, int e_) : m(m_), b(b_), e(LOCAL ? 0 : e_), S(LOCAL ? 1u : (uint32_t)b_.stride) {}

  // element i of a field for this env
  OX_HD T& at(T* f, int i) const { return LOCAL ? f[i] : f[(uint32_t)i * S + (uint32_t)e]; }
  OX_HD const T& at(const T* f, int i) const { return LOCAL ? f[i] : f[(uint32_t)i * S + (uint32_t)e]; }
  OX_HD int32_t& ati(int32_t* f, int i) const { return LOCAL ? f[i] : f[(uint32_t)i * S + (uint32_t)e]; }
  template <int N> OX_HD void ld(T* dst, const T* f, int first) const {
#pragma unroll
    for (int k = 0; k < N; k++) dst[k] = at(f, first + k);
  }
  template <int N> OX_HD void st(T* f, int first, const T* src) const {
#pragma unroll
    for (int k = 0; k < N; k++) at(f, first + k) = src[k];
  }
  OX_HD bool dis(int bit) const { return (m.h().disableflags & bit) != 0; }

  // ============================================================ A.1 kinematics (+ geoms, sites)
  // The stage functions below are loops over per-item pieces (kin_body, kin_geom, cinert_body, mass_row, vel_body, ...).
  // The thread-per-env kernels call the loops; the cooperative kernel (ox_coop.cu, one lane group per env) hands the items
  // of one loop to different lanes and synchronises between tree levels - the arithmetic is the same code.
  OX_HD void kin_body(int i) const {
    {
      T pos[3], quat[4], mat[9];
      if (i == 0) {
        pos[0] = pos[1] = pos[2] = 0;
        quat[0] = 1; quat[1] = quat[2] = quat[3] = 0;
      } else {
        const int jntadr = m.body_jntadr(i), jntnum = m.body_jntnum(i);
        const int mocapid = m.h().nmocap > 0 ? m.body_mocapid(i) : -1;
        if (mocapid >= 0) {   // mocap body: pose straight from mjData (mj_kinematics), quaternion normalised on the fly
          ld<3>(pos, b.mocap_pos, 3 * mocapid);
          ld<4>(quat, b.mocap_quat, 4 * mocapid);
          normalize4(quat);
        } else if (jntnum == 1 && m.jnt_type(jntadr) == OX_JNT_FREE) {
          const int qadr = m.jnt_qposadr(jntadr);
          ld<3>(pos, b.qpos, qadr);
          ld<4>(quat, b.qpos, qadr + 3);
          normalize4(quat);
          st<3>(b.xanchor, 3 * jntadr, pos);
          T ax[3];
          OX_LDM(3, ax, jnt_axis, 3 * jntadr);
          st<3>(b.xaxis, 3 * jntadr, ax);
        } else {
          const int pid = m.body_parentid(i);
          T ppos[3], pquat[4], pmat[9], bp[3], bq[4];
          ld<3>(ppos, b.xpos, 3 * pid);
          ld<4>(pquat, b.xquat, 4 * pid);
          ld<9>(pmat, b.xmat, 9 * pid);
          OX_LDM(3, bp, body_pos, 3 * i);
          OX_LDM(4, bq, body_quat, 4 * i);
          mat_vec3(pos, pmat, bp);
          pos[0] += ppos[0]; pos[1] += ppos[1]; pos[2] += ppos[2];
          mul_quat(quat, pquat, bq);
          OX_MLOOP
          for (int j = 0; j < jntnum; j++) {
            const int jid = jntadr + j, qadr = m.jnt_qposadr(jid), jt = m.jnt_type(jid);
            T jaxis[3], jpos[3], anchor[3], axis[3];
            OX_LDM(3, jaxis, jnt_axis, 3 * jid);
            OX_LDM(3, jpos, jnt_pos, 3 * jid);
            rot_vec_quat(axis, jaxis, quat);
            rot_vec_quat(anchor, jpos, quat);
            anchor[0] += pos[0]; anchor[1] += pos[1]; anchor[2] += pos[2];
            if (jt == OX_JNT_SLIDE) {
              T dq = at(b.qpos, qadr) - m.qpos0(qadr);
              pos[0] += axis[0] * dq; pos[1] += axis[1] * dq; pos[2] += axis[2] * dq;
            } else if (jt == OX_JNT_BALL || jt == OX_JNT_HINGE) {
              T qloc[4], qn[4], vec[3];
              if (jt == OX_JNT_BALL) {
                ld<4>(qloc, b.qpos, qadr);
                normalize4(qloc);
              } else {
                axis_angle2quat(qloc, jaxis, at(b.qpos, qadr) - m.qpos0(qadr));
              }
              mul_quat(qn, quat, qloc);
              quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];
              rot_vec_quat(vec, jpos, quat);
              pos[0] = anchor[0] - vec[0]; pos[1] = anchor[1] - vec[1]; pos[2] = anchor[2] - vec[2];
            }
            st<3>(b.xanchor, 3 * jid, anchor);
            st<3>(b.xaxis, 3 * jid, axis);
          }
        }
        normalize4(quat);
      }
      quat2mat(mat, quat);
      st<3>(b.xpos, 3 * i, pos);
      st<4>(b.xquat, 4 * i, quat);
      st<9>(b.xmat, 9 * i, mat);
      {  // inertial frame
        T ip[3], iq[4], v[3], q[4], im[9];
        OX_LDM(3, ip, body_ipos, 3 * i);
        OX_LDM(4, iq, body_iquat, 4 * i);
        mat_vec3(v, mat, ip);
        v[0] += pos[0]; v[1] += pos[1]; v[2] += pos[2];
        mul_quat(q, quat, iq);
        quat2mat(im, q);
        st<3>(b.xipos, 3 * i, v);
        st<9>(b.ximat, 9 * i, im);
      }
    }
  }
  OX_HD void kin_geom(int g) const {
    const int i = m.geom_bodyid(g);
    T pos[3], quat[4], mat[9], gp[3], gq[4], v[3], q[4], gm[9];
    ld<3>(pos, b.xpos, 3 * i); ld<4>(quat, b.xquat, 4 * i); ld<9>(mat, b.xmat, 9 * i);
    OX_LDM(3, gp, geom_pos, 3 * g);
    OX_LDM(4, gq, geom_quat, 4 * g);
    mat_vec3(v, mat, gp);
    v[0] += pos[0]; v[1] += pos[1]; v[2] += pos[2];
    mul_quat(q, quat, gq);
    quat2mat(gm, q);
    st<3>(b.geom_xpos, 3 * g, v);
    st<9>(b.geom_xmat, 9 * g, gm);
  }
  OX_HD void kin_site(int s) const {
    const int i = m.site_bodyid(s);
    T pos[3], quat[4], mat[9], sp[3], sq[4], v[3], q[4], sm[9];
    ld<3>(pos, b.xpos, 3 * i); ld<4>(quat, b.xquat, 4 * i); ld<9>(mat, b.xmat, 9 * i);
    OX_LDM(3, sp, site_pos, 3 * s);
    OX_LDM(4, sq, site_quat, 4 * s);
    mat_vec3(v, mat, sp);
    v[0] += pos[0]; v[1] += pos[1]; v[2] += pos[2];
    mul_quat(q, quat, sq);
    quat2mat(sm, q);
    st<3>(b.site_xpos, 3 * s, v);
    st<9>(b.site_xmat, 9 * s, sm);
  }
  OX_HDN void kinematics() const {
    const auto& h = m.h();
    OX_MLOOP
    for (int i = 0; i < h.nbody; i++) kin_body(i);
    OX_MLOOP
    for (int g = 0; g < h.ngeom; g++) kin_geom(g);
    OX_MLOOP
    for (int s = 0; s < h.nsite; s++) kin_site(s);
  }

  // ============================================================ A.2 subtree com, cinert, cdof
  OX_HDN void com_pos() const {
    const auto& h = m.h();
    const int nbody = h.nbody, njnt = h.njnt;
    OX_MLOOP
    for (int i = 0; i < 3 * nbody; i++) at(b.subtree_com, i) = 0;
    OX_MLOOP
    for (int i = nbody - 1; i >= 0; i--) {
      T acc[3], xi[3];
      ld<3>(acc, b.subtree_com, 3 * i);
      ld<3>(xi, b.xipos, 3 * i);
      const T mi = m.body_mass(i);
      acc[0] += xi[0] * mi; acc[1] += xi[1] * mi; acc[2] += xi[2] * mi;
      if (i) {
        const int j = m.body_parentid(i);
        T pa[3];
        ld<3>(pa, b.subtree_com, 3 * j);
        pa[0] += acc[0]; pa[1] += acc[1]; pa[2] += acc[2];
        st<3>(b.subtree_com, 3 * j, pa);
      }
      if (m.body_subtreemass(i) < (T)OX_MINVAL) { acc[0] = xi[0]; acc[1] = xi[1]; acc[2] = xi[2]; }
      else { const T sm = m.body_subtreemass(i); acc[0] /= sm; acc[1] /= sm; acc[2] /= sm; }
      st<3>(b.subtree_com, 3 * i, acc);
    }
    OX_MLOOP
    for (int k = 0; k < 10; k++) at(b.cinert, k) = 0;
    OX_MLOOP
    for (int i = 1; i < nbody; i++) cinert_body(i);
    OX_MLOOP
    for (int j = 0; j < njnt; j++) cdof_joint(j);
  }
  OX_HD void cinert_body(int i) const {
    {
      T mat[9], xi[3], sc[3], inert[3], dif[3], res[10], tmp[9];
      ld<9>(mat, b.ximat, 9 * i);
      ld<3>(xi, b.xipos, 3 * i);
      ld<3>(sc, b.subtree_com, 3 * m.body_rootid(i));
      OX_LDM(3, inert, body_inertia, 3 * i);
      const T ms = m.body_mass(i);
      dif[0] = xi[0] - sc[0]; dif[1] = xi[1] - sc[1]; dif[2] = xi[2] - sc[2];
      tmp[0] = mat[0] * inert[0]; tmp[3] = mat[1] * inert[1]; tmp[6] = mat[2] * inert[2];
      tmp[1] = mat[3] * inert[0]; tmp[4] = mat[4] * inert[1]; tmp[7] = mat[5] * inert[2];
      tmp[2] = mat[6] * inert[0]; tmp[5] = mat[7] * inert[1]; tmp[8] = mat[8] * inert[2];
      res[0] = mat[0] * tmp[0] + mat[1] * tmp[3] + mat[2] * tmp[6];
      res[1] = mat[3] * tmp[1] + mat[4] * tmp[4] + mat[5] * tmp[7];
      res[2] = mat[6] * tmp[2] + mat[7] * tmp[5] + mat[8] * tmp[8];
      res[3] = mat[0] * tmp[1] + mat[1] * tmp[4] + mat[2] * tmp[7];
      res[4] = mat[0] * tmp[2] + mat[1] * tmp[5] + mat[2] * tmp[8];
      res[5] = mat[3] * tmp[2] + mat[4] * tmp[5] + mat[5] * tmp[8];
      res[0] += ms * (dif[1] * dif[1] + dif[2] * dif[2]);
      res[1] += ms * (dif[0] * dif[0] + dif[2] * dif[2]);
      res[2] += ms * (dif[0] * dif[0] + dif[1] * dif[1]);
      res[3] -= ms * dif[0] * dif[1];
      res[4] -= ms * dif[0] * dif[2];
      res[5] -= ms * dif[1] * dif[2];
      res[6] = ms * dif[0]; res[7] = ms * dif[1]; res[8] = ms * dif[2];
      res[9] = ms;
      st<10>(b.cinert, 10 * i, res);
    }
  }
  OX_HD void cdof_joint(int j) const {
    {
      const int bi = m.jnt_bodyid(j), da = m.jnt_dofadr(j), jt = m.jnt_type(j);
      T sc[3], anchor[3], offset[3];
      ld<3>(sc, b.subtree_com, 3 * m.body_rootid(bi));
      ld<3>(anchor, b.xanchor, 3 * j);
      offset[0] = sc[0] - anchor[0]; offset[1] = sc[1] - anchor[1]; offset[2] = sc[2] - anchor[2];
      if (jt == OX_JNT_FREE || jt == OX_JNT_BALL) {
        int skip = 0;
        if (jt == OX_JNT_FREE) {
          OX_MLOOP
          for (int i = 0; i < 3; i++) {
            T cd[6] = {0, 0, 0, 0, 0, 0};
            cd[3 + i] = 1;
            st<6>(b.cdof, 6 * (da + i), cd);
          }
          skip = 3;
        }
        T xm[9];
        ld<9>(xm, b.xmat, 9 * bi);
        OX_MLOOP
        for (int i = 0; i < 3; i++) {
          T cd[6];
          cd[0] = xm[i]; cd[1] = xm[i + 3]; cd[2] = xm[i + 6];
          cross3(cd + 3, cd, offset);
          st<6>(b.cdof, 6 * (da + i + skip), cd);
        }
      } else {
        T ax[3], cd[6];
        ld<3>(ax, b.xaxis, 3 * j);
        if (jt == OX_JNT_SLIDE) {
          cd[0] = cd[1] = cd[2] = 0;
          cd[3] = ax[0]; cd[4] = ax[1]; cd[5] = ax[2];
        } else {
          cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
          cross3(cd + 3, ax, offset);
        }
        st<6>(b.cdof, 6 * da, cd);
      }
    }
  }

  // ============================================================ A.3 composite rigid body -> qM
  OX_HDN void crb() const {
    const auto& h = m.h();
    const int nbody = h.nbody, nv = h.nv;
    OX_MLOOP
    for (int i = 0; i < 10 * nbody; i++) at(b.crb, i) = at(b.cinert, i);
    OX_MLOOP
    for (int i = nbody - 1; i > 0; i--) {
      const int p = m.body_parentid(i);
      if (p > 0) {
        T a[10], c[10];
        ld<10>(a, b.crb, 10 * p);
        ld<10>(c, b.crb, 10 * i);
#pragma unroll
        for (int k = 0; k < 10; k++) a[k] += c[k];
        st<10>(b.crb, 10 * p, a);
      }
    }
    OX_MLOOP
    for (int i = 0; i < nv; i++) mass_row(i);
  }
  OX_HD void mass_row(int i) const {
    {
      T in[10], cd[6], buf[6];
      ld<10>(in, b.crb, 10 * m.dof_bodyid(i));
      ld<6>(cd, b.cdof, 6 * i);
      mul_inert_vec(buf, in, cd);
      int adr = m.dof_Madr(i);
      at(b.qM, adr++) = m.dof_armature(i) + dot6(cd, buf);
      OX_MLOOP
      for (int d_ = 1, j = m.dof_parentid(i); d_ < m.dof_depth(i); d_++, j = m.dof_parentid(j)) {
        T cj[6];
        ld<6>(cj, b.cdof, 6 * j);
        at(b.qM, adr++) = dot6(cj, buf);
      }
    }
  }

  // ============================================================ A.4 sparse L'DL factor / solve / M*v
  OX_HDN void factor_ld(T* qLD, T* qLDiagInv) const {
    const auto& h = m.h();
    const int nv = h.nv;
    OX_MLOOP
    for (int k = nv - 1; k >= 0; k--) {
      const int Madr_kk = m.dof_Madr(k);
      int Madr_ki = Madr_kk + 1, i = m.dof_parentid(k);
      const T inv = (T)1 / at(qLD, Madr_kk);  // one reciprocal per pivot; the row scalings multiply by it
      OX_MLOOP
      for (int d_ = 1; d_ < m.dof_depth(k); d_++) {
        const T tmp = at(qLD, Madr_ki) * inv;
        const int rowi = m.dof_Madr(i), n = m.dof_depth(i);
        OX_MLOOP
        for (int c = 0; c < n; c++) at(qLD, rowi + c) -= tmp * at(qLD, Madr_ki + c);
        at(qLD, Madr_ki) = tmp;
        i = m.dof_parentid(i);
        Madr_ki++;
      }
      at(qLDiagInv, k) = inv;
    }
  }
  OX_HDN void factor_m() const {
    const int nM = m.h().nM;
    OX_MLOOP
    for (int i = 0; i < nM; i++) at(b.qLD, i) = at(b.qM, i);
    factor_ld(b.qLD, b.qLDiagInv);
  }
  OX_HDN void solve_ld(T* x) const {
    const int nv = m.h().nv;
    OX_NVLOOP
    for (int i = nv - 1; i >= 0; i--) {
      int adr = m.dof_Madr(i) + 1;
      const T xi = at(x, i);
      OX_NVLOOP
      for (int d_ = 1, j = m.dof_parentid(i); d_ < m.dof_depth(i); d_++, j = m.dof_parentid(j)) at(x, j) -= at(b.qLD, adr++) * xi;
    }
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(x, i) *= at(b.qLDiagInv, i);
    OX_NVLOOP
    for (int i = 0; i < nv; i++) {
      int adr = m.dof_Madr(i) + 1;
      T xi = at(x, i);
      OX_NVLOOP
      for (int d_ = 1, j = m.dof_parentid(i); d_ < m.dof_depth(i); d_++, j = m.dof_parentid(j)) xi -= at(b.qLD, adr++) * at(x, j);
      at(x, i) = xi;
    }
  }
  OX_HDN void mul_m(T* res, const T* v) const {
    const int nv = m.h().nv;
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(res, i) = 0;
    OX_NVLOOP
    for (int i = 0; i < nv; i++) {
      int adr = m.dof_Madr(i);
      const T vi = at(v, i);
      T ri = at(res, i) + at(b.qM, adr++) * vi;
      OX_NVLOOP
      for (int d_ = 1, j = m.dof_parentid(i); d_ < m.dof_depth(i); d_++, j = m.dof_parentid(j)) {
        const T mij = at(b.qM, adr++);
        ri += mij * at(v, j);
        at(res, j) += mij * vi;
      }
      at(res, i) = ri;
    }
  }

  // ============================================================ A.7 com velocities
  OX_HDN void com_vel() const {
    const auto& h = m.h();
    const int nbody = h.nbody;
    OX_MLOOP
    for (int k = 0; k < 6; k++) at(b.cvel, k) = 0;
    OX_MLOOP
    for (int i = 1; i < nbody; i++) vel_body(i);
  }
  OX_HD void vel_body(int i) const {
    {
      T cvel[6];
      ld<6>(cvel, b.cvel, 6 * m.body_parentid(i));
      const int jntadr = m.body_jntadr(i), jntnum = m.body_jntnum(i);
      OX_MLOOP
      for (int jj = 0; jj < jntnum; jj++) {
        const int jid = jntadr + jj, jt = m.jnt_type(jid);
        int da = m.jnt_dofadr(jid);
        if (jt == OX_JNT_FREE) {  // translational dofs: cdof_dot = 0, velocity added first
          OX_MLOOP
          for (int k = 0; k < 3; k++) {
            T z[6] = {0, 0, 0, 0, 0, 0}, cd[6];
            st<6>(b.cdof_dot, 6 * (da + k), z);
            ld<6>(cd, b.cdof, 6 * (da + k));
            const T qv = at(b.qvel, da + k);
#pragma unroll
            for (int c = 0; c < 6; c++) cvel[c] += cd[c] * qv;
          }
          da += 3;
        }
        if (jt == OX_JNT_FREE || jt == OX_JNT_BALL) {  // all three cdof_dot with the same incoming velocity
          T cd[3][6];
          OX_MLOOP
          for (int k = 0; k < 3; k++) {
            T cdd[6];
            ld<6>(cd[k], b.cdof, 6 * (da + k));
            cross_motion(cdd, cvel, cd[k]);
            st<6>(b.cdof_dot, 6 * (da + k), cdd);
          }
          OX_MLOOP
          for (int k = 0; k < 3; k++) {
            const T qv = at(b.qvel, da + k);
#pragma unroll
            for (int c = 0; c < 6; c++) cvel[c] += cd[k][c] * qv;
          }
        } else {
          T cd[6], cdd[6];
          ld<6>(cd, b.cdof, 6 * da);
          cross_motion(cdd, cvel, cd);
          st<6>(b.cdof_dot, 6 * da, cdd);
          const T qv = at(b.qvel, da);
#pragma unroll
          for (int c = 0; c < 6; c++) cvel[c] += cd[c] * qv;
        }
      }
      st<6>(b.cvel, 6 * i, cvel);
    }
  }

  // ============================================================ passive forces
  // mj_tendon: length and Jacobian row (ten_J) of tendon i. Fixed: length = sum of coef * joint coordinate, J = the coefficients.
  // Spatial: straight segments through sites, length = sum |p_{k+1} - p_k|, J = sum dir_k' (Jp(site_{k+1}) - Jp(site_k)).
  OX_HD void tendon_length(int i) const {
    const int nv = m.h().nv, adr = m.tendon_adr(i);
    T L = 0;
    OX_MLOOP
    for (int k = 0; k < nv; k++) at(b.ten_J, i * nv + k) = 0;
    if (m.tendon_type(i) == OX_TEN_FIXED) {
      OX_MLOOP
      for (int w = 0; w < m.tendon_num(i); w++) {
        const int j = m.wrap_objid(adr + w);
        L += m.wrap_prm(adr + w) * at(b.qpos, m.jnt_qposadr(j));
        at(b.ten_J, i * nv + m.jnt_dofadr(j)) += m.wrap_prm(adr + w);
      }
    } else {
      OX_MLOOP
      for (int w = 0; w + 1 < m.tendon_num(i); w++) {
        const int sites[2] = {m.wrap_objid(adr + w), m.wrap_objid(adr + w + 1)};
        T pa[3], pb[3], dir[3];
        ld<3>(pa, b.site_xpos, 3 * sites[0]);
        ld<3>(pb, b.site_xpos, 3 * sites[1]);
        dir[0] = pb[0] - pa[0]; dir[1] = pb[1] - pa[1]; dir[2] = pb[2] - pa[2];
        const T len = ox_sqrt(dot3(dir, dir));
        L += len;
        if (len < (T)OX_MINVAL) continue;
        dir[0] /= len; dir[1] /= len; dir[2] /= len;
        OX_MLOOP
        for (int s = 0; s < 2; s++) {   // + Jp(site b) - Jp(site a), projected on the segment direction
          const T sign = s == 0 ? (T)-1 : (T)1;
          const T* pt = s == 0 ? pa : pb;
          const int sb = m.site_bodyid(sites[s]);
          const int body = m.body_weldid(sb);
          if (!body) continue;
          T sc[3], offset[3];
          ld<3>(sc, b.subtree_com, 3 * m.body_rootid(sb));
          offset[0] = pt[0] - sc[0]; offset[1] = pt[1] - sc[1]; offset[2] = pt[2] - sc[2];
          const int last_ = m.body_dofadr(body) + m.body_dofnum(body) - 1;
          OX_MLOOP
          for (int d_ = 0, dof = last_; d_ < m.dof_depth(last_); d_++, dof = m.dof_parentid(dof)) {
            T cd[6], jp[3];
            ld<6>(cd, b.cdof, 6 * dof);
            cross3(jp, cd, offset);
            jp[0] += cd[3]; jp[1] += cd[4]; jp[2] += cd[5];
            at(b.ten_J, i * nv + dof) += sign * dot3(dir, jp);
          }
        }
      }
    }
    at(b.ten_length, i) = L;
  }
  OX_HD T tendon_velocity(int i) const {
    const int nv = m.h().nv;
    T v = 0;
    OX_MLOOP
    for (int k = 0; k < nv; k++) v += at(b.ten_J, i * nv + k) * at(b.qvel, k);
    return v;
  }
  // spring with dead band [lengthspring0, lengthspring1] and damper, mapped to the joints through J'
  OX_HD void passive_tendon(int i) const {
    const T L = at(b.ten_length, i), lo = m.tendon_lengthspring(2 * i), hi = m.tendon_lengthspring(2 * i + 1);
    T f = 0;
    if (L > hi) f = m.tendon_stiffness(i) * (hi - L);
    else if (L < lo) f = m.tendon_stiffness(i) * (lo - L);
    f -= m.tendon_damping(i) * tendon_velocity(i);
    if (f == 0) return;
    const int nv = m.h().nv;
    OX_MLOOP
    for (int k = 0; k < nv; k++) at(b.qfrc_passive, k) += at(b.ten_J, i * nv + k) * f;
  }
  OX_HDN void passive() const {
    const auto& h = m.h();
    const int nv = h.nv, njnt = h.njnt;
    OX_MLOOP
    for (int i = 0; i < nv; i++) at(b.qfrc_passive, i) = 0;
    OX_MLOOP
    for (int i = 0; i < h.ntendon; i++) tendon_length(i);
    if (dis(OX_DSBL_PASSIVE)) return;
    OX_MLOOP
    for (int j = 0; j < njnt; j++) passive_joint(j);
    OX_MLOOP
    for (int i = 0; i < nv; i++) at(b.qfrc_passive, i) -= m.dof_damping(i) * at(b.qvel, i);
    OX_MLOOP
    for (int i = 0; i < h.ntendon; i++) passive_tendon(i);
    OX_MLOOP
    for (int bd = 1; bd < h.nfluid; bd++) passive_fluid(bd);
    if (!dis(OX_DSBL_GRAVITY)) {
      OX_MLOOP
      for (int bd = 1; bd < h.ngravcomp; bd++) passive_gravcomp(bd);
    }
  }
  // body gravcomp: the force -gravity * mass * gravcomp at the body's com, mapped to the joints (mj_passive, qfrc_gravcomp)
  OX_HD void passive_gravcomp(int bd) const {
    const T gc = m.body_gravcomp(bd) * m.body_mass(bd);
    const int body = m.body_weldid(bd);
    if (gc == 0 || !body) return;
    const auto& h = m.h();
    T xi[3], sc[3], offset[3], f[3] = {-(T)h.grav(0) * gc, -(T)h.grav(1) * gc, -(T)h.grav(2) * gc};
    ld<3>(xi, b.xipos, 3 * bd);
    ld<3>(sc, b.subtree_com, 3 * m.body_rootid(bd));
    offset[0] = xi[0] - sc[0]; offset[1] = xi[1] - sc[1]; offset[2] = xi[2] - sc[2];
    const int last_ = m.body_dofadr(body) + m.body_dofnum(body) - 1;
    OX_MLOOP
    for (int d_ = 0, i = last_; d_ < m.dof_depth(last_); d_++, i = m.dof_parentid(i)) {
      T cd[6], jp[3];
      ld<6>(cd, b.cdof, 6 * i);
      cross3(jp, cd, offset);
      jp[0] += cd[3]; jp[1] += cd[4]; jp[2] += cd[5];
      at(b.qfrc_passive, i) += dot3(jp, f);
    }
  }
  // mj_inertiaBoxFluidModel: drag of the medium on the equivalent inertia box of body bd. The body's velocity at its com in
  // the inertial frame (ximat), minus the wind; a viscous term linear in it (equivalent sphere) and a quadratic term face by
  // face (coefficients: body_fluid, compiled from density / viscosity / the box sizes); the wrench is applied at xipos
  OX_HD void passive_fluid(int bd) const {
    if (m.body_mass(bd) < (T)OX_MINVAL) return;
    const int body = m.body_weldid(bd);
    if (!body) return;
    T cv[6], xi[3], sc[3], offset[3], R[9], vel[3], lw[3], lv[3], lt[3], lf[3], f[6];
    ld<6>(cv, b.cvel, 6 * bd);
    ld<3>(xi, b.xipos, 3 * bd);
    ld<3>(sc, b.subtree_com, 3 * m.body_rootid(bd));
    ld<9>(R, b.ximat, 9 * bd);
    offset[0] = xi[0] - sc[0]; offset[1] = xi[1] - sc[1]; offset[2] = xi[2] - sc[2];
    cross3(vel, cv, offset);   // cvel is expressed at the subtree com: v(xipos) = v + w x offset
    OX_MLOOP
    for (int k = 0; k < 3; k++) vel[k] += cv[3 + k] - m.body_fluid(11 * bd + 8 + k);
    OX_MLOOP
    for (int k = 0; k < 3; k++) {   // into the inertial frame: R' w, R' v
      lw[k] = R[k] * cv[0] + R[3 + k] * cv[1] + R[6 + k] * cv[2];
      lv[k] = R[k] * vel[0] + R[3 + k] * vel[1] + R[6 + k] * vel[2];
    }
    OX_MLOOP
    for (int k = 0; k < 3; k++) {
      lt[k] = -m.body_fluid(11 * bd) * lw[k] - m.body_fluid(11 * bd + 5 + k) * ox_abs(lw[k]) * lw[k];
      lf[k] = -m.body_fluid(11 * bd + 1) * lv[k] - m.body_fluid(11 * bd + 2 + k) * ox_abs(lv[k]) * lv[k];
    }
    OX_MLOOP
    for (int k = 0; k < 3; k++) {   // back to the world frame: (force, torque)
      f[k] = R[3 * k] * lf[0] + R[3 * k + 1] * lf[1] + R[3 * k + 2] * lf[2];
      f[3 + k] = R[3 * k] * lt[0] + R[3 * k + 1] * lt[1] + R[3 * k + 2] * lt[2];
    }
    const int last_ = m.body_dofadr(body) + m.body_dofnum(body) - 1;
    OX_MLOOP
    for (int d_ = 0, i = last_; d_ < m.dof_depth(last_); d_++, i = m.dof_parentid(i)) {
      T cd[6], jp[3];
      ld<6>(cd, b.cdof, 6 * i);
      cross3(jp, cd, offset);
      jp[0] += cd[3]; jp[1] += cd[4]; jp[2] += cd[5];
      at(b.qfrc_passive, i) += dot3(jp, f) + dot3(cd, f + 3);
    }
  }
  OX_HD void passive_joint(int j) const {  // joint spring: touches only this joint's dofs
    {
      const T k = m.jnt_stiffness(j);
      if (k == 0) return;
      int pa = m.jnt_qposadr(j), da = m.jnt_dofadr(j);
      const int jt = m.jnt_type(j);
      if (jt == OX_JNT_FREE) {
        OX_MLOOP
        for (int c = 0; c < 3; c++) at(b.qfrc_passive, da + c) -= k * (at(b.qpos, pa + c) - m.qpos_spring(pa + c));
        pa += 3; da += 3;
      }
      if (jt == OX_JNT_FREE || jt == OX_JNT_BALL) {
        T q[4], qs[4], dif[3];
        ld<4>(q, b.qpos, pa);
        normalize4(q);
        OX_LDM(4, qs, qpos_spring, pa);
        sub_quat(dif, q, qs);
        OX_MLOOP
        for (int c = 0; c < 3; c++) at(b.qfrc_passive, da + c) -= k * dif[c];
      } else {
        at(b.qfrc_passive, da) -= k * (at(b.qpos, pa) - m.qpos_spring(pa));
      }
    }
  }

  // ============================================================ A.8 bias forces (RNE, no acceleration)
  OX_HDN void rne() const {
    const auto& h = m.h();
    const int nbody = h.nbody, nv = h.nv;
    {
      T a0[6] = {0, 0, 0, 0, 0, 0}, z[6] = {0, 0, 0, 0, 0, 0};
      if (!dis(OX_DSBL_GRAVITY)) { a0[3] = -(T)h.grav(0); a0[4] = -(T)h.grav(1); a0[5] = -(T)h.grav(2); }
      st<6>(b.cacc, 0, a0);
      st<6>(b.cfrc, 0, z);
    }
    OX_MLOOP
    for (int i = 1; i < nbody; i++) rne_fwd_body(i);
    OX_MLOOP
    for (int i = nbody - 1; i > 0; i--) {
      const int p = m.body_parentid(i);
      if (p) {
        T a[6], c[6];
        ld<6>(a, b.cfrc, 6 * p);
        ld<6>(c, b.cfrc, 6 * i);
#pragma unroll
        for (int k = 0; k < 6; k++) a[k] += c[k];
        st<6>(b.cfrc, 6 * p, a);
      }
    }
    OX_MLOOP
    for (int i = 0; i < nv; i++) rne_bias_dof(i);
  }
  OX_HD void rne_bias_dof(int i) const {
    T cd[6], f[6];
    ld<6>(cd, b.cdof, 6 * i);
    ld<6>(f, b.cfrc, 6 * m.dof_bodyid(i));
    at(b.qfrc_bias, i) = dot6(cd, f);
  }
  OX_HD void rne_fwd_body(int i) const {
    {
      const int bda = m.body_dofadr(i), nd = m.body_dofnum(i);
      T cacc[6], in[10], cv[6], f[6], tmp[6], tmp1[6];
      ld<6>(cacc, b.cacc, 6 * m.body_parentid(i));
      OX_MLOOP
      for (int j = 0; j < nd; j++) {
        T cdd[6];
        ld<6>(cdd, b.cdof_dot, 6 * (bda + j));
        const T qv = at(b.qvel, bda + j);
#pragma unroll
        for (int c = 0; c < 6; c++) cacc[c] += cdd[c] * qv;
      }
      st<6>(b.cacc, 6 * i, cacc);
      ld<10>(in, b.cinert, 10 * i);
      ld<6>(cv, b.cvel, 6 * i);
      mul_inert_vec(f, in, cacc);
      mul_inert_vec(tmp, in, cv);
      cross_force(tmp1, cv, tmp);
#pragma unroll
      for (int c = 0; c < 6; c++) f[c] += tmp1[c];
      st<6>(b.cfrc, 6 * i, f);
    }
  }

  // ============================================================ A.9 actuation (joint transmission)
  OX_HDN void actuation() const {
    const auto& h = m.h();
    const int nv = h.nv, nu = h.nu;
    OX_MLOOP
    for (int i = 0; i < nv; i++) at(b.qfrc_actuator, i) = 0;
    OX_MLOOP
    for (int i = 0; i < nu; i++) {
      const T gf = actuator_one(i);   // gear * force (0 when actuation is disabled)
      if (m.actuator_trntype(i) == OX_TRN_TENDON) {   // moment = gear * tendon Jacobian (the coefficient vector)
        const int tn = m.actuator_trnid(i);
        OX_MLOOP
        for (int k = 0; k < nv; k++) at(b.qfrc_actuator, k) += gf * at(b.ten_J, tn * nv + k);
      } else {
        at(b.qfrc_actuator, m.jnt_dofadr(m.actuator_trnid(i))) += gf;
      }
    }
  }
  // transmission length / velocity of actuator i without the gear (joint coordinate, or fixed-tendon length)
  OX_HD T act_length(int i) const {
    if (m.actuator_trntype(i) == OX_TRN_TENDON) return at(b.ten_length, m.actuator_trnid(i));
    return at(b.qpos, m.jnt_qposadr(m.actuator_trnid(i)));
  }
  OX_HD T act_velocity(int i) const {
    if (m.actuator_trntype(i) == OX_TRN_TENDON) return tendon_velocity(m.actuator_trnid(i));
    return at(b.qvel, m.jnt_dofadr(m.actuator_trnid(i)));
  }
  // force of one actuator -> actuator_force[i] (and act_dot for stateful ones); returns its generalised force gear * force
  OX_HD T actuator_one(int i) const {
    const bool off = dis(OX_DSBL_ACTUATION);
    const bool clamp = !dis(OX_DSBL_CLAMPCTRL);
    {
      if (off) { at(b.actuator_force, i) = 0; return (T)0; }
      const T gear = m.actuator_gear(i);
      const T length = gear * act_length(i), velocity = gear * act_velocity(i);
      T ctrl = at(b.ctrl, i);
      if (m.actuator_ctrllimited(i) && clamp) ctrl = ox_clip(ctrl, m.actuator_ctrlrange(2 * i), m.actuator_ctrlrange(2 * i + 1));
      // stateful actuators (mjtDyn integrator / filter / filterexact): the force is driven by the activation state, and the
      // control sets its time derivative (mj_fwdActuation); mj_advance integrates it (next_activation below)
      const int dyn = m.actuator_dyntype(i);
      if (dyn != OX_DYN_NONE) {
        const int aa = m.actuator_actadr(i);
        const T act = at(b.act, aa);
        at(b.act_dot, aa) = dyn == OX_DYN_INTEGRATOR ? ctrl : (ctrl - act) / ox_max((T)OX_MINVAL, m.actuator_dynprm(3 * i));
        ctrl = act;
      }
      T gp[3], bp[3];
      OX_LDM(3, gp, actuator_gainprm, 3 * i);
      OX_LDM(3, bp, actuator_biasprm, 3 * i);
      T gain = gp[0];
      if (m.actuator_gaintype(i) == OX_GAIN_AFFINE) gain += gp[1] * length + gp[2] * velocity;
      T bias = 0;
      if (m.actuator_biastype(i) == OX_BIAS_AFFINE) bias = bp[0] + bp[1] * length + bp[2] * velocity;
      T force = gain * ctrl + bias;
      if (m.actuator_forcelimited(i)) force = ox_clip(force, m.actuator_forcerange(2 * i), m.actuator_forcerange(2 * i + 1));
      at(b.actuator_force, i) = force;
      return gear * force;
    }
  }

  // ============================================================ A.10 smooth acceleration
  OX_HDN void fwd_acceleration() const {
    const auto& h = m.h();
    const int nv = h.nv;
    OX_MLOOP
    for (int i = 0; i < nv; i++)
      at(b.qfrc_smooth, i) = at(b.qfrc_passive, i) - at(b.qfrc_bias, i) + at(b.qfrc_applied, i) + at(b.qfrc_actuator, i);
    apply_xfrc();
    OX_MLOOP
    for (int i = 0; i < nv; i++) at(b.qacc_smooth, i) = at(b.qfrc_smooth, i);
    solve_ld(b.qacc_smooth);
  }
  // qfrc_smooth += J' xfrc_applied (Cartesian forces / torques at the body com)
  OX_HD void apply_xfrc() const {
    const auto& h = m.h();
    const int nbody = h.nbody;
    // Branch hygiene (here and in make_constraint / collision): a skipped block is a taken branch to a far target, which
    // in these large straight-line kernels is an instruction-cache miss served from L2 (~500 cycles, profiles/r1_notes.md).
    // So test the common "nothing to do" case once for the whole loop instead of once per body / joint / contact site.
    bool any_xfrc = false;
    OX_MLOOP
    for (int i = 6; i < 6 * nbody; i++) any_xfrc |= at(b.xfrc_applied, i) != 0;
    OX_MLOOP
    for (int bd = 1; any_xfrc && bd < nbody; bd++) {
      T f[6];
      ld<6>(f, b.xfrc_applied, 6 * bd);
      if (f[0] == 0 && f[1] == 0 && f[2] == 0 && f[3] == 0 && f[4] == 0 && f[5] == 0) continue;
      T xi[3], sc[3], offset[3];
      ld<3>(xi, b.xipos, 3 * bd);
      ld<3>(sc, b.subtree_com, 3 * m.body_rootid(bd));
      offset[0] = xi[0] - sc[0]; offset[1] = xi[1] - sc[1]; offset[2] = xi[2] - sc[2];
      int body = bd;
      body = m.body_weldid(body);  // nearest ancestor-or-self that has dofs (0 = static)
      if (!body) continue;
      const int last_ = m.body_dofadr(body) + m.body_dofnum(body) - 1;
      OX_MLOOP
      for (int d_ = 0, i = last_; d_ < m.dof_depth(last_); d_++, i = m.dof_parentid(i)) {
        T cd[6], jp[3];
        ld<6>(cd, b.cdof, 6 * i);
        cross3(jp, cd, offset);
        jp[0] += cd[3]; jp[1] += cd[4]; jp[2] += cd[5];
        at(b.qfrc_smooth, i) += dot3(jp, f) + dot3(cd, f + 3);
      }
    }
  }

  // ============================================================ A.5 collision (precompiled pair list)
  struct Con { T dist, pos[3], frame[9]; };

  static OX_HD int plane_sphere(Con& c, T margin, const T* pos1, const T* n, const T* pos2, T radius) {
    T tmp[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
    T cdist = dot3(tmp, n);
    if (cdist > margin + radius) return 0;
    c.dist = cdist - radius;
    c.frame[0] = n[0]; c.frame[1] = n[1]; c.frame[2] = n[2];
    c.frame[3] = 0; c.frame[4] = 0; c.frame[5] = 0;
    const T s = -c.dist / 2 - radius;
    c.pos[0] = pos2[0] + n[0] * s; c.pos[1] = pos2[1] + n[1] * s; c.pos[2] = pos2[2] + n[2] * s;
    return 1;
  }
  static OX_HD int sphere_sphere(Con& c, T margin, const T* pos1, T r1, const T* pos2, T r2) {
    T dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
    T cd2 = dot3(dif, dif), mind = margin + r1 + r2;
    if (cd2 > mind * mind) return 0;
    c.dist = ox_sqrt(cd2) - r1 - r2;
    c.frame[0] = dif[0]; c.frame[1] = dif[1]; c.frame[2] = dif[2];
    normalize3(c.frame);
    c.frame[3] = 0; c.frame[4] = 0; c.frame[5] = 0;
    const T s = r1 + c.dist / 2;
    c.pos[0] = pos1[0] + c.frame[0] * s; c.pos[1] = pos1[1] + c.frame[1] * s; c.pos[2] = pos1[2] + c.frame[2] * s;
    return 1;
  }
  // sphere against a box (mjc_SphereBox): closest box point = the centre clamped to the box in the box frame; a centre inside the
  // box leaves through the nearest face. Normal from the sphere (geom1) to the box (geom2).
  static OX_HD int sphere_box(Con& c, T margin, const T* centre, T radius, const T* pos2, const T* mat2, const T* size2) {
    const T tmp[3] = {centre[0] - pos2[0], centre[1] - pos2[1], centre[2] - pos2[2]};
    T lc[3], clamped[3], dir[3], lpos[3], lnorm[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      lc[k] = mat2[k] * tmp[0] + mat2[3 + k] * tmp[1] + mat2[6 + k] * tmp[2];
      clamped[k] = ox_clip(lc[k], -size2[k], size2[k]);
      dir[k] = clamped[k] - lc[k];
    }
    const T dist = ox_sqrt(dot3(dir, dir));
    if (dist - radius > margin) return 0;
    if (dist <= (T)OX_MINVAL) {
      T closest = 2 * ox_max(size2[0], ox_max(size2[1], size2[2]));
      int face = 0;
#pragma unroll
      for (int i = 0; i < 6; i++) {
        const T fd = ox_abs((i % 2 ? (T)1 : (T)-1) * size2[i / 2] - lc[i / 2]);
        if (fd < closest) { closest = fd; face = i; }
      }
#pragma unroll
      for (int k = 0; k < 3; k++) lnorm[k] = (k == face / 2) ? (face % 2 ? (T)-1 : (T)1) : (T)0;
#pragma unroll
      for (int k = 0; k < 3; k++) lpos[k] = lc[k] + lnorm[k] * (radius - closest) / 2;
      c.dist = -closest - radius;
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) lnorm[k] = dir[k] / dist;
#pragma unroll
      for (int k = 0; k < 3; k++) lpos[k] = (T)0.5 * (clamped[k] + lc[k] + lnorm[k] * radius);
      c.dist = dist - radius;
    }
    mat_vec3(c.frame, mat2, lnorm);
    c.frame[3] = 0; c.frame[4] = 0; c.frame[5] = 0;
    T w[3];
    mat_vec3(w, mat2, lpos);
    c.pos[0] = w[0] + pos2[0]; c.pos[1] = w[1] + pos2[1]; c.pos[2] = w[2] + pos2[2];
    return 1;
  }
  // parameter t in [-1, 1] of the point of the segment lc + t la (box frame) closest to the box: zero crossing of the piecewise
  // linear, non-decreasing derivative g(t) of half the squared distance (ORACLE_DECISIONS.md #12)
  static OX_HD T segment_box_closest(const T* lc, const T* la, const T* size2) {
    auto g = [&](T t) {
      T s = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) { const T pk = lc[k] + t * la[k]; s += la[k] * (pk - ox_clip(pk, -size2[k], size2[k])); }
      return s;
    };
    {  // the axis itself cuts through the box (deep penetration): the middle of the cut, so that sphere_box's nearest-face rule
       // sees a point well inside
      T tin = -1, tout = 1;
      bool cut = true;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (ox_abs(la[k]) > (T)OX_MINVAL) {
          T t1 = (-size2[k] - lc[k]) / la[k], t2 = (size2[k] - lc[k]) / la[k];
          if (t1 > t2) { const T sw = t1; t1 = t2; t2 = sw; }
          tin = ox_max(tin, t1); tout = ox_min(tout, t2);
        } else if (ox_abs(lc[k]) > size2[k]) cut = false;
      }
      if (cut && tin <= tout) return (T)0.5 * (tin + tout);
    }
    const T gm = g((T)-1), gp = g((T)1);
    if (gm >= 0) return -1;
    if (gp <= 0) return 1;
    T ta = -1, ga = gm, tb = 1, gb = gp;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (!(ox_abs(la[k]) > (T)OX_MINVAL)) continue;
#pragma unroll
      for (int sg = -1; sg <= 1; sg += 2) {
        const T t = ((T)sg * size2[k] - lc[k]) / la[k];
        if (!(t > -1 && t < 1)) continue;
        const T gi = g(t);
        if (gi < 0) { if (t > ta) { ta = t; ga = gi; } }
        else if (t < tb) { tb = t; gb = gi; }
      }
    }
    return gb - ga > (T)OX_MINVAL ? ta - ga * (tb - ta) / (gb - ga) : ta;
  }
  static OX_HD void make_frame(T* frame) {
    normalize3(frame);
    if (ox_sqrt(dot3(frame + 3, frame + 3)) < (T)0.5) {
      frame[3] = 0; frame[4] = 0; frame[5] = 0;
      if (frame[1] < (T)0.5 && frame[1] > (T)-0.5) frame[4] = 1; else frame[5] = 1;
    }
    const T t = dot3(frame, frame + 3);
    frame[3] -= t * frame[0]; frame[4] -= t * frame[1]; frame[5] -= t * frame[2];
    normalize3(frame + 3);
    cross3(frame + 6, frame, frame + 3);
  }
  // Store one contact. Generic kernels compact (MuJoCo's contact list); the specialised kernels give every narrowphase
  // test site k of pair p its own slot pair_conadr(p) + k, so that all contact indices are compile-time constants and the
  // contact arrays live in registers; walking the slots in order visits the contacts in the same order as the compact list.
  OX_HD void emit(Con& c, int p, int k, int& ncon) const {
    make_frame(c.frame);
    const int idx = (STATIC_CON || slots) ? m.pair_conadr(p) + k : ncon;
    at(b.con_dist, idx) = c.dist;
    st<3>(b.con_pos, 3 * idx, c.pos);
    st<9>(b.con_frame, 9 * idx, c.frame);
    ati(b.con_pair, idx) = p;
    if (STATIC_CON || slots) ati(b.con_active, idx) = 1;
    ncon++;
  }

  OX_HDN void collision() const {
    const auto& h = m.h();
    int ncon = 0;
    if (STATIC_CON) {
      OX_MLOOP
      for (int i = 0; i < h.nconmax; i++) ati(b.con_active, i) = 0;
    }
    if (!(dis(OX_DSBL_CONTACT) || dis(OX_DSBL_CONSTRAINT))) {
      const int npair = h.npair;
      OX_MLOOP
      for (int p = 0; p < npair; p++) collide_pair(p, ncon);
    }
    ati(b.ncon, 0) = ncon;
  }
  // box against box (ORACLE_DECISIONS.md #32; not mjc_BoxBox): separating-axis test over 6 face normals and 9 edge cross products;
  // face axis -> the opposed face of the other box clipped against the reference face's side planes, every clipped vertex within
  // the margin is a contact (<= 8); edge axis -> one contact between the closest points of the supporting edges. Faces win ties.
  OX_HD void box_box(int p, T margin, const T* pos1, const T* size1, const T* pos2, const T* size2, int g1, int g2, int& ncon) const {
    T A[3][3], B[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) { A[i][k] = at(b.geom_xmat, 9 * g1 + 3 * k + i); B[i][k] = at(b.geom_xmat, 9 * g2 + 3 * k + i); }
    const T dvec[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
    auto radius = [&](const T (*ax)[3], const T* sz, const T* L) {
      return sz[0] * ox_abs(dot3(ax[0], L)) + sz[1] * ox_abs(dot3(ax[1], L)) + sz[2] * ox_abs(dot3(ax[2], L));
    };
    T bestFace = (T)-1e30, bestEdge = (T)-1e30;
    int faceIdx = -1, edgeI = -1, edgeJ = -1;
#pragma unroll 1
    for (int f = 0; f < 6; f++) {
      const T* L = f < 3 ? A[f] : B[f - 3];
      const T sep = ox_abs(dot3(dvec, L)) - radius(A, size1, L) - radius(B, size2, L);
      if (sep > margin) return;
      if (sep > bestFace) { bestFace = sep; faceIdx = f; }
    }
#pragma unroll 1
    for (int ij = 0; ij < 9; ij++) {
      const int i = ij / 3, j = ij % 3;
      T L[3];
      cross3(L, A[i], B[j]);
      const T n = ox_sqrt(dot3(L, L));
      if (n < (T)1e-6) continue;
      L[0] /= n; L[1] /= n; L[2] /= n;
      const T sep = ox_abs(dot3(dvec, L)) - radius(A, size1, L) - radius(B, size2, L);
      if (sep > margin) return;
      if (sep > bestEdge) { bestEdge = sep; edgeI = i; edgeJ = j; }
    }
    const T scale = ox_max(ox_max(size1[0], size1[1]), ox_max(size1[2], ox_max(size2[0], ox_max(size2[1], size2[2]))));
    if (edgeI >= 0 && bestEdge > bestFace + (T)1e-3 * scale) {
      T L[3];
      cross3(L, A[edgeI], B[edgeJ]);
      normalize3(L);
      if (dot3(L, dvec) < 0) { L[0] = -L[0]; L[1] = -L[1]; L[2] = -L[2]; }
      T c1[3] = {pos1[0], pos1[1], pos1[2]}, c2[3] = {pos2[0], pos2[1], pos2[2]};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (k != edgeI) { const T sg = dot3(A[k], L) > 0 ? (T)1 : (T)-1; for (int c = 0; c < 3; c++) c1[c] += sg * size1[k] * A[k][c]; }
        if (k != edgeJ) { const T sg = dot3(B[k], L) > 0 ? (T)-1 : (T)1; for (int c = 0; c < 3; c++) c2[c] += sg * size2[k] * B[k][c]; }
      }
      const T* u = A[edgeI]; const T* v = B[edgeJ];
      const T w[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
      const T bb = dot3(u, v), dd = dot3(u, w), e = dot3(v, w), den = 1 - bb * bb;
      T sA = den > (T)1e-12 ? (bb * e - dd) / den : (T)0, tB = den > (T)1e-12 ? (e - bb * dd) / den : (T)0;
      sA = ox_clip(sA, -size1[edgeI], size1[edgeI]); tB = ox_clip(tB, -size2[edgeJ], size2[edgeJ]);
      T pA[3], pB[3], diff[3];
      for (int c = 0; c < 3; c++) { pA[c] = c1[c] + sA * u[c]; pB[c] = c2[c] + tB * v[c]; diff[c] = pB[c] - pA[c]; }
      Con cc;
      cc.dist = dot3(diff, L);
      if (cc.dist > margin) return;
      for (int c = 0; c < 3; c++) { cc.pos[c] = (T)0.5 * (pA[c] + pB[c]); cc.frame[c] = L[c]; cc.frame[3 + c] = 0; }
      emit(cc, p, 0, ncon);
      return;
    }
    const bool refA = faceIdx < 3;
    const int ra = refA ? faceIdx : faceIdx - 3;
    const T (*Rax)[3] = refA ? A : B; const T (*Iax)[3] = refA ? B : A;
    const T* Rpos = refA ? pos1 : pos2; const T* Ipos = refA ? pos2 : pos1;
    const T* Rsz = refA ? size1 : size2; const T* Isz = refA ? size2 : size1;
    const T toI[3] = {Ipos[0] - Rpos[0], Ipos[1] - Rpos[1], Ipos[2] - Rpos[2]};
    const T sgn = dot3(toI, Rax[ra]) >= 0 ? (T)1 : (T)-1;
    const T n[3] = {sgn * Rax[ra][0], sgn * Rax[ra][1], sgn * Rax[ra][2]};
    int ia = 0;
    T best = -1;
    for (int k = 0; k < 3; k++) { const T a = ox_abs(dot3(Iax[k], n)); if (a > best) { best = a; ia = k; } }
    const T isg = dot3(Iax[ia], n) > 0 ? (T)-1 : (T)1;
    const int iu = (ia + 1) % 3, iv = (ia + 2) % 3;
    T poly[16][3], tmp[16][3];
    int np = 4;
    for (int q = 0; q < 4; q++) {
      const T su = (q == 0 || q == 3) ? (T)1 : (T)-1, sv = q < 2 ? (T)1 : (T)-1;
      for (int c = 0; c < 3; c++) poly[q][c] = Ipos[c] + isg * Isz[ia] * Iax[ia][c] + su * Isz[iu] * Iax[iu][c] + sv * Isz[iv] * Iax[iv][c];
    }
#pragma unroll 1
    for (int side = 0; side < 4 && np > 0; side++) {
      const int ta = (ra + 1 + side / 2) % 3;
      const T ps = side % 2 ? (T)-1 : (T)1;
      auto inside = [&](const T* pt) {
        const T rel[3] = {pt[0] - Rpos[0], pt[1] - Rpos[1], pt[2] - Rpos[2]};
        return Rsz[ta] - ps * dot3(rel, Rax[ta]);
      };
      int no = 0;
      for (int q = 0; q < np; q++) {
        const T* P = poly[q]; const T* Q = poly[(q + 1) % np];
        const T dp = inside(P), dq = inside(Q);
        if (dp >= 0) { for (int c = 0; c < 3; c++) tmp[no][c] = P[c]; no++; }
        if ((dp >= 0) != (dq >= 0)) { const T t = dp / (dp - dq); for (int c = 0; c < 3; c++) tmp[no][c] = P[c] + t * (Q[c] - P[c]); no++; }
      }
      np = no;
      for (int q = 0; q < np; q++) for (int c = 0; c < 3; c++) poly[q][c] = tmp[q][c];
    }
    int cnt = 0;
    for (int q = 0; q < np && cnt < 8; q++) {
      const T rel[3] = {poly[q][0] - Rpos[0], poly[q][1] - Rpos[1], poly[q][2] - Rpos[2]};
      const T depth = dot3(rel, n) - Rsz[ra];
      if (depth > margin) continue;
      Con cc;
      cc.dist = depth;
      for (int c = 0; c < 3; c++) {
        cc.pos[c] = poly[q][c] - (T)0.5 * depth * n[c];
        cc.frame[c] = refA ? n[c] : -n[c];
        cc.frame[3 + c] = 0;
      }
      emit(cc, p, cnt, ncon);
      cnt++;
    }
  }
  // narrowphase of one candidate pair; emits 0..pair_maxcon(p) contacts
  OX_HD void collide_pair(int p, int& ncon) const {
    {
      {
        const int g1 = m.pair_geom1(p), g2 = m.pair_geom2(p), t1 = m.geom_type(g1), t2 = m.geom_type(g2);
        const T margin = m.pair_margin(p);
        T size1[3], size2[3];
        OX_LDM(3, size1, geom_size, 3 * g1);
        OX_LDM(3, size2, geom_size, 3 * g2);
        T pos1[3], pos2[3];
        ld<3>(pos1, b.geom_xpos, 3 * g1);
        ld<3>(pos2, b.geom_xpos, 3 * g2);
        if (t1 != OX_GEOM_PLANE) {
          // bounding-sphere cull (what mj_collision's broadphase does): every point of a sphere / capsule lies within
          // rbound = radius (+ half length) of its centre, so centres further apart than rbound1 + rbound2 + margin cannot
          // produce a contact. Conservative (a hair of slack for round-off; the exact test follows when it does not fire),
          // so the contact set is unchanged - but the capsule-capsule narrowphase is ~1 k instructions per pair, and for a
          // humanoid's 20 self-collision pairs, almost always far apart, it was 43 % of the PRE kernel's instruction stream.
          const T rb2 = t2 == OX_GEOM_BOX ? ox_sqrt(dot3(size2, size2)) : size2[0] + (t2 == OX_GEOM_CAPSULE ? size2[1] : (T)0);
          const T rb1 = t1 == OX_GEOM_BOX ? ox_sqrt(dot3(size1, size1)) : size1[0] + (t1 == OX_GEOM_CAPSULE ? size1[1] : (T)0);
          const T rb = rb1 + rb2 + margin;
          const T dx = pos2[0] - pos1[0], dy = pos2[1] - pos1[1], dz = pos2[2] - pos1[2];
          if (dx * dx + dy * dy + dz * dz > rb * rb * (T)1.0005 + (T)1e-9) return;
        }
        if (t1 == OX_GEOM_PLANE) {
          T n[3] = {at(b.geom_xmat, 9 * g1 + 2), at(b.geom_xmat, 9 * g1 + 5), at(b.geom_xmat, 9 * g1 + 8)};
          if (t2 == OX_GEOM_SPHERE) {
            Con c;
            if (plane_sphere(c, margin, pos1, n, pos2, size2[0])) emit(c, p, 0, ncon);
          } else if (t2 == OX_GEOM_CAPSULE) {
            T axis[3] = {at(b.geom_xmat, 9 * g2 + 2), at(b.geom_xmat, 9 * g2 + 5), at(b.geom_xmat, 9 * g2 + 8)};
            const T hl = size2[1];
            {  // neither end cap within reach of the plane (the common case): one branch for the pair
              const T dc = (pos2[0] - pos1[0]) * n[0] + (pos2[1] - pos1[1]) * n[1] + (pos2[2] - pos1[2]) * n[2];
              const T da = ox_abs(dot3(axis, n)) * hl;
              if (dc - da > (margin + size2[0]) * (T)1.0001 + (T)1e-6) return;   // conservative: the exact tests follow
            }
#pragma unroll
            for (int sgn = 1; sgn >= -1; sgn -= 2) {
              T pt[3] = {pos2[0] + sgn * axis[0] * hl, pos2[1] + sgn * axis[1] * hl, pos2[2] + sgn * axis[2] * hl};
              Con c;
              if (plane_sphere(c, margin, pos1, n, pt, size2[0])) {
                c.frame[3] = axis[0]; c.frame[4] = axis[1]; c.frame[5] = axis[2];
                emit(c, p, sgn > 0 ? 0 : 1, ncon);
              }
            }
          } else if (t2 == OX_GEOM_BOX) {
            T mat2[9];
            ld<9>(mat2, b.geom_xmat, 9 * g2);
            T dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
            const T dist = dot3(dif, n);
            int cnt = 0;
            OX_MLOOP
            for (int i = 0; i < 8 && cnt < 4; i++) {
              T vec[3] = {(i & 1 ? size2[0] : -size2[0]), (i & 2 ? size2[1] : -size2[1]), (i & 4 ? size2[2] : -size2[2])}, corner[3];
              mat_vec3(corner, mat2, vec);
              const T ldist = dot3(n, corner);
              if (dist + ldist > margin || ldist > 0) continue;
              Con c;
              c.dist = dist + ldist;
              c.frame[0] = n[0]; c.frame[1] = n[1]; c.frame[2] = n[2]; c.frame[3] = 0; c.frame[4] = 0; c.frame[5] = 0;
              const T s = -c.dist / 2;
              c.pos[0] = corner[0] + pos2[0] + n[0] * s; c.pos[1] = corner[1] + pos2[1] + n[1] * s; c.pos[2] = corner[2] + pos2[2] + n[2] * s;
              if (cnt == 0) emit(c, p, 0, ncon); else if (cnt == 1) emit(c, p, 1, ncon); else if (cnt == 2) emit(c, p, 2, ncon); else emit(c, p, 3, ncon);
              cnt++;
            }
          }
        } else if (t1 == OX_GEOM_SPHERE && t2 == OX_GEOM_SPHERE) {
          Con c;
          if (sphere_sphere(c, margin, pos1, size1[0], pos2, size2[0])) emit(c, p, 0, ncon);
        } else if (t1 == OX_GEOM_SPHERE && t2 == OX_GEOM_CAPSULE) {
          T axis[3] = {at(b.geom_xmat, 9 * g2 + 2), at(b.geom_xmat, 9 * g2 + 5), at(b.geom_xmat, 9 * g2 + 8)};
          T vec[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
          const T x = ox_clip(dot3(axis, vec), -size2[1], size2[1]);
          vec[0] = pos2[0] + axis[0] * x; vec[1] = pos2[1] + axis[1] * x; vec[2] = pos2[2] + axis[2] * x;
          Con c;
          if (sphere_sphere(c, margin, pos1, size1[0], vec, size2[0])) emit(c, p, 0, ncon);
        } else if (t1 == OX_GEOM_BOX && t2 == OX_GEOM_BOX) {
          box_box(p, margin, pos1, size1, pos2, size2, g1, g2, ncon);
        } else if (t1 == OX_GEOM_SPHERE && t2 == OX_GEOM_BOX) {
          T mat2[9];
          ld<9>(mat2, b.geom_xmat, 9 * g2);
          Con c;
          if (sphere_box(c, margin, pos1, size1[0], pos2, mat2, size2)) emit(c, p, 0, ncon);
        } else if (t1 == OX_GEOM_CAPSULE && t2 == OX_GEOM_BOX) {
          T mat2[9];
          ld<9>(mat2, b.geom_xmat, 9 * g2);
          const T axis[3] = {at(b.geom_xmat, 9 * g1 + 2), at(b.geom_xmat, 9 * g1 + 5), at(b.geom_xmat, 9 * g1 + 8)};
          const T hl = size1[1];
          const T tmp[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
          T lc[3], la[3];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            lc[k] = mat2[k] * tmp[0] + mat2[3 + k] * tmp[1] + mat2[6 + k] * tmp[2];
            la[k] = (mat2[k] * axis[0] + mat2[3 + k] * axis[1] + mat2[6 + k] * axis[2]) * hl;
          }
          const T tstar = segment_box_closest(lc, la, size2);
          int n = 0;
#pragma unroll
          for (int which = 0; which < 2; which++) {   // the closest point of the axis, then the far end cap
            const T t = which == 0 ? tstar : (tstar <= 0 ? (T)1 : (T)-1);
            const T pt[3] = {pos1[0] + t * hl * axis[0], pos1[1] + t * hl * axis[1], pos1[2] + t * hl * axis[2]};
            Con c;
            if (sphere_box(c, margin, pt, size1[0], pos2, mat2, size2)) {
              c.frame[3] = axis[0]; c.frame[4] = axis[1]; c.frame[5] = axis[2];
              if (n == 0) emit(c, p, 0, ncon); else emit(c, p, 1, ncon);
              n++;
            }
          }
        } else if (t1 == OX_GEOM_CAPSULE && t2 == OX_GEOM_CAPSULE) {
          T axis1[3] = {at(b.geom_xmat, 9 * g1 + 2) * size1[1], at(b.geom_xmat, 9 * g1 + 5) * size1[1], at(b.geom_xmat, 9 * g1 + 8) * size1[1]};
          T axis2[3] = {at(b.geom_xmat, 9 * g2 + 2) * size2[1], at(b.geom_xmat, 9 * g2 + 5) * size2[1], at(b.geom_xmat, 9 * g2 + 8) * size2[1]};
          T dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
          const T ma = dot3(axis1, axis1), mb = -dot3(axis1, axis2), mc = dot3(axis2, axis2);
          const T u = -dot3(axis1, dif), v = dot3(axis2, dif), det = ma * mc - mb * mb;
          T vec1[3], vec2[3];
          Con c;
          if (ox_abs(det) >= (T)OX_MINVAL) {
            T x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
            if (x1 > 1) { x1 = 1; x2 = (v - mb) / mc; } else if (x1 < -1) { x1 = -1; x2 = (v + mb) / mc; }
            if (x2 > 1) { x2 = 1; x1 = (u - mb) / ma; } else if (x2 < -1) { x2 = -1; x1 = (u + mb) / ma; }
            x1 = ox_clip(x1, (T)-1, (T)1);
            OX_MLOOP
            for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k] * x1; vec2[k] = pos2[k] + axis2[k] * x2; }
            if (sphere_sphere(c, margin, vec1, size1[0], vec2, size2[0])) emit(c, p, 0, ncon);
          } else {
            int n = 0;
            T x2 = ox_clip((v - mb) / mc, (T)-1, (T)1);
            OX_MLOOP
            for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k]; vec2[k] = pos2[k] + axis2[k] * x2; }
            if (sphere_sphere(c, margin, vec1, size1[0], vec2, size2[0])) { if (n == 0) emit(c, p, 0, ncon); else emit(c, p, 1, ncon); n++; }
            x2 = ox_clip((v + mb) / mc, (T)-1, (T)1);
            OX_MLOOP
            for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] - axis1[k]; vec2[k] = pos2[k] + axis2[k] * x2; }
            if (sphere_sphere(c, margin, vec1, size1[0], vec2, size2[0])) { if (n == 0) emit(c, p, 0, ncon); else emit(c, p, 1, ncon); n++; }
            if (n < 2) {
              T x1 = ox_clip((u - mb) / ma, (T)-1, (T)1);
              OX_MLOOP
              for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k] * x1; vec2[k] = pos2[k] + axis2[k]; }
              if (sphere_sphere(c, margin, vec1, size1[0], vec2, size2[0])) { if (n == 0) emit(c, p, 0, ncon); else emit(c, p, 1, ncon); n++; }
            }
            if (n < 2) {
              T x1 = ox_clip((u + mb) / ma, (T)-1, (T)1);
              OX_MLOOP
              for (int k = 0; k < 3; k++) { vec1[k] = pos1[k] + axis1[k] * x1; vec2[k] = pos2[k] - axis2[k]; }
              if (sphere_sphere(c, margin, vec1, size1[0], vec2, size2[0])) { if (n == 0) emit(c, p, 0, ncon); else emit(c, p, 1, ncon); n++; }
            }
          }
        }
      }
    }
  }

  // ============================================================ A.6 constraint assembly
  // impedance d(x), reference acceleration aref, regulariser R for one row; returns R
  OX_HD T row_params(const T* solref, const T* solimp, T pos, T margin, T diagApprox, T vel, T* aref) const {
    const auto& h = m.h();
    const T dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
    T imp;
    if (dmin == dmax || width <= (T)OX_MINVAL) imp = (T)0.5 * (dmin + dmax);
    else {
      const T x = ox_abs(pos - margin) / width;
      if (x >= 1) imp = dmax;
      else if (x <= 0) imp = dmin;
      else {
        T y;
        if (power == 1) y = x;
        else if (power == 2) y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);  // MuJoCo's default; no pow()
        else if (x <= mid) y = ox_pow(x, power) / ox_pow(mid, power - 1);
        else y = 1 - ox_pow(1 - x, power) / ox_pow(1 - mid, power - 1);
        imp = dmin + y * (dmax - dmin);
      }
    }
    imp = ox_clip(imp, (T)OX_MINIMP, (T)OX_MAXIMP);
    T K, Bd;
    if (solref[0] > 0) {
      T tc = solref[0];
      const T dr = solref[1];
      if (!dis(OX_DSBL_REFSAFE)) tc = ox_max(tc, 2 * (T)h.timestep);
      K = 1 / ox_max((T)OX_MINVAL, dmax * dmax * tc * tc * dr * dr);
      Bd = 2 / ox_max((T)OX_MINVAL, dmax * tc);
    } else {
      K = -solref[0] / ox_max((T)OX_MINVAL, dmax * dmax);
      Bd = -solref[1] / ox_max((T)OX_MINVAL, dmax);
    }
    *aref = -Bd * vel - K * imp * (pos - margin);
    return ox_max((T)OX_MINVAL, (1 - imp) * diagApprox / imp);
  }

  // rows of one contact (pyramidal cone or frictionless): Jacobian, impedance, reference acceleration. condim 3: normal + two
  // tangents; condim 4 adds torsion about the normal, condim 6 rolling about the two tangents - those friction directions act
  // on the RELATIVE ANGULAR velocity of the two bodies (rotational Jacobian), every direction k contributes the edge pair
  // J_n +/- mu_k J_k.
  OX_HD void contact_rows(int c, int p, int& nefc) const {
    const auto& h = m.h();
    const int nv = h.nv;
      const T includemargin = m.pair_margin(p) - m.pair_gap(p);
      const T dist = at(b.con_dist, c);
      ati(b.con_efcadr, c) = -1;
      if (dist >= includemargin) return;
      const int dim = m.pair_dim(p);
      const int nrow = dim == 1 ? 1 : 2 * (dim - 1);
      const int r0 = nefc;
      ati(b.con_efcadr, c) = r0;   // contact.efc_address: where this contact's rows start (touch sensor, contact forces)
      nefc += nrow;
      T fri[5];
      OX_LDM(5, fri, pair_friction, 5 * p);
      T cpos[3], frame[9];
      ld<3>(cpos, b.con_pos, 3 * c);
      ld<9>(frame, b.con_frame, 9 * c);
#pragma unroll
      for (int rr = 0; rr < 10; rr++) {
        if (rr >= nrow) break;
        const int r = r0 + rr;
        OX_MLOOP
        for (int i = 0; i < nv; i++) at(b.efc_J, r * nv + i) = 0;
      }
      T veln = 0, velt[5] = {0, 0, 0, 0, 0};
      const int bodies[2] = {m.geom_bodyid(m.pair_geom2(p)), m.geom_bodyid(m.pair_geom1(p))};
      OX_MLOOP
      for (int sidx = 0; sidx < 2; sidx++) {
        const T sign = sidx == 0 ? (T)1 : (T)-1;
        int body = bodies[sidx];
        T sc[3], offset[3];
        ld<3>(sc, b.subtree_com, 3 * m.body_rootid(body));
        offset[0] = cpos[0] - sc[0]; offset[1] = cpos[1] - sc[1]; offset[2] = cpos[2] - sc[2];
        body = m.body_weldid(body);  // nearest ancestor-or-self that has dofs (0 = static)
        if (!body) continue;  // static body: no dofs on this side
        const int last_ = m.body_dofadr(body) + m.body_dofnum(body) - 1;
        OX_MLOOP
        for (int d_ = 0, i = last_; d_ < m.dof_depth(last_); d_++, i = m.dof_parentid(i)) {
          T cd[6], jp[3];
          ld<6>(cd, b.cdof, 6 * i);
          cross3(jp, cd, offset);
          jp[0] = sign * (jp[0] + cd[3]); jp[1] = sign * (jp[1] + cd[4]); jp[2] = sign * (jp[2] + cd[5]);
          const T jn = dot3(frame, jp);
          const T qv = at(b.qvel, i);
          veln += jn * qv;
          if (dim == 1) {
            at(b.efc_J, r0 * nv + i) += jn;
          } else {
#pragma unroll
            for (int k = 1; k < 6; k++) {  // static bound so that fri / velt stay in registers; k >= 3: torsion / rolling (angular part of cdof)
              if (k >= dim) break;
              const T jr[3] = {sign * cd[0], sign * cd[1], sign * cd[2]};
              const T jt = (k < 3 ? dot3(frame + 3 * k, jp) : dot3(frame + 3 * (k - 3), jr)) * fri[k - 1];
              velt[k - 1] += jt * qv;
              at(b.efc_J, (r0 + 2 * (k - 1)) * nv + i) += jn + jt;
              at(b.efc_J, (r0 + 2 * (k - 1) + 1) * nv + i) += jn - jt;
            }
          }
        }
      }
      const T tran = m.body_invweight0(2 * bodies[1]) + m.body_invweight0(2 * bodies[0]);
      T solref[2], solimp[5];
      OX_LDM(2, solref, pair_solref, 2 * p);
      OX_LDM(5, solimp, pair_solimp, 5 * p);
      if (dim == 1) {
        T aref;
        const T R = row_params(solref, solimp, dist, includemargin, tran, veln, &aref);
        at(b.efc_pos, r0) = dist; at(b.efc_margin, r0) = includemargin; at(b.efc_D, r0) = 1 / R; at(b.efc_aref, r0) = aref;
      } else {
        T arefs[10], Rfirst = 0;
        const T rot = dim > 3 ? m.body_invweight0(2 * bodies[1] + 1) + m.body_invweight0(2 * bodies[0] + 1) : (T)0;
#pragma unroll
        for (int k = 1; k < 6; k++)
#pragma unroll
          for (int s = 0; s < 2; s++) {
            if (k >= dim) continue;
            const T vel = veln + (s ? -velt[k - 1] : velt[k - 1]);
            const T R = row_params(solref, solimp, dist, includemargin, tran + fri[k - 1] * fri[k - 1] * (k < 3 ? tran : rot), vel, &arefs[2 * (k - 1) + s]);
            if (k == 1 && s == 0) Rfirst = R;
          }
        const T mu = fri[0] * ox_sqrt(1 / (T)h.impratio);
        const T D = 1 / (2 * mu * mu * Rfirst);
#pragma unroll
        for (int r = 0; r < 10; r++) {
          if (r >= nrow) break;
          at(b.efc_pos, r0 + r) = dist; at(b.efc_margin, r0 + r) = includemargin; at(b.efc_D, r0 + r) = D;
          at(b.efc_aref, r0 + r) = arefs[r];
        }
      }
  }

  // rows of one contact under the ELLIPTIC cone: normal row, then one row per friction direction (tangents; torsion / rolling
  // for condim 4 / 6) - the contact-frame components themselves, not pyramid edges. Only the normal row has a position term;
  // friction row k gets D_k = D_n friction_k^2 / mu^2 with the regularised friction mu = friction_1 / sqrt(impratio), which makes
  // the cone circular in the scaled coordinates of elliptic_cost().
  OX_HD void contact_rows_elliptic(int c, int p, int& nefc) const {
    const auto& h = m.h();
    const int nv = h.nv;
    const T includemargin = m.pair_margin(p) - m.pair_gap(p);
    const T dist = at(b.con_dist, c);
    ati(b.con_efcadr, c) = -1;
    if (dist >= includemargin) return;
    const int dim = m.pair_dim(p);
    const int r0 = nefc;
    ati(b.con_efcadr, c) = r0;
    nefc += dim;
    T fri[5], cpos[3], frame[9], vel[6] = {0, 0, 0, 0, 0, 0};
    OX_LDM(5, fri, pair_friction, 5 * p);
    ld<3>(cpos, b.con_pos, 3 * c);
    ld<9>(frame, b.con_frame, 9 * c);
#pragma unroll
    for (int rr = 0; rr < 6; rr++) {
      if (rr >= dim) break;
      OX_MLOOP
      for (int i = 0; i < nv; i++) at(b.efc_J, (r0 + rr) * nv + i) = 0;
    }
    const int bodies[2] = {m.geom_bodyid(m.pair_geom2(p)), m.geom_bodyid(m.pair_geom1(p))};
    OX_MLOOP
    for (int sidx = 0; sidx < 2; sidx++) {
      const T sign = sidx == 0 ? (T)1 : (T)-1;
      int body = bodies[sidx];
      T sc[3], offset[3];
      ld<3>(sc, b.subtree_com, 3 * m.body_rootid(body));
      offset[0] = cpos[0] - sc[0]; offset[1] = cpos[1] - sc[1]; offset[2] = cpos[2] - sc[2];
      body = m.body_weldid(body);
      if (!body) continue;
      const int last_ = m.body_dofadr(body) + m.body_dofnum(body) - 1;
      OX_MLOOP
      for (int d_ = 0, i = last_; d_ < m.dof_depth(last_); d_++, i = m.dof_parentid(i)) {
        T cd[6], jp[3];
        ld<6>(cd, b.cdof, 6 * i);
        cross3(jp, cd, offset);
        jp[0] = sign * (jp[0] + cd[3]); jp[1] = sign * (jp[1] + cd[4]); jp[2] = sign * (jp[2] + cd[5]);
        const T jr[3] = {sign * cd[0], sign * cd[1], sign * cd[2]};
        const T qv = at(b.qvel, i);
#pragma unroll
        for (int k = 0; k < 6; k++) {
          if (k >= dim) break;
          const T j = k < 3 ? dot3(frame + 3 * k, jp) : dot3(frame + 3 * (k - 3), jr);
          at(b.efc_J, (r0 + k) * nv + i) += j;
          vel[k] += j * qv;
        }
      }
    }
    const T tran = m.body_invweight0(2 * bodies[1]) + m.body_invweight0(2 * bodies[0]);
    T solref[2], solimp[5], aref;
    OX_LDM(2, solref, pair_solref, 2 * p);
    OX_LDM(5, solimp, pair_solimp, 5 * p);
    const T Rn = row_params(solref, solimp, dist, includemargin, tran, vel[0], &aref);
    at(b.efc_pos, r0) = dist; at(b.efc_margin, r0) = includemargin; at(b.efc_D, r0) = 1 / Rn; at(b.efc_aref, r0) = aref;
    const T mu = fri[0] * ox_sqrt(1 / (T)h.impratio);
#pragma unroll
    for (int k = 1; k < 6; k++) {
      if (k >= dim) break;
      (void)row_params(solref, solimp, (T)0, (T)0, tran, vel[k], &aref);   // no position term: aref = -B vel
      at(b.efc_pos, r0 + k) = 0; at(b.efc_margin, r0 + k) = 0; at(b.efc_aref, r0 + k) = aref;
      at(b.efc_D, r0 + k) = fri[k - 1] * fri[k - 1] / (Rn * mu * mu);
    }
  }
  OX_HD void contact_rows_any(int c, int p, int& nefc) const {
    if (m.h().cone == OX_CONE_ELLIPTIC && m.pair_dim(p) > 1) contact_rows_elliptic(c, p, nefc);
    else contact_rows(c, p, nefc);
  }
  // visit every contact that has constraint rows: f(contact slot, pair, first row, dim)
  template <typename F>
  OX_HD void for_each_contact(F&& f) const {
    const auto& h = m.h();
    if (STATIC_CON || slots) {
      OX_MLOOP
      for (int p = 0; p < h.npair; p++) {
        OX_MLOOP
        for (int k = 0; k < m.pair_maxcon(p); k++) {
          const int c = m.pair_conadr(p) + k;
          if (ati(b.con_active, c) && ati(b.con_efcadr, c) >= 0) f(c, p, ati(b.con_efcadr, c), m.pair_dim(p));
        }
      }
    } else {
      const int ncon = ati(b.ncon, 0);
      for (int c = 0; c < ncon; c++) {
        const int p = ati(b.con_pair, c);
        if (ati(b.con_efcadr, c) >= 0) f(c, p, ati(b.con_efcadr, c), m.pair_dim(p));
      }
    }
  }
  // Cost of one elliptic contact as a function of x = J a - aref of its rows (mj_constraintUpdate, elliptic). In the scaled
  // coordinates N = mu x_0, T = |(friction_j x_j)|: top zone (N >= mu T) cost 0; bottom zone (mu N + T <= 0) every row an ordinary
  // quadratic row; middle zone 1/2 Dm (N - mu T)^2, Dm = D_n / (mu^2 (1 + mu^2)). force = -d cost/dx (may be null), Hc = d2 cost/dx2
  // (dim x dim, may be null). Returns the zone (0 top, 1 bottom, 2 middle).
  OX_HD int elliptic_cost(int p, int dim, const T* D, const T* x, T* cost, T* force, T* Hc) const {
    T U[6] = {0, 0, 0, 0, 0, 0}, scl[6] = {0, 0, 0, 0, 0, 0};
    const T mu = m.pair_friction(5 * p) * ox_sqrt(1 / (T)m.h().impratio);
    scl[0] = mu;
#pragma unroll
    for (int j = 1; j < 6; j++) { if (j >= dim) break; scl[j] = m.pair_friction(5 * p + j - 1); }
    T T2 = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) { if (j >= dim) break; U[j] = x[j] * scl[j]; if (j) T2 += U[j] * U[j]; }
    const T N = U[0], Tn = ox_sqrt(T2);
    *cost = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) { if (j >= dim) break; if (force) force[j] = 0; }
    if (Hc) {
#pragma unroll
      for (int j = 0; j < 36; j++) { if (j >= dim * dim) break; Hc[j] = 0; }
    }
    if (N >= mu * Tn || (Tn <= 0 && N >= 0)) return 0;
    if (mu * N + Tn <= 0 || (Tn <= 0 && N < 0)) {
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (j >= dim) break;
        *cost += (T)0.5 * D[j] * x[j] * x[j];
        if (force) force[j] = -D[j] * x[j];
        if (Hc) Hc[j * dim + j] = D[j];
      }
      return 1;
    }
    const T Dm = D[0] / (mu * mu * (1 + mu * mu)), NT = N - mu * Tn;
    *cost = (T)0.5 * Dm * NT * NT;
    if (force) {
      force[0] = -Dm * NT * scl[0];
#pragma unroll
      for (int j = 1; j < 6; j++) { if (j >= dim) break; force[j] = Dm * NT * mu * U[j] / Tn * scl[j]; }
    }
    if (Hc) {
      Hc[0] = Dm * scl[0] * scl[0];
#pragma unroll
      for (int j = 1; j < 6; j++) {
        if (j >= dim) break;
        const T h0 = -Dm * mu * U[j] / Tn * scl[0] * scl[j];
        Hc[j] = h0; Hc[j * dim] = h0;
#pragma unroll
        for (int k = 1; k < 6; k++) {
          if (k >= dim) break;
          Hc[j * dim + k] = (Dm * mu * mu * U[j] * U[k] / T2 - Dm * NT * mu * ((j == k ? (T)1 : (T)0) / Tn - U[j] * U[k] / (T2 * Tn))) * scl[j] * scl[k];
        }
      }
    }
    return 2;
  }

  OX_HDN void make_constraint() const {
    const auto& h = m.h();
    const int nv = h.nv, njnt = h.njnt;
    int nefc = 0;
    ati(b.ne, 0) = 0;
    ati(b.nf, 0) = 0;
    if (!dis(OX_DSBL_CONSTRAINT)) {
      if (!dis(OX_DSBL_EQUALITY)) {   // equality rows come first (mj_makeConstraint order) and are always active in the solver
        OX_MLOOP
        for (int i = 0; i < h.neq; i++) equality_rows(i, nefc);
      }
      ati(b.ne, 0) = nefc;
      if (h.nfloss > 0 && !dis(OX_DSBL_FRICTIONLOSS)) {   // dry joint friction: one row per dof with frictionloss > 0
        OX_MLOOP
        for (int i = 0; i < nv; i++) friction_row(i, nefc);
      }
      ati(b.nf, 0) = nefc - ati(b.ne, 0);
      bool any_limit = false;  // one test for "no joint is near a limit" (the common case), see fwd_acceleration
      if (!dis(OX_DSBL_LIMIT)) {
        OX_MLOOP
        for (int j = 0; j < njnt; j++) {
          if (!m.jnt_limited(j)) continue;
          const int jt = m.jnt_type(j);
          if (jt == OX_JNT_BALL) { any_limit |= limit_count(j) > 0; continue; }
          if (jt != OX_JNT_SLIDE && jt != OX_JNT_HINGE) continue;
          const T value = at(b.qpos, m.jnt_qposadr(j)), margin = m.jnt_margin(j);
          any_limit |= (value - m.jnt_range(2 * j) < margin) | (m.jnt_range(2 * j + 1) - value < margin);
        }
      }
      if (any_limit) {
        OX_MLOOP
        for (int j = 0; j < njnt; j++) limit_rows(j, nefc);
      }
      if (!dis(OX_DSBL_LIMIT)) {
        OX_MLOOP
        for (int i = 0; i < h.ntendon; i++) tendon_limit_rows(i, nefc);
      }
      if (STATIC_CON) {  // static contact slots: every index below is a compile-time constant after unrolling
        OX_MLOOP
        for (int p = 0; p < h.npair; p++) {
          int pair_active = 0;
          OX_MLOOP
          for (int k = 0; k < m.pair_maxcon(p); k++) pair_active |= ati(b.con_active, m.pair_conadr(p) + k);
          if (!pair_active) continue;
          OX_MLOOP
          for (int k = 0; k < m.pair_maxcon(p); k++) {
            const int c = m.pair_conadr(p) + k;
            if (ati(b.con_active, c)) contact_rows_any(c, p, nefc);
          }
        }
      } else {
        const int ncon = ati(b.ncon, 0);
        for (int c = 0; c < ncon; c++) contact_rows_any(c, ati(b.con_pair, c), nefc);
      }
    }
    ati(b.nefc, 0) = nefc;
    if (h.cone == OX_CONE_ELLIPTIC && nefc > 0) {   // mark the rows of elliptic contacts (efc_floss < 0): the solver's row loops skip them
      OX_ROWLOOP
      for (int r = ati(b.ne, 0) + ati(b.nf, 0); r < nefc; r++) at(b.efc_floss, r) = 0;
      for_each_contact([&](int, int, int ea, int dim) {
        if (dim > 1)
          for (int j = 0; j < dim; j++) at(b.efc_floss, ea + j) = -1;
      });
    }
  }
  // rows of equality constraint i (mj_instantiateEquality): connect = the two anchors coincide (3 rows), joint = q1 tracks a
  // quartic polynomial of q2 (1 row). pos = residual, margin = 0.
  OX_HD int equality_count(int i) const {
    if (!(at(b.eq_active, i) != 0)) return 0;
    return m.eq_type(i) == OX_EQ_CONNECT ? 3 : (m.eq_type(i) == OX_EQ_WELD ? 6 : 1);
  }
  // weld orientation rows: residual = ts * imag(conj(q2) q1 qrel); G (3x3, row-major) maps the relative angular velocity w1 - w2 to
  // its time derivative: column c = ts/2 * imag(conj(q2) [0, e_c] q1 qrel)
  OX_HD void weld_rot_map(int i, int b1, int b2, T* G, T* residual) const {
    T q1[4], q2[4], qrel[4], quat[4], e[4], t1[4], t2[4];
    ld<4>(q1, b.xquat, 4 * b1);
    ld<4>(q2, b.xquat, 4 * b2);
    OX_LDM(4, qrel, eq_data, 11 * i + 6);
    const T ts = m.eq_data(11 * i + 10);
    const T q2c[4] = {q2[0], -q2[1], -q2[2], -q2[3]};
    mul_quat(quat, q1, qrel);
    mul_quat(e, q2c, quat);
    residual[0] = ts * e[1]; residual[1] = ts * e[2]; residual[2] = ts * e[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const T ax[4] = {0, c == 0 ? (T)1 : (T)0, c == 1 ? (T)1 : (T)0, c == 2 ? (T)1 : (T)0};
      mul_quat(t1, q2c, ax);
      mul_quat(t2, t1, quat);
      G[c] = (T)0.5 * ts * t2[1]; G[3 + c] = (T)0.5 * ts * t2[2]; G[6 + c] = (T)0.5 * ts * t2[3];
    }
  }
  OX_HD void eq_row_finish(int i, int r, T pos, T diag) const {
    const int nv = m.h().nv;
    T vel = 0;
    OX_MLOOP
    for (int k = 0; k < nv; k++) vel += at(b.efc_J, r * nv + k) * at(b.qvel, k);
    T sr[2], si[5], aref;
    OX_LDM(2, sr, eq_solref, 2 * i);
    OX_LDM(5, si, eq_solimp, 5 * i);
    const T R = row_params(sr, si, pos, (T)0, diag, vel, &aref);
    at(b.efc_pos, r) = pos; at(b.efc_margin, r) = 0; at(b.efc_D, r) = 1 / R; at(b.efc_aref, r) = aref;
  }
  OX_HD void equality_rows(int i, int& nefc) const {
    const int nv = m.h().nv;
    if (!(at(b.eq_active, i) != 0)) return;
    if (m.eq_type(i) == OX_EQ_CONNECT || m.eq_type(i) == OX_EQ_WELD) {
      const int b1 = m.eq_obj1id(i), b2 = m.eq_obj2id(i);
      T p[2][3];
      const int bodies[2] = {b1, b2};
      OX_MLOOP
      for (int s = 0; s < 2; s++) {
        T mat[9], xp[3], a[3], w[3];
        ld<9>(mat, b.xmat, 9 * bodies[s]);
        ld<3>(xp, b.xpos, 3 * bodies[s]);
        OX_LDM(3, a, eq_data, 11 * i + 3 * s);
        mat_vec3(w, mat, a);
        p[s][0] = xp[0] + w[0]; p[s][1] = xp[1] + w[1]; p[s][2] = xp[2] + w[2];
      }
      const int r0 = nefc;
      nefc += 3;
      OX_MLOOP
      for (int k = 0; k < 3 * nv; k++) at(b.efc_J, r0 * nv + k) = 0;
      OX_MLOOP
      for (int s = 0; s < 2; s++) {   // J = Jp(body1, p1) - Jp(body2, p2), rows = world x, y, z
        const T sign = s == 0 ? (T)1 : (T)-1;
        int body = m.body_weldid(bodies[s]);
        if (!body) continue;
        T sc[3], offset[3];
        ld<3>(sc, b.subtree_com, 3 * m.body_rootid(bodies[s]));
        offset[0] = p[s][0] - sc[0]; offset[1] = p[s][1] - sc[1]; offset[2] = p[s][2] - sc[2];
        const int last_ = m.body_dofadr(body) + m.body_dofnum(body) - 1;
        OX_MLOOP
        for (int d_ = 0, dof = last_; d_ < m.dof_depth(last_); d_++, dof = m.dof_parentid(dof)) {
          T cd[6], jp[3];
          ld<6>(cd, b.cdof, 6 * dof);
          cross3(jp, cd, offset);
          OX_MLOOP
          for (int k = 0; k < 3; k++) at(b.efc_J, (r0 + k) * nv + dof) += sign * (jp[k] + cd[3 + k]);
        }
      }
      const T diag = m.body_invweight0(2 * b1) + m.body_invweight0(2 * b2);
      OX_MLOOP
      for (int k = 0; k < 3; k++) eq_row_finish(i, r0 + k, p[0][k] - p[1][k], diag);
      if (m.eq_type(i) == OX_EQ_WELD) {   // three orientation rows: J = G (jacr(body1) - jacr(body2))
        T G[9], res[3];
        weld_rot_map(i, b1, b2, G, res);
        const int q0 = nefc;
        nefc += 3;
        OX_MLOOP
        for (int k = 0; k < 3 * nv; k++) at(b.efc_J, q0 * nv + k) = 0;
        OX_MLOOP
        for (int s = 0; s < 2; s++) {
          const T sign = s == 0 ? (T)1 : (T)-1;
          const int body = m.body_weldid(bodies[s]);
          if (!body) continue;
          const int last_ = m.body_dofadr(body) + m.body_dofnum(body) - 1;
          OX_MLOOP
          for (int d_ = 0, dof = last_; d_ < m.dof_depth(last_); d_++, dof = m.dof_parentid(dof)) {
            T cd[6];
            ld<6>(cd, b.cdof, 6 * dof);
            OX_MLOOP
            for (int k = 0; k < 3; k++) at(b.efc_J, (q0 + k) * nv + dof) += sign * (G[3 * k] * cd[0] + G[3 * k + 1] * cd[1] + G[3 * k + 2] * cd[2]);
          }
        }
        const T rdiag = m.body_invweight0(2 * b1 + 1) + m.body_invweight0(2 * b2 + 1);
        OX_MLOOP
        for (int k = 0; k < 3; k++) eq_row_finish(i, q0 + k, res[k], rdiag);
      }
    } else {
      const int j1 = m.eq_obj1id(i), j2 = m.eq_obj2id(i);
      const int d1 = m.jnt_dofadr(j1), q1 = m.jnt_qposadr(j1);
      const int r = nefc++;
      OX_MLOOP
      for (int k = 0; k < nv; k++) at(b.efc_J, r * nv + k) = 0;
      T c[5];
      OX_LDM(5, c, eq_data, 11 * i);
      T pos = at(b.qpos, q1) - m.qpos0(q1), diag = m.dof_invweight0(d1);
      at(b.efc_J, r * nv + d1) = 1;
      if (j2 >= 0) {
        const int d2 = m.jnt_dofadr(j2), q2 = m.jnt_qposadr(j2);
        const T dq = at(b.qpos, q2) - m.qpos0(q2);
        pos -= c[0] + dq * (c[1] + dq * (c[2] + dq * (c[3] + dq * c[4])));
        at(b.efc_J, r * nv + d2) = -(c[1] + dq * (2 * c[2] + dq * (3 * c[3] + dq * 4 * c[4])));
        diag += m.dof_invweight0(d2);
      } else {
        pos -= c[0];
      }
      eq_row_finish(i, r, pos, diag);
    }
  }
  // dry friction of dof i (mj_instantiateFriction): J = e_i, pos = margin = 0, |force| <= frictionloss (efc_floss marks the row)
  OX_HD void friction_row(int i, int& nefc) const {
    const T fl = m.dof_frictionloss(i);
    if (!(fl > 0)) return;
    const int nv = m.h().nv, r = nefc++;
    OX_MLOOP
    for (int k = 0; k < nv; k++) at(b.efc_J, r * nv + k) = 0;
    at(b.efc_J, r * nv + i) = 1;
    T aref, sr[2], si[5];
    OX_LDM(2, sr, dof_solref_fri, 2 * i);
    OX_LDM(5, si, dof_solimp_fri, 5 * i);
    const T R = row_params(sr, si, (T)0, (T)0, m.dof_invweight0(i), at(b.qvel, i), &aref);
    at(b.efc_pos, r) = 0; at(b.efc_margin, r) = 0; at(b.efc_D, r) = 1 / R; at(b.efc_aref, r) = aref; at(b.efc_floss, r) = fl;
  }
  // number of limit rows joint j contributes (0, 1 or 2) and the rows themselves
  // ball joint: rotation angle and unit axis of the joint quaternion (mju_quat2Vel with dt = 1, then normalised); the limit is on
  // the angle, range = (0, max angle)
  OX_HD T ball_angle_axis(int j, T* axis) const {
    T q[4];
    ld<4>(q, b.qpos, m.jnt_qposadr(j));
    normalize4(q);
    const T sn = ox_sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    T ang = 2 * ox_atan2(sn, q[0]);
    if (ang > (T)3.14159265358979323846) ang -= (T)(2 * 3.14159265358979323846);
    const T inv = sn > (T)OX_MINVAL ? 1 / sn : (T)0;
    const T sg = ang < 0 ? (T)-1 : (T)1;
    axis[0] = sg * q[1] * inv; axis[1] = sg * q[2] * inv; axis[2] = sg * q[3] * inv;
    return ox_abs(ang);
  }
  OX_HD int limit_count(int j) const {
    if (!m.jnt_limited(j)) return 0;
    const int jt = m.jnt_type(j);
    if (jt == OX_JNT_BALL) {
      T axis[3];
      return ox_max(m.jnt_range(2 * j), m.jnt_range(2 * j + 1)) - ball_angle_axis(j, axis) < m.jnt_margin(j) ? 1 : 0;
    }
    if (jt != OX_JNT_SLIDE && jt != OX_JNT_HINGE) return 0;
    const T value = at(b.qpos, m.jnt_qposadr(j)), margin = m.jnt_margin(j);
    return (value - m.jnt_range(2 * j) < margin ? 1 : 0) + (m.jnt_range(2 * j + 1) - value < margin ? 1 : 0);
  }
  OX_HD void limit_rows(int j, int& nefc) const {
    const int nv = m.h().nv;
    {
      {
          if (!m.jnt_limited(j)) return;
          const int jt = m.jnt_type(j);
          if (jt == OX_JNT_BALL) {   // one row: J = -axis on the joint's three dofs, pos = max angle - angle
            T axis[3];
            const T margin = m.jnt_margin(j), dist = ox_max(m.jnt_range(2 * j), m.jnt_range(2 * j + 1)) - ball_angle_axis(j, axis);
            if (!(dist < margin)) return;
            const int da = m.jnt_dofadr(j), r = nefc++;
            OX_MLOOP
            for (int i = 0; i < nv; i++) at(b.efc_J, r * nv + i) = 0;
            T vel = 0;
            OX_MLOOP
            for (int k = 0; k < 3; k++) { at(b.efc_J, r * nv + da + k) = -axis[k]; vel -= axis[k] * at(b.qvel, da + k); }
            T aref, sr[2], si[5];
            OX_LDM(2, sr, jnt_solref, 2 * j);
            OX_LDM(5, si, jnt_solimp, 5 * j);
            const T R = row_params(sr, si, dist, margin, m.dof_invweight0(da), vel, &aref);
            at(b.efc_pos, r) = dist; at(b.efc_margin, r) = margin; at(b.efc_D, r) = 1 / R; at(b.efc_aref, r) = aref;
            return;
          }
          if (jt != OX_JNT_SLIDE && jt != OX_JNT_HINGE) return;
          const int da = m.jnt_dofadr(j);
          const T value = at(b.qpos, m.jnt_qposadr(j)), margin = m.jnt_margin(j);
          OX_MLOOP
          for (int side = -1; side <= 1; side += 2) {
            const T dist = side * (m.jnt_range(2 * j + (side + 1) / 2) - value);
            if (dist < margin) {
              const int r = nefc++;
              OX_MLOOP
              for (int i = 0; i < nv; i++) at(b.efc_J, r * nv + i) = 0;
              at(b.efc_J, r * nv + da) = (T)(-side);
              const T vel = (T)(-side) * at(b.qvel, da);
              T aref;
              T sr[2], si[5];
              OX_LDM(2, sr, jnt_solref, 2 * j);
              OX_LDM(5, si, jnt_solimp, 5 * j);
              const T R = row_params(sr, si, dist, margin, m.dof_invweight0(da), vel, &aref);
              at(b.efc_pos, r) = dist; at(b.efc_margin, r) = margin; at(b.efc_D, r) = 1 / R; at(b.efc_aref, r) = aref;
            }
          }
      }
    }
  }

  // tendon limits: the joint-limit row with J = -side * (tendon coefficients), diagApprox = tendon_invweight0
  OX_HD int tendon_limit_count(int i) const {
    if (!m.tendon_limited(i)) return 0;
    const T value = at(b.ten_length, i), margin = m.tendon_margin(i);
    return (value - m.tendon_range(2 * i) < margin ? 1 : 0) + (m.tendon_range(2 * i + 1) - value < margin ? 1 : 0);
  }
  OX_HD void tendon_limit_rows(int i, int& nefc) const {
    const int nv = m.h().nv;
    if (!m.tendon_limited(i)) return;
    const T value = at(b.ten_length, i), margin = m.tendon_margin(i);
    OX_MLOOP
    for (int side = -1; side <= 1; side += 2) {
      const T dist = side * (m.tendon_range(2 * i + (side + 1) / 2) - value);
      if (dist < margin) {
        const int r = nefc++;
        OX_MLOOP
        for (int k = 0; k < nv; k++) at(b.efc_J, r * nv + k) = 0;
        T vel = 0;
        OX_MLOOP
        for (int k = 0; k < nv; k++) {
          const T c = (T)(-side) * at(b.ten_J, i * nv + k);
          at(b.efc_J, r * nv + k) = c;
          vel += c * at(b.qvel, k);
        }
        T aref, sr[2], si[5];
        OX_LDM(2, sr, tendon_solref_lim, 2 * i);
        OX_LDM(5, si, tendon_solimp_lim, 5 * i);
        const T R = row_params(sr, si, dist, margin, m.tendon_invweight0(i), vel, &aref);
        at(b.efc_pos, r) = dist; at(b.efc_margin, r) = margin; at(b.efc_D, r) = 1 / R; at(b.efc_aref, r) = aref;
      }
    }
  }

  // ============================================================ A.11 primal solver (Newton / CG)
  struct LsPt { T alpha, cost, d0, d1, s0; };  // s0 = sum of |terms| of d0: the resolution of the derivative

  OX_HD LsPt ls_eval(T a, int nefc, T qg0, T qg1, T qg2) const {
    const int ne_ = ati(b.ne, 0), nfe_ = ne_ + ati(b.nf, 0);
    LsPt p;
    p.alpha = a;
    p.cost = a * a * qg2 + a * qg1 + qg0;
    p.d0 = 2 * a * qg2 + qg1;
    p.d1 = 2 * qg2;
    p.s0 = ox_abs(2 * a * qg2) + ox_abs(qg1);
    OX_ROWLOOP
    const bool ell_ = m.h().cone == OX_CONE_ELLIPTIC;
    if (ell_)   // elliptic contacts: exact cost, slope and curvature of the three-zone function along the search direction
      for_each_contact([&](int, int pp, int ea, int dim) {
        if (dim == 1) return;
        T x[6], D[6], jvv[6], f[6], Hc[36], ck;
#pragma unroll
        for (int j = 0; j < 6; j++) { if (j >= dim) break; jvv[j] = at(b.s_Jv, ea + j); x[j] = at(b.s_Jaref, ea + j) + a * jvv[j]; D[j] = at(b.efc_D, ea + j); }
        elliptic_cost(pp, dim, D, x, &ck, f, Hc);
        p.cost += ck;
#pragma unroll
        for (int j = 0; j < 6; j++) {
          if (j >= dim) break;
          p.d0 -= f[j] * jvv[j];
          p.s0 += ox_abs(f[j] * jvv[j]);
#pragma unroll
          for (int l = 0; l < 6; l++) { if (l >= dim) break; p.d1 += jvv[j] * Hc[j * dim + l] * jvv[l]; }
        }
      });
    for (int r = 0; r < nefc; r++) {
      if (ell_ && r >= nfe_ && at(b.efc_floss, r) < 0) continue;
      const T ja = at(b.s_Jaref, r), jv = at(b.s_Jv, r);
      const T x = ja + a * jv;
      if (r >= ne_ && r < nfe_) {  // dry friction: quadratic inside |x| < R floss, linear (force pinned at floss) outside
        const T fl = at(b.efc_floss, r), D = at(b.efc_D, r);
        if (ox_abs(x) * D >= fl) {
          const T sg = x > 0 ? (T)1 : (T)-1;
          p.cost += fl * (sg * x - (T)0.5 * fl / D);
          p.d0 += sg * fl * jv;
          p.s0 += ox_abs(fl * jv);
          continue;
        }
      }
      if (x < 0 || r < nfe_) {  // active at alpha (equality and quadratic-zone friction rows always): 1/2 D x^2 and its first / second derivative in alpha
        const T Dx = at(b.efc_D, r) * x, Dj = at(b.efc_D, r) * jv;
        p.cost += (T)0.5 * Dx * x;
        p.d0 += Dx * jv;
        p.d1 += Dj * jv;
        p.s0 += ox_abs(Dx * jv);
      }
    }
    if (p.d1 < (T)OX_MINVAL) p.d1 = (T)OX_MINVAL;
    return p;
  }

  // efc_force, qfrc_constraint and total cost at the current (qacc, Ma, Jaref); returns cost, gauss via pointer
  OX_HD T update_constraint(int nv, int nefc, T* gauss_out) const {
    const int ne_ = ati(b.ne, 0), nfe_ = ne_ + ati(b.nf, 0);
    T c = 0;
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(b.qfrc_constraint, i) = 0;
    const bool ell_ = m.h().cone == OX_CONE_ELLIPTIC;
    if (ell_)
      for_each_contact([&](int, int pp, int ea, int dim) {
        if (dim == 1) return;
        T x[6], D[6], f[6], ck;
#pragma unroll
        for (int j = 0; j < 6; j++) { if (j >= dim) break; x[j] = at(b.s_Jaref, ea + j); D[j] = at(b.efc_D, ea + j); }
        elliptic_cost(pp, dim, D, x, &ck, f, (T*)nullptr);
        c += ck;
#pragma unroll
        for (int j = 0; j < 6; j++) {
          if (j >= dim) break;
          at(b.efc_force, ea + j) = f[j];
          if (f[j] != 0) {
            OX_NVLOOP
            for (int i = 0; i < nv; i++) at(b.qfrc_constraint, i) += at(b.efc_J, (ea + j) * nv + i) * f[j];
          }
        }
      });
    OX_ROWLOOP
    for (int r = 0; r < nefc; r++) {
      if (ell_ && r >= nfe_ && at(b.efc_floss, r) < 0) continue;
      const T ja = at(b.s_Jaref, r);
      T f = 0;
      if (r >= ne_ && r < nfe_ && ox_abs(ja) * at(b.efc_D, r) >= at(b.efc_floss, r)) {   // dry friction, linear zone
        const T fl = at(b.efc_floss, r), sg = ja > 0 ? (T)1 : (T)-1;
        f = -sg * fl;
        c += fl * (sg * ja - (T)0.5 * fl / at(b.efc_D, r));
        OX_NVLOOP
        for (int i = 0; i < nv; i++) at(b.qfrc_constraint, i) += at(b.efc_J, r * nv + i) * f;
      } else if (ja < 0 || r < nfe_) {
        const T D = at(b.efc_D, r);
        f = -D * ja;
        c += (T)0.5 * D * ja * ja;
        OX_NVLOOP
        for (int i = 0; i < nv; i++) at(b.qfrc_constraint, i) += at(b.efc_J, r * nv + i) * f;
      }
      at(b.efc_force, r) = f;
    }
    T g = 0;
    OX_NVLOOP
    for (int i = 0; i < nv; i++) g += (T)0.5 * (at(b.s_Ma, i) - at(b.qfrc_smooth, i)) * (at(b.qacc, i) - at(b.qacc_smooth, i));
    *gauss_out = g;
    return c + g;
  }

  // grad, Mgrad (Newton: H^-1 grad with H = M + J' D_active J; CG: M^-1 grad); returns |grad|
  OX_HD T update_gradient(int nv, int nefc, bool newton) const {
    T gn = 0;
    OX_NVLOOP
    for (int i = 0; i < nv; i++) {
      const T g = at(b.s_Ma, i) - at(b.qfrc_smooth, i) - at(b.qfrc_constraint, i);
      at(b.s_grad, i) = g;
      gn += g * g;
    }
    if (!newton) {
      OX_NVLOOP
      for (int i = 0; i < nv; i++) at(b.s_Mgrad, i) = at(b.s_grad, i);
      solve_ld(b.s_Mgrad);
      return ox_sqrt(gn);
    }
    T* H = b.s_H;
    const int ne_ = ati(b.ne, 0), nfe_ = ne_ + ati(b.nf, 0);
    OX_NVLOOP
    for (int i = 0; i < nv; i++) {
      OX_NVLOOP
      for (int j = 0; j <= i; j++) at(H, i * nv + j) = 0;
      int adr = m.dof_Madr(i);
      OX_NVLOOP
      for (int d_ = 0, j = i; d_ < m.dof_depth(i); d_++, j = m.dof_parentid(j)) at(H, i * nv + j) = at(b.qM, adr++);
    }
    OX_ROWLOOP
    for (int r = 0; r < nefc; r++) {
      if (r >= ne_ && r < nfe_) { if (ox_abs(at(b.s_Jaref, r)) * at(b.efc_D, r) >= at(b.efc_floss, r)) continue; }   // friction row in its linear zone: no curvature
      else if (m.h().cone == OX_CONE_ELLIPTIC && r >= nfe_ && at(b.efc_floss, r) < 0) continue;   // rows of elliptic contacts: their block follows
      else if (!(at(b.s_Jaref, r) < 0 || r < ne_)) continue;
      const T D = at(b.efc_D, r);
      if constexpr (UNROLL_NV) {  // the row once into registers, then the rank-1 update on register-resident H
        constexpr int NV = M::Hdr::nv;
        T Jr[NV > 0 ? NV : 1];
#pragma unroll
        for (int i = 0; i < NV; i++) Jr[i] = at(b.efc_J, r * NV + i);
#pragma unroll
        for (int i = 0; i < NV; i++) {
          const T s = D * Jr[i];
#pragma unroll
          for (int j = 0; j <= i; j++) at(H, i * NV + j) += s * Jr[j];
        }
      } else {
        for (int i = 0; i < nv; i++) {
          const T ji = at(b.efc_J, r * nv + i);
          if (ji == 0) continue;
          const T s = D * ji;
          for (int j = 0; j <= i; j++) at(H, i * nv + j) += s * at(b.efc_J, r * nv + j);
        }
      }
    }
    if (m.h().cone == OX_CONE_ELLIPTIC)   // J_c' Hc J_c of every elliptic contact outside the top zone (bottom: diag(D); middle: the cone Hessian)
      for_each_contact([&](int, int pp, int ea, int dim) {
        if (dim == 1) return;
        T x[6], D[6], Hc[36], ck;
#pragma unroll
        for (int j = 0; j < 6; j++) { if (j >= dim) break; x[j] = at(b.s_Jaref, ea + j); D[j] = at(b.efc_D, ea + j); }
        if (elliptic_cost(pp, dim, D, x, &ck, (T*)nullptr, Hc) == 0) return;
        for (int a = 0; a < dim; a++)
          for (int c2 = 0; c2 < dim; c2++) {
            const T hh = Hc[a * dim + c2];
            if (hh == 0) continue;
            for (int i = 0; i < nv; i++) {
              const T sa = hh * at(b.efc_J, (ea + a) * nv + i);
              if (sa == 0) continue;
              for (int j = 0; j <= i; j++) at(H, i * nv + j) += sa * at(b.efc_J, (ea + c2) * nv + j);
            }
          }
      });
    OX_NVLOOP
    for (int j = 0; j < nv; j++) {  // Cholesky, lower
      T s = at(H, j * nv + j);
      OX_NVLOOP
      for (int k = 0; k < j; k++) { const T l = at(H, j * nv + k); s -= l * l; }
      const T inv = ox_rsqrt(ox_max(s, (T)OX_MINVAL));
      at(H, j * nv + j) = inv;  // the diagonal holds 1 / L[j][j]: the substitutions below multiply instead of dividing
      OX_NVLOOP
      for (int i = j + 1; i < nv; i++) {
        T v = at(H, i * nv + j);
        OX_NVLOOP
        for (int k = 0; k < j; k++) v -= at(H, i * nv + k) * at(H, j * nv + k);
        at(H, i * nv + j) = v * inv;
      }
    }
    OX_NVLOOP
    for (int i = 0; i < nv; i++) {
      T v = at(b.s_grad, i);
      OX_NVLOOP
      for (int k = 0; k < i; k++) v -= at(H, i * nv + k) * at(b.s_Mgrad, k);
      at(b.s_Mgrad, i) = v * at(H, i * nv + i);
    }
    OX_NVLOOP
    for (int i = nv - 1; i >= 0; i--) {
      T v = at(b.s_Mgrad, i);
      OX_NVLOOP
      for (int k = i + 1; k < nv; k++) v -= at(H, k * nv + i) * at(b.s_Mgrad, k);
      at(b.s_Mgrad, i) = v * at(H, i * nv + i);
    }
    return ox_sqrt(gn);
  }

  OX_HD T cost_at(const T* qacc, int nv, int nefc) const {  // warm-start selection; uses s_Mv as scratch
    const int ne_ = ati(b.ne, 0), nfe_ = ne_ + ati(b.nf, 0);
    const bool ell_ = m.h().cone == OX_CONE_ELLIPTIC;
    mul_m(b.s_Mv, qacc);
    T c = 0;
    OX_NVLOOP
    for (int i = 0; i < nv; i++) c += (T)0.5 * (at(b.s_Mv, i) - at(b.qfrc_smooth, i)) * (at(qacc, i) - at(b.qacc_smooth, i));
    OX_ROWLOOP
    for (int r = 0; r < nefc; r++) {
      T v = -at(b.efc_aref, r);
      OX_NVLOOP
      for (int i = 0; i < nv; i++) v += at(b.efc_J, r * nv + i) * at(qacc, i);
      if (ell_) at(b.s_Jv, r) = v;   // elliptic contacts need all their rows at once (s_Jv is free before the first line search)
      if (ell_ && r >= nfe_ && at(b.efc_floss, r) < 0) continue;
      if (r >= ne_ && r < nfe_ && ox_abs(v) * at(b.efc_D, r) >= at(b.efc_floss, r)) c += at(b.efc_floss, r) * (ox_abs(v) - (T)0.5 * at(b.efc_floss, r) / at(b.efc_D, r));
      else if (v < 0 || r < nfe_) c += (T)0.5 * at(b.efc_D, r) * v * v;
    }
    if (ell_)
      for_each_contact([&](int, int pp, int ea, int dim) {
        if (dim == 1) return;
        T x[6], D[6], ck;
#pragma unroll
        for (int j = 0; j < 6; j++) { if (j >= dim) break; x[j] = at(b.s_Jv, ea + j); D[j] = at(b.efc_D, ea + j); }
        elliptic_cost(pp, dim, D, x, &ck, (T*)nullptr, (T*)nullptr);
        c += ck;
      });
    return c;
  }

  // ---- dual solvers: PGS (mj_solPGS) and the noslip post-pass (mj_solNoSlip). They work on the forces f with
  // A = J M^-1 J' (+ diag R for PGS), b = J qacc_smooth - aref, WITHOUT forming the nefc x nefc matrix: s_Ma carries
  // w = qacc_smooth + M^-1 J' f, so row r's residual is J_r w - aref_r (+ R_r f_r), and a change of f_r by delta moves w by
  // delta * M^-1 J_r' (one sparse L'DL solve into s_Mv). Memory stays O(nefc nv) per env - what the primal solvers use.
  OX_HD T row_dot(int r, const T* v, int nv) const {
    T s = 0;
    OX_NVLOOP
    for (int i = 0; i < nv; i++) s += at(b.efc_J, r * nv + i) * at(v, i);
    return s;
  }
  // qfrc_constraint = J' f;  s_Ma = w = qacc_smooth + M^-1 qfrc_constraint  (mj dual2Primal)
  OX_HD void dual_to_primal(int nv, int nefc) const {
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(b.qfrc_constraint, i) = 0;
    OX_ROWLOOP
    for (int r = 0; r < nefc; r++) {
      const T f = at(b.efc_force, r);
      if (f == 0) continue;
      OX_NVLOOP
      for (int i = 0; i < nv; i++) at(b.qfrc_constraint, i) += at(b.efc_J, r * nv + i) * f;
    }
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(b.s_Ma, i) = at(b.qfrc_constraint, i);
    solve_ld(b.s_Ma);
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(b.s_Ma, i) += at(b.qacc_smooth, i);
  }
  OX_HD void dual_pgs(int nv, int nefc) const {
    const auto& h = m.h();
    const int ne_ = ati(b.ne, 0), nfe_ = ne_ + ati(b.nf, 0);
    const T scale = 1 / ((T)h.meaninertia * (T)(nv > 1 ? nv : 1));
    bool warm = false;
    if (!dis(OX_DSBL_WARMSTART)) {  // forces implied by qacc_warmstart, kept only if their dual cost beats f = 0 (cost 0)
      OX_ROWLOOP
      for (int r = 0; r < nefc; r++) {
        const T jar = row_dot(r, b.qacc_warmstart, nv) - at(b.efc_aref, r);
        T f = (jar < 0 || r < nfe_) ? -at(b.efc_D, r) * jar : (T)0;
        if (r >= ne_ && r < nfe_) f = ox_clip(f, -at(b.efc_floss, r), at(b.efc_floss, r));
        at(b.efc_force, r) = f;
      }
      dual_to_primal(nv, nefc);
      T cost = 0;
      OX_ROWLOOP
      for (int r = 0; r < nefc; r++) {
        const T f = at(b.efc_force, r), jw = row_dot(r, b.s_Ma, nv), js = row_dot(r, b.qacc_smooth, nv);
        cost += (T)0.5 * f * (jw - js + f / at(b.efc_D, r)) + f * (js - at(b.efc_aref, r));
      }
      warm = cost < 0;
    }
    if (!warm) {
      OX_ROWLOOP
      for (int r = 0; r < nefc; r++) at(b.efc_force, r) = 0;
      OX_NVLOOP
      for (int i = 0; i < nv; i++) at(b.s_Ma, i) = at(b.qacc_smooth, i);
    }
    int iter = 0;
    const int maxiter = h.iterations;
    const T tol = (T)h.tolerance;
#pragma unroll 1
    while (iter < maxiter) {
      T improvement = 0;
#pragma unroll 1
      for (int r = 0; r < nefc; r++) {
        OX_NVLOOP
        for (int i = 0; i < nv; i++) at(b.s_Mv, i) = at(b.efc_J, r * nv + i);
        solve_ld(b.s_Mv);
        const T R = 1 / at(b.efc_D, r), ARrr = row_dot(r, b.s_Mv, nv) + R;
        const T old = at(b.efc_force, r);
        const T res = row_dot(r, b.s_Ma, nv) - at(b.efc_aref, r) + R * old;
        T f = old - res / ARrr;
        if (r >= ne_ && r < nfe_) f = ox_clip(f, -at(b.efc_floss, r), at(b.efc_floss, r));   // dry friction: a box
        else if (r >= ne_ && f < 0) f = 0;   // limits and pyramidal contact edges push only; equalities pull both ways
        const T delta = f - old;
        if (delta != 0) {
          at(b.efc_force, r) = f;
          OX_NVLOOP
          for (int i = 0; i < nv; i++) at(b.s_Ma, i) += delta * at(b.s_Mv, i);
          improvement -= (T)0.5 * delta * delta * ARrr + delta * res;
        }
      }
      iter++;
      if (improvement * scale < tol) break;
    }
    ati(b.solver_niter, 0) = iter;
  }
  // one friction direction of one pyramidal contact: edges ra = '+', rb = '-' keep their sum, their difference moves
  OX_HD T noslip_pair(int ra, int rb, int nv) const {
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(b.s_Mv, i) = at(b.efc_J, ra * nv + i) - at(b.efc_J, rb * nv + i);
    solve_ld(b.s_Mv);
    const T K = row_dot(ra, b.s_Mv, nv) - row_dot(rb, b.s_Mv, nv);
    if (K < (T)OX_MINVAL) return 0;
    const T fa = at(b.efc_force, ra), fb = at(b.efc_force, rb), mid = (T)0.5 * (fa + fb), yold = (T)0.5 * (fa - fb);
    const T dres = (row_dot(ra, b.s_Ma, nv) - at(b.efc_aref, ra)) - (row_dot(rb, b.s_Ma, nv) - at(b.efc_aref, rb));
    T y = yold - dres / K;
    y = ox_max(-mid, ox_min(mid, y));
    const T delta = y - yold;
    if (delta == 0) return 0;
    at(b.efc_force, ra) = mid + y; at(b.efc_force, rb) = mid - y;
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(b.s_Ma, i) += delta * at(b.s_Mv, i);
    return -((T)0.5 * delta * delta * K + delta * dres);
  }
  OX_HD T noslip_contact(int c, int p, int nv) const {
    const int ea = ati(b.con_efcadr, c), dim = m.pair_dim(p);
    T imp = 0;
    if (ea < 0 || dim < 3) return imp;
#pragma unroll 1
    for (int k = 0; k < dim - 1; k++) imp += noslip_pair(ea + 2 * k, ea + 2 * k + 1, nv);
    return imp;
  }
  OX_HD void dual_noslip(int nv, int nefc) const {
    const auto& h = m.h();
    const T scale = 1 / ((T)h.meaninertia * (T)(nv > 1 ? nv : 1));
    dual_to_primal(nv, nefc);
#pragma unroll 1
    const int ne_ = ati(b.ne, 0), nfe_ = ne_ + ati(b.nf, 0);
#pragma unroll 1
    for (int iter = 0; iter < h.noslip_iterations; iter++) {
      T improvement = 0;
#pragma unroll 1
      for (int r = ne_; r < nfe_; r++) {   // dry-friction rows: the PGS update without the regulariser
        OX_NVLOOP
        for (int i = 0; i < nv; i++) at(b.s_Mv, i) = at(b.efc_J, r * nv + i);
        solve_ld(b.s_Mv);
        const T Arr = row_dot(r, b.s_Mv, nv);
        if (Arr < (T)OX_MINVAL) continue;
        const T old = at(b.efc_force, r), res = row_dot(r, b.s_Ma, nv) - at(b.efc_aref, r);
        const T f = ox_clip(old - res / Arr, -at(b.efc_floss, r), at(b.efc_floss, r)), delta = f - old;
        if (delta != 0) {
          at(b.efc_force, r) = f;
          OX_NVLOOP
          for (int i = 0; i < nv; i++) at(b.s_Ma, i) += delta * at(b.s_Mv, i);
          improvement -= (T)0.5 * delta * delta * Arr + delta * res;
        }
      }
      if (STATIC_CON || slots) {
#pragma unroll 1
        for (int p = 0; p < h.npair; p++)
#pragma unroll 1
          for (int k = 0; k < m.pair_maxcon(p); k++) {
            const int c = m.pair_conadr(p) + k;
            if (ati(b.con_active, c)) improvement += noslip_contact(c, p, nv);
          }
      } else {
        const int ncon = ati(b.ncon, 0);
#pragma unroll 1
        for (int c = 0; c < ncon; c++) improvement += noslip_contact(c, ati(b.con_pair, c), nv);
      }
      if (improvement * scale < (T)h.noslip_tolerance) break;
    }
  }
  OX_HD void dual_finish(int nv, int nefc) const {
    dual_to_primal(nv, nefc);
    OX_NVLOOP
    for (int i = 0; i < nv; i++) { const T a = at(b.s_Ma, i); at(b.qacc, i) = a; at(b.qacc_warmstart, i) = a; }
  }

  OX_HDN void fwd_constraint() const {
    const auto& h = m.h();
    const int nv = h.nv, nefc = ati(b.nefc, 0);
    if (nefc == 0) {
      OX_NVLOOP
      for (int i = 0; i < nv; i++) {
        const T a = at(b.qacc_smooth, i);
        at(b.qacc, i) = a; at(b.qacc_warmstart, i) = a; at(b.qfrc_constraint, i) = 0;
      }
      ati(b.solver_niter, 0) = 0;
      return;
    }
    if (h.solver == OX_SOL_PGS) {
      dual_pgs(nv, nefc);
      if (h.noslip_iterations > 0) dual_noslip(nv, nefc);
      dual_finish(nv, nefc);
      return;
    }
    const bool newton = h.solver == OX_SOL_NEWTON;
    bool use_smooth = true;
    if (!dis(OX_DSBL_WARMSTART)) {
      T cand[2];
#pragma unroll 1
      for (int k = 0; k < 2; k++) {  // cost(qacc_warmstart), cost(qacc_smooth) through one evaluation site
        OX_NVLOOP
        for (int i = 0; i < nv; i++) at(b.qacc, i) = k ? at(b.qacc_smooth, i) : at(b.qacc_warmstart, i);
        cand[k] = cost_at(b.qacc, nv, nefc);
      }
      use_smooth = cand[0] > cand[1];
    }
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(b.qacc, i) = use_smooth ? at(b.qacc_smooth, i) : at(b.qacc_warmstart, i);
    // initial state
    mul_m(b.s_Ma, b.qacc);
    OX_ROWLOOP
    for (int r = 0; r < nefc; r++) {
      T v = -at(b.efc_aref, r);
      OX_NVLOOP
      for (int i = 0; i < nv; i++) v += at(b.efc_J, r * nv + i) * at(b.qacc, i);
      at(b.s_Jaref, r) = v;
    }
    // One loop, one call site of update_constraint / update_gradient (the first pass evaluates the starting point, every
    // later pass follows a line search and a move) and one of ls_eval (its first trip evaluates alpha = 0): the iteration
    // body is what the instruction cache has to hold, and duplicated inlined copies are pure fetch traffic.
    T gauss = 0, cost = 0, gnorm = 0;
    const T tol = (T)h.tolerance;
    const T mscale = (T)h.meaninertia * (T)(nv > 1 ? nv : 1);
    const T scale = 1 / mscale;
    const int maxiter = h.iterations;
    int iter = 0;
    bool init = true;
#pragma unroll 1
    for (;;) {
      if (!init) {
        if (iter >= maxiter) break;
        // ---- exact line search on the convex piecewise-quadratic phi(alpha)
        T snorm = 0;
        OX_NVLOOP
        for (int i = 0; i < nv; i++) { const T s = at(b.s_search, i); snorm += s * s; }
        snorm = ox_sqrt(snorm);
        if (snorm < (T)OX_MINVAL) break;
        const T gtol = tol * (T)h.ls_tolerance * snorm * mscale;
        mul_m(b.s_Mv, b.s_search);
        OX_ROWLOOP
        for (int r = 0; r < nefc; r++) {
          T v = 0;
          OX_NVLOOP
          for (int i = 0; i < nv; i++) v += at(b.efc_J, r * nv + i) * at(b.s_search, i);
          at(b.s_Jv, r) = v;
        }
        T qg1 = 0, qg2 = 0;
        OX_NVLOOP
        for (int i = 0; i < nv; i++) {
          const T s = at(b.s_search, i);
          qg1 += s * (at(b.s_Ma, i) - at(b.qfrc_smooth, i));
          qg2 += (T)0.5 * s * at(b.s_Mv, i);
        }
        LsPt p0, lo, hi, cur;
        bool have_hi = false, descent = true;
        T a = 0;
#pragma unroll 1
        for (int it = -1; it < h.ls_iterations; it++) {  // it = -1 evaluates alpha = 0
          if (it >= 0) {
            a = cur.alpha - cur.d0 / cur.d1;
            if (have_hi && !(a > lo.alpha && a < hi.alpha)) a = (T)0.5 * (lo.alpha + hi.alpha);
            if (ox_abs(a - cur.alpha) <= Eps<T>::v() * ox_abs(a)) break;
          }
          cur = ls_eval(a, nefc, gauss, qg1, qg2);
          if (it < 0) {
            p0 = cur; lo = cur; hi = cur;
            if (!(p0.d0 < 0)) { descent = false; break; }
          } else {
            if (ox_abs(cur.d0) < gtol || ox_abs(cur.d0) <= 8 * Eps<T>::v() * cur.s0) break;  // converged, or phi' below its own round-off
            if (cur.d0 < 0) lo = cur; else { hi = cur; have_hi = true; }
          }
        }
        if (!descent) break;
        const T alpha = cur.cost <= p0.cost ? cur.alpha : 0;
        if (alpha == 0) break;
        // ---- move
        OX_NVLOOP
        for (int i = 0; i < nv; i++) {
          at(b.qacc, i) += alpha * at(b.s_search, i);
          at(b.s_Ma, i) += alpha * at(b.s_Mv, i);
          if (!newton) { at(b.s_gradold, i) = at(b.s_grad, i); at(b.s_Mgradold, i) = at(b.s_Mgrad, i); }
        }
        OX_ROWLOOP
        for (int r = 0; r < nefc; r++) at(b.s_Jaref, r) += alpha * at(b.s_Jv, r);
      }
      const T oldcost = cost;
      cost = update_constraint(nv, nefc, &gauss);
      gnorm = update_gradient(nv, nefc, newton);
      bool first = init;
      if (init) {
        init = false;
        if (scale * gnorm < tol) break;
      } else {
        iter++;
        const T improvement = scale * (oldcost - cost), gradient = scale * gnorm;
        if (improvement < tol || gradient < tol) break;
        // floating-point floor: a decrease below the resolution of the cost itself is round-off, not progress
        // (inert in fp64 at MuJoCo's tolerances; in fp32 it removes the noise-driven iteration tail)
        if (oldcost - cost <= OX_FLOOR_MULT * Eps<T>::v() * (ox_abs(oldcost) + ox_abs(cost))) break;
      }
      if (newton || first) {
        OX_NVLOOP
        for (int i = 0; i < nv; i++) at(b.s_search, i) = -at(b.s_Mgrad, i);
      } else {
        T num = 0, den = 0;
        OX_NVLOOP
        for (int i = 0; i < nv; i++) {
          num += at(b.s_grad, i) * (at(b.s_Mgrad, i) - at(b.s_Mgradold, i));
          den += at(b.s_gradold, i) * at(b.s_Mgradold, i);
        }
        T beta = num / ox_max((T)OX_MINVAL, den);
        if (beta < 0) beta = 0;
        OX_NVLOOP
        for (int i = 0; i < nv; i++) at(b.s_search, i) = -at(b.s_Mgrad, i) + beta * at(b.s_search, i);
      }
    }
    ati(b.solver_niter, 0) = iter;
    if (h.noslip_iterations > 0) {
      dual_noslip(nv, nefc);
      dual_finish(nv, nefc);
      return;
    }
    OX_NVLOOP
    for (int i = 0; i < nv; i++) at(b.qacc_warmstart, i) = at(b.qacc, i);
  }

  // ============================================================ ray - site volume (mju_rayGeom for sphere / capsule / box)
  // smallest non-negative root of a x^2 + 2 b x + c = 0 (both roots in xx), -1 if none
  static OX_HD T ray_quad(T a, T bq, T c, T* xx) {
    const T det = bq * bq - a * c;
    if (det < (T)OX_MINVAL) { xx[0] = -1; xx[1] = -1; return -1; }
    const T sq = ox_sqrt(det);
    xx[0] = (-bq - sq) / a; xx[1] = (-bq + sq) / a;
    if (xx[0] >= 0) return xx[0];
    if (xx[1] >= 0) return xx[1];
    return -1;
  }
  static OX_HD T ray_geom(const T* pos, const T* mat, const T* size, const T* pnt, const T* vec, int type) {
    const T dif[3] = {pnt[0] - pos[0], pnt[1] - pos[1], pnt[2] - pos[2]};
    T lp[3], lv[3];   // ray in the site frame
#pragma unroll
    for (int k = 0; k < 3; k++) {
      lp[k] = mat[k] * dif[0] + mat[3 + k] * dif[1] + mat[6 + k] * dif[2];
      lv[k] = mat[k] * vec[0] + mat[3 + k] * vec[1] + mat[6 + k] * vec[2];
    }
    T xx[2];
    if (type == OX_GEOM_SPHERE) return ray_quad(dot3(lv, lv), dot3(lv, lp), dot3(lp, lp) - size[0] * size[0], xx);
    T x = -1;
    if (type == OX_GEOM_BOX) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        if (ox_abs(lv[i]) > (T)OX_MINVAL) {
#pragma unroll
          for (int side = -1; side <= 1; side += 2) {
            const T sol = ((T)side * size[i] - lp[i]) / lv[i];
            if (sol >= 0) {
              const int i0 = (i + 1) % 3, i1 = (i + 2) % 3;
              const T p0 = lp[i0] + sol * lv[i0], p1 = lp[i1] + sol * lv[i1];
              if (ox_abs(p0) <= size[i0] && ox_abs(p1) <= size[i1] && (x < 0 || sol < x)) x = sol;
            }
          }
        }
      }
      return x;
    }
    // capsule: the cylinder between the flat ends, then the outer halves of the two cap spheres
    T sol = ray_quad(lv[0] * lv[0] + lv[1] * lv[1], lv[0] * lp[0] + lv[1] * lp[1], lp[0] * lp[0] + lp[1] * lp[1] - size[0] * size[0], xx);
    if (sol >= 0 && ox_abs(lp[2] + sol * lv[2]) <= size[1]) x = sol;
#pragma unroll
    for (int side = -1; side <= 1; side += 2) {
      const T ld_[3] = {lp[0], lp[1], lp[2] - (T)side * size[1]};
      ray_quad(dot3(lv, lv), dot3(lv, ld_), dot3(ld_, ld_) - size[0] * size[0], xx);
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const T z = lp[2] + xx[i] * lv[2];
        if (xx[i] >= 0 && (side > 0 ? z >= size[1] : z <= -size[1]) && (x < 0 || xx[i] < x)) x = xx[i];
      }
    }
    return x;
  }

  // ============================================================ sensors (N2 subset)
  OX_HD void obj_frame(int objtype, int id, T* pos, T* mat, int* body) const {
    if (objtype == OX_OBJ_BODY) { ld<3>(pos, b.xipos, 3 * id); ld<9>(mat, b.ximat, 9 * id); *body = id; }
    else if (objtype == OX_OBJ_XBODY) { ld<3>(pos, b.xpos, 3 * id); ld<9>(mat, b.xmat, 9 * id); *body = id; }
    else if (objtype == OX_OBJ_GEOM) { ld<3>(pos, b.geom_xpos, 3 * id); ld<9>(mat, b.geom_xmat, 9 * id); *body = m.geom_bodyid(id); }
    else { ld<3>(pos, b.site_xpos, 3 * id); ld<9>(mat, b.site_xmat, 9 * id); *body = m.site_bodyid(id); }
  }
  // body accelerations in the com frame (forward pass of mj_rnePostConstraint): cacc[0] = (0, -gravity),
  // cacc[b] = cacc[parent] + sum over the body's dofs of cdof_dot*qvel + cdof*qacc (b.cacc is free after rne())
  OX_HD void body_acc() const {
    const auto& h = m.h();
    OX_MLOOP
    for (int k = 0; k < 3; k++) { at(b.cacc, k) = 0; at(b.cacc, 3 + k) = dis(OX_DSBL_GRAVITY) ? (T)0 : -(T)h.grav(k); }
    OX_MLOOP
    for (int bd = 1; bd < h.nbody; bd++) {
      T a[6];
      ld<6>(a, b.cacc, 6 * m.body_parentid(bd));
      OX_MLOOP
      for (int d_ = 0; d_ < m.body_dofnum(bd); d_++) {
        const int i = m.body_dofadr(bd) + d_;
        T cd[6], cdd[6];
        ld<6>(cd, b.cdof, 6 * i);
        ld<6>(cdd, b.cdof_dot, 6 * i);
        const T v = at(b.qvel, i), qa = at(b.qacc, i);
        OX_MLOOP
        for (int k = 0; k < 6; k++) a[k] += cdd[k] * v + cd[k] * qa;
      }
      st<6>(b.cacc, 6 * bd, a);
    }
  }
  // b.cfrc[body] -= sign * (Cartesian force at `point` [+ torque]) moved to the com-frame origin subtree_com[root]
  OX_HD void sub_ext_force(int body, const T* point, const T* force, const T* torque, T sign) const {
    if (body == 0) return;
    T sc[3], dif[3], t[3];
    ld<3>(sc, b.subtree_com, 3 * m.body_rootid(body));
    dif[0] = point[0] - sc[0]; dif[1] = point[1] - sc[1]; dif[2] = point[2] - sc[2];
    cross3(t, dif, force);
    OX_MLOOP
    for (int k = 0; k < 3; k++) {
      at(b.cfrc, 6 * body + k) -= sign * (t[k] + (torque ? torque[k] : (T)0));
      at(b.cfrc, 6 * body + 3 + k) -= sign * force[k];
    }
  }
  // mj_rnePostConstraint: b.cfrc = cfrc_int, the com-based [torque; force] each body exchanges with its parent
  // (needs body_acc() first; b.cfrc is free after rne())
  OX_HD void cfrc_int() const {
    const auto& h = m.h();
    const int nbody = h.nbody;
    OX_MLOOP
    for (int bd = 1; bd < nbody; bd++) {   // inertial part: cinert cacc + cvel x* (cinert cvel)
      T in[10], ca[6], cv[6], ia[6], iv[6], cf[6];
      ld<10>(in, b.cinert, 10 * bd);
      ld<6>(ca, b.cacc, 6 * bd);
      ld<6>(cv, b.cvel, 6 * bd);
      mul_inert_vec(ia, in, ca);
      mul_inert_vec(iv, in, cv);
      cross_force(cf, cv, iv);
      OX_MLOOP
      for (int k = 0; k < 6; k++) ia[k] += cf[k];
      st<6>(b.cfrc, 6 * bd, ia);
    }
    OX_MLOOP
    for (int bd = 1; bd < nbody; bd++) {   // minus the applied Cartesian forces
      T f[6], xi[3];
      ld<6>(f, b.xfrc_applied, 6 * bd);
      if (f[0] == 0 && f[1] == 0 && f[2] == 0 && f[3] == 0 && f[4] == 0 && f[5] == 0) continue;
      ld<3>(xi, b.xipos, 3 * bd);
      sub_ext_force(bd, xi, f, f + 3, (T)1);
    }
    // minus the contact forces (mj_contactForce: pyramid normal = sum of the edge forces, tangent k = (f+ - f-) mu_k): geom2's body
    // is pushed along +normal, geom1's body the other way
    auto one = [&](int c, int p) {
      const int ea = ati(b.con_efcadr, c);
      if (ea < 0) return;
      const int dim = m.pair_dim(p);
      T lf[6] = {0, 0, 0, 0, 0, 0};   // contact-frame force (normal, tangents) and torque (torsion, rolling)
      if (dim == 1) lf[0] = at(b.efc_force, ea);
      else if (h.cone == OX_CONE_ELLIPTIC) {   // the rows ARE the contact-frame components
#pragma unroll
        for (int k = 0; k < 6; k++) { if (k >= dim) break; lf[k] = at(b.efc_force, ea + k); }
      } else {
#pragma unroll
        for (int k = 0; k < 5; k++) {
          if (k >= dim - 1) break;
          const T fp = at(b.efc_force, ea + 2 * k), fm = at(b.efc_force, ea + 2 * k + 1);
          lf[0] += fp + fm;
          lf[1 + k] = (fp - fm) * m.pair_friction(5 * p + k);
        }
      }
      T fr[9], cp[3], wf[3], wt[3];
      ld<9>(fr, b.con_frame, 9 * c);
      ld<3>(cp, b.con_pos, 3 * c);
      OX_MLOOP
      for (int k = 0; k < 3; k++) {
        wf[k] = fr[k] * lf[0] + fr[3 + k] * lf[1] + fr[6 + k] * lf[2];
        wt[k] = fr[k] * lf[3] + fr[3 + k] * lf[4] + fr[6 + k] * lf[5];
      }
      sub_ext_force(m.geom_bodyid(m.pair_geom1(p)), cp, wf, dim > 3 ? wt : nullptr, (T)-1);
      sub_ext_force(m.geom_bodyid(m.pair_geom2(p)), cp, wf, dim > 3 ? wt : nullptr, (T)1);
    };
    if (STATIC_CON || slots) {
      OX_MLOOP
      for (int p = 0; p < h.npair; p++) {
        OX_MLOOP
        for (int k = 0; k < m.pair_maxcon(p); k++) {
          const int c = m.pair_conadr(p) + k;
          if (ati(b.con_active, c)) one(c, p);
        }
      }
    } else {
      const int ncon = ati(b.ncon, 0);
      for (int c = 0; c < ncon; c++) one(c, ati(b.con_pair, c));
    }
    // minus the connect-equality forces: a world-frame force on body1 at its anchor, the opposite on body2 at its own
    if (h.neq > 0 && ati(b.ne, 0) > 0) {
      int r = 0;
      OX_MLOOP
      for (int i = 0; i < h.neq; i++) {
        if (!(at(b.eq_active, i) != 0)) continue;
        if (m.eq_type(i) == OX_EQ_JOINT) { r += 1; continue; }
        T f[3] = {at(b.efc_force, r), at(b.efc_force, r + 1), at(b.efc_force, r + 2)};
        r += 3;
        const int bodies[2] = {m.eq_obj1id(i), m.eq_obj2id(i)};
        OX_MLOOP
        for (int s = 0; s < 2; s++) {
          T mat[9], xp[3], a[3], w[3], pt[3];
          ld<9>(mat, b.xmat, 9 * bodies[s]);
          ld<3>(xp, b.xpos, 3 * bodies[s]);
          OX_LDM(3, a, eq_data, 11 * i + 3 * s);
          mat_vec3(w, mat, a);
          pt[0] = xp[0] + w[0]; pt[1] = xp[1] + w[1]; pt[2] = xp[2] + w[2];
          sub_ext_force(bodies[s], pt, f, nullptr, s == 0 ? (T)1 : (T)-1);
        }
        if (m.eq_type(i) == OX_EQ_WELD) {   // the three orientation rows are a torque pair: tau = G' f on body1, -tau on body2
          T G[9], res[3], tau[3], xp[3];
          weld_rot_map(i, bodies[0], bodies[1], G, res);
          const T fr[3] = {at(b.efc_force, r), at(b.efc_force, r + 1), at(b.efc_force, r + 2)};
          r += 3;
          const T zero[3] = {0, 0, 0};
          OX_MLOOP
          for (int c = 0; c < 3; c++) tau[c] = G[c] * fr[0] + G[3 + c] * fr[1] + G[6 + c] * fr[2];
          OX_MLOOP
          for (int s = 0; s < 2; s++) {
            ld<3>(xp, b.xpos, 3 * bodies[s]);
            sub_ext_force(bodies[s], xp, zero, tau, s == 0 ? (T)1 : (T)-1);
          }
        }
      }
    }
    OX_MLOOP
    for (int bd = nbody - 1; bd > 0; bd--) {   // leaves to root
      const int p = m.body_parentid(bd);
      if (p == 0) continue;
      T c[6], pc[6];
      ld<6>(c, b.cfrc, 6 * bd);
      ld<6>(pc, b.cfrc, 6 * p);
      OX_MLOOP
      for (int k = 0; k < 6; k++) pc[k] += c[k];
      st<6>(b.cfrc, 6 * p, pc);
    }
  }
  OX_HDN void sensors() const {
    const auto& h = m.h();
    const int ns = h.nsensor, nbody = h.nbody;
    bool have_slv = false, have_cacc = false, have_cfrc = false;
    OX_MLOOP
    for (int s = 0; s < ns; s++) {
      const int adr = m.sensor_adr(s), id = m.sensor_objid(s), ot = m.sensor_objtype(s), ty = m.sensor_type(s);
      switch (ty) {
        case OX_SENS_JOINTPOS: at(b.sensordata, adr) = at(b.qpos, m.jnt_qposadr(id)); break;
        case OX_SENS_JOINTVEL: at(b.sensordata, adr) = at(b.qvel, m.jnt_dofadr(id)); break;
        case OX_SENS_ACTUATORPOS: at(b.sensordata, adr) = m.actuator_gear(id) * act_length(id); break;
        case OX_SENS_ACTUATORVEL: at(b.sensordata, adr) = m.actuator_gear(id) * act_velocity(id); break;
        case OX_SENS_ACTUATORFRC: at(b.sensordata, adr) = at(b.actuator_force, id); break;
        case OX_SENS_JOINTACTFRC: at(b.sensordata, adr) = at(b.qfrc_actuator, m.jnt_dofadr(id)); break;
        case OX_SENS_BALLQUAT: {   // a normalised copy of the joint's quaternion
          T q[4];
          ld<4>(q, b.qpos, m.jnt_qposadr(id));
          normalize4(q);
          st<4>(b.sensordata, adr, q);
          break;
        }
        case OX_SENS_BALLANGVEL: for (int k = 0; k < 3; k++) at(b.sensordata, adr + k) = at(b.qvel, m.jnt_dofadr(id) + k); break;
        case OX_SENS_TENDONPOS: at(b.sensordata, adr) = at(b.ten_length, id); break;
        case OX_SENS_TENDONVEL: at(b.sensordata, adr) = tendon_velocity(id); break;
        case OX_SENS_SUBTREECOM: for (int k = 0; k < 3; k++) at(b.sensordata, adr + k) = at(b.subtree_com, 3 * id + k); break;
        case OX_SENS_SUBTREELINVEL: {
          if (!have_slv) {
            OX_MLOOP
            for (int i = 0; i < 3 * nbody; i++) at(b.subtree_linvel, i) = 0;
            OX_MLOOP
            for (int bd = nbody - 1; bd > 0; bd--) {
              T xi[3], sc[3], cv[6], dif[3], v[3], acc[3], pa[3];
              ld<3>(xi, b.xipos, 3 * bd);
              ld<3>(sc, b.subtree_com, 3 * m.body_rootid(bd));
              ld<6>(cv, b.cvel, 6 * bd);
              dif[0] = xi[0] - sc[0]; dif[1] = xi[1] - sc[1]; dif[2] = xi[2] - sc[2];
              cross3(v, cv, dif);
              ld<3>(acc, b.subtree_linvel, 3 * bd);
              const T ms = m.body_mass(bd);
              OX_MLOOP
              for (int k = 0; k < 3; k++) acc[k] += ms * (cv[3 + k] + v[k]);
              st<3>(b.subtree_linvel, 3 * bd, acc);
              const int p = m.body_parentid(bd);
              ld<3>(pa, b.subtree_linvel, 3 * p);
              OX_MLOOP
              for (int k = 0; k < 3; k++) pa[k] += acc[k];
              st<3>(b.subtree_linvel, 3 * p, pa);
            }
            OX_MLOOP
            for (int bd = 0; bd < nbody; bd++) {
              const T inv = 1 / ox_max((T)OX_MINVAL, m.body_subtreemass(bd));
              OX_MLOOP
              for (int k = 0; k < 3; k++) at(b.subtree_linvel, 3 * bd + k) *= inv;
            }
            have_slv = true;
          }
          OX_MLOOP
          for (int k = 0; k < 3; k++) at(b.sensordata, adr + k) = at(b.subtree_linvel, 3 * id + k);
          break;
        }
        case OX_SENS_FRAMEPOS: case OX_SENS_FRAMEQUAT: case OX_SENS_FRAMELINVEL: case OX_SENS_FRAMEANGVEL:
        case OX_SENS_FRAMEXAXIS: case OX_SENS_FRAMEYAXIS: case OX_SENS_FRAMEZAXIS:
        case OX_SENS_VELOCIMETER: case OX_SENS_GYRO: {
          T pos[3], mat[9];
          int body;
          obj_frame(ot, id, pos, mat, &body);
          if (ty == OX_SENS_FRAMEPOS) { st<3>(b.sensordata, adr, pos); break; }
          if (ty == OX_SENS_FRAMEQUAT) { T q[4]; mat2quat(q, mat); st<4>(b.sensordata, adr, q); break; }
          if (ty >= OX_SENS_FRAMEXAXIS && ty <= OX_SENS_FRAMEZAXIS) {   // a column of the frame's rotation matrix
            const int c = ty - OX_SENS_FRAMEXAXIS;
            T ax[3] = {mat[c], mat[3 + c], mat[6 + c]};
            st<3>(b.sensordata, adr, ax);
            break;
          }
          T cv[6], sc[3], dif[3], tmp[3], lin[3], out[3];
          ld<6>(cv, b.cvel, 6 * body);
          ld<3>(sc, b.subtree_com, 3 * m.body_rootid(body));
          dif[0] = pos[0] - sc[0]; dif[1] = pos[1] - sc[1]; dif[2] = pos[2] - sc[2];
          cross3(tmp, cv, dif);
          lin[0] = cv[3] + tmp[0]; lin[1] = cv[4] + tmp[1]; lin[2] = cv[5] + tmp[2];
          const T* src = (ty == OX_SENS_FRAMELINVEL || ty == OX_SENS_VELOCIMETER) ? lin : cv;
          if (ty == OX_SENS_VELOCIMETER || ty == OX_SENS_GYRO) {
            OX_MLOOP
            for (int k = 0; k < 3; k++) out[k] = mat[k] * src[0] + mat[3 + k] * src[1] + mat[6 + k] * src[2];
          } else { out[0] = src[0]; out[1] = src[1]; out[2] = src[2]; }
          st<3>(b.sensordata, adr, out);
          break;
        }
        case OX_SENS_FRAMELINACC: case OX_SENS_FRAMEANGACC: case OX_SENS_ACCELEROMETER: {
          // framelinacc / frameangacc: the same object acceleration, of any frame object, left in world axes
          // mj_objectAcceleration(local) on top of mj_rnePostConstraint's forward pass: cacc[0] = (0, -gravity),
          // cacc[b] = cacc[parent] + sum over the body's dofs of cdof_dot*qvel + cdof*qacc (b.cacc is free after rne())
          if (!have_cacc) { body_acc(); have_cacc = true; }
          T pos[3], mat[9];
          int body;
          obj_frame(ty == OX_SENS_ACCELEROMETER ? (int)OX_OBJ_SITE : ot, id, pos, mat, &body);
          T cv[6], ca[6], sc[3], dif[3], lv[3], la[3], t1[3], t2[3], t3[3];
          ld<6>(cv, b.cvel, 6 * body);
          ld<6>(ca, b.cacc, 6 * body);
          ld<3>(sc, b.subtree_com, 3 * m.body_rootid(body));
          dif[0] = pos[0] - sc[0]; dif[1] = pos[1] - sc[1]; dif[2] = pos[2] - sc[2];
          cross3(t1, cv, dif);
          cross3(t2, ca, dif);
          OX_MLOOP
          for (int k = 0; k < 3; k++) { lv[k] = cv[3 + k] + t1[k]; la[k] = ca[3 + k] + t2[k]; }
          cross3(t3, cv, lv);
          T out[3];
          OX_MLOOP
          for (int k = 0; k < 3; k++) la[k] += t3[k];
          if (ty == OX_SENS_FRAMELINACC) { st<3>(b.sensordata, adr, la); break; }
          if (ty == OX_SENS_FRAMEANGACC) { st<3>(b.sensordata, adr, ca); break; }
          OX_MLOOP
          for (int k = 0; k < 3; k++) out[k] = mat[k] * la[0] + mat[3 + k] * la[1] + mat[6 + k] * la[2];
          st<3>(b.sensordata, adr, out);
          break;
        }
        case OX_SENS_TOUCH: {
          // mj_sensorAcc, mjSENS_TOUCH: sum of the normal forces of the contacts that involve the site's body and whose
          // point lies in the site's volume - decided by a ray from the contact point along the contact normal (flipped
          // when the sensorised body is the second one), which always hits when the point is inside
          const int sbody = m.site_bodyid(id);
          T spos[3], smat[9], ssize[3], total = 0;
          ld<3>(spos, b.site_xpos, 3 * id);
          ld<9>(smat, b.site_xmat, 9 * id);
          OX_LDM(3, ssize, site_size, 3 * id);
          const int stype = m.site_type(id);
          auto one = [&](int c, int p) {
            const int b1 = m.geom_bodyid(m.pair_geom1(p)), b2 = m.geom_bodyid(m.pair_geom2(p));
            if (sbody != b1 && sbody != b2) return;
            const int ea = ati(b.con_efcadr, c);
            if (ea < 0) return;
            const int dim = m.pair_dim(p);
            T fn = at(b.efc_force, ea);                                   // frictionless / elliptic: the (first) row force; pyramid: sum over the edges
            if (dim > 1 && h.cone != OX_CONE_ELLIPTIC) {
#pragma unroll
              for (int r = 1; r < 10; r++) { if (r >= 2 * (dim - 1)) break; fn += at(b.efc_force, ea + r); }
            }
            if (!(fn > 0)) return;
            T ray[3], cp[3];
            ld<3>(ray, b.con_frame, 9 * c);
            ld<3>(cp, b.con_pos, 3 * c);
            if (sbody == b2) { ray[0] = -ray[0]; ray[1] = -ray[1]; ray[2] = -ray[2]; }
            if (ray_geom(spos, smat, ssize, cp, ray, stype) >= 0) total += fn;
          };
          if (STATIC_CON || slots) {
            OX_MLOOP
            for (int p = 0; p < h.npair; p++) {
              OX_MLOOP
              for (int k = 0; k < m.pair_maxcon(p); k++) {
                const int c = m.pair_conadr(p) + k;
                if (ati(b.con_active, c)) one(c, p);
              }
            }
          } else {
            const int ncon = ati(b.ncon, 0);
            for (int c = 0; c < ncon; c++) one(c, ati(b.con_pair, c));
          }
          at(b.sensordata, adr) = total;
          break;
        }
        case OX_SENS_FORCE: case OX_SENS_TORQUE: {
          // interaction force / torque between the site's body and its parent, at the site, in the site frame (mj_sensorAcc)
          if (!have_cacc) { body_acc(); have_cacc = true; }
          if (!have_cfrc) { cfrc_int(); have_cfrc = true; }
          T pos[3], mat[9], ci[6], sc[3], dif[3], t[3], v[3], out[3];
          int body;
          obj_frame(OX_OBJ_SITE, id, pos, mat, &body);
          ld<6>(ci, b.cfrc, 6 * body);
          ld<3>(sc, b.subtree_com, 3 * m.body_rootid(body));
          dif[0] = pos[0] - sc[0]; dif[1] = pos[1] - sc[1]; dif[2] = pos[2] - sc[2];
          cross3(t, dif, ci + 3);
          OX_MLOOP
          for (int k = 0; k < 3; k++) v[k] = ty == OX_SENS_FORCE ? ci[3 + k] : ci[k] - t[k];
          OX_MLOOP
          for (int k = 0; k < 3; k++) out[k] = mat[k] * v[0] + mat[3 + k] * v[1] + mat[6 + k] * v[2];
          st<3>(b.sensordata, adr, out);
          break;
        }
        case OX_SENS_CLOCK: at(b.sensordata, adr) = at(b.time, 0); break;
        default: break;
      }
    }
  }

  // ============================================================ forward (mj_forwardSkip)
  OX_HD void fwd_position() const { kinematics(); com_pos(); crb(); factor_m(); collision(); }
  OX_HD void fwd_velocity() const { com_vel(); passive(); rne(); }
  OX_HDN void forward(bool skipsensor) const {
    fwd_position();
    fwd_velocity();
    make_constraint();
    actuation();
    fwd_acceleration();
    fwd_constraint();
    if (!skipsensor) sensors();
  }

  // ============================================================ A.12 integration
  OX_HD void integrate_pos(T* qpos, const T* qvel, T dt) const {
    const int njnt = m.h().njnt;
    OX_MLOOP
    for (int j = 0; j < njnt; j++) integrate_pos_joint(qpos, qvel, dt, j);
  }
  OX_HD void integrate_pos_joint(T* qpos, const T* qvel, T dt, int j) const {
    {
      int pa = m.jnt_qposadr(j), va = m.jnt_dofadr(j);
      const int jt = m.jnt_type(j);
      if (jt == OX_JNT_FREE) {
        OX_MLOOP
        for (int i = 0; i < 3; i++) at(qpos, pa + i) += dt * at(qvel, va + i);
        pa += 3; va += 3;
      }
      if (jt == OX_JNT_FREE || jt == OX_JNT_BALL) {
        T q[4], w[3] = {at(qvel, va), at(qvel, va + 1), at(qvel, va + 2)};
        ld<4>(q, qpos, pa);
        quat_integrate(q, w, dt);
        st<4>(qpos, pa, q);
      } else {
        at(qpos, pa) += dt * at(qvel, va);
      }
    }
  }
  // mj_nextActivation: explicit Euler, or the exact solution of the first-order filter over one step; clamped to actrange
  OX_HD T next_activation(int i, T act, T act_dot) const {
    const T dt = (T)m.h().timestep;
    if (m.actuator_dyntype(i) == OX_DYN_FILTEREXACT) {
      const T tau = ox_max((T)OX_MINVAL, m.actuator_dynprm(3 * i));
      act += act_dot * tau * ((T)1 - ox_exp(-dt / tau));
    } else {
      act += act_dot * dt;
    }
    if (m.actuator_actlimited(i)) act = ox_clip(act, m.actuator_actrange(2 * i), m.actuator_actrange(2 * i + 1));
    return act;
  }
  OX_HD void advance_act(const T* act_dot) const {
    const int nu = m.h().nu;
    if (m.h().na == 0) return;
    OX_MLOOP
    for (int i = 0; i < nu; i++) {
      if (m.actuator_dyntype(i) == OX_DYN_NONE) continue;
      const int aa = m.actuator_actadr(i);
      at(b.act, aa) = next_activation(i, at(b.act, aa), at(act_dot, aa));
    }
  }
  OX_HD void advance(const T* qacc, const T* qvel_override, const T* act_dot) const {
    const auto& h = m.h();
    const T dt = (T)h.timestep;
    advance_act(act_dot);
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) at(b.qvel, i) += dt * at(qacc, i);
    integrate_pos(b.qpos, qvel_override ? qvel_override : b.qvel, dt);
    at(b.time, 0) += dt;
  }
  OX_HDN void euler() const {
    const auto& h = m.h();
    const int nv = h.nv, nM = h.nM, nu = h.nu;
    const bool fast = h.integrator == OX_INT_IMPLICITFAST;
    if (!fast && (!h.any_damping || dis(OX_DSBL_EULERDAMP))) { advance(b.qacc, nullptr, b.act_dot); return; }
    // implicit-in-velocity joint damping: (M + h B) qacc' = qfrc_smooth + qfrc_constraint; like MuJoCo
    // the factor of M in qLD is overwritten by the factor of M + h B.
    // implicitfast (mj_implicit with the RNE derivatives dropped): M - h dqfrc_smooth/dqvel, which for this subset (joint
    // transmissions, joint damping) is diagonal: B_i = damping_i - sum over actuators on dof i of gear^2 dforce/dvelocity,
    // so the same sparse L'DL serves; an actuator whose force sits on its forcerange bound contributes nothing.
    const T dt = (T)h.timestep;
    OX_MLOOP
    for (int i = 0; i < nM; i++) at(b.qLD, i) = at(b.qM, i);
    OX_MLOOP
    for (int i = 0; i < nv; i++) at(b.qLD, m.dof_Madr(i)) += dt * m.dof_damping(i);
    if (fast && !dis(OX_DSBL_ACTUATION)) {
      const bool clamp = !dis(OX_DSBL_CLAMPCTRL);
      OX_MLOOP
      for (int i = 0; i < nu; i++) {
        const bool gaff = m.actuator_gaintype(i) == OX_GAIN_AFFINE, baff = m.actuator_biastype(i) == OX_BIAS_AFFINE;
        if (!gaff && !baff) continue;
        if (m.actuator_forcelimited(i)) {
          const T f = at(b.actuator_force, i);
          if (f <= m.actuator_forcerange(2 * i) || f >= m.actuator_forcerange(2 * i + 1)) continue;
        }
        T input = at(b.ctrl, i);
        if (m.actuator_ctrllimited(i) && clamp) input = ox_clip(input, m.actuator_ctrlrange(2 * i), m.actuator_ctrlrange(2 * i + 1));
        if (m.actuator_dyntype(i) != OX_DYN_NONE) input = at(b.act, m.actuator_actadr(i));  // not advanced yet: what the force was computed from
        const T gear = m.actuator_gear(i);
        const T dfdv = (gaff ? m.actuator_gainprm(3 * i + 2) * input : (T)0) + (baff ? m.actuator_biasprm(3 * i + 2) : (T)0);
        at(b.qLD, m.dof_Madr(m.jnt_dofadr(m.actuator_trnid(i)))) -= dt * gear * gear * dfdv;
      }
    }
    factor_ld(b.qLD, b.qLDiagInv);
    OX_MLOOP
    for (int i = 0; i < nv; i++) at(b.i_qacc, i) = at(b.qfrc_smooth, i) + at(b.qfrc_constraint, i);
    solve_ld(b.i_qacc);
    advance(b.i_qacc, nullptr, b.act_dot);
  }
  // classic RK4; the Butcher matrix has one entry per row, so X_i = X_0 (+) h a_i F_{i-1}. Split into pieces so that
  // step() can drive all four forward evaluations through ONE call site (code size: forward() is inlined once).
  OX_HD void rk4_begin() const {
    const auto& h = m.h();
    OX_MLOOP
    for (int i = 0; i < h.nq; i++) at(b.rk_q0, i) = at(b.qpos, i);
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) {
      at(b.rk_v0, i) = at(b.qvel, i);
      at(b.rk_sv, i) = (T)(1.0 / 6) * at(b.qvel, i);
      at(b.rk_sa, i) = (T)(1.0 / 6) * at(b.qacc, i);
    }
    OX_MLOOP
    for (int i = 0; i < h.na; i++) { at(b.rk_a0, i) = at(b.act, i); at(b.rk_sad, i) = (T)(1.0 / 6) * at(b.act_dot, i); }
    at(b.rk_t0, 0) = at(b.time, 0);
  }
  OX_HD void rk4_prepare(int st_) const {  // state for stage st_ = 1..3 from F_{st_-1} = (current qvel, current qacc)
    const auto& h = m.h();
    const T dt = (T)h.timestep;
    const T a = st_ == 3 ? (T)1 : (T)0.5;
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) { at(b.s_gradold, i) = a * at(b.qvel, i); at(b.s_Mgradold, i) = a * at(b.qacc, i); }
    OX_MLOOP
    for (int i = 0; i < h.nq; i++) at(b.qpos, i) = at(b.rk_q0, i);
    integrate_pos(b.qpos, b.s_gradold, dt);
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) at(b.qvel, i) = at(b.rk_v0, i) + dt * at(b.s_Mgradold, i);
    OX_MLOOP
    for (int i = 0; i < h.na; i++) at(b.act, i) = at(b.rk_a0, i) + a * dt * at(b.act_dot, i);   // intermediate stages: plain Euler, no clamp
    at(b.time, 0) = at(b.rk_t0, 0) + a * dt;
  }
  OX_HD void rk4_accumulate(int st_) const {
    const auto& h = m.h();
    const T w = st_ == 3 ? (T)(1.0 / 6) : (T)(1.0 / 3);
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) {
      at(b.rk_sv, i) += w * at(b.qvel, i);
      at(b.rk_sa, i) += w * at(b.qacc, i);
    }
    OX_MLOOP
    for (int i = 0; i < h.na; i++) at(b.rk_sad, i) += w * at(b.act_dot, i);
  }
  OX_HD void rk4_finish() const {
    const auto& h = m.h();
    at(b.time, 0) = at(b.rk_t0, 0);
    OX_MLOOP
    for (int i = 0; i < h.nq; i++) at(b.qpos, i) = at(b.rk_q0, i);
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) at(b.qvel, i) = at(b.rk_v0, i);
    OX_MLOOP
    for (int i = 0; i < h.na; i++) at(b.act, i) = at(b.rk_a0, i);
    advance(b.rk_sa, b.rk_sv, b.rk_sad);
  }
  OX_HDN void rk4() const {  // staged mode: stages 2..4 after the forward the step already ran
    rk4_begin();
#pragma unroll 1
    for (int st_ = 1; st_ < 4; st_++) {
      rk4_prepare(st_);
      forward(true);
      rk4_accumulate(st_);
    }
    rk4_finish();
  }

  // ============================================================ reset / checks / step
  OX_HDN void reset_data() const {
    const auto& h = m.h();
    OX_MLOOP
    for (int i = 0; i < h.nq; i++) at(b.qpos, i) = m.qpos0(i);
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) { at(b.qvel, i) = 0; at(b.qfrc_applied, i) = 0; at(b.qacc_warmstart, i) = 0; at(b.qacc, i) = 0; }
    OX_MLOOP
    for (int i = 0; i < h.nu; i++) at(b.ctrl, i) = 0;
    OX_MLOOP
    for (int i = 0; i < h.na; i++) { at(b.act, i) = 0; at(b.act_dot, i) = 0; }
    if (h.nmocap > 0) {
      OX_MLOOP
      for (int i = 1; i < h.nbody; i++) {
        const int id = m.body_mocapid(i);
        if (id < 0) continue;
        OX_MLOOP
        for (int k = 0; k < 3; k++) at(b.mocap_pos, 3 * id + k) = m.body_pos(3 * i + k);
        OX_MLOOP
        for (int k = 0; k < 4; k++) at(b.mocap_quat, 4 * id + k) = m.body_quat(4 * i + k);
      }
    }
    OX_MLOOP
    for (int i = 0; i < h.neq; i++) at(b.eq_active, i) = (T)m.eq_active0(i);
    OX_MLOOP
    for (int i = 0; i < 6 * h.nbody; i++) at(b.xfrc_applied, i) = 0;
    at(b.time, 0) = 0;
    ati(b.ncon, 0) = 0; ati(b.nefc, 0) = 0; ati(b.solver_niter, 0) = 0; ati(b.ne, 0) = 0;
  }
  OX_HD bool bad_state() const {
    const auto& h = m.h();
    bool bad = false;
    OX_MLOOP
    for (int i = 0; i < h.nq; i++) bad |= ox_bad(at(b.qpos, i));
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) bad |= ox_bad(at(b.qvel, i));
    OX_MLOOP
    for (int i = 0; i < h.na; i++) bad |= ox_bad(at(b.act, i));
    return bad;
  }
  OX_HD bool bad_acc() const {
    const auto& h = m.h();
    bool bad = false;
    OX_MLOOP
    for (int i = 0; i < h.nv; i++) bad |= ox_bad(at(b.qacc, i));
    return bad;
  }
  OX_HD void fill_ctrl_philox(uint64_t seed, int64_t genv, int64_t stepno, T scale = (T)1) const {
    const int nu = m.h().nu;
    OX_MLOOP
    for (int g = 0; g * 4 < nu; g++) {
      uint32_t out[4];
      philox4x32_10((uint32_t)genv, (uint32_t)((uint64_t)genv >> 32), (uint32_t)stepno, (uint32_t)g, (uint32_t)seed,
                    (uint32_t)(seed >> 32), out);
      OX_MLOOP
      for (int k = 0; k < 4 && g * 4 + k < nu; k++)
        at(b.ctrl, g * 4 + k) = scale * ((T)(int32_t)((out[k] >> 9) * 2u + 1u) * (T)(1.0 / 8388608.0) - (T)1);
    }
  }
  OX_HD void accumulate_stats() const {
    ati(b.acc_ncon, 0) += ati(b.ncon, 0);
    ati(b.acc_nefc, 0) += ati(b.nefc, 0);
    ati(b.acc_niter, 0) += ati(b.solver_niter, 0);
  }
  OX_HDN void step() const {
    if (bad_state()) { reset_data(); ati(b.diverged, 0) += 1; }
    const bool rk = m.h().integrator == OX_INT_RK4;
    int stage = 0;
    bool retried = false;
#pragma unroll 1
    for (;;) {  // the only call site of forward(): first evaluation, its retry after a bad-qacc reset, and RK4 stages 2..4
      forward(stage > 0);
      if (stage == 0) {
        if (!retried && bad_acc()) { reset_data(); ati(b.diverged, 0) += 1; retried = true; continue; }
        accumulate_stats();
        if (!rk) break;
        rk4_begin();
      } else {
        rk4_accumulate(stage);
        if (stage == 3) break;
      }
      stage++;
      rk4_prepare(stage);
    }
    if (rk) rk4_finish(); else euler();
  }
};

}  // namespace ox

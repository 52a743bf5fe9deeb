// Device-side view of the compiled model ("constant tables") and of the SoA batch arena.
//
// Model blob  : [BlobHeader][int tables][real tables in the batch precision], 16-byte padded, built once
//               per (model, precision) on the host, uploaded once, and staged into shared memory by every
//               kernel with one TMA bulk copy (cp.async.bulk, see ox_batch.cu: stage_model()).
// Batch arena : one allocation; every per-env field is stored [element][env] with env fastest
//               (stride = nenv rounded up to 32) so that a warp of 32 consecutive envs reads/writes
//               one 128-byte (fp32) or two (fp64) fully coalesced lines per element.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/ox_b200.h"

#if defined(__CUDACC__)
#define OX_HD __host__ __device__ __forceinline__
#define OX_HDN __host__ __device__
#else
#define OX_HD inline
#define OX_HDN
#endif

namespace ox {

struct BlobHeader {
  int32_t nq, nv, nu, na, nbody, njnt, ngeom, nsite, nM, npair, nsensor, nsensordata, nconmax, nefcmax, nvv, nmocap, neq, noslip_iterations, ntendon, nwrap, nfloss, nfluid, ngravcomp, padn_[1];
  int32_t integrator, solver, cone, iterations, ls_iterations, disableflags;
  int32_t total_bytes, any_damping;
  double timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia, noslip_tolerance;
#define OX_X(name, n, w) int32_t off_##name;
  OX_MODEL_INT_TABLES(OX_X)
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  int32_t pad_[2];
  OX_HD double grav(int k) const { return gravity[k]; }
};

template <typename T>
struct DevModel {
  struct Hdr { static constexpr int nv = 1 << 30, nconmax = 1 << 30; };  // runtime model: no compile-time sizes
  const unsigned char* base;
  OX_HD const BlobHeader& h() const { return *reinterpret_cast<const BlobHeader*>(base); }
#define OX_X(name, n, w) \
  OX_HD int32_t name(int i) const { return reinterpret_cast<const int32_t*>(base + h().off_##name)[i]; }
  OX_MODEL_INT_TABLES(OX_X)
#undef OX_X
#define OX_X(name, n, w) \
  OX_HD T name(int i) const { return reinterpret_cast<const T*>(base + h().off_##name)[i]; }
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
};

// Solver tolerance a batch runs with: the caller's if given, else the model's, but never below the resolution of the
// arithmetic type (8 eps): a relative gradient of 1e-8 (MuJoCo's default, meant for doubles) cannot be resolved in fp32
// and only buys extra iterations on round-off. No effect in fp64 (8 eps = 1.8e-15). ORACLE_DECISIONS #8, DESIGN 4.
template <typename T>
inline double effective_tolerance(const ox_model_tables& t, double user) {
  if (user >= 0) return user;
  const double floor_t = 8.0 * (sizeof(T) == 4 ? 1.1920928955078125e-7 : 2.220446049250313e-16);
  return t.tolerance > floor_t ? t.tolerance : floor_t;
}

template <typename T>
inline std::vector<unsigned char> build_blob(const ox_model_tables& t, int iterations, int ls_iterations, double tolerance) {
  BlobHeader h;
  std::memset(&h, 0, sizeof h);
  h.nq = t.nq; h.nv = t.nv; h.nu = t.nu; h.na = t.na; h.nbody = t.nbody; h.njnt = t.njnt; h.ngeom = t.ngeom; h.nsite = t.nsite;
  h.nM = t.nM; h.npair = t.npair; h.nsensor = t.nsensor; h.nsensordata = t.nsensordata; h.nconmax = t.nconmax; h.nefcmax = t.nefcmax; h.nvv = t.nvv; h.nmocap = t.nmocap; h.neq = t.neq; h.ntendon = t.ntendon; h.nwrap = t.nwrap; h.nfloss = t.nfloss; h.nfluid = t.nfluid; h.ngravcomp = t.ngravcomp; h.noslip_iterations = t.noslip_iterations; h.noslip_tolerance = t.noslip_tolerance;
  h.integrator = t.integrator; h.solver = t.solver; h.cone = t.cone;
  h.iterations = iterations > 0 ? iterations : t.iterations;
  h.ls_iterations = ls_iterations > 0 ? ls_iterations : t.ls_iterations;
  h.disableflags = t.disableflags;
  h.timestep = t.timestep;
  for (int k = 0; k < 3; k++) h.gravity[k] = t.gravity[k];
  h.tolerance = effective_tolerance<T>(t, tolerance);
  h.ls_tolerance = t.ls_tolerance; h.impratio = t.impratio; h.meaninertia = t.meaninertia;
  h.any_damping = 0;
  for (int i = 0; i < t.nv; i++) if (t.dof_damping[i] > 0) h.any_damping = 1;
  size_t off = (sizeof(BlobHeader) + 15) / 16 * 16;
#define OX_X(name, n, w)                                   \
  h.off_##name = (int32_t)off;                             \
  off += ((size_t)t.n * (w) * sizeof(int32_t) + 15) / 16 * 16;
  OX_MODEL_INT_TABLES(OX_X)
#undef OX_X
#define OX_X(name, n, w)                                   \
  h.off_##name = (int32_t)off;                             \
  off += ((size_t)t.n * (w) * sizeof(T) + 15) / 16 * 16;
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  h.total_bytes = (int32_t)off;
  std::vector<unsigned char> blob(off, 0);
  std::memcpy(blob.data(), &h, sizeof h);
#define OX_X(name, n, w) \
  if (t.n * (w) > 0) std::memcpy(blob.data() + h.off_##name, t.name, (size_t)t.n * (w) * sizeof(int32_t));
  OX_MODEL_INT_TABLES(OX_X)
#undef OX_X
#define OX_X(name, n, w)                                                  \
  {                                                                       \
    T* dst = reinterpret_cast<T*>(blob.data() + h.off_##name);            \
    for (long i = 0; i < (long)t.n * (w); i++) dst[i] = (T)t.name[i];     \
  }
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  return blob;
}

// ---- batch arena fields: X(name, elements-per-env) in terms of nq nv nu nb nj ng ns nM ncm nem nsd ----
// SMALL: O(nbody + nv + nM) words per env - what the cooperative kernel (ox_coop.cu) keeps in shared memory.
// ROWS : the contact list, the constraint rows and the dense solver scratch, O(ncon + nefc * nv) - always in the arena.
#define OX_BATCH_REAL_FIELDS_SMALL(X)                                                                  \
  /* state */                                                                                          \
  X(qpos, nq) X(qvel, nv) X(ctrl, nu) X(qfrc_applied, nv) X(xfrc_applied, 6 * nb) X(qacc_warmstart, nv) \
  X(time, 1) X(act, na) X(act_dot, na) X(mocap_pos, 3 * nmc) X(mocap_quat, 4 * nmc) X(eq_active, neq) X(ten_length, nten) X(ten_J, nten * nv)  \
  /* position stage */                                                                                 \
  X(xpos, 3 * nb) X(xquat, 4 * nb) X(xmat, 9 * nb) X(xipos, 3 * nb) X(ximat, 9 * nb)                   \
  X(xanchor, 3 * nj) X(xaxis, 3 * nj) X(geom_xpos, 3 * ng) X(geom_xmat, 9 * ng)                        \
  X(site_xpos, 3 * ns) X(site_xmat, 9 * ns) X(subtree_com, 3 * nb) X(cinert, 10 * nb) X(cdof, 6 * nv)  \
  X(crb, 10 * nb) X(qM, nM) X(qLD, nM) X(qLDiagInv, nv)                                                \
  /* velocity / actuation / acceleration */                                                            \
  X(cvel, 6 * nb) X(cdof_dot, 6 * nv) X(cacc, 6 * nb) X(cfrc, 6 * nb) X(qfrc_bias, nv)                 \
  X(qfrc_passive, nv) X(actuator_force, nu) X(qfrc_actuator, nv) X(qfrc_smooth, nv) X(qacc_smooth, nv) \
  /* solver vectors */                                                                                 \
  X(qacc, nv) X(qfrc_constraint, nv) X(s_Ma, nv) X(s_grad, nv) X(s_Mgrad, nv)                          \
  X(s_search, nv) X(s_Mv, nv) X(s_gradold, nv) X(s_Mgradold, nv)                                       \
  /* integrator scratch */                                                                             \
  X(rk_q0, nq) X(rk_v0, nv) X(rk_sv, nv) X(rk_sa, nv) X(rk_t0, 1) X(i_qacc, nv) X(rk_a0, na) X(rk_sad, na) \
  X(sensordata, nsd) X(subtree_linvel, 3 * nb)
#define OX_BATCH_REAL_FIELDS_ROWS(X)                                                                   \
  X(con_dist, ncm) X(con_pos, 3 * ncm) X(con_frame, 9 * ncm)                                           \
  X(efc_J, nem * nv) X(efc_pos, nem) X(efc_margin, nem) X(efc_D, nem) X(efc_aref, nem) X(efc_force, nem) X(efc_floss, nem) \
  X(s_Jaref, nem) X(s_Jv, nem) X(s_H, nv * nv)
#define OX_BATCH_REAL_FIELDS(X) OX_BATCH_REAL_FIELDS_SMALL(X) OX_BATCH_REAL_FIELDS_ROWS(X)

#define OX_BATCH_INT_FIELDS_SCALAR(X) X(ncon, 1) X(nefc, 1) X(solver_niter, 1) X(diverged, 1) X(acc_ncon, 1) X(acc_nefc, 1) X(acc_niter, 1) X(ne, 1) X(nf, 1)
#define OX_BATCH_INT_FIELDS_ROWS(X) X(con_pair, ncm) X(con_active, ncm) X(con_efcadr, ncm)
#define OX_BATCH_INT_FIELDS(X) OX_BATCH_INT_FIELDS_SCALAR(X) OX_BATCH_INT_FIELDS_ROWS(X)

template <typename T>
struct DevBatch {
  int32_t nenv;
  int32_t stride;  // env stride of every field (nenv rounded up to a multiple of 32)
  int32_t lanes;   // active lanes per warp (32 = full warps)
#define OX_X(name, cnt) T* name;
  OX_BATCH_REAL_FIELDS(OX_X)
#undef OX_X
#define OX_X(name, cnt) int32_t* name;
  OX_BATCH_INT_FIELDS(OX_X)
#undef OX_X
};

}  // namespace ox

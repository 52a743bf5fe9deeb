// Model-specialised step: the same stage code as the generic path (ox_stages.cuh), instantiated with
//   * a generated model policy (ox_specgen -> Spec_<name><T>) whose tables are compile-time constants, so every
//     loop over bodies / joints / dofs / geoms / pairs unrolls and the tree walks disappear, and
//   * LOCAL = true storage: every intermediate mjData field is a per-thread array that the unrolled code turns into
//     registers; only the algorithmic state is read from / written to the SoA batch in HBM.
// One thread = one environment, one launch = nsteps steps.
#pragma once
#include <string>

#include "ox_stages.cuh"

namespace ox {

struct SpecRuntime {  // options that stay runtime so ox_batch_config can override them without a rebuild
  int iterations, ls_iterations;
  double tolerance;
};

struct StepArgs {
  int nsteps;
  int philox;
  uint64_t seed;
  int64_t env_id_offset;
  // Philox step index of the first step of this launch: a kernel argument from the batch's host-side counter, so that a
  // step is ONE launch. Only a captured CUDA graph (staged mode, frozen arguments) reads the device counter d_step
  // instead, which the last stage kernel of the captured step advances itself.
  long long step0 = 0;
  long long* d_step = nullptr;
  // 0: no user of this batch has written qfrc_applied / xfrc_applied since the last full reset, so the step neither loads
  // nor tests them (they are 57 of the cheetah's 128 words read per env-step, all zero in an RL loop)
  int applied = 1;
  float ctrl_scale = 1.0f;  // Philox controls are ctrl_scale * U(-1,1); powers of two keep CPU fp64 and GPU fp32 bit-identical
  // ox_batch_step_io: env-major user buffers (device memory or pinned host memory addressed over PCIe) that the step
  // kernel itself reads the controls from and writes the new state to - no separate layout-conversion launches
  const void* io_ctrl = nullptr;
  void* io_qpos = nullptr;
  void* io_qvel = nullptr;
  int io_f64 = 0;  // element type of the user buffers
};

// FNV-1a over everything the generated code depends on
inline uint64_t model_hash(const ox_model_tables& t) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
  };
  const int32_t sizes[] = {t.ngravcomp, t.nfluid, t.nfloss, t.ntendon, t.nwrap, t.nmocap, t.neq, t.nq, t.nv, t.nu, t.na, t.nbody, t.njnt, t.ngeom, t.nsite, t.nM, t.npair, t.nsensor, t.nsensordata,
                           t.nconmax, t.nefcmax, t.integrator, t.solver, t.cone, t.disableflags, t.noslip_iterations};
  mix(sizes, sizeof sizes);
  const double opts[] = {t.timestep, t.gravity[0], t.gravity[1], t.gravity[2], t.ls_tolerance, t.impratio, t.meaninertia, t.noslip_tolerance};
  mix(opts, sizeof opts);
#define OX_X(name, n, w) mix(t.name, (size_t)t.n * (w) * sizeof(int32_t));
  OX_MODEL_INT_TABLES(OX_X)
#undef OX_X
#define OX_X(name, n, w) mix(t.name, (size_t)t.n * (w) * sizeof(double));
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  return h;
}

template <int N> struct AtLeast1 { static constexpr int v = N > 0 ? N : 1; };

// PHASE 0: the whole life of one environment inside one launch: load state, step nsteps times on-chip, store state.
// PHASE 1 / 2: the same step split around the constraint solve, for models whose solve runs in the warp-cooperative
// kernel (ox_solve_coop.cu): PRE = checks + forward up to the smooth acceleration and the constraint rows, written to
// the SoA batch for the solver; POST = bad-qacc check, statistics, integration. Only Euler models are split.
template <class S, typename T, int PHASE>
OX_HDN void spec_step_env(const DevBatch<T>& g, int e, const StepArgs& a, const SpecRuntime& rt, long long step0) {
  using H = typename S::Hdr;
  // per-thread copy of every batch field, sized by the spec's compile-time dimensions. One local array per field (not one
  // struct): each array whose indices all fold to constants after unrolling is promoted to registers independently;
  // only the runtime-indexed ones (contact / constraint rows and what the contact Jacobian walk reads) stay in local memory.
  constexpr int nq = H::nq, nv = H::nv, nu = H::nu, na = H::na, nb = H::nbody, nj = H::njnt, ng = H::ngeom, ns = H::nsite, nM = H::nM,
                ncm = AtLeast1<H::nconmax>::v, nem = AtLeast1<H::nefcmax>::v, nsd = H::nsensordata, nmc = H::nmocap, neq = H::neq, nten = H::ntendon;
#define OX_X(name, cnt) T loc_##name[AtLeast1<(cnt)>::v];
  OX_BATCH_REAL_FIELDS(OX_X)
#undef OX_X
#define OX_X(name, cnt) int32_t loc_##name[AtLeast1<(cnt)>::v];
  OX_BATCH_INT_FIELDS(OX_X)
#undef OX_X
  DevBatch<T> lb;
  lb.nenv = 1;
  lb.stride = 1;
  lb.lanes = 32;
#define OX_X(name, cnt) lb.name = loc_##name;
  OX_BATCH_REAL_FIELDS(OX_X)
  OX_BATCH_INT_FIELDS(OX_X)
#undef OX_X
  S m;
  m.hdr.iterations = rt.iterations;
  m.hdr.ls_iterations = rt.ls_iterations;
  m.hdr.tolerance = rt.tolerance;
  Env<T, S, true> env(m, lb, 0);
  const uint32_t st = (uint32_t)g.stride, ue = (uint32_t)e;
#define G(field, i) g.field[(uint32_t)(i) * st + ue]
  // ---- algorithmic reads: qpos, qvel, ctrl, qacc_warmstart, time (+ applied forces)
#pragma unroll
  for (int i = 0; i < H::nq; i++) loc_qpos[i] = G(qpos, i);
#pragma unroll
  for (int i = 0; i < H::nv; i++) loc_qvel[i] = G(qvel, i);
#pragma unroll
  for (int i = 0; i < H::na; i++) loc_act[i] = G(act, i);
  loc_time[0] = G(time, 0);
  loc_diverged[0] = G(diverged, 0);
  const int32_t div0 = loc_diverged[0];
  (void)div0;
  auto load_inputs = [&]() {
#pragma unroll
    for (int i = 0; i < H::nv; i++) loc_qacc_warmstart[i] = G(qacc_warmstart, i);
#pragma unroll
    for (int i = 0; i < H::nu; i++) loc_ctrl[i] = G(ctrl, i);
#pragma unroll
    for (int i = 0; i < 3 * H::nmocap; i++) loc_mocap_pos[i] = G(mocap_pos, i);
#pragma unroll
    for (int i = 0; i < 4 * H::nmocap; i++) loc_mocap_quat[i] = G(mocap_quat, i);
#pragma unroll
    for (int i = 0; i < H::neq; i++) loc_eq_active[i] = G(eq_active, i);
    if (a.applied) {
#pragma unroll
      for (int i = 0; i < H::nv; i++) loc_qfrc_applied[i] = G(qfrc_applied, i);
#pragma unroll
      for (int i = 0; i < 6 * H::nbody; i++) loc_xfrc_applied[i] = G(xfrc_applied, i);
    } else {
#pragma unroll
      for (int i = 0; i < H::nv; i++) loc_qfrc_applied[i] = 0;
#pragma unroll
      for (int i = 0; i < 6 * H::nbody; i++) loc_xfrc_applied[i] = 0;
    }
  };
  auto store_state = [&](bool inputs_too) {
#pragma unroll
    for (int i = 0; i < H::nq; i++) G(qpos, i) = loc_qpos[i];
#pragma unroll
    for (int i = 0; i < H::nv; i++) G(qvel, i) = loc_qvel[i];
#pragma unroll
    for (int i = 0; i < H::na; i++) { G(act, i) = loc_act[i]; G(act_dot, i) = loc_act_dot[i]; }
    G(time, 0) = loc_time[0];
    G(diverged, 0) = loc_diverged[0];
    if (inputs_too) {  // an auto-reset cleared them
#pragma unroll
      for (int i = 0; i < 3 * H::nmocap; i++) G(mocap_pos, i) = loc_mocap_pos[i];
#pragma unroll
      for (int i = 0; i < 4 * H::nmocap; i++) G(mocap_quat, i) = loc_mocap_quat[i];
#pragma unroll
      for (int i = 0; i < H::neq; i++) G(eq_active, i) = loc_eq_active[i];
#pragma unroll
      for (int i = 0; i < H::nv; i++) { G(qacc_warmstart, i) = loc_qacc_warmstart[i]; G(qfrc_applied, i) = loc_qfrc_applied[i]; }
#pragma unroll
      for (int i = 0; i < 6 * H::nbody; i++) G(xfrc_applied, i) = loc_xfrc_applied[i];
    }
  };
  if constexpr (PHASE == 0) {
    load_inputs();
    loc_acc_ncon[0] = G(acc_ncon, 0); loc_acc_nefc[0] = G(acc_nefc, 0); loc_acc_niter[0] = G(acc_niter, 0);
    loc_ncon[0] = 0; loc_nefc[0] = 0; loc_solver_niter[0] = 0; loc_ne[0] = 0;
#pragma unroll
    for (int i = 0; i < H::nv; i++) loc_qacc[i] = 0;
    for (int s = 0; s < a.nsteps; s++) {
      if (a.philox) env.fill_ctrl_philox(a.seed, a.env_id_offset + e, step0 + s, (T)a.ctrl_scale);
      env.step();
    }
    // ---- algorithmic writes: qpos, qvel, qacc, qacc_warmstart, time (+ sensordata, ctrl actually applied, counters)
    // an auto-reset (mj_checkPos/Vel/Acc -> mj_resetData) inside this launch also cleared ctrl and the applied forces: the
    // batch copy must see that too, as it does with the generic kernels
    const bool did_reset = loc_diverged[0] != div0;
    store_state(did_reset);
#pragma unroll
    for (int i = 0; i < H::nv; i++) { G(qacc_warmstart, i) = loc_qacc_warmstart[i]; G(qacc, i) = loc_qacc[i]; }
    if (a.philox || did_reset) {
#pragma unroll
      for (int i = 0; i < H::nu; i++) G(ctrl, i) = loc_ctrl[i];
    }
#pragma unroll
    for (int i = 0; i < H::nsensordata; i++) G(sensordata, i) = loc_sensordata[i];
    G(ncon, 0) = loc_ncon[0]; G(nefc, 0) = loc_nefc[0]; G(solver_niter, 0) = loc_solver_niter[0];
    G(acc_ncon, 0) = loc_acc_ncon[0]; G(acc_nefc, 0) = loc_acc_nefc[0]; G(acc_niter, 0) = loc_acc_niter[0];
  } else if constexpr (PHASE == 1) {
    load_inputs();
    if (a.philox) env.fill_ctrl_philox(a.seed, a.env_id_offset + e, step0, (T)a.ctrl_scale);
    const bool did_reset = env.bad_state();
    if (did_reset) { env.reset_data(); loc_diverged[0] += 1; }
    env.fwd_position();
    env.fwd_velocity();
    env.make_constraint();
    env.actuation();
    env.fwd_acceleration();
    env.sensors();
    // ---- what the solver and the POST phase need, into the SoA batch
    if (did_reset) store_state(true);
    if (a.philox || did_reset) {
#pragma unroll
      for (int i = 0; i < H::nu; i++) G(ctrl, i) = loc_ctrl[i];
    }
#pragma unroll
    for (int i = 0; i < H::nM; i++) G(qM, i) = loc_qM[i];
#pragma unroll
    for (int i = 0; i < H::na; i++) G(act_dot, i) = loc_act_dot[i];   // POST integrates the activations
#pragma unroll
    for (int i = 0; i < H::nv; i++) { G(qfrc_smooth, i) = loc_qfrc_smooth[i]; G(qacc_smooth, i) = loc_qacc_smooth[i]; }
#pragma unroll
    for (int i = 0; i < H::nsensordata; i++) G(sensordata, i) = loc_sensordata[i];
    const int nefc = loc_nefc[0];
    G(ncon, 0) = loc_ncon[0]; G(nefc, 0) = nefc; G(ne, 0) = loc_ne[0];
    for (int r = 0; r < nefc; r++) {
      G(efc_D, r) = loc_efc_D[r]; G(efc_aref, r) = loc_efc_aref[r]; G(efc_pos, r) = loc_efc_pos[r]; G(efc_margin, r) = loc_efc_margin[r];
#pragma unroll
      for (int i = 0; i < H::nv; i++) G(efc_J, r * H::nv + i) = loc_efc_J[r * H::nv + i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < H::nv; i++) { loc_qacc[i] = G(qacc, i); loc_qfrc_constraint[i] = G(qfrc_constraint, i); loc_qfrc_smooth[i] = G(qfrc_smooth, i); }
#pragma unroll
    for (int i = 0; i < H::nM; i++) loc_qM[i] = G(qM, i);
#pragma unroll
    for (int i = 0; i < H::na; i++) loc_act_dot[i] = G(act_dot, i);
    loc_acc_ncon[0] = G(acc_ncon, 0); loc_acc_nefc[0] = G(acc_nefc, 0); loc_acc_niter[0] = G(acc_niter, 0);
    loc_ncon[0] = G(ncon, 0); loc_nefc[0] = G(nefc, 0); loc_solver_niter[0] = G(solver_niter, 0);
    const bool redo = env.bad_acc();
    if (redo) {  // mj_checkAcc: reset and redo the forward (thread-serial solver; this path is rare)
      load_inputs();
      env.reset_data();
      loc_diverged[0] += 1;
      env.forward(false);
    }
    env.accumulate_stats();
    env.euler();
    store_state(redo);
    if (redo) {
#pragma unroll
      for (int i = 0; i < H::nv; i++) { G(qacc, i) = loc_qacc[i]; G(qacc_warmstart, i) = loc_qacc_warmstart[i]; }
#pragma unroll
      for (int i = 0; i < H::nu; i++) G(ctrl, i) = loc_ctrl[i];
#pragma unroll
      for (int i = 0; i < H::nsensordata; i++) G(sensordata, i) = loc_sensordata[i];
      G(ncon, 0) = loc_ncon[0]; G(nefc, 0) = loc_nefc[0]; G(solver_niter, 0) = loc_solver_niter[0];
    }
    G(acc_ncon, 0) = loc_acc_ncon[0]; G(acc_nefc, 0) = loc_acc_nefc[0]; G(acc_niter, 0) = loc_acc_niter[0];
  }
#undef G
}

#if defined(__CUDACC__)
// Fused I/O of ox_batch_step_io: a full warp owns one contiguous [32][cnt] block of an env-major user buffer and walks it
// linearly (coalesced on the user side, which may be pinned host memory: whole PCIe lines), scattering to / gathering
// from a few rows of the SoA field that the same warp's threads read right after / wrote right before (L1/L2 hits).
template <typename T, typename TU>
__device__ __forceinline__ void spec_io_in(T* field, int cnt, const TU* user, const DevBatch<T>& g, int e) {
  const int lane = threadIdx.x & 31;
  if (g.lanes == 32 && __activemask() == 0xffffffffu) {
    const int e0 = e - lane, n = 32 * cnt;
    for (int base = lane; base < n; base += 32 * 8) {  // 8 loads in flight per lane: a pinned source is a PCIe round trip each
      TU tmp[8];
#pragma unroll
      for (int j = 0; j < 8; j++) { const int idx = base + 32 * j; tmp[j] = idx < n ? user[(size_t)e0 * cnt + idx] : (TU)0; }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int idx = base + 32 * j;
        if (idx < n) { const int el = idx / cnt, k = idx - el * cnt; field[(uint32_t)k * (uint32_t)g.stride + (uint32_t)(e0 + el)] = (T)tmp[j]; }
      }
    }
    __syncwarp();
  } else {
    for (int k = 0; k < cnt; k++) field[(uint32_t)k * (uint32_t)g.stride + (uint32_t)e] = (T)user[(size_t)e * cnt + k];
  }
}
template <typename T, typename TU>
__device__ __forceinline__ void spec_io_out(const T* field, int cnt, TU* user, const DevBatch<T>& g, int e) {
  const int lane = threadIdx.x & 31;
  if (g.lanes == 32 && __activemask() == 0xffffffffu) {
    __syncwarp();
    const int e0 = e - lane;
    for (int idx = lane; idx < 32 * cnt; idx += 32) {
      const int el = idx / cnt, k = idx - el * cnt;
      user[(size_t)e0 * cnt + idx] = (TU)field[(uint32_t)k * (uint32_t)g.stride + (uint32_t)(e0 + el)];
    }
  } else {
    for (int k = 0; k < cnt; k++) user[(size_t)e * cnt + k] = (TU)field[(uint32_t)k * (uint32_t)g.stride + (uint32_t)e];
  }
}

// body shared by the compiled-in kernels (k_step_spec below) and the run-time compiled ones (extern "C" ox_jit_*, ox_specsrc.h)
template <class S, typename T, int PHASE>
__device__ __forceinline__ void spec_kernel_body(const DevBatch<T>& g, const StepArgs& a, const SpecRuntime& rt) {
  using H = typename S::Hdr;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (lane >= g.lanes) return;  // deliberately under-filled warps at small batch sizes (see env_index in ox_kernels.cuh)
  const int e = warp * g.lanes + lane;
  if (e >= g.nenv) return;
  if (PHASE != 2 && a.io_ctrl && H::nu > 0) {
    if (a.io_f64) spec_io_in<T, double>(g.ctrl, H::nu, (const double*)a.io_ctrl, g, e);
    else spec_io_in<T, float>(g.ctrl, H::nu, (const float*)a.io_ctrl, g, e);
  }
  const long long step0 = a.d_step ? *a.d_step : a.step0;
  spec_step_env<S, T, PHASE>(g, e, a, rt, step0);
  if (PHASE != 1) {
    if (a.io_qpos) {
      if (a.io_f64) spec_io_out<T, double>(g.qpos, H::nq, (double*)a.io_qpos, g, e);
      else spec_io_out<T, float>(g.qpos, H::nq, (float*)a.io_qpos, g, e);
    }
    if (a.io_qvel) {
      if (a.io_f64) spec_io_out<T, double>(g.qvel, H::nv, (double*)a.io_qvel, g, e);
      else spec_io_out<T, float>(g.qvel, H::nv, (float*)a.io_qvel, g, e);
    }
  }
}

template <class S, typename T, int PHASE>
__global__ void k_step_spec(DevBatch<T> g, StepArgs a, SpecRuntime rt) { spec_kernel_body<S, T, PHASE>(g, a, rt); }
#endif

// ---- registry of the specialisations: compiled into this library (ox_specgen at build time) or compiled at run time for
// the model of a batch (ox_jit.cpp; jit_* hold cudaKernel_t handles of the extern "C" kernels, PHASE 0 / 1 / 2)
struct SpecEntry {
  uint64_t hash = 0;
  const char* name = nullptr;
  void (*launch_f32)(int grid, int block, void* stream, const DevBatch<float>& g, const StepArgs& a, const SpecRuntime& rt) = nullptr;
  void (*launch_f64)(int grid, int block, void* stream, const DevBatch<double>& g, const StepArgs& a, const SpecRuntime& rt) = nullptr;
  // split pipeline around the warp-cooperative solver (null when not generated for this model): [0] = PRE, [1] = POST
  void (*launch_split_f32[2])(int grid, int block, void* stream, const DevBatch<float>& g, const StepArgs& a, const SpecRuntime& rt) = {nullptr, nullptr};
  void (*launch_split_f64[2])(int grid, int block, void* stream, const DevBatch<double>& g, const StepArgs& a, const SpecRuntime& rt) = {nullptr, nullptr};
  void (*host_f32)(const DevBatch<float>& g, int e, const StepArgs& a, const SpecRuntime& rt, long long step0) = nullptr;   // tests/native only
  void (*host_f64)(const DevBatch<double>& g, int e, const StepArgs& a, const SpecRuntime& rt, long long step0) = nullptr;
  void (*host_split_f32[2])(const DevBatch<float>& g, int e, const StepArgs& a, const SpecRuntime& rt, long long step0) = {nullptr, nullptr};
  void (*host_split_f64[2])(const DevBatch<double>& g, int e, const StepArgs& a, const SpecRuntime& rt, long long step0) = {nullptr, nullptr};
  void* jit_f32[3] = {nullptr, nullptr, nullptr};
  void* jit_f64[3] = {nullptr, nullptr, nullptr};
  bool has_phase(bool f64, int phase) const {
    if (f64) return jit_f64[phase] || (phase == 0 ? launch_f64 != nullptr : launch_split_f64[phase - 1] != nullptr);
    return jit_f32[phase] || (phase == 0 ? launch_f32 != nullptr : launch_split_f32[phase - 1] != nullptr);
  }
};
void register_spec(const SpecEntry& e);
const SpecEntry* find_spec(uint64_t hash);
int spec_count();
const SpecEntry* spec_at(int i);
// run-time specialisation (ox_jit.cpp): a SpecEntry whose jit_f32 / jit_f64 kernels exist for `f64`, compiling the model's
// unit with nvcc if the on-disk cache has none and `allow_compile`; nullptr + *why otherwise. Needs no GPU unless `load`.
const SpecEntry* jit_spec(const ox_model_tables& t, bool f64, bool allow_compile, bool load, std::string* why, std::string* cubin_path);
bool jit_enabled();

}  // namespace ox

// Batch side of the C ABI: SoA arena on one B200, sm_100a kernels for every mj_step stage, launch
// orchestration (fused / staged, optional CUDA graph), bulk and per-env I/O.
// Replaces mj_makeData / mj_step / mj_forward / mj_resetData as called by the reference at
// src/physics.rs:14,22,44-54 for nenv independent copies of one model.
//
// There is no CPU fallback in this file: every compute entry point launches CUDA kernels or fails
// with OX_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ox_internal.h"
#include "ox_model.h"
#include "ox_arena.h"
#include "ox_kernels.cuh"
#include "ox_spec.cuh"
#include "ox_stages.cuh"
#include "ox_batch_internal.cuh"

namespace ox {

enum Stage { ST_CTRL = 0, ST_CHECK, ST_KIN, ST_CRB, ST_COLLIDE, ST_VEL, ST_EFC, ST_ACC, ST_SOLVE, ST_SENSE, ST_INTEGRATE, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"ctrl_rng", "check_pos_vel", "kin_com", "crb_ldl", "collide", "vel_bias",
                                             "make_efc", "act_smooth_acc", "solve", "sensors_check_acc", "integrate"};

template <typename T>
__global__ void k_step_fused(const unsigned char* __restrict__ gblob, int bytes, DevBatch<T> b, StepArgs a) {
  DevModel<T> m{stage_model(gblob, bytes)};
  const int e = env_index(b);
  if (e < 0) return;
  Env<T> env(m, b, e);
  const long long step0 = a.d_step ? *a.d_step : a.step0;
  for (int s = 0; s < a.nsteps; s++) {
    if (a.philox) env.fill_ctrl_philox(a.seed, a.env_id_offset + e, step0 + s, (T)a.ctrl_scale);
    env.step();
  }
}

template <typename T>
__global__ void k_forward_fused(const unsigned char* __restrict__ gblob, int bytes, DevBatch<T> b) {
  DevModel<T> m{stage_model(gblob, bytes)};
  const int e = env_index(b);
  if (e < 0) return;
  Env<T> env(m, b, e);
  env.forward(false);
}

template <typename T, int STAGE>
__global__ void k_stage(const unsigned char* __restrict__ gblob, int bytes, DevBatch<T> b, StepArgs a) {
  DevModel<T> m{stage_model(gblob, bytes)};
  const int e = env_index(b);
  if (e < 0) return;
  Env<T> env(m, b, e);
  if (STAGE == ST_CTRL) env.fill_ctrl_philox(a.seed, a.env_id_offset + e, a.d_step ? *a.d_step : a.step0, (T)a.ctrl_scale);
  if (STAGE == ST_CHECK) {
    if (env.bad_state()) { env.reset_data(); env.ati(b.diverged, 0) += 1; }
  }
  if (STAGE == ST_KIN) { env.kinematics(); env.com_pos(); }
  if (STAGE == ST_CRB) { env.crb(); env.factor_m(); }
  if (STAGE == ST_COLLIDE) env.collision();
  if (STAGE == ST_VEL) { env.com_vel(); env.passive(); env.rne(); }
  if (STAGE == ST_EFC) env.make_constraint();
  if (STAGE == ST_ACC) { env.actuation(); env.fwd_acceleration(); }
  if (STAGE == ST_SOLVE) env.fwd_constraint();
  if (STAGE == ST_SENSE) {
    env.sensors();
    if (env.bad_acc()) { env.reset_data(); env.ati(b.diverged, 0) += 1; env.forward(false); }
    env.accumulate_stats();
  }
  if (STAGE == ST_INTEGRATE) {
    if (m.h().integrator == OX_INT_RK4) env.rk4(); else env.euler();
    // captured CUDA graph only (frozen arguments): the last kernel of the step advances the device-side Philox step counter.
    // Nothing else in this launch reads it, and the next launch is stream-ordered after this one.
    if (a.d_step && a.philox && e == 0) *a.d_step += 1;
  }
}

template <typename T>
__global__ void k_reset(const unsigned char* __restrict__ gblob, int bytes, DevBatch<T> b, const uint8_t* __restrict__ mask) {
  DevModel<T> m{stage_model(gblob, bytes)};
  const int e = env_index(b);
  if (e < 0) return;
  if (mask && !mask[e]) return;
  Env<T> env(m, b, e);
  env.reset_data();
}

// layout / dtype conversion between a user buffer and a native SoA field
// dir 0: user -> field, 1: field -> user
// Threads walk the USER buffer linearly, so that side is always fully coalesced - it may be pinned host memory accessed
// straight over PCIe (zero-copy, see host_mapped()); the SoA side is coalesced for the element-major layout and a gather
// over a few rows of consecutive envs (L2-resident) for the env-major one.
template <typename TF, typename TU>
__global__ void k_pack(TF* __restrict__ field, TU* __restrict__ user, int nenv, int cnt, int stride, int layout, int dir) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nenv * cnt) return;
  int e, i;
  if (layout == OX_LAYOUT_ELEM_MAJOR) { i = (int)(idx / nenv); e = (int)(idx - (long long)i * nenv); }
  else { e = (int)(idx / cnt); i = (int)(idx - (long long)e * cnt); }
  const size_t fi = (size_t)i * stride + e;
  if (dir == 0) field[fi] = (TF)user[idx];
  else user[idx] = (TU)field[fi];
}

// one field of the packed per-env state record (ox_batch_get_state / set_state): user[e * ustride + uoff + i] <-> field[i][e]
template <typename TF, typename TU>
__global__ void k_pack_record(TF* __restrict__ field, TU* __restrict__ user, int nenv, int cnt, int stride, int ustride, int uoff, int dir) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nenv * cnt) return;
  const int e = (int)(idx / cnt), i = (int)(idx - (long long)e * cnt);
  const size_t fi = (size_t)i * stride + e, ui = (size_t)e * ustride + uoff + i;
  if (dir == 0) field[fi] = (TF)user[ui];
  else user[ui] = (TU)field[fi];
}

}  // namespace ox

using namespace ox;

namespace ox {
cudaError_t launch_solve_coop_f32(cudaStream_t s, const unsigned char* blob, int bytes, const DevBatch<float>& b, int resident_ctas, int* ctr);
cudaError_t launch_solve_coop_f64(cudaStream_t s, const unsigned char* blob, int bytes, const DevBatch<double>& b, int resident_ctas, int* ctr);
int solve_coop_resident_ctas(int blob_bytes, bool f64);
cudaError_t solve_coop_prepare(int blob_bytes, bool f64);  // per-device opt-in to > 48 KB dynamic shared memory
// cooperative whole-step kernel (ox_coop.cu)
bool step_coop_eligible(const ox_model_tables& t);
size_t step_coop_smem(const ox_model_tables& t, int blob_bytes, bool f64);
cudaError_t step_coop_prepare(const ox_model_tables& t, int blob_bytes, bool f64);
cudaError_t launch_step_coop_f32(cudaStream_t s, const ox_model_tables& t, const unsigned char* blob, int bytes, const DevBatch<float>& g, const StepArgs& a);
cudaError_t launch_step_coop_f64(cudaStream_t s, const ox_model_tables& t, const unsigned char* blob, int bytes, const DevBatch<double>& g, const StepArgs& a);
bool solve_coop_eligible(const ox_model_tables& t);
size_t solve_coop_smem(int blob_bytes, bool f64);
}  // namespace ox

namespace {

template <typename T> DevBatch<T>& dev(ox_batch* b);
template <> DevBatch<float>& dev<float>(ox_batch* b) { return b->bf; }
template <> DevBatch<double>& dev<double>(ox_batch* b) { return b->bd; }

template <typename T>
cudaError_t launch_solve_coop(ox_batch* b) {
  b->launches++;
  if (sizeof(T) == 8) return launch_solve_coop_f64(b->stream, b->d_blob, b->blob_bytes, b->bd, b->coop_resident, b->d_coop_ctr);
  return launch_solve_coop_f32(b->stream, b->d_blob, b->blob_bytes, b->bf, b->coop_resident, b->d_coop_ctr);
}

// launch one phase (0 = whole step, 1 = PRE, 2 = POST) of the batch's model-specialised kernel: a compiled-in launcher, or a
// kernel compiled at run time (cudaKernel_t from ox_jit.cpp)
template <typename T>
cudaError_t launch_spec(ox_batch* b, int phase, const StepArgs& a) {
  const SpecEntry* sp = b->spec;
  b->launches++;
  void* k = sizeof(T) == 8 ? sp->jit_f64[phase] : sp->jit_f32[phase];
  if (k) {
    void* args[3] = {(void*)&dev<T>(b), (void*)&a, (void*)&b->spec_rt};
    return cudaLaunchKernel((const void*)k, dim3(b->grid), dim3(b->block), args, 0, b->stream);
  }
  if (sizeof(T) == 8) (phase == 0 ? sp->launch_f64 : sp->launch_split_f64[phase - 1])(b->grid, b->block, b->stream, b->bd, a, b->spec_rt);
  else (phase == 0 ? sp->launch_f32 : sp->launch_split_f32[phase - 1])(b->grid, b->block, b->stream, b->bf, a, b->spec_rt);
  return cudaPeekAtLastError();
}

StepArgs make_args(ox_batch* b, int nsteps) {
  StepArgs a;
  a.nsteps = nsteps; a.philox = b->philox; a.seed = b->seed; a.env_id_offset = b->cfg.env_id_offset;
  a.step0 = b->h_step; a.d_step = nullptr; a.applied = b->applied; a.ctrl_scale = b->ctrl_scale;
  a.io_ctrl = b->io_ctrl; a.io_qpos = b->io_qpos; a.io_qvel = b->io_qvel; a.io_f64 = b->io_f64;
  return a;
}

template <typename T, int STAGE>
void launch_stage(ox_batch* b, const StepArgs& a) {
  k_stage<T, STAGE><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, dev<T>(b), a);
  b->launches++;
}

template <typename T>
ox_status launch_staged_step(ox_batch* b, bool capturing = false) {
  StepArgs a = make_args(b, 1);
  if (capturing) a.d_step = b->d_step;
  if (b->philox) launch_stage<T, ST_CTRL>(b, a);
  launch_stage<T, ST_CHECK>(b, a);
  launch_stage<T, ST_KIN>(b, a);
  launch_stage<T, ST_CRB>(b, a);
  launch_stage<T, ST_COLLIDE>(b, a);
  launch_stage<T, ST_VEL>(b, a);
  launch_stage<T, ST_EFC>(b, a);
  launch_stage<T, ST_ACC>(b, a);
  if (b->coop) {
    CU_TRY(launch_solve_coop<T>(b));
  } else {
    launch_stage<T, ST_SOLVE>(b, a);
  }
  launch_stage<T, ST_SENSE>(b, a);
  launch_stage<T, ST_INTEGRATE>(b, a);
  if (b->philox) b->h_step++;
  return OX_OK;
}

template <typename T>
ox_status do_step(ox_batch* b, int nsteps) {
  if (b->cfg.mode == OX_MODE_COOP) {
    const StepArgs a = make_args(b, nsteps);
    b->launches++;
    b->derived_stale = true;   // intermediates stay in shared memory; rows / contacts use an env-contiguous layout in this mode
    if (sizeof(T) == 8) CU_TRY(launch_step_coop_f64(b->stream, b->model->t, b->d_blob, b->blob_bytes, b->bd, a));
    else CU_TRY(launch_step_coop_f32(b->stream, b->model->t, b->d_blob, b->blob_bytes, b->bf, a));
    if (b->philox) b->h_step += nsteps;
    return OX_OK;
  }
  if (b->cfg.mode == OX_MODE_FUSED) {
    if (b->spec && b->split) {
      // specialised PRE -> warp-cooperative Newton solve -> specialised POST, one step at a time
      for (int s = 0; s < nsteps; s++) {
        const StepArgs a1 = make_args(b, 1);
        CU_TRY(launch_spec<T>(b, 1, a1));
        CU_TRY(launch_solve_coop<T>(b));
        CU_TRY(launch_spec<T>(b, 2, a1));
        if (b->philox) b->h_step++;
      }
      b->derived_stale = true;
      CU_TRY(cudaGetLastError());
      return OX_OK;
    } else if (b->spec) {
      b->derived_stale = true;
      CU_TRY(launch_spec<T>(b, 0, make_args(b, nsteps)));
    } else {
      k_step_fused<T><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, dev<T>(b), make_args(b, nsteps));
      b->launches++;
    }
    if (b->philox) b->h_step += nsteps;
  } else if (b->cfg.use_graph) {
    if (!b->graph_exec) {
      // the captured kernels read the Philox step index from the device counter (their arguments are frozen): bring it
      // up to date with the host counter first
      const long long saved = b->launches, saved_step = b->h_step;
      CU_TRY(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
      ox_status cs = launch_staged_step<T>(b, true);
      cudaError_t ce = cudaStreamEndCapture(b->stream, &b->graph);
      b->launches = saved; b->h_step = saved_step;
      if (cs) return cs;
      CU_TRY(ce);
      CU_TRY(cudaGraphInstantiate(&b->graph_exec, b->graph, 0));
    }
    const int per = ST_COUNT - 1 + (b->philox ? 1 : 0);
    if (b->d_step_val != b->h_step) {  // set_step_counter / stage_times moved the host counter
      CU_TRY(cudaMemcpyAsync(b->d_step, &b->h_step, sizeof(long long), cudaMemcpyHostToDevice, b->stream));
      b->d_step_val = b->h_step;
    }
    for (int s = 0; s < nsteps; s++) {
      CU_TRY(cudaGraphLaunch(b->graph_exec, b->stream));
      b->launches += per;
      if (b->philox) { b->h_step++; b->d_step_val++; }
    }
  } else {
    for (int s = 0; s < nsteps; s++) { ox_status st = launch_staged_step<T>(b); if (st) return st; }
  }
  CU_TRY(cudaGetLastError());
  return OX_OK;
}

void drop_graph(ox_batch* b) {
  if (b->graph_exec) { cudaGraphExecDestroy(b->graph_exec); b->graph_exec = nullptr; }
  if (b->graph) { cudaGraphDestroy(b->graph); b->graph = nullptr; }
}

ox_status ensure_tmp(ox_batch* b, size_t bytes) {
  if (bytes <= b->d_tmp_bytes) return OX_OK;
  if (b->d_tmp) { CU_TRY(cudaStreamSynchronize(b->stream)); CU_TRY(cudaFree(b->d_tmp)); b->d_tmp = nullptr; }
  CU_TRY(cudaMalloc(&b->d_tmp, bytes));
  b->d_tmp_bytes = bytes;
  return OX_OK;
}
ox_status ensure_htmp(ox_batch* b, size_t bytes) {
  if (bytes <= b->h_tmp_bytes) return OX_OK;
  if (b->h_tmp) { CU_TRY(cudaStreamSynchronize(b->stream)); CU_TRY(cudaFreeHost(b->h_tmp)); b->h_tmp = nullptr; }
  CU_TRY(cudaMallocHost(&b->h_tmp, bytes));
  b->h_tmp_bytes = bytes;
  return OX_OK;
}

// fields of features the model does not have answer OX_ABSENT (the reference's Option::None, src/physics.rs:96-102,154-170)
static ox_status field_absent(const ox_batch* b, int field) {
  const ox_model_tables& t = b->model->t;
  if ((field == OX_F_ACT || field == OX_F_ACT_DOT) && t.na == 0) { ox::set_error("model has no stateful actuators (na = 0)"); return OX_ABSENT; }
  if ((field == OX_F_MOCAP_POS || field == OX_F_MOCAP_QUAT) && t.nmocap == 0) { ox::set_error("model has no mocap bodies (nmocap = 0)"); return OX_ABSENT; }
  if (field == OX_F_EQ_ACTIVE && t.neq == 0) { ox::set_error("model has no equality constraints (neq = 0)"); return OX_ABSENT; }
  return OX_OK;
}

template <typename TF, typename TU>
void launch_pack(ox_batch* b, void* field, void* user, int cnt, int layout, int dir) {
  const long long n = (long long)b->nenv * cnt;
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads);
  k_pack<TF, TU><<<blocks, threads, 0, b->stream>>>((TF*)field, (TU*)user, b->nenv, cnt, b->stride, layout, dir);
  b->launches++;
}

ox_status bulk_io(ox_batch* b, int field, void* buf, int dtype, int mem, int layout, int dir, size_t tmp_off = 0, bool sync = true) {
  if (!b || !buf) { ox::set_error("bulk I/O: null argument"); return OX_ERR_INVALID; }
  auto it = b->fields.find(field);
  if (it == b->fields.end()) { ox::set_error("bulk I/O: unknown field id " + std::to_string(field)); return OX_ERR_INVALID; }
  const FieldInfo& fi = it->second;
  if (ox_status absent = field_absent(b, field)) return absent;
  if (fi.count == 0) return OX_OK;
  if (dir == 1 && !field_live(b, field)) { ox::set_error("ox_batch_get: " + OX_STALE_MSG(field)); return OX_ERR_INVALID; }
  if (layout != OX_LAYOUT_ENV_MAJOR && layout != OX_LAYOUT_ELEM_MAJOR) { ox::set_error("bulk I/O: bad layout"); return OX_ERR_INVALID; }
  if (!fi.is_int && dtype != OX_F32 && dtype != OX_F64) { ox::set_error("bulk I/O: bad dtype"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  const size_t usz = fi.is_int ? 4 : (dtype == OX_F64 ? 8 : 4);
  const size_t bytes = (size_t)b->nenv * fi.count * usz;
  void* dbuf = buf;
  bool staged = false;
  if (mem == OX_MEM_HOST) {
    if (void* alias = host_mapped(buf)) dbuf = alias;
    else {
      ox_status s = ensure_tmp(b, tmp_off + bytes);
      if (s) return s;
      dbuf = (unsigned char*)b->d_tmp + tmp_off;
      staged = true;
      if (dir == 0) CU_TRY(cudaMemcpyAsync(dbuf, buf, bytes, cudaMemcpyHostToDevice, b->stream));
    }
  } else if (mem != OX_MEM_DEVICE) { ox::set_error("bulk I/O: bad mem"); return OX_ERR_INVALID; }
  if (fi.is_int) launch_pack<int32_t, int32_t>(b, fi.ptr, dbuf, fi.count, layout, dir);
  else if (b->f64) {
    if (dtype == OX_F64) launch_pack<double, double>(b, fi.ptr, dbuf, fi.count, layout, dir);
    else launch_pack<double, float>(b, fi.ptr, dbuf, fi.count, layout, dir);
  } else {
    if (dtype == OX_F64) launch_pack<float, double>(b, fi.ptr, dbuf, fi.count, layout, dir);
    else launch_pack<float, float>(b, fi.ptr, dbuf, fi.count, layout, dir);
  }
  CU_TRY(cudaGetLastError());
  if (dir == 0 && (field == OX_F_QFRC_APPLIED || field == OX_F_XFRC_APPLIED)) b->applied = 1;
  if (mem == OX_MEM_HOST && dir == 1) {
    if (staged) CU_TRY(cudaMemcpyAsync(buf, dbuf, bytes, cudaMemcpyDeviceToHost, b->stream));
    if (sync) CU_TRY(cudaStreamSynchronize(b->stream));
  }
  return OX_OK;
}

ox_status slice_io(ox_batch* b, int field, int env, int offset, int count, double* dout, const double* din, int32_t* iout) {
  if (!b) { ox::set_error("per-env I/O: null batch"); return OX_ERR_INVALID; }
  auto it = b->fields.find(field);
  if (it == b->fields.end()) { ox::set_error("per-env I/O: unknown field id " + std::to_string(field)); return OX_ERR_INVALID; }
  const FieldInfo& fi = it->second;
  if (ox_status absent = field_absent(b, field)) return absent;
  if (env < 0 || env >= b->nenv || offset < 0 || count < 0 || offset + count > fi.count) {
    ox::set_error("per-env I/O: index out of range");
    return OX_ERR_INVALID;
  }
  if ((iout != nullptr) != fi.is_int) { ox::set_error("per-env I/O: int/real field mismatch"); return OX_ERR_INVALID; }
  if (count == 0) return OX_OK;
  if (!din && !field_live(b, field)) { ox::set_error("ox_batch_get1: " + OX_STALE_MSG(field)); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  const size_t esz = fi.is_int ? 4 : (b->f64 ? 8 : 4);
  ox_status s = ensure_htmp(b, (size_t)count * 8);
  if (s) return s;
  unsigned char* base = (unsigned char*)fi.ptr + ((size_t)offset * b->stride + env) * esz;
  const size_t pitch = (size_t)b->stride * esz;
  if (din) {
    if (field == OX_F_QFRC_APPLIED || field == OX_F_XFRC_APPLIED) b->applied = 1;
    if (b->f64) for (int i = 0; i < count; i++) ((double*)b->h_tmp)[i] = din[i];
    else for (int i = 0; i < count; i++) ((float*)b->h_tmp)[i] = (float)din[i];
    CU_TRY(cudaMemcpy2DAsync(base, pitch, b->h_tmp, esz, esz, count, cudaMemcpyHostToDevice, b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
  } else {
    CU_TRY(cudaMemcpy2DAsync(b->h_tmp, esz, base, pitch, esz, count, cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
    if (fi.is_int) for (int i = 0; i < count; i++) iout[i] = ((int32_t*)b->h_tmp)[i];
    else if (b->f64) for (int i = 0; i < count; i++) dout[i] = ((double*)b->h_tmp)[i];
    else for (int i = 0; i < count; i++) dout[i] = ((float*)b->h_tmp)[i];
  }
  return OX_OK;
}

}  // namespace

template <typename T>
static ox_status stage_times_impl(ox_batch* b, int reps, double* out_ms) {
  // Steps the batch `reps` times in staged order (it DOES advance the simulation, exactly like ox_batch_step(reps) in
  // staged mode: user controls are kept unless the Philox source is on), with an event after every stage.
  struct Events {
    std::vector<cudaEvent_t> ev;
    ~Events() { for (auto e : ev) if (e) cudaEventDestroy(e); }
  } E;
  E.ev.assign(ST_COUNT + 1, nullptr);
  auto& ev = E.ev;
  for (auto& e : ev) CU_TRY(cudaEventCreate(&e));
  for (int s = 0; s < ST_COUNT; s++) out_ms[s] = 0;
  for (int r = 0; r < reps; r++) {
    StepArgs a = make_args(b, 1);
    CU_TRY(cudaEventRecord(ev[0], b->stream));
#define RUN(S)                                          \
  launch_stage<T, S>(b, a);                             \
  CU_TRY(cudaEventRecord(ev[S + 1], b->stream));
    if (b->philox) launch_stage<T, ST_CTRL>(b, a);
    CU_TRY(cudaEventRecord(ev[ST_CTRL + 1], b->stream));
    RUN(ST_CHECK) RUN(ST_KIN) RUN(ST_CRB) RUN(ST_COLLIDE) RUN(ST_VEL) RUN(ST_EFC) RUN(ST_ACC)
    if (b->coop) {
      CU_TRY(launch_solve_coop<T>(b));
      CU_TRY(cudaEventRecord(ev[ST_SOLVE + 1], b->stream));
    } else {
      RUN(ST_SOLVE)
    }
    RUN(ST_SENSE) RUN(ST_INTEGRATE)
#undef RUN
    if (b->philox) b->h_step++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(b->stream));
    for (int s = 0; s < ST_COUNT; s++) {
      float ms = 0;
      CU_TRY(cudaEventElapsedTime(&ms, ev[s], ev[s + 1]));
      out_ms[s] += ms;
    }
  }
  for (int s = 0; s < ST_COUNT; s++) out_ms[s] /= std::max(1, reps);
  return OX_OK;
}

extern "C" {

void ox_batch_config_default(ox_batch_config* cfg) {
  if (!cfg) return;
  cfg->nenv = 1; cfg->device = 0; cfg->precision = OX_F32; cfg->mode = OX_MODE_FUSED; cfg->iterations = 0; cfg->ls_iterations = 0;
  cfg->use_graph = 0; cfg->block_threads = 0; cfg->env_id_offset = 0; cfg->tolerance = -1; cfg->specialize = 1; cfg->lanes_per_warp = 0; cfg->coop_solver = -1; cfg->reserved_ = 0;
}

ox_status ox_batch_create(const ox_model* m, const ox_batch_config* cfg, ox_batch** out) {
  if (!m || !cfg || !out) { ox::set_error("ox_batch_create: null argument"); return OX_ERR_INVALID; }
  *out = nullptr;
  if (cfg->nenv < 1) { ox::set_error("ox_batch_create: nenv must be >= 1"); return OX_ERR_INVALID; }
  if (cfg->precision != OX_F32 && cfg->precision != OX_F64) { ox::set_error("ox_batch_create: bad precision"); return OX_ERR_INVALID; }
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    ox::set_error(std::string("no CUDA device available (") + cudaGetErrorString(err) +
                  "); libox_b200 has no CPU fallback");
    cudaGetLastError();
    return OX_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { ox::set_error("ox_batch_create: bad device ordinal"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(cfg->device));
  std::unique_ptr<ox_batch, void (*)(ox_batch*)> b(new ox_batch(), ox_batch_free);
  b->model = m;
  b->cfg = *cfg;
  b->nenv = cfg->nenv;
  b->stride = (cfg->nenv + 31) / 32 * 32;
  b->f64 = cfg->precision == OX_F64;
  int block = cfg->block_threads;
  if (block <= 0) {
    const int warps = (b->nenv + 31) / 32;
    block = warps <= 8 * 148 ? 32 : (warps <= 16 * 148 ? 64 : 128);
  }
  if (block % 32 != 0 || block > 1024) { ox::set_error("ox_batch_create: block_threads must be a multiple of 32, <= 1024"); return OX_ERR_INVALID; }
  b->block = block;
  int lanes = cfg->lanes_per_warp;
  if (lanes <= 0) lanes = 32;  // measured on B200 (profiles/r1_notes.md): thinner warps lose to L1 pressure from per-thread storage
  if (lanes > 32) { ox::set_error("ox_batch_create: lanes_per_warp must be in 1..32"); return OX_ERR_INVALID; }
  b->lanes = lanes;
  const int nwarps = (b->nenv + lanes - 1) / lanes, wpb = block / 32;
  b->grid = (nwarps + wpb - 1) / wpb;
  CU_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  const ox_model_tables& t = m->t;
  std::vector<unsigned char> blob = b->f64 ? build_blob<double>(t, cfg->iterations, cfg->ls_iterations, cfg->tolerance)
                                           : build_blob<float>(t, cfg->iterations, cfg->ls_iterations, cfg->tolerance);
  b->blob_bytes = (int)blob.size();
  if (b->blob_bytes > 200 * 1024) { ox::set_error("model constant tables exceed shared memory (200 KB)"); return OX_ERR_INVALID; }
  CU_TRY(cudaMalloc(&b->d_blob, blob.size()));
  CU_TRY(cudaMemcpy(b->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  {  // device code indexes every field with 32-bit arithmetic: element * stride + env must fit
    const long long nem = std::max(1, t.nefcmax), big = std::max<long long>({nem * (long long)t.nv, (long long)t.nv * t.nv,
                                                                             10LL * t.nbody, 9LL * std::max(1, t.nconmax), 9LL * t.ngeom});
    if ((big + 1) * (long long)b->stride >= (1LL << 32)) {
      ox::set_error("ox_batch_create: nenv too large for 32-bit field indexing (largest field x nenv >= 2^32); split the batch");
      return OX_ERR_INVALID;
    }
  }
  b->arena_bytes = b->f64 ? layout_arena<double>(t, b->stride, nullptr, nullptr, nullptr)
                          : layout_arena<float>(t, b->stride, nullptr, nullptr, nullptr);
  CU_TRY(cudaMalloc(&b->arena, b->arena_bytes));
  CU_TRY(cudaMemset(b->arena, 0, b->arena_bytes));
  if (b->f64) { layout_arena<double>(t, b->stride, b->arena, &b->bd, &b->fields); b->bd.nenv = b->nenv; b->bd.stride = b->stride; b->bd.lanes = b->lanes; }
  else { layout_arena<float>(t, b->stride, b->arena, &b->bf, &b->fields); b->bf.nenv = b->nenv; b->bf.stride = b->stride; b->bf.lanes = b->lanes; }
  // the warp-cooperative solver also needs its shared-memory footprint (model tables + 8 warps of row scratch) to fit
  const bool coop_ok = ox::solve_coop_eligible(t) && ox::solve_coop_smem(b->blob_bytes, b->f64) <= 200 * 1024;
  if (cfg->mode == OX_MODE_STAGED && coop_ok)
    b->coop = cfg->coop_solver > 0 || (cfg->coop_solver < 0 && t.nv > 12);
  if (cfg->coop_solver > 0 && !b->coop) {
    ox::set_error("ox_batch_create: coop_solver needs mode=staged, the Newton solver, nv <= 32 and model tables + solver scratch within 200 KB of shared memory");
    return OX_ERR_INVALID;
  }
  if (cfg->mode == OX_MODE_COOP) {
    if (!ox::step_coop_eligible(t)) { ox::set_error("ox_batch_create: mode = coop needs nv <= 32, the Newton solver with the pyramidal cone, joint transmissions only, no noslip or frictionloss, and the Euler or implicitfast integrator"); return OX_ERR_INVALID; }
    if (ox::step_coop_smem(t, b->blob_bytes, b->f64) > 227 * 1024) { ox::set_error("ox_batch_create: mode = coop: model tables + per-env intermediates exceed shared memory"); return OX_ERR_INVALID; }
    CU_TRY(ox::step_coop_prepare(t, b->blob_bytes, b->f64));
  }
  if (cfg->mode == OX_MODE_FUSED && cfg->specialize != 0) {
    b->spec = ox::find_spec(ox::model_hash(t));
    if (b->spec && !b->spec->has_phase(b->f64, 0)) b->spec = nullptr;
    if (!b->spec) {
      // not one of the models compiled into the library: specialise at run time (ox_jit.cpp). specialize = 1 does so for
      // batches where throughput is the point (>= 1024 envs; a cached cubin is a file read, a miss is one nvcc run) and
      // silently keeps the generic kernel otherwise or when the JIT is unavailable; specialize = 2 insists.
      const bool want = cfg->specialize >= 2 || b->nenv >= 1024;
      if (want) {
        std::string why;
        b->spec = ox::jit_spec(t, b->f64, true, true, &why, nullptr);
        if (!b->spec) {
          if (cfg->specialize >= 2) { ox::set_error("ox_batch_create: specialize = 2 but no specialised kernel could be built: " + why); return OX_ERR_CUDA; }
          b->jit_note = why;
        }
      }
    }
    b->spec_rt.iterations = cfg->iterations > 0 ? cfg->iterations : t.iterations;
    b->spec_rt.ls_iterations = cfg->ls_iterations > 0 ? cfg->ls_iterations : t.ls_iterations;
    b->spec_rt.tolerance = b->f64 ? effective_tolerance<double>(t, cfg->tolerance) : effective_tolerance<float>(t, cfg->tolerance);
    bool acc_sensor = false;  // acceleration-stage sensors read the solved qacc: the PRE phase of the split pipeline runs too early for them
    for (int i = 0; i < t.nsensor; i++) acc_sensor |= t.sensor_type[i] == OX_SENS_ACCELEROMETER || t.sensor_type[i] == OX_SENS_TOUCH || t.sensor_type[i] == OX_SENS_FORCE || t.sensor_type[i] == OX_SENS_TORQUE || t.sensor_type[i] == OX_SENS_FRAMELINACC || t.sensor_type[i] == OX_SENS_FRAMEANGACC;
    b->split = b->spec && b->spec->has_phase(b->f64, 1) && b->spec->has_phase(b->f64, 2) && cfg->coop_solver != 0 && coop_ok && t.integrator == OX_INT_EULER && !acc_sensor;
  }
  if (b->coop || b->split) {
    CU_TRY(ox::solve_coop_prepare(b->blob_bytes, b->f64));
    b->coop_resident = ox::solve_coop_resident_ctas(b->blob_bytes, b->f64);   // persistent grid of the ticket-drawing solver
    CU_TRY(cudaMalloc(&b->d_coop_ctr, 2 * sizeof(int)));
    CU_TRY(cudaMemset(b->d_coop_ctr, 0, 2 * sizeof(int)));
  }
  CU_TRY(cudaMalloc(&b->d_step, sizeof(long long)));
  CU_TRY(cudaMemset(b->d_step, 0, sizeof(long long)));
  CU_TRY(cudaMalloc(&b->d_mask, b->stride));
  if (b->blob_bytes > 48 * 1024) {
#define SETATTR(K) CU_TRY(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, b->blob_bytes));
    SETATTR(k_step_fused<float>) SETATTR(k_step_fused<double>) SETATTR(k_forward_fused<float>) SETATTR(k_forward_fused<double>)
    SETATTR(k_reset<float>) SETATTR(k_reset<double>)
#define SETST(S) SETATTR((k_stage<float, S>)) SETATTR((k_stage<double, S>))
    SETST(ST_CTRL) SETST(ST_CHECK) SETST(ST_KIN) SETST(ST_CRB) SETST(ST_COLLIDE) SETST(ST_VEL) SETST(ST_EFC) SETST(ST_ACC)
    SETST(ST_SOLVE) SETST(ST_SENSE) SETST(ST_INTEGRATE)
#undef SETST
#undef SETATTR
  }
  CU_TRY(cudaDeviceSynchronize());  // the uploads / memsets above ran on the legacy stream; the batch's stream is non-blocking
  ox_batch* raw = b.release();
  ox_status s = ox_batch_reset(raw, nullptr);
  if (s == OX_OK) s = ox_batch_sync(raw);
  if (s != OX_OK) { ox_batch_free(raw); return s; }
  raw->launches = 0;
  *out = raw;
  return OX_OK;
}

void ox_batch_free(ox_batch* b) {
  if (!b) return;
  cudaSetDevice(b->cfg.device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  drop_graph(b);
  cudaFree(b->arena); cudaFree(b->d_blob); cudaFree(b->d_step); cudaFree(b->d_tmp); cudaFree(b->d_mask); cudaFree(b->d_coop_ctr);
  if (b->h_tmp) cudaFreeHost(b->h_tmp);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int32_t ox_batch_nenv(const ox_batch* b) { return b ? b->nenv : -1; }
void* ox_batch_stream(const ox_batch* b) { return b ? (void*)b->stream : nullptr; }
int64_t ox_batch_launch_count(const ox_batch* b) { return b ? b->launches : -1; }
const char* ox_batch_kernel_name(const ox_batch* b) {
  if (!b) return nullptr;
  if (b->cfg.mode == OX_MODE_COOP) return "k_step_coop (lane group per env, generic)";
  if (b->cfg.mode != OX_MODE_FUSED) return b->coop ? "k_stage + k_solve_coop (warp-per-env Newton)" : "k_stage (generic, one kernel per stage)";
  if (b->spec && b->split) { static thread_local std::string nm; nm = std::string(b->spec->name) + " (split: spec PRE + k_solve_coop + spec POST)"; return nm.c_str(); }
  return b->spec ? b->spec->name : "k_step_fused (generic)";
}
/* why a batch that asked for a specialised kernel runs the generic one ("" when it does not) */
const char* ox_batch_jit_note(const ox_batch* b) { return b ? b->jit_note.c_str() : nullptr; }

/* Compile (or find in the on-disk cache) the run-time specialised step kernel of a model WITHOUT creating a batch: needs
 * nvcc but no GPU, so images can be warmed at build / deploy time. The cubin's path is copied into path_out. */
ox_status ox_jit_compile(const ox_model* m, int32_t precision, char* path_out, int32_t path_cap) {
  if (!m || (precision != OX_F32 && precision != OX_F64)) { ox::set_error("ox_jit_compile: bad argument"); return OX_ERR_INVALID; }
  std::string why, path;
  if (!ox::jit_spec(m->t, precision == OX_F64, true, false, &why, &path)) { ox::set_error("ox_jit_compile: " + why); return OX_ERR_COMPILE; }
  if (path_out && path_cap > 0) { snprintf(path_out, (size_t)path_cap, "%s", path.c_str()); }
  return OX_OK;
}
int32_t ox_spec_count(void) { return ox::spec_count(); }
const char* ox_spec_name(int32_t i) { const ox::SpecEntry* e = ox::spec_at(i); return e ? e->name : nullptr; }

ox_status ox_batch_step(ox_batch* b, int32_t nsteps) {
  if (!b || nsteps < 0) { ox::set_error("ox_batch_step: bad argument"); return OX_ERR_INVALID; }
  if (nsteps == 0) return OX_OK;
  CU_TRY(cudaSetDevice(b->cfg.device));
  return b->f64 ? do_step<double>(b, nsteps) : do_step<float>(b, nsteps);
}

// Action::apply + Physics::step + Observation::generate of the reference's Environment::step (src/lib.rs:63-66) as ONE call
// for the whole batch: ctrl[nenv][nu] in, one mj_step, qpos[nenv][nq] / qvel[nenv][nv] out (env-major, `dtype`, in `mem`).
// With a model-specialised kernel and device-addressable buffers (device memory, or pinned host memory) the step kernel
// does the I/O itself; otherwise the call is ox_batch_set + ox_batch_step + ox_batch_get_many.
ox_status ox_batch_step_io(ox_batch* b, const void* ctrl, void* qpos, void* qvel, int32_t dtype, int32_t mem) {
  if (!b) { ox::set_error("ox_batch_step_io: null batch"); return OX_ERR_INVALID; }
  if ((dtype != OX_F32 && dtype != OX_F64) || (mem != OX_MEM_HOST && mem != OX_MEM_DEVICE)) { ox::set_error("ox_batch_step_io: bad dtype / mem"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  const void* dctrl = ctrl; void* dqpos = qpos; void* dqvel = qvel;
  bool fused = b->spec != nullptr && b->cfg.mode == OX_MODE_FUSED;
  if (fused && mem == OX_MEM_HOST) {
    dctrl = ctrl ? ox::host_mapped(ctrl) : nullptr; dqpos = qpos ? ox::host_mapped(qpos) : nullptr; dqvel = qvel ? ox::host_mapped(qvel) : nullptr;
    fused = (!ctrl || dctrl) && (!qpos || dqpos) && (!qvel || dqvel);
  }
  if (fused) {
    b->io_ctrl = dctrl; b->io_qpos = dqpos; b->io_qvel = dqvel; b->io_f64 = dtype == OX_F64;
    ox_status s = b->f64 ? do_step<double>(b, 1) : do_step<float>(b, 1);
    b->io_ctrl = nullptr; b->io_qpos = nullptr; b->io_qvel = nullptr;
    if (s) return s;
    if (mem == OX_MEM_HOST && (qpos || qvel)) CU_TRY(cudaStreamSynchronize(b->stream));
    return OX_OK;
  }
  if (ctrl && b->model->t.nu > 0) { ox_status s = ox_batch_set(b, OX_F_CTRL, ctrl, dtype, mem, OX_LAYOUT_ENV_MAJOR); if (s) return s; }
  ox_status s = ox_batch_step(b, 1);
  if (s) return s;
  int32_t fields[2]; void* bufs[2]; int n = 0;
  if (qpos) { fields[n] = OX_F_QPOS; bufs[n++] = qpos; }
  if (qvel) { fields[n] = OX_F_QVEL; bufs[n++] = qvel; }
  return n ? ox_batch_get_many(b, n, fields, bufs, dtype, mem, OX_LAYOUT_ENV_MAJOR) : OX_OK;
}

ox_status ox_batch_forward(ox_batch* b) {
  if (!b) { ox::set_error("ox_batch_forward: null batch"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  if (b->f64) k_forward_fused<double><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, b->bd);
  else k_forward_fused<float><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, b->bf);
  b->launches++;
  b->derived_stale = false;
  CU_TRY(cudaGetLastError());
  return OX_OK;
}

ox_status ox_batch_reset(ox_batch* b, const uint8_t* host_mask) {
  if (!b) { ox::set_error("ox_batch_reset: null batch"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  const uint8_t* dm = nullptr;
  if (host_mask) {
    CU_TRY(cudaMemcpyAsync(b->d_mask, host_mask, b->nenv, cudaMemcpyHostToDevice, b->stream));
    dm = b->d_mask;
  }
  if (b->f64) k_reset<double><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, b->bd, dm);
  else k_reset<float><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, b->bf, dm);
  b->launches++;
  CU_TRY(cudaGetLastError());
  if (!host_mask) b->applied = 0;  // mj_resetData on every env zeroed the applied forces
  return OX_OK;
}

ox_status ox_batch_sync(ox_batch* b) {
  if (!b) { ox::set_error("ox_batch_sync: null batch"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  CU_TRY(cudaStreamSynchronize(b->stream));
  return OX_OK;
}

ox_status ox_batch_ctrl_philox(ox_batch* b, int32_t enable, uint64_t seed) {
  if (!b) { ox::set_error("ox_batch_ctrl_philox: null batch"); return OX_ERR_INVALID; }
  if ((enable != 0) != (b->philox != 0)) drop_graph(b);
  b->philox = enable ? 1 : 0;
  if (b->seed != seed) drop_graph(b);
  b->seed = seed;
  return OX_OK;
}

ox_status ox_batch_ctrl_philox_scale(ox_batch* b, double scale) {
  if (!b || !(scale > 0)) { ox::set_error("ox_batch_ctrl_philox_scale: bad argument"); return OX_ERR_INVALID; }
  if ((float)scale != b->ctrl_scale) drop_graph(b);
  b->ctrl_scale = (float)scale;
  return OX_OK;
}

ox_status ox_batch_set_step_counter(ox_batch* b, int64_t step) {
  if (!b) { ox::set_error("ox_batch_set_step_counter: null batch"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  b->h_step = step;  // a captured graph picks it up at its next launch (do_step compares with d_step_val)
  return OX_OK;
}

// ---- checkpoint / resume: the integration state of mj_step plus the inputs that persist between steps
static const int kStateFields[] = {OX_F_TIME, OX_F_QPOS, OX_F_QVEL, OX_F_ACT, OX_F_CTRL, OX_F_QFRC_APPLIED, OX_F_XFRC_APPLIED, OX_F_QACC_WARMSTART,
                                   OX_F_MOCAP_POS, OX_F_MOCAP_QUAT, OX_F_EQ_ACTIVE};

int32_t ox_batch_state_size(const ox_batch* b) {
  if (!b) return -1;
  int n = 0;
  for (int f : kStateFields) n += b->fields.at(f).count;
  return n;
}
int64_t ox_batch_get_step_counter(const ox_batch* b) { return b ? b->h_step : -1; }

static ox_status state_io(ox_batch* b, void* buf, int32_t dtype, int32_t mem, int dir) {
  if (!b || !buf) { ox::set_error("ox_batch_get_state / set_state: null argument"); return OX_ERR_INVALID; }
  if ((dtype != OX_F32 && dtype != OX_F64) || (mem != OX_MEM_HOST && mem != OX_MEM_DEVICE)) { ox::set_error("ox_batch_get_state / set_state: bad dtype / mem"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  const int rec = ox_batch_state_size(b);
  const size_t usz = dtype == OX_F64 ? 8 : 4, bytes = (size_t)b->nenv * rec * usz;
  void* dbuf = buf;
  bool staged = false;
  if (mem == OX_MEM_HOST) {
    if (void* alias = host_mapped(buf)) dbuf = alias;
    else {
      ox_status s = ensure_tmp(b, bytes);
      if (s) return s;
      dbuf = b->d_tmp; staged = true;
      if (dir == 0) CU_TRY(cudaMemcpyAsync(dbuf, buf, bytes, cudaMemcpyHostToDevice, b->stream));
    }
  }
  int off = 0;
  for (int f : kStateFields) {
    const FieldInfo& fi = b->fields.at(f);
    if (fi.count == 0) continue;
    const long long n = (long long)b->nenv * fi.count;
    const int threads = 256, blocks = (int)((n + threads - 1) / threads);
#define PACK(TF, TU) k_pack_record<TF, TU><<<blocks, threads, 0, b->stream>>>((TF*)fi.ptr, (TU*)dbuf, b->nenv, fi.count, b->stride, rec, off, dir)
    if (b->f64) { if (dtype == OX_F64) PACK(double, double); else PACK(double, float); }
    else { if (dtype == OX_F64) PACK(float, double); else PACK(float, float); }
#undef PACK
    b->launches++;
    off += fi.count;
  }
  CU_TRY(cudaGetLastError());
  if (dir == 0) b->applied = 1;  // the record carries qfrc_applied / xfrc_applied
  if (mem == OX_MEM_HOST && dir == 1) {
    if (staged) CU_TRY(cudaMemcpyAsync(buf, dbuf, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
  }
  return OX_OK;
}
ox_status ox_batch_get_state(ox_batch* b, void* buf, int32_t dtype, int32_t mem) { return state_io(b, buf, dtype, mem, 1); }
ox_status ox_batch_set_state(ox_batch* b, const void* buf, int32_t dtype, int32_t mem) { return state_io(b, const_cast<void*>(buf), dtype, mem, 0); }

int32_t ox_batch_field_size(const ox_batch* b, int32_t field) {
  if (!b) return -1;
  auto it = b->fields.find(field);
  return it == b->fields.end() ? -1 : it->second.count;
}

ox_status ox_batch_get(ox_batch* b, int32_t field, void* buf, int32_t dtype, int32_t mem, int32_t layout) {
  return bulk_io(b, field, buf, dtype, mem, layout, 1);
}
ox_status ox_batch_get_many(ox_batch* b, int32_t nfields, const int32_t* fields, void* const* bufs, int32_t dtype, int32_t mem, int32_t layout) {
  if (!b || !fields || !bufs || nfields < 1) { ox::set_error("ox_batch_get_many: bad argument"); return OX_ERR_INVALID; }
  // size the staging buffer once for all fields (a growth between fields would free memory an earlier copy still reads)
  size_t total = 0;
  for (int i = 0; i < nfields; i++) {
    auto it = b->fields.find(fields[i]);
    if (it == b->fields.end()) { ox::set_error("ox_batch_get_many: unknown field id " + std::to_string(fields[i])); return OX_ERR_INVALID; }
    total += ((size_t)b->nenv * it->second.count * 8 + 255) / 256 * 256;
  }
  if (mem == OX_MEM_HOST) { ox_status s = ensure_tmp(b, total); if (s) return s; }
  size_t off = 0;
  for (int i = 0; i < nfields; i++) {
    ox_status s = bulk_io(b, fields[i], bufs[i], dtype, mem, layout, 1, off, false);
    if (s) return s;
    off += ((size_t)b->nenv * b->fields[fields[i]].count * 8 + 255) / 256 * 256;
  }
  if (mem == OX_MEM_HOST) CU_TRY(cudaStreamSynchronize(b->stream));
  return OX_OK;
}
ox_status ox_batch_set(ox_batch* b, int32_t field, const void* buf, int32_t dtype, int32_t mem, int32_t layout) {
  return bulk_io(b, field, const_cast<void*>(buf), dtype, mem, layout, 0);
}
ox_status ox_batch_get1(ox_batch* b, int32_t field, int32_t env, int32_t offset, int32_t count, double* out) {
  if (!out) { ox::set_error("ox_batch_get1: null output"); return OX_ERR_INVALID; }
  return slice_io(b, field, env, offset, count, out, nullptr, nullptr);
}
ox_status ox_batch_set1(ox_batch* b, int32_t field, int32_t env, int32_t offset, int32_t count, const double* in) {
  if (!in) { ox::set_error("ox_batch_set1: null input"); return OX_ERR_INVALID; }
  return slice_io(b, field, env, offset, count, nullptr, in, nullptr);
}
ox_status ox_batch_get1_int(ox_batch* b, int32_t field, int32_t env, int32_t offset, int32_t count, int32_t* out) {
  if (!out) { ox::set_error("ox_batch_get1_int: null output"); return OX_ERR_INVALID; }
  return slice_io(b, field, env, offset, count, nullptr, nullptr, out);
}

ox_status ox_batch_stats(ox_batch* b, double* out4) {
  if (!b || !out4) { ox::set_error("ox_batch_stats: null argument"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  std::vector<int32_t> h((size_t)b->nenv);
  int32_t* ptrs[4] = {b->f64 ? b->bd.acc_ncon : b->bf.acc_ncon, b->f64 ? b->bd.acc_nefc : b->bf.acc_nefc,
                      b->f64 ? b->bd.acc_niter : b->bf.acc_niter, b->f64 ? b->bd.diverged : b->bf.diverged};
  for (int k = 0; k < 4; k++) {
    CU_TRY(cudaMemcpyAsync(h.data(), ptrs[k], (size_t)b->nenv * 4, cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
    double s = 0;
    for (int32_t v : h) s += v;
    out4[k] = s;
    if (k < 3) CU_TRY(cudaMemsetAsync(ptrs[k], 0, (size_t)b->stride * 4, b->stream));
  }
  return OX_OK;
}

const char* ox_stage_name(int32_t i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : nullptr; }

ox_status ox_batch_stage_times(ox_batch* b, int32_t reps, double* out_ms, int32_t* nstage) {
  if (!b || !out_ms || !nstage || reps < 1) { ox::set_error("ox_batch_stage_times: bad argument"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(b->cfg.device));
  *nstage = ST_COUNT;
  return b->f64 ? stage_times_impl<double>(b, reps, out_ms) : stage_times_impl<float>(b, reps, out_ms);
}

}  // extern "C"

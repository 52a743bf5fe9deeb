// N1: batched Environment / Task evaluated on device, the direct caller of the batched step.
// Mirrors reference src/lib.rs:8-26 (Task / Observation / Action traits) and :50-88 (Environment::reset / step and
// enum TimeStep) for nenv environments at once; see include/ox_b200.h for how the user-written trait bodies map to
// the declarative ox_task_spec. One thread per environment: k_env_post gathers the observation from the SoA arena,
// evaluates reward and the finish test, keeps episode return / length, and re-initialises finished environments
// (Task::init_episode) after their terminal observation has been written.
//
// No CPU fallback: every entry point launches kernels on the batch's stream or fails with OX_ERR_CUDA.
#include <cuda_runtime.h>

#include <cmath>
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

struct EnvTerm { long long off; int kind; double w; };     // reward term: arena element offset (in reals) of [index][env 0]
struct EnvCond { long long off; double lo, hi; };

struct EnvArgs {
  const long long* obs_off;  // [obs_dim] element offsets
  const EnvTerm* terms;
  const EnvCond* conds;
  int obs_dim, nterm, ncond;
  double bias, time_limit, discount, qpos_noise, qvel_noise;
  uint64_t seed;
  int64_t env_id_offset;
  int auto_reset;
  int* episode;        // [nenv] episode index (Philox counter word)
  int* last_div;       // [nenv] diverged counter seen at the previous call
  double* ep_return;   // [nenv]
  int* ep_len;         // [nenv]
  double* totals;      // [3] finished episodes, sum of returns, sum of lengths
};

// U(-1,1) on the same 2^-23 lattice as the control stream (exact in fp32 and fp64)
OX_HD double env_uniform(uint32_t bits) { return (double)(int32_t)((bits >> 9) * 2u + 1u) * (1.0 / 8388608.0) - 1.0; }

// Task::init_episode for one env: mj_resetData, then seeded noise. Word w of the stream is Philox(counter = (gid.lo,
// gid.hi, episode, w/4), key = seed ^ 0x0E9150DE)[w%4]; qpos coordinate a of a hinge/slide joint uses w = a, dof i uses
// w = 65536 + i (tests/test_env_layer.py restates this on the host).
template <typename T>
__device__ void env_init_episode(const DevModel<T>& m, const DevBatch<T>& b, int e, const EnvArgs& a) {
  Env<T> env(m, b, e);
  env.reset_data();
  const auto& h = m.h();
  const int64_t gid = a.env_id_offset + e;
  const uint64_t key = a.seed ^ 0x0E9150DEull;
  const uint32_t ep = (uint32_t)a.episode[e];
  auto word = [&](uint32_t w) {
    uint32_t out[4];
    philox4x32_10((uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), ep, w >> 2, (uint32_t)key, (uint32_t)(key >> 32), out);
    return env_uniform(out[w & 3]);
  };
  if (a.qpos_noise != 0) {
    for (int j = 0; j < h.njnt; j++) {
      const int t = m.jnt_type(j);
      if (t != OX_JNT_HINGE && t != OX_JNT_SLIDE) continue;
      const int adr = m.jnt_qposadr(j);
      env.at(b.qpos, adr) = (T)((double)m.qpos0(adr) + a.qpos_noise * word((uint32_t)adr));
    }
  }
  if (a.qvel_noise != 0)
    for (int i = 0; i < h.nv; i++) env.at(b.qvel, i) = (T)(a.qvel_noise * word(65536u + (uint32_t)i));
}

template <typename T>
__global__ void k_env_init(const unsigned char* __restrict__ gblob, int bytes, DevBatch<T> b, EnvArgs a, int next_episode) {
  DevModel<T> m{stage_model(gblob, bytes)};
  const int e = env_index(b);
  if (e < 0) return;
  if (next_episode) a.episode[e] += 1;  // a reset after the first one abandons the running episode and draws fresh noise
  env_init_episode(m, b, e, a);
  a.ep_return[e] = 0; a.ep_len[e] = 0; a.last_div[e] = b.diverged[e];
}

// Action::apply for the whole batch: ctrl[i][e] = action[e][i]
template <typename T, typename TU>
__global__ void k_env_apply(DevBatch<T> b, const TU* __restrict__ action, int nu) {
  const int e = env_index(b);
  if (e < 0) return;
  if (b.lanes == 32 && __activemask() == 0xffffffffu) {  // coalesced reads of the warp's contiguous [32][nu] block
    const int e0 = e - (threadIdx.x & 31), n = 32 * nu;
    for (int base = threadIdx.x & 31; base < n; base += 32 * 8) {  // 8 loads in flight: a pinned source is a PCIe round trip each
      TU tmp[8];
#pragma unroll
      for (int j = 0; j < 8; j++) { const int idx = base + 32 * j; tmp[j] = idx < n ? action[(size_t)e0 * nu + idx] : (TU)0; }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int idx = base + 32 * j;
        if (idx < n) { const int el = idx / nu, k = idx - el * nu; b.ctrl[(size_t)k * b.stride + e0 + el] = (T)tmp[j]; }
      }
    }
  } else {
    for (int i = 0; i < nu; i++) b.ctrl[(size_t)i * b.stride + e] = (T)action[(size_t)e * nu + i];
  }
}

// Observation::generate (+ reward, finish, bookkeeping, auto-reset when `post`)
template <typename T, typename TU>
__global__ void k_env_post(const unsigned char* __restrict__ gblob, int bytes, DevBatch<T> b, EnvArgs a, const T* __restrict__ arena,
                           TU* __restrict__ obs, TU* __restrict__ reward, TU* __restrict__ discount, uint8_t* __restrict__ finished, int post) {
  DevModel<T> m{stage_model(gblob, bytes)};
  const int e = env_index(b);
  if (e < 0) return;
  if (obs) {
    // the 32 envs of a full warp own one contiguous [32][obs_dim] block of the output: lanes walk it linearly (coalesced
    // stores; the gathers hit at most a few arena rows of consecutive envs per instruction)
    if (b.lanes == 32 && __activemask() == 0xffffffffu) {
      const int e0 = e - (threadIdx.x & 31), n = 32 * a.obs_dim;
      for (int idx = threadIdx.x & 31; idx < n; idx += 32) {
        const int el = idx / a.obs_dim, k = idx - el * a.obs_dim;
        obs[(size_t)e0 * a.obs_dim + idx] = (TU)arena[a.obs_off[k] + e0 + el];
      }
      __syncwarp();  // the gather read OTHER lanes' arena values; an auto-reset below rewrites this lane's qpos / qvel
    } else {
      for (int k = 0; k < a.obs_dim; k++) obs[(size_t)e * a.obs_dim + k] = (TU)arena[a.obs_off[k] + e];
    }
  }
  if (!post) return;
  double r = a.bias;
  for (int k = 0; k < a.nterm; k++) {
    const double x = (double)arena[a.terms[k].off + e];
    r += a.terms[k].w * (a.terms[k].kind == OX_REWARD_SQUARE ? x * x : a.terms[k].kind == OX_REWARD_ABS ? fabs(x) : x);
  }
  bool fin = false;
  for (int k = 0; k < a.ncond; k++) {
    const double x = (double)arena[a.conds[k].off + e];
    fin |= !(x >= a.conds[k].lo && x <= a.conds[k].hi);  // NaN finishes
  }
  if (a.time_limit > 0) fin |= (double)b.time[e] >= a.time_limit - 0.5 * (double)m.h().timestep;
  const int div = b.diverged[e];
  fin |= div != a.last_div[e];  // mj_checkPos/Vel/Acc reset this env during the step
  a.last_div[e] = div;
  const double ret = a.ep_return[e] + r;
  const int len = a.ep_len[e] + 1;
  if (reward) reward[e] = (TU)r;
  if (discount) discount[e] = (TU)(fin ? 0.0 : a.discount);
  if (finished) finished[e] = fin ? 1 : 0;
  if (fin) {
    atomicAdd(a.totals + 0, 1.0); atomicAdd(a.totals + 1, ret); atomicAdd(a.totals + 2, (double)len);
    a.ep_return[e] = 0; a.ep_len[e] = 0;
    if (a.auto_reset) { a.episode[e] += 1; env_init_episode(m, b, e, a); }
  } else {
    a.ep_return[e] = ret; a.ep_len[e] = len;
  }
}

}  // namespace ox

using namespace ox;

struct ox_env {
  ox_batch* b = nullptr;
  ox_task_spec spec{};
  int obs_dim = 0;
  EnvArgs args{};
  void* d_tables = nullptr;   // obs offsets | terms | conds
  void* d_state = nullptr;    // episode | last_div | ep_len | ep_return | totals
  void* d_io = nullptr;       // staging for host buffers: action | obs | reward | discount | finished
  size_t io_action = 0, io_obs = 0, io_reward = 0, io_discount = 0, io_finished = 0, io_bytes = 0;
  long long resets = 0;
  ~ox_env() { cudaFree(d_tables); cudaFree(d_state); cudaFree(d_io); }
};

namespace {

ox_status resolve(const ox_batch* b, int field, int index, const char* what, long long* off) {
  auto it = b->fields.find(field);
  if (it == b->fields.end() || it->second.is_int || !it->second.ptr) {
    ox::set_error(std::string("ox_env_create: ") + what + " names field " + std::to_string(field) + ", which is not a real-valued batch field");
    return OX_ERR_INVALID;
  }
  {  // a task may only read what the batch's step kernel keeps current (see field_live in ox_batch_internal.cuh)
    if (!field_live(b->spec != nullptr && b->cfg.mode == OX_MODE_FUSED, b->split, field)) { ox::set_error(std::string("ox_env_create: ") + what + ": " + OX_STALE_MSG(field)); return OX_ERR_INVALID; }
  }
  if (index < 0 || index >= it->second.count) {
    ox::set_error(std::string("ox_env_create: ") + what + " index " + std::to_string(index) + " out of range for field " + std::to_string(field) +
                  " (size " + std::to_string(it->second.count) + ")");
    return OX_ERR_INVALID;
  }
  const size_t esz = b->f64 ? 8 : 4;
  *off = (long long)(((unsigned char*)it->second.ptr - b->arena) / esz) + (long long)index * b->stride;
  return OX_OK;
}

template <typename T, typename TU>
void launch_post(ox_env* e, void* obs, void* reward, void* discount, uint8_t* finished, int post) {
  ox_batch* b = e->b;
  const DevBatch<T>& db = *reinterpret_cast<const DevBatch<T>*>(b->f64 ? (const void*)&b->bd : (const void*)&b->bf);
  k_env_post<T, TU><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, db, e->args, (const T*)b->arena, (TU*)obs, (TU*)reward,
                                                                     (TU*)discount, finished, post);
  b->launches++;
}

void dispatch_post(ox_env* e, int dtype, void* obs, void* reward, void* discount, uint8_t* finished, int post) {
  if (e->b->f64) {
    if (dtype == OX_F64) launch_post<double, double>(e, obs, reward, discount, finished, post);
    else launch_post<double, float>(e, obs, reward, discount, finished, post);
  } else {
    if (dtype == OX_F64) launch_post<float, double>(e, obs, reward, discount, finished, post);
    else launch_post<float, float>(e, obs, reward, discount, finished, post);
  }
}

}  // namespace

extern "C" {

void ox_task_spec_default(ox_task_spec* s) {
  if (!s) return;
  *s = ox_task_spec{};
  s->discount = 1.0;
  s->frame_skip = 1;
  s->auto_reset = 1;
}

ox_status ox_env_create(ox_batch* b, const ox_task_spec* spec, ox_env** out) {
  if (!b || !spec || !out) { ox::set_error("ox_env_create: null argument"); return OX_ERR_INVALID; }
  *out = nullptr;
  if (spec->nobs < 0 || spec->nreward < 0 || spec->nfinish < 0 || (spec->nobs && !spec->obs) || (spec->nreward && !spec->reward) ||
      (spec->nfinish && !spec->finish)) { ox::set_error("ox_env_create: inconsistent task arrays"); return OX_ERR_INVALID; }
  if (spec->frame_skip < 1) { ox::set_error("ox_env_create: frame_skip must be >= 1"); return OX_ERR_INVALID; }
  if (!(spec->init_qpos_noise >= 0) || !(spec->init_qvel_noise >= 0)) { ox::set_error("ox_env_create: noise amplitudes must be >= 0"); return OX_ERR_INVALID; }
  std::vector<long long> obs_off;
  for (int s = 0; s < spec->nobs; s++) {
    const ox_obs_segment& g = spec->obs[s];
    if (g.count < 0) { ox::set_error("ox_env_create: negative observation segment"); return OX_ERR_INVALID; }
    for (int i = 0; i < g.count; i++) {
      long long off;
      ox_status st = resolve(b, g.field, g.first + i, "observation segment", &off);
      if (st) return st;
      obs_off.push_back(off);
    }
  }
  std::vector<EnvTerm> terms(spec->nreward);
  for (int k = 0; k < spec->nreward; k++) {
    const ox_reward_term& t = spec->reward[k];
    if (t.kind < OX_REWARD_LINEAR || t.kind > OX_REWARD_ABS) { ox::set_error("ox_env_create: bad reward term kind"); return OX_ERR_INVALID; }
    ox_status st = resolve(b, t.field, t.index, "reward term", &terms[k].off);
    if (st) return st;
    terms[k].kind = t.kind; terms[k].w = t.weight;
  }
  std::vector<EnvCond> conds(spec->nfinish);
  for (int k = 0; k < spec->nfinish; k++) {
    const ox_finish_cond& c = spec->finish[k];
    if (!(c.lo <= c.hi)) { ox::set_error("ox_env_create: finish condition needs lo <= hi"); return OX_ERR_INVALID; }
    ox_status st = resolve(b, c.field, c.index, "finish condition", &conds[k].off);
    if (st) return st;
    conds[k].lo = c.lo; conds[k].hi = c.hi;
  }
  CU_TRY(cudaSetDevice(b->cfg.device));
  std::unique_ptr<ox_env> e(new ox_env);
  e->b = b; e->spec = *spec; e->spec.obs = nullptr; e->spec.reward = nullptr; e->spec.finish = nullptr;
  e->obs_dim = (int)obs_off.size();
  const size_t n = (size_t)b->nenv;
  const size_t b_obs = (obs_off.size() * sizeof(long long) + 255) / 256 * 256, b_terms = (terms.size() * sizeof(EnvTerm) + 255) / 256 * 256;
  const size_t b_conds = (conds.size() * sizeof(EnvCond) + 255) / 256 * 256;
  CU_TRY(cudaMalloc(&e->d_tables, b_obs + b_terms + b_conds + 256));
  unsigned char* t = (unsigned char*)e->d_tables;
  if (!obs_off.empty()) CU_TRY(cudaMemcpyAsync(t, obs_off.data(), obs_off.size() * sizeof(long long), cudaMemcpyHostToDevice, b->stream));
  if (!terms.empty()) CU_TRY(cudaMemcpyAsync(t + b_obs, terms.data(), terms.size() * sizeof(EnvTerm), cudaMemcpyHostToDevice, b->stream));
  if (!conds.empty()) CU_TRY(cudaMemcpyAsync(t + b_obs + b_terms, conds.data(), conds.size() * sizeof(EnvCond), cudaMemcpyHostToDevice, b->stream));
  const size_t b_int = (n * sizeof(int) + 255) / 256 * 256, b_dbl = (n * sizeof(double) + 255) / 256 * 256;
  CU_TRY(cudaMalloc(&e->d_state, 3 * b_int + b_dbl + 256));
  CU_TRY(cudaMemsetAsync(e->d_state, 0, 3 * b_int + b_dbl + 256, b->stream));  // the batch's stream is non-blocking: order everything on it
  unsigned char* s = (unsigned char*)e->d_state;
  EnvArgs& a = e->args;
  a.obs_off = (const long long*)t; a.terms = (const EnvTerm*)(t + b_obs); a.conds = (const EnvCond*)(t + b_obs + b_terms);
  a.obs_dim = e->obs_dim; a.nterm = (int)terms.size(); a.ncond = (int)conds.size();
  a.bias = spec->reward_bias; a.time_limit = spec->time_limit; a.discount = spec->discount;
  a.qpos_noise = spec->init_qpos_noise; a.qvel_noise = spec->init_qvel_noise; a.seed = spec->seed;
  a.env_id_offset = b->cfg.env_id_offset; a.auto_reset = spec->auto_reset ? 1 : 0;
  a.episode = (int*)s; a.last_div = (int*)(s + b_int); a.ep_len = (int*)(s + 2 * b_int); a.ep_return = (double*)(s + 3 * b_int);
  a.totals = (double*)(s + 3 * b_int + b_dbl);
  // staging for host-side buffers (sized for fp64 users)
  const size_t nu = (size_t)b->model->t.nu;
  auto take = [&](size_t bytes) { size_t o = e->io_bytes; e->io_bytes += (bytes + 255) / 256 * 256; return o; };
  e->io_action = take(n * nu * 8); e->io_obs = take(n * (size_t)e->obs_dim * 8); e->io_reward = take(n * 8); e->io_discount = take(n * 8);
  e->io_finished = take(n);
  CU_TRY(cudaMalloc(&e->d_io, e->io_bytes + 256));
  CU_TRY(cudaStreamSynchronize(b->stream));  // the host-side tables above go out of scope
  *out = e.release();
  return OX_OK;
}

void ox_env_free(ox_env* e) {
  if (!e) return;
  cudaSetDevice(e->b->cfg.device);
  cudaStreamSynchronize(e->b->stream);
  delete e;
}

int32_t ox_env_obs_dim(const ox_env* e) { return e ? e->obs_dim : -1; }

ox_status ox_env_reset(ox_env* e, void* obs, int32_t dtype, int32_t mem) {
  if (!e) { ox::set_error("ox_env_reset: null env"); return OX_ERR_INVALID; }
  if ((dtype != OX_F32 && dtype != OX_F64) || (mem != OX_MEM_HOST && mem != OX_MEM_DEVICE)) { ox::set_error("ox_env_reset: bad dtype / mem"); return OX_ERR_INVALID; }
  ox_batch* b = e->b;
  CU_TRY(cudaSetDevice(b->cfg.device));
  const int next = e->resets++ > 0;
  if (b->f64) k_env_init<double><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, b->bd, e->args, next);
  else k_env_init<float><<<b->grid, b->block, b->blob_bytes, b->stream>>>(b->d_blob, b->blob_bytes, b->bf, e->args, next);
  b->launches++;
  CU_TRY(cudaGetLastError());
  ox_status st = ox_batch_forward(b);
  if (st) return st;
  if (obs && e->obs_dim) {
    const size_t usz = dtype == OX_F64 ? 8 : 4, bytes = (size_t)b->nenv * e->obs_dim * usz;
    void* alias = mem == OX_MEM_HOST ? ox::host_mapped(obs) : nullptr;  // pinned host buffer: the kernel writes it directly
    void* dobs = mem == OX_MEM_HOST ? (alias ? alias : (void*)((unsigned char*)e->d_io + e->io_obs)) : obs;
    dispatch_post(e, dtype, dobs, nullptr, nullptr, nullptr, 0);
    CU_TRY(cudaGetLastError());
    if (mem == OX_MEM_HOST) {
      if (!alias) CU_TRY(cudaMemcpyAsync(obs, dobs, bytes, cudaMemcpyDeviceToHost, b->stream));
      CU_TRY(cudaStreamSynchronize(b->stream));
    }
  }
  return OX_OK;
}

ox_status ox_env_step(ox_env* e, const void* action, void* obs, void* reward, void* discount, uint8_t* finished, int32_t dtype, int32_t mem) {
  if (!e) { ox::set_error("ox_env_step: null env"); return OX_ERR_INVALID; }
  if ((dtype != OX_F32 && dtype != OX_F64) || (mem != OX_MEM_HOST && mem != OX_MEM_DEVICE)) { ox::set_error("ox_env_step: bad dtype / mem"); return OX_ERR_INVALID; }
  ox_batch* b = e->b;
  CU_TRY(cudaSetDevice(b->cfg.device));
  const size_t usz = dtype == OX_F64 ? 8 : 4, n = (size_t)b->nenv;
  const int nu = b->model->t.nu;
  const bool host = mem == OX_MEM_HOST;
  unsigned char* io = (unsigned char*)e->d_io;
  if (action && nu > 0) {
    const void* dact = action;
    if (host) {
      if (void* alias = ox::host_mapped(action)) dact = alias;  // pinned: read straight over PCIe, stream-ordered
      else {
        CU_TRY(cudaMemcpyAsync(io + e->io_action, action, n * nu * usz, cudaMemcpyHostToDevice, b->stream));
        dact = io + e->io_action;
      }
    }
    if (b->spec && b->cfg.mode == OX_MODE_FUSED) {
      // the model-specialised step kernel reads the actions itself (StepArgs::io_ctrl): no separate launch
      b->io_ctrl = dact; b->io_f64 = dtype == OX_F64;
    } else if (b->f64) {
      if (dtype == OX_F64) k_env_apply<double, double><<<b->grid, b->block, 0, b->stream>>>(b->bd, (const double*)dact, nu);
      else k_env_apply<double, float><<<b->grid, b->block, 0, b->stream>>>(b->bd, (const float*)dact, nu);
    } else {
      if (dtype == OX_F64) k_env_apply<float, double><<<b->grid, b->block, 0, b->stream>>>(b->bf, (const double*)dact, nu);
      else k_env_apply<float, float><<<b->grid, b->block, 0, b->stream>>>(b->bf, (const float*)dact, nu);
    }
    if (!b->io_ctrl) { b->launches++; CU_TRY(cudaGetLastError()); }
  }
  ox_status st = ox_batch_step(b, e->spec.frame_skip);
  b->io_ctrl = nullptr;
  if (st) return st;
  // host outputs: pinned buffers are written by the kernel itself (zero-copy), pageable ones through the staging area
  void* aobs = host ? ox::host_mapped(obs) : nullptr;
  void* arew = host ? ox::host_mapped(reward) : nullptr;
  void* adis = host ? ox::host_mapped(discount) : nullptr;
  void* afin = host ? ox::host_mapped(finished) : nullptr;
  void* dobs = obs && host ? (aobs ? aobs : (void*)(io + e->io_obs)) : obs;
  void* drew = reward && host ? (arew ? arew : (void*)(io + e->io_reward)) : reward;
  void* ddis = discount && host ? (adis ? adis : (void*)(io + e->io_discount)) : discount;
  uint8_t* dfin = finished && host ? (afin ? (uint8_t*)afin : (uint8_t*)(io + e->io_finished)) : finished;
  dispatch_post(e, dtype, e->obs_dim ? dobs : nullptr, drew, ddis, dfin, 1);
  CU_TRY(cudaGetLastError());
  if (host) {
    if (obs && e->obs_dim && !aobs) CU_TRY(cudaMemcpyAsync(obs, dobs, n * e->obs_dim * usz, cudaMemcpyDeviceToHost, b->stream));
    if (reward && !arew) CU_TRY(cudaMemcpyAsync(reward, drew, n * usz, cudaMemcpyDeviceToHost, b->stream));
    if (discount && !adis) CU_TRY(cudaMemcpyAsync(discount, ddis, n * usz, cudaMemcpyDeviceToHost, b->stream));
    if (finished && !afin) CU_TRY(cudaMemcpyAsync(finished, dfin, n, cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
  }
  return OX_OK;
}

ox_status ox_env_stats(ox_env* e, double* out3) {
  if (!e || !out3) { ox::set_error("ox_env_stats: null argument"); return OX_ERR_INVALID; }
  CU_TRY(cudaSetDevice(e->b->cfg.device));
  CU_TRY(cudaMemcpyAsync(out3, e->args.totals, 3 * sizeof(double), cudaMemcpyDeviceToHost, e->b->stream));
  CU_TRY(cudaStreamSynchronize(e->b->stream));
  return OX_OK;
}

}  // extern "C"

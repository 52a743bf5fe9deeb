// TEST-ONLY harness: instantiates the product's templated stage code (oxide_control_b200/csrc/ox_stages.cuh)
// on the HOST so that the exact arithmetic the CUDA kernels run can be compared with the oracle in a
// container that has no GPU. It is never linked into libox_b200.so and is not a fallback: the product
// has no host path. Built by tests/native/Makefile into tests/native/libox_hostcheck.so.
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../../oxide_control_b200/csrc/ox_arena.h"
#include "../../oxide_control_b200/csrc/ox_spec.cuh"
#include "../../oxide_control_b200/csrc/ox_stages.cuh"

using namespace ox;

struct hc_batch {
  int nenv, stride;
  bool f64;
  std::vector<unsigned char> blob, arena;
  DevBatch<float> bf{};
  DevBatch<double> bd{};
  std::map<int, FieldInfo> fields;
  uint64_t hash = 0;
  SpecRuntime rt{};
};

template <typename T, typename F>
static void for_envs(hc_batch* b, DevBatch<T>& db, F f) {
  DevModel<T> m{b->blob.data()};
  for (int e = 0; e < b->nenv; e++) {
    Env<T> env(m, db, e);
    f(env, e);
  }
}
extern "C" {
#define HC_API __attribute__((visibility("default")))

HC_API hc_batch* hc_create(const ox_model_tables* t, int nenv, int precision, int iterations, int ls_iterations, double tolerance) {
  hc_batch* b = new hc_batch();
  b->nenv = nenv;
  b->stride = (nenv + 31) / 32 * 32;
  b->f64 = precision == OX_F64;
  b->hash = model_hash(*t);
  b->rt.iterations = iterations > 0 ? iterations : t->iterations;
  b->rt.ls_iterations = ls_iterations > 0 ? ls_iterations : t->ls_iterations;
  b->rt.tolerance = precision == 1 ? effective_tolerance<double>(*t, tolerance) : effective_tolerance<float>(*t, tolerance);
  b->blob = b->f64 ? build_blob<double>(*t, iterations, ls_iterations, tolerance) : build_blob<float>(*t, iterations, ls_iterations, tolerance);
  size_t bytes = b->f64 ? layout_arena<double>(*t, b->stride, nullptr, nullptr, nullptr) : layout_arena<float>(*t, b->stride, nullptr, nullptr, nullptr);
  b->arena.assign(bytes + 256, 0);
  unsigned char* base = (unsigned char*)(((uintptr_t)b->arena.data() + 255) / 256 * 256);
  if (b->f64) { layout_arena<double>(*t, b->stride, base, &b->bd, &b->fields); b->bd.nenv = nenv; b->bd.stride = b->stride; b->bd.lanes = 32; }
  else { layout_arena<float>(*t, b->stride, base, &b->bf, &b->fields); b->bf.nenv = nenv; b->bf.stride = b->stride; b->bf.lanes = 32; }
  return b;
}
HC_API void hc_free(hc_batch* b) { delete b; }
HC_API int hc_stride(hc_batch* b) { return b->stride; }

#define DISPATCH(body)                                                     \
  if (b->f64) for_envs<double>(b, b->bd, [&](Env<double>& env, int e) { (void)e; body; }); \
  else for_envs<float>(b, b->bf, [&](Env<float>& env, int e) { (void)e; body; });

HC_API void hc_reset(hc_batch* b) { DISPATCH(env.reset_data()) }
HC_API void hc_forward(hc_batch* b) { DISPATCH(env.forward(false)) }
HC_API void hc_step(hc_batch* b, int nsteps, int philox, uint64_t seed, int64_t env_off, int64_t step0) {
  DISPATCH(for (int s = 0; s < nsteps; s++) {
    if (philox) env.fill_ctrl_philox(seed, env_off + e, step0 + s);
    env.step();
  })
}
// the model-specialised step (csrc/ox_spec.cuh + generated spec_<name>.cu) run on the host; returns 0 if no spec matches
HC_API int hc_step_spec(hc_batch* b, int nsteps, int philox, uint64_t seed, int64_t env_off, int64_t step0) {
  const SpecEntry* sp = find_spec(b->hash);
  if (!sp || !sp->host_f32) return 0;
  StepArgs a;
  a.nsteps = nsteps; a.philox = philox; a.seed = seed; a.env_id_offset = env_off; a.d_step = nullptr;
  for (int e = 0; e < b->nenv; e++) {
    if (b->f64) sp->host_f64(b->bd, e, a, b->rt, step0);
    else sp->host_f32(b->bf, e, a, b->rt, step0);
  }
  return 1;
}
// split pipeline on the host: spec PRE, then the generic thread-serial solver on the arena (the warp-cooperative kernel
// cannot run on a CPU), then spec POST. Checks the data exchange of the PRE/POST phases.
HC_API int hc_step_split(hc_batch* b, int nsteps, int philox, uint64_t seed, int64_t env_off, int64_t step0) {
  const SpecEntry* sp = find_spec(b->hash);
  if (!sp || !sp->host_split_f32[0]) return 0;
  StepArgs a;
  a.nsteps = 1; a.philox = philox; a.seed = seed; a.env_id_offset = env_off; a.d_step = nullptr;
  for (int s = 0; s < nsteps; s++)
    for (int e = 0; e < b->nenv; e++) {
      if (b->f64) {
        sp->host_split_f64[0](b->bd, e, a, b->rt, step0 + s);
        DevModel<double> m{b->blob.data()};
        Env<double> env(m, b->bd, e);
        env.fwd_constraint();
        sp->host_split_f64[1](b->bd, e, a, b->rt, step0 + s);
      } else {
        sp->host_split_f32[0](b->bf, e, a, b->rt, step0 + s);
        DevModel<float> m{b->blob.data()};
        Env<float> env(m, b->bf, e);
        env.fwd_constraint();
        sp->host_split_f32[1](b->bf, e, a, b->rt, step0 + s);
      }
    }
  return 1;
}
HC_API int hc_spec_count() { return spec_count(); }
// raw SoA access: element i of env e of field id is ptr[i*stride + e]
HC_API void* hc_field(hc_batch* b, int field, int* count, int* is_int) {
  auto it = b->fields.find(field);
  if (it == b->fields.end()) { *count = -1; return nullptr; }
  *count = it->second.count;
  *is_int = it->second.is_int;
  return it->second.ptr;
}
HC_API void hc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}
}

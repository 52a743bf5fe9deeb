// Internal: the batch object behind the opaque ox_batch handle, shared by ox_batch.cu and ox_env.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <string>

#include "ox_internal.h"
#include "ox_model.h"
#include "ox_arena.h"
#include "ox_spec.cuh"

struct ox_batch {
  const ox_model* model = nullptr;
  ox_batch_config cfg{};
  int nenv = 0, stride = 0, block = 32, grid = 0, lanes = 32;
  bool f64 = false;
  cudaStream_t stream = nullptr;
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  unsigned char* d_blob = nullptr;
  int blob_bytes = 0;
  ox::DevBatch<float> bf{};
  ox::DevBatch<double> bd{};
  std::map<int, ox::FieldInfo> fields;
  long long launches = 0;
  int philox = 0;
  uint64_t seed = 0;
  long long* d_step = nullptr;
  void* d_tmp = nullptr;  // staging for bulk I/O
  size_t d_tmp_bytes = 0;
  void* h_tmp = nullptr;  // pinned staging for per-env I/O
  size_t h_tmp_bytes = 0;
  uint8_t* d_mask = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  cudaGraph_t graph = nullptr;
  bool split = false;  // fused mode: specialised PRE/POST kernels around the warp-cooperative solver
  bool coop = false;  // staged mode: warp-cooperative Newton solver (ox_solve_coop.cu) instead of the thread-per-env solve stage
  const ox::SpecEntry* spec = nullptr;  // model-specialised step kernel, when one was compiled in for this model
  ox::SpecRuntime spec_rt{};
  // transient: set by ox_batch_step_io around one do_step (device-addressable env-major buffers, see StepArgs)
  const void* io_ctrl = nullptr;
  void* io_qpos = nullptr;
  void* io_qvel = nullptr;
  int io_f64 = 0;
};


#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t err__ = (expr);                                                                        \
    if (err__ != cudaSuccess) {                                                                        \
      ox::set_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " #expr);           \
      return OX_ERR_CUDA;                                                                              \
    }                                                                                                  \
  } while (0)


namespace ox {
// Pinned (page-locked, mapped) host memory is addressable from the device: the pack kernels then read / write the
// caller's buffer directly over PCIe instead of staging through a device buffer plus a cudaMemcpyAsync. Returns the
// device-side alias, or nullptr for pageable memory (staged path). OX_ZERO_COPY=0 in the environment disables it.
inline void* host_mapped(const void* p) {
  static const bool enabled = [] { const char* v = getenv("OX_ZERO_COPY"); return !(v && v[0] == '0'); }();
  if (!enabled || !p) return nullptr;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

}  // namespace ox

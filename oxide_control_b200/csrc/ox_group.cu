// ox_group: N batches of one model, one per GPU of the box, driven from ONE host process (north_star: "one CUDA
// context per GPU ... NCCL over NVLink is used only to gather episode statistics, never inside the step").
//
// The reference owns a single Physics (src/lib.rs:28-31) and has no notion of devices; a Rust (or C) host that wants all
// 8 GPUs gets it here without Python or torch.distributed:
//   * envs shard by global env id: rank r owns ids [offset + r*E, offset + (r+1)*E), so trajectories do not depend on the
//     number of GPUs (the Philox control stream is keyed on the global id);
//   * one persistent host thread per device issues that device's launches (8 devices launch concurrently instead of one
//     thread paying 8 launch latencies per step);
//   * no data-path collective exists: a step is N independent launches;
//   * ox_group_stats reduces the per-env counters to 4 doubles on each device (one kernel) and sums them across devices
//     with ncclAllReduce over NVLink when libnccl.so.2 can be loaded (dlopen - the library has no link-time dependency on
//     NCCL), falling back to a host-side sum of N x 32 bytes otherwise. ox_group_stats_backend() says which ran.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ox_batch_internal.cuh"

namespace ox {

__global__ void k_stats_reduce(const int32_t* __restrict__ ncon, const int32_t* __restrict__ nefc, const int32_t* __restrict__ niter,
                               const int32_t* __restrict__ diverged, int nenv, double* __restrict__ out4) {
  double v[4] = {0, 0, 0, 0};
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nenv; e += gridDim.x * blockDim.x) {
    v[0] += ncon[e]; v[1] += nefc[e]; v[2] += niter[e]; v[3] += diverged[e];
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    if ((threadIdx.x & 31) == 0 && v[k] != 0) atomicAdd(out4 + k, v[k]);
  }
}

// ---- NCCL through dlopen: only the five entry points the statistics reduction needs
struct Nccl {
  void* lib = nullptr;
  int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t stream) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok() const { return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd; }
};
static Nccl& nccl() {
  static Nccl n = [] {
    Nccl x;
    const char* env = getenv("OX_B200_NCCL");  // explicit path, or "0" to disable
    if (env && env[0] == '0' && env[1] == 0) return x;
    for (const char* name : {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"}) {
      x.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (x.lib) break;
    }
    if (!x.lib) return x;
    x.CommInitAll = (decltype(x.CommInitAll))dlsym(x.lib, "ncclCommInitAll");
    x.CommDestroy = (decltype(x.CommDestroy))dlsym(x.lib, "ncclCommDestroy");
    x.AllReduce = (decltype(x.AllReduce))dlsym(x.lib, "ncclAllReduce");
    x.GroupStart = (decltype(x.GroupStart))dlsym(x.lib, "ncclGroupStart");
    x.GroupEnd = (decltype(x.GroupEnd))dlsym(x.lib, "ncclGroupEnd");
    x.GetErrorString = (decltype(x.GetErrorString))dlsym(x.lib, "ncclGetErrorString");
    return x;
  }();
  return n;
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values of NCCL 2.x

// one persistent worker per device: runs the jobs posted to it in order
class Worker {
 public:
  Worker() : th_([this] { loop(); }) {}
  ~Worker() {
    { std::lock_guard<std::mutex> l(mu_); quit_ = true; }
    cv_.notify_all();
    th_.join();
  }
  void post(std::function<ox_status()> job) {
    { std::lock_guard<std::mutex> l(mu_); job_ = std::move(job); busy_ = true; }
    cv_.notify_all();
  }
  ox_status wait(std::string* err) {
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return !busy_; });
    if (status_ != OX_OK && err) *err = error_;
    return status_;
  }
 private:
  void loop() {
    std::unique_lock<std::mutex> l(mu_);
    for (;;) {
      cv_.wait(l, [this] { return quit_ || (busy_ && job_); });
      if (quit_) return;
      auto job = std::move(job_);
      job_ = nullptr;
      l.unlock();
      ox_status s = job();
      std::string e = s != OX_OK ? std::string(ox_last_error_message()) : std::string();  // the message is thread-local: carry it over
      l.lock();
      status_ = s; error_ = e; busy_ = false;
      done_.notify_all();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::function<ox_status()> job_;
  bool busy_ = false, quit_ = false;
  ox_status status_ = OX_OK;
  std::string error_;
  std::thread th_;
};

}  // namespace ox

struct ox_group {
  std::vector<ox_batch*> batches;
  std::vector<std::unique_ptr<ox::Worker>> workers;
  std::vector<double*> d_stats;      // 4 doubles per device
  std::vector<void*> comms;          // ncclComm_t per device (empty: host-side sum)
  std::string backend = "host";
};

using namespace ox;

namespace {
ox_status run_all(ox_group* g, const std::function<ox_status(int)>& fn) {
  const int n = (int)g->batches.size();
  for (int r = 0; r < n; r++) g->workers[r]->post([&fn, r] { return fn(r); });
  ox_status first = OX_OK;
  std::string err;
  for (int r = 0; r < n; r++) {
    std::string e;
    ox_status s = g->workers[r]->wait(&e);
    if (s != OX_OK && first == OX_OK) { first = s; err = "device rank " + std::to_string(r) + ": " + e; }
  }
  if (first != OX_OK) ox::set_error(err);
  return first;
}
}  // namespace

extern "C" {

ox_status ox_group_create(const ox_model* m, const ox_batch_config* cfg, int32_t ndevices, const int32_t* devices, ox_group** out) {
  if (!m || !cfg || !out || ndevices < 1) { ox::set_error("ox_group_create: bad argument"); return OX_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); ox::set_error("ox_group_create: no CUDA device available; libox_b200 has no CPU fallback"); return OX_ERR_CUDA; }
  if (!devices && ndevices > ndev) { ox::set_error("ox_group_create: " + std::to_string(ndevices) + " devices requested, " + std::to_string(ndev) + " visible"); return OX_ERR_INVALID; }
  for (int r = 0; devices && r < ndevices; r++)
    if (devices[r] < 0 || devices[r] >= ndev) { ox::set_error("ox_group_create: bad device ordinal " + std::to_string(devices[r])); return OX_ERR_INVALID; }
  std::unique_ptr<ox_group, void (*)(ox_group*)> g(new ox_group, ox_group_free);
  g->batches.assign(ndevices, nullptr);
  g->d_stats.assign(ndevices, nullptr);
  for (int r = 0; r < ndevices; r++) g->workers.emplace_back(new Worker);
  std::vector<int> devs(ndevices);
  for (int r = 0; r < ndevices; r++) devs[r] = devices ? devices[r] : r;
  ox_group* gp = g.get();
  const ox_batch_config base = *cfg;
  ox_status s = run_all(gp, [gp, m, base, &devs](int r) -> ox_status {
    ox_batch_config c = base;
    c.device = devs[r];
    c.env_id_offset = base.env_id_offset + (int64_t)r * base.nenv;   // global env ids: the trajectory of env g is the same for any GPU count
    ox_status st = ox_batch_create(m, &c, &gp->batches[r]);
    if (st) return st;
    CU_TRY(cudaSetDevice(c.device));
    CU_TRY(cudaMalloc(&gp->d_stats[r], 4 * sizeof(double)));
    return OX_OK;
  });
  if (s) return s;
  bool distinct = true;   // an explicit device list may repeat an ordinal (several shards on one GPU): NCCL needs one rank per device
  for (int r = 0; r < ndevices; r++) for (int q = 0; q < r; q++) distinct &= devs[q] != devs[r];
  if (ndevices > 1 && distinct && nccl().ok()) {
    g->comms.assign(ndevices, nullptr);
    const int rc = nccl().CommInitAll(g->comms.data(), ndevices, devs.data());
    if (rc != 0) g->comms.clear();   // fall back to the host sum; not an error
    else g->backend = "nccl";
  }
  *out = g.release();
  return OX_OK;
}

void ox_group_free(ox_group* g) {
  if (!g) return;
  for (size_t r = 0; r < g->batches.size(); r++) {
    if (!g->batches[r]) continue;
    cudaSetDevice(g->batches[r]->cfg.device);
    cudaStreamSynchronize(g->batches[r]->stream);
    if (r < g->comms.size() && g->comms[r]) nccl().CommDestroy(g->comms[r]);
    cudaFree(g->d_stats[r]);
    ox_batch_free(g->batches[r]);
  }
  g->workers.clear();
  delete g;
}

int32_t ox_group_size(const ox_group* g) { return g ? (int32_t)g->batches.size() : -1; }
ox_batch* ox_group_batch(ox_group* g, int32_t rank) { return (g && rank >= 0 && rank < (int)g->batches.size()) ? g->batches[rank] : nullptr; }
const char* ox_group_stats_backend(const ox_group* g) { return g ? g->backend.c_str() : nullptr; }

ox_status ox_group_step(ox_group* g, int32_t nsteps) {
  if (!g) { ox::set_error("ox_group_step: null group"); return OX_ERR_INVALID; }
  return run_all(g, [g, nsteps](int r) { return ox_batch_step(g->batches[r], nsteps); });
}
ox_status ox_group_sync(ox_group* g) {
  if (!g) { ox::set_error("ox_group_sync: null group"); return OX_ERR_INVALID; }
  return run_all(g, [g](int r) { return ox_batch_sync(g->batches[r]); });
}
ox_status ox_group_reset(ox_group* g) {
  if (!g) { ox::set_error("ox_group_reset: null group"); return OX_ERR_INVALID; }
  return run_all(g, [g](int r) { return ox_batch_reset(g->batches[r], nullptr); });
}
ox_status ox_group_ctrl_philox(ox_group* g, int32_t enable, uint64_t seed) {
  if (!g) { ox::set_error("ox_group_ctrl_philox: null group"); return OX_ERR_INVALID; }
  for (ox_batch* b : g->batches) { ox_status s = ox_batch_ctrl_philox(b, enable, seed); if (s) return s; }
  return OX_OK;
}

ox_status ox_group_stats(ox_group* g, double* out4) {
  if (!g || !out4) { ox::set_error("ox_group_stats: null argument"); return OX_ERR_INVALID; }
  const int n = (int)g->batches.size();
  // per device: per-env counters -> 4 doubles on the device, accumulators cleared (stream-ordered)
  ox_status s = run_all(g, [g](int r) -> ox_status {
    ox_batch* b = g->batches[r];
    CU_TRY(cudaSetDevice(b->cfg.device));
    CU_TRY(cudaMemsetAsync(g->d_stats[r], 0, 4 * sizeof(double), b->stream));
    int32_t *nc = b->f64 ? b->bd.acc_ncon : b->bf.acc_ncon, *ne = b->f64 ? b->bd.acc_nefc : b->bf.acc_nefc;
    int32_t *ni = b->f64 ? b->bd.acc_niter : b->bf.acc_niter, *dv = b->f64 ? b->bd.diverged : b->bf.diverged;
    k_stats_reduce<<<64, 256, 0, b->stream>>>(nc, ne, ni, dv, b->nenv, g->d_stats[r]);
    b->launches++;
    CU_TRY(cudaGetLastError());
    for (int32_t* p : {nc, ne, ni}) CU_TRY(cudaMemsetAsync(p, 0, (size_t)b->stride * 4, b->stream));
    return OX_OK;
  });
  if (s) return s;
  if (!g->comms.empty()) {
    // the one collective of the whole design: 32 bytes per GPU, summed in place over NVLink / NVSwitch
    nccl().GroupStart();
    int rc = 0;
    for (int r = 0; r < n; r++) {
      cudaSetDevice(g->batches[r]->cfg.device);
      rc |= nccl().AllReduce(g->d_stats[r], g->d_stats[r], 4, kNcclFloat64, kNcclSum, g->comms[r], g->batches[r]->stream);
    }
    rc |= nccl().GroupEnd();
    if (rc != 0) { ox::set_error(std::string("ox_group_stats: ncclAllReduce failed: ") + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "?")); return OX_ERR_CUDA; }
    ox_batch* b0 = g->batches[0];
    CU_TRY(cudaSetDevice(b0->cfg.device));
    CU_TRY(cudaMemcpyAsync(out4, g->d_stats[0], 4 * sizeof(double), cudaMemcpyDeviceToHost, b0->stream));
    CU_TRY(cudaStreamSynchronize(b0->stream));
    for (int r = 1; r < n; r++) { CU_TRY(cudaSetDevice(g->batches[r]->cfg.device)); CU_TRY(cudaStreamSynchronize(g->batches[r]->stream)); }
    return OX_OK;
  }
  for (int k = 0; k < 4; k++) out4[k] = 0;
  for (int r = 0; r < n; r++) {
    double v[4];
    ox_batch* b = g->batches[r];
    CU_TRY(cudaSetDevice(b->cfg.device));
    CU_TRY(cudaMemcpyAsync(v, g->d_stats[r], sizeof v, cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
    for (int k = 0; k < 4; k++) out4[k] += v[k];
  }
  return OX_OK;
}

}  // extern "C"

// ---- FMA peak probe
template <typename T>
__global__ void k_fma_probe(T* out, int iters, T a, T b) {
  T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
      x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <typename T>
static ox_status fma_peak(int device, double* tflops) {
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = sizeof(T) == 8 ? 2000 : 20000;
  T* d = nullptr;
  CU_TRY(cudaMalloc(&d, (size_t)blocks * threads * sizeof(T)));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CU_TRY(cudaEventRecord(e0));
    k_fma_probe<T><<<blocks, threads>>>(d, iters, (T)0.999999, (T)1e-6);
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *tflops = best;
  return OX_OK;
}

extern "C" {

ox_status ox_measure_fma_peak(int32_t device, int32_t precision, double* tflops) {
  if (!tflops || (precision != OX_F32 && precision != OX_F64)) { ox::set_error("ox_measure_fma_peak: bad argument"); return OX_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); ox::set_error("ox_measure_fma_peak: no such CUDA device"); return OX_ERR_CUDA; }
  return precision == OX_F64 ? fma_peak<double>(device, tflops) : fma_peak<float>(device, tflops);
}

}  // extern "C"

// Run-time specialisation: compile the model-specialised step kernel for the model of a batch at ox_batch_create time.
//
// Physics::from_xml_string (reference src/physics.rs:18-24) accepts any model at run time, so the throughput kernel cannot be
// limited to the XML files known when libox_b200.so was built. ox_specsrc.h generates the same translation unit ox_specgen
// emits at build time (model policy with compile-time tables + the kernels of ox_spec.cuh); this file compiles it with
// `nvcc -cubin` for sm_100a (the stage headers it needs are embedded in the library, so nothing but the CUDA toolkit has to
// be present), keeps the cubin in an on-disk cache keyed by (model hash, precision, hash of the embedded headers + flags), and
// loads it with cudaLibraryLoadData. A cache hit costs a file read; a miss costs one nvcc run (20-60 s for cheetah-class
// models, minutes for humanoid-class ones).
//
// Environment: OX_B200_JIT=0 disables it (generic kernels are used instead); OX_B200_CACHE_DIR overrides the cache
// location ($XDG_CACHE_HOME/ox_b200, $HOME/.cache/ox_b200, /tmp/ox_b200_cache); OX_B200_NVCC / CUDA_HOME locate nvcc.
#include <cuda_runtime_api.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "ox_specsrc.h"

// the headers a generated unit includes, embedded verbatim (.incbin) so that the JIT does not depend on the source tree
#define OX_INCBIN(sym, file)                                                                                       \
  __asm__(".section .rodata\n.hidden " #sym "\n.global " #sym "\n.balign 16\n" #sym ":\n.incbin \"" file "\"\n"   \
          ".hidden " #sym "_end\n.global " #sym "_end\n" #sym "_end:\n.byte 0\n.previous\n");                      \
  extern "C" const char sym[];                                                                                     \
  extern "C" const char sym##_end[];
OX_INCBIN(ox_embed_spec_cuh, "ox_spec.cuh")
OX_INCBIN(ox_embed_stages_cuh, "ox_stages.cuh")
OX_INCBIN(ox_embed_blob_h, "ox_blob.h")
OX_INCBIN(ox_embed_abi_h, "../../include/ox_b200.h")

namespace ox {
namespace {

const char* kFlags = "-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -cubin";

struct Embedded { const char* rel; const char* begin; const char* end; };
const Embedded* embedded(int* n) {
  static const Embedded e[] = {
      {"src/csrc/ox_spec.cuh", ox_embed_spec_cuh, ox_embed_spec_cuh_end},
      {"src/csrc/ox_stages.cuh", ox_embed_stages_cuh, ox_embed_stages_cuh_end},
      {"src/csrc/ox_blob.h", ox_embed_blob_h, ox_embed_blob_h_end},
      {"include/ox_b200.h", ox_embed_abi_h, ox_embed_abi_h_end},  // "../../include/ox_b200.h" as seen from src/csrc
  };
  *n = 4;
  return e;
}

uint64_t fnv(uint64_t h, const void* p, size_t n) {
  const unsigned char* c = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
  return h;
}
uint64_t abi_hash() {
  static const uint64_t h = [] {
    uint64_t x = 1469598103934665603ull;
    int n;
    const Embedded* e = embedded(&n);
    for (int i = 0; i < n; i++) x = fnv(x, e[i].begin, (size_t)(e[i].end - e[i].begin));
    x = fnv(x, kFlags, strlen(kFlags));
    return x;
  }();
  return h;
}

bool mkdirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); i++) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty() && mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) return false;
    }
    if (i < path.size()) cur += path[i];
  }
  return true;
}
bool write_file(const std::string& path, const char* data, size_t n) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  f.write(data, (std::streamsize)n);
  return (bool)f;
}
bool read_file(const std::string& path, std::vector<char>* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  out->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return !out->empty();
}
std::string cache_dir() {
  if (const char* d = getenv("OX_B200_CACHE_DIR")) if (*d) return d;
  if (const char* d = getenv("XDG_CACHE_HOME")) if (*d) return std::string(d) + "/ox_b200";
  if (const char* d = getenv("HOME")) if (*d) return std::string(d) + "/.cache/ox_b200";
  return "/tmp/ox_b200_cache";
}
std::string find_nvcc() {
  auto ok = [](const std::string& p) { return !p.empty() && access(p.c_str(), X_OK) == 0; };
  if (const char* p = getenv("OX_B200_NVCC")) if (ok(p)) return p;
  for (const char* v : {"CUDA_HOME", "CUDA_PATH"})
    if (const char* d = getenv(v)) { std::string p = std::string(d) + "/bin/nvcc"; if (ok(p)) return p; }
  if (ok("/usr/local/cuda/bin/nvcc")) return "/usr/local/cuda/bin/nvcc";
  if (const char* path = getenv("PATH")) {
    std::stringstream ss(path);
    std::string dir;
    while (std::getline(ss, dir, ':')) { std::string p = dir + "/nvcc"; if (ok(p)) return p; }
  }
  return "";
}
std::string shq(const std::string& s) {  // single-quote for /bin/sh
  std::string r = "'";
  for (char c : s) { if (c == '\'') r += "'\\''"; else r += c; }
  return r + "'";
}
std::string tail(const std::string& path, size_t n) {
  std::vector<char> d;
  if (!read_file(path, &d)) return "";
  return std::string(d.size() > n ? d.end() - (long)n : d.begin(), d.end());
}

// nvcc run: sources into a private directory, cubin moved into the cache atomically
bool compile_unit(const ox_model_tables& t, bool f64, const std::string& cubin, std::string* why) {
  const std::string nvcc = find_nvcc();
  if (nvcc.empty()) { *why = "nvcc not found (set OX_B200_NVCC or CUDA_HOME)"; return false; }
  static std::atomic<int> counter{0};
  const std::string dir = cache_dir();
  char tmpname[64];
  snprintf(tmpname, sizeof tmpname, "/build_%ld_%d", (long)getpid(), counter++);
  const std::string work = dir + tmpname;
  if (!mkdirs(work + "/src/csrc") || !mkdirs(work + "/include")) { *why = "cannot create " + work; return false; }
  int n;
  const Embedded* e = embedded(&n);
  for (int i = 0; i < n; i++)
    if (!write_file(work + "/" + e[i].rel, e[i].begin, (size_t)(e[i].end - e[i].begin))) { *why = "cannot write " + work + "/" + e[i].rel; return false; }
  const std::string src = spec_source_jit(t, f64);
  if (!write_file(work + "/src/csrc/spec.cu", src.data(), src.size())) { *why = "cannot write spec.cu"; return false; }
  const std::string cmd = shq(nvcc) + " " + kFlags + " -I" + shq(work + "/src/csrc") + " -o " + shq(work + "/spec.cubin") + " " +
                          shq(work + "/src/csrc/spec.cu") + " > " + shq(work + "/nvcc.log") + " 2>&1";
  const int rc = system(cmd.c_str());
  std::vector<char> probe;
  if (rc != 0 || !read_file(work + "/spec.cubin", &probe)) {
    *why = "nvcc failed (rc " + std::to_string(rc) + "): " + tail(work + "/nvcc.log", 600) + " [sources kept in " + work + "]";
    return false;
  }
  if (rename((work + "/spec.cubin").c_str(), cubin.c_str()) != 0) { *why = "cannot move the cubin into " + cubin; return false; }
  const std::string rm = "rm -rf " + shq(work);
  if (system(rm.c_str()) != 0) { /* leftovers are harmless */ }
  return true;
}

struct JitEntry {
  SpecEntry spec;
  std::string name;
  cudaLibrary_t lib[2] = {nullptr, nullptr};
  std::string cubin[2];
};
std::mutex g_mu;
std::map<uint64_t, std::unique_ptr<JitEntry>>& table() {
  static std::map<uint64_t, std::unique_ptr<JitEntry>> t;
  return t;
}

}  // namespace

bool jit_enabled() {
  const char* v = getenv("OX_B200_JIT");
  return !(v && v[0] == '0');
}

const SpecEntry* jit_spec(const ox_model_tables& t, bool f64, bool allow_compile, bool load, std::string* why, std::string* cubin_path) {
  std::string dummy;
  if (!why) why = &dummy;
  if (!jit_enabled()) { *why = "run-time specialisation disabled (OX_B200_JIT=0)"; return nullptr; }
  const uint64_t hash = model_hash(t);
  std::lock_guard<std::mutex> lock(g_mu);
  auto& slot = table()[hash];
  if (!slot) {
    slot.reset(new JitEntry);
    char nm[64];
    snprintf(nm, sizeof nm, "jit_%016" PRIx64, hash);
    slot->name = nm;
    slot->spec.hash = hash;
    slot->spec.name = slot->name.c_str();
  }
  JitEntry& je = *slot;
  const int p = f64 ? 1 : 0;
  void** kern = f64 ? je.spec.jit_f64 : je.spec.jit_f32;
  if (je.cubin[p].empty()) {
    const std::string dir = cache_dir();
    if (!mkdirs(dir)) { *why = "cannot create the cache directory " + dir; return nullptr; }
    char fn[128];
    snprintf(fn, sizeof fn, "/ox_%016" PRIx64 "_%s_%016" PRIx64 ".cubin", hash, f64 ? "f64" : "f32", abi_hash());
    const std::string path = dir + fn;
    if (access(path.c_str(), R_OK) != 0) {
      if (!allow_compile) { *why = "no cached cubin for this model (" + path + ")"; return nullptr; }
      if (!compile_unit(t, f64, path, why)) return nullptr;
    }
    je.cubin[p] = path;
  }
  if (cubin_path) *cubin_path = je.cubin[p];
  if (!load) return &je.spec;
  if (!kern[0]) {
    std::vector<char> image;
    if (!read_file(je.cubin[p], &image)) { *why = "cannot read " + je.cubin[p]; return nullptr; }
    cudaLibrary_t lib = nullptr;
    cudaError_t err = cudaLibraryLoadData(&lib, image.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (err != cudaSuccess) { *why = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(err); cudaGetLastError(); return nullptr; }
    cudaKernel_t k0 = nullptr, k1 = nullptr, k2 = nullptr;
    err = cudaLibraryGetKernel(&k0, lib, "ox_jit_step");
    if (err != cudaSuccess) { *why = std::string("cudaLibraryGetKernel(ox_jit_step): ") + cudaGetErrorString(err); cudaGetLastError(); cudaLibraryUnload(lib); return nullptr; }
    if (spec_wants_split(t)) {
      if (cudaLibraryGetKernel(&k1, lib, "ox_jit_pre") != cudaSuccess || cudaLibraryGetKernel(&k2, lib, "ox_jit_post") != cudaSuccess) {
        cudaGetLastError(); k1 = k2 = nullptr;
      }
    }
    je.lib[p] = lib;
    kern[1] = (void*)k1; kern[2] = (void*)k2;
    kern[0] = (void*)k0;
  }
  return &je.spec;
}

}  // namespace ox

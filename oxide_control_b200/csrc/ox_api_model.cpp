// C-ABI entry points for the model (setup) side of the boundary + the error convention.
// Mirrors Physics::from_xml / from_xml_string / object_id / object_name, reference
// src/physics.rs:12-24,56-62 and the Error mapping of src/error.rs:3-21.
#include <cstring>
#include <fstream>
#include <memory>
#include <vector>
#include <sstream>

#include "ox_internal.h"
#include "ox_model.h"
#include "ox_xml.h"

namespace ox {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
}  // namespace ox

extern "C" {

const char* ox_last_error_message(void) { return ox::g_err.c_str(); }
const char* ox_version(void) { return "ox_b200 0.1.0 (sm_100a)"; }

ox_status ox_model_from_xml_string(const char* xml, ox_model** out) {
  if (!xml || !out) {
    ox::set_error("ox_model_from_xml_string: null argument");
    return OX_ERR_INVALID;
  }
  *out = nullptr;
  try {
    *out = ox::compile_mjcf(xml);
    return OX_OK;
  } catch (const ox::XmlError& e) {
    ox::set_error(e.what());
    return OX_ERR_PARSE;
  } catch (const ox::CompileError& e) {
    ox::set_error(e.what());
    return OX_ERR_COMPILE;
  } catch (const std::exception& e) {
    ox::set_error(std::string("internal error: ") + e.what());
    return OX_ERR_COMPILE;
  }
}

ox_status ox_model_from_xml_path(const char* path, ox_model** out) {
  if (!path || !out) {
    ox::set_error("ox_model_from_xml_path: null argument");
    return OX_ERR_INVALID;
  }
  *out = nullptr;
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    ox::set_error(std::string("could not open XML file '") + path + "'");
    return OX_ERR_IO;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  return ox_model_from_xml_string(ss.str().c_str(), out);
}

void ox_model_free(ox_model* m) { delete m; }

}  // extern "C"

// ---- binary model format -------------------------------------------------------------------------------------------
namespace {
constexpr char kMagic[8] = {'O', 'X', 'B', '2', 'M', 'D', 'L', 0};
constexpr uint32_t kFormatVersion = 1;

uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* c = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
  return h;
}
// fingerprint of the table layout: a file written by a library with different tables must be refused, not misread
uint64_t layout_fingerprint() {
  std::string s = std::to_string(sizeof(ox_model_tables)) + ";";
#define OX_X(name, n, w) s += std::string("i:") + #name + ":" + #n + ":" + std::to_string(w) + ";";
  OX_MODEL_INT_TABLES(OX_X)
#undef OX_X
#define OX_X(name, n, w) s += std::string("r:") + #name + ":" + #n + ":" + std::to_string(w) + ";";
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  return fnv1a(s.data(), s.size());
}
struct Writer {
  std::vector<unsigned char> out;
  void raw(const void* p, size_t n) { const unsigned char* c = static_cast<const unsigned char*>(p); out.insert(out.end(), c, c + n); }
  template <typename V> void pod(const V& v) { raw(&v, sizeof v); }
  void str(const std::string& s) { pod<uint32_t>((uint32_t)s.size()); raw(s.data(), s.size()); }
};
struct Reader {
  const unsigned char* p; size_t n, off = 0;
  void raw(void* dst, size_t k) { if (off + k > n) throw std::runtime_error("truncated model file"); std::memcpy(dst, p + off, k); off += k; }
  template <typename V> V pod() { V v; raw(&v, sizeof v); return v; }
  std::string str() { uint32_t k = pod<uint32_t>(); if (off + k > n) throw std::runtime_error("truncated model file"); std::string s((const char*)p + off, k); off += k; return s; }
};
std::vector<unsigned char> serialize(const ox_model* m) {
  Writer w;
  w.raw(kMagic, 8);
  w.pod(kFormatVersion);
  w.pod(layout_fingerprint());
  ox_model_tables t = m->t;   // sizes and options by value; the pointers are re-made on load
#define OX_X(name, n, w_) t.name = nullptr;
  OX_MODEL_INT_TABLES(OX_X)
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  w.pod(t);
#define OX_X(name, n, w_) w.pod<uint64_t>(m->v_##name.size()); w.raw(m->v_##name.data(), m->v_##name.size() * sizeof(m->v_##name[0]));
  OX_MODEL_INT_TABLES(OX_X)
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  w.str(m->model_name);
  w.pod<uint32_t>((uint32_t)m->names.size());
  for (const auto& kv : m->names) {
    w.pod<int32_t>(kv.first);
    w.pod<uint32_t>((uint32_t)kv.second.size());
    for (const auto& s : kv.second) w.str(s);
  }
  w.pod(fnv1a(w.out.data(), w.out.size()));
  return w.out;
}
}  // namespace

extern "C" {

int64_t ox_model_serialize(const ox_model* m, void* buf, int64_t capacity) {
  if (!m) { ox::set_error("ox_model_serialize: null model"); return -1; }
  const std::vector<unsigned char> out = serialize(m);
  if (buf && capacity >= (int64_t)out.size()) std::memcpy(buf, out.data(), out.size());
  return (int64_t)out.size();
}

ox_status ox_model_deserialize(const void* buf, int64_t size, ox_model** out) {
  if (!buf || !out || size < 0) { ox::set_error("ox_model_deserialize: bad argument"); return OX_ERR_INVALID; }
  *out = nullptr;
  try {
    Reader r{static_cast<const unsigned char*>(buf), (size_t)size};
    char magic[8];
    r.raw(magic, 8);
    if (std::memcmp(magic, kMagic, 8)) throw std::runtime_error("not an ox_b200 binary model (bad magic)");
    const uint32_t ver = r.pod<uint32_t>();
    if (ver != kFormatVersion) throw std::runtime_error("binary model format version " + std::to_string(ver) + " is not supported (this library reads version " + std::to_string(kFormatVersion) + ")");
    if (r.pod<uint64_t>() != layout_fingerprint()) throw std::runtime_error("binary model was written by a library with a different table layout; recompile it from the XML");
    if ((size_t)size < sizeof(uint64_t) || fnv1a(buf, (size_t)size - sizeof(uint64_t)) != *reinterpret_cast<const uint64_t*>(static_cast<const unsigned char*>(buf) + size - sizeof(uint64_t)))
      throw std::runtime_error("binary model checksum mismatch (corrupt or truncated file)");
    std::unique_ptr<ox_model> m(new ox_model);
    m->t = r.pod<ox_model_tables>();
#define OX_X(name, n, w_) { const uint64_t k = r.pod<uint64_t>(); if (k > (uint64_t)size) throw std::runtime_error("truncated model file"); m->v_##name.resize(k); r.raw(m->v_##name.data(), k * sizeof(m->v_##name[0])); }
    OX_MODEL_INT_TABLES(OX_X)
    OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
    m->model_name = r.str();
    const uint32_t ntypes = r.pod<uint32_t>();
    for (uint32_t i = 0; i < ntypes; i++) {
      const int32_t type = r.pod<int32_t>();
      const uint32_t cnt = r.pod<uint32_t>();
      auto& v = m->names[type];
      for (uint32_t k = 0; k < cnt; k++) v.push_back(r.str());
    }
    m->finalize();   // re-points the table struct at the vectors and validates every length against the sizes
    *out = m.release();
    return OX_OK;
  } catch (const std::exception& e) {
    ox::set_error(e.what());
    return OX_ERR_IO;
  }
}

ox_status ox_model_save(const ox_model* m, const char* path) {
  if (!m || !path) { ox::set_error("ox_model_save: null argument"); return OX_ERR_INVALID; }
  const std::vector<unsigned char> out = serialize(m);
  std::ofstream f(path, std::ios::binary);
  if (!f || !f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size())) {
    ox::set_error(std::string("could not write binary model '") + path + "'");
    return OX_ERR_IO;
  }
  return OX_OK;
}

ox_status ox_model_load(const char* path, ox_model** out) {
  if (!path || !out) { ox::set_error("ox_model_load: null argument"); return OX_ERR_INVALID; }
  *out = nullptr;
  std::ifstream f(path, std::ios::binary);
  if (!f) { ox::set_error(std::string("could not open binary model '") + path + "'"); return OX_ERR_IO; }
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string s = ss.str();
  return ox_model_deserialize(s.data(), (int64_t)s.size(), out);
}

const ox_model_tables* ox_model_get_tables(const ox_model* m) { return m ? &m->t : nullptr; }

ox_status ox_model_int_table(const ox_model* m, const char* name, const int32_t** ptr, int32_t* count) {
  if (!m || !name || !ptr || !count) {
    ox::set_error("ox_model_int_table: null argument");
    return OX_ERR_INVALID;
  }
#define OX_X(nm, n, w)                    \
  if (!std::strcmp(name, #nm)) {          \
    *ptr = m->t.nm;                       \
    *count = m->t.n * (w);                \
    return OX_OK;                         \
  }
  OX_MODEL_INT_TABLES(OX_X)
#undef OX_X
  ox::set_error(std::string("unknown int table '") + name + "'");
  return OX_ERR_INVALID;
}

ox_status ox_model_real_table(const ox_model* m, const char* name, const double** ptr, int32_t* count) {
  if (!m || !name || !ptr || !count) {
    ox::set_error("ox_model_real_table: null argument");
    return OX_ERR_INVALID;
  }
#define OX_X(nm, n, w)                    \
  if (!std::strcmp(name, #nm)) {          \
    *ptr = m->t.nm;                       \
    *count = m->t.n * (w);                \
    return OX_OK;                         \
  }
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  if (!std::strcmp(name, "gravity")) { *ptr = m->t.gravity; *count = 3; return OX_OK; }
  if (!std::strcmp(name, "timestep")) { *ptr = &m->t.timestep; *count = 1; return OX_OK; }
  if (!std::strcmp(name, "tolerance")) { *ptr = &m->t.tolerance; *count = 1; return OX_OK; }
  if (!std::strcmp(name, "ls_tolerance")) { *ptr = &m->t.ls_tolerance; *count = 1; return OX_OK; }
  if (!std::strcmp(name, "impratio")) { *ptr = &m->t.impratio; *count = 1; return OX_OK; }
  if (!std::strcmp(name, "noslip_tolerance")) { *ptr = &m->t.noslip_tolerance; *count = 1; return OX_OK; }
  if (!std::strcmp(name, "meaninertia")) { *ptr = &m->t.meaninertia; *count = 1; return OX_OK; }
  if (!std::strcmp(name, "density")) { *ptr = &m->t.density; *count = 1; return OX_OK; }
  if (!std::strcmp(name, "viscosity")) { *ptr = &m->t.viscosity; *count = 1; return OX_OK; }
  if (!std::strcmp(name, "wind")) { *ptr = m->t.wind; *count = 3; return OX_OK; }
  ox::set_error(std::string("unknown real table '") + name + "'");
  return OX_ERR_INVALID;
}

int32_t ox_model_size(const ox_model* m, const char* name) {
  if (!m || !name) return -1;
  const ox_model_tables& t = m->t;
#define S(f) if (!std::strcmp(name, #f)) return t.f;
  S(nq) S(nv) S(nu) S(na) S(nbody) S(njnt) S(ngeom) S(nsite) S(nM) S(npair) S(nsensor) S(nsensordata) S(nconmax) S(nefcmax) S(nvv) S(nmocap) S(neq) S(ntendon) S(nwrap)
  S(integrator) S(solver) S(cone) S(iterations) S(ls_iterations) S(disableflags) S(noslip_iterations) S(nfloss) S(nfluid) S(ngravcomp)
#undef S
  return -1;
}

int32_t ox_model_name2id(const ox_model* m, int32_t objtype, const char* name) {
  if (!m || !name || !*name) return -1;
  if (objtype == OX_OBJ_XBODY) objtype = OX_OBJ_BODY;
  auto it = m->names.find(objtype);
  if (it == m->names.end()) return -1;
  for (size_t i = 0; i < it->second.size(); i++)
    if (it->second[i] == name) return (int32_t)i;
  return -1;
}

const char* ox_model_id2name(const ox_model* m, int32_t objtype, int32_t id) {
  if (!m) return nullptr;
  if (objtype == OX_OBJ_XBODY) objtype = OX_OBJ_BODY;
  auto it = m->names.find(objtype);
  if (it == m->names.end() || id < 0 || id >= (int32_t)it->second.size()) return nullptr;
  return it->second[id].c_str();
}

}  // extern "C"

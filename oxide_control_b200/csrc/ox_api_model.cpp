// C-ABI entry points for the model (setup) side of the boundary + the error convention.
// Mirrors Physics::from_xml / from_xml_string / object_id / object_name, reference
// src/physics.rs:12-24,56-62 and the Error mapping of src/error.rs:3-21.
#include <cstring>
#include <fstream>
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
  if (!std::strcmp(name, "meaninertia")) { *ptr = &m->t.meaninertia; *count = 1; return OX_OK; }
  ox::set_error(std::string("unknown real table '") + name + "'");
  return OX_ERR_INVALID;
}

int32_t ox_model_size(const ox_model* m, const char* name) {
  if (!m || !name) return -1;
  const ox_model_tables& t = m->t;
#define S(f) if (!std::strcmp(name, #f)) return t.f;
  S(nq) S(nv) S(nu) S(na) S(nbody) S(njnt) S(ngeom) S(nsite) S(nM) S(npair) S(nsensor) S(nsensordata) S(nconmax) S(nefcmax) S(nvv) S(nmocap) S(neq)
  S(integrator) S(solver) S(cone) S(iterations) S(ls_iterations) S(disableflags)
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

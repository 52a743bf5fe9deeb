// Internal representation of a compiled model: owns the vectors behind ox_model_tables.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ox_b200.h"

struct ox_model {
  ox_model_tables t{};
#define OX_X(name, n, w) std::vector<int32_t> v_##name;
  OX_MODEL_INT_TABLES(OX_X)
#undef OX_X
#define OX_X(name, n, w) std::vector<double> v_##name;
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
  std::map<int, std::vector<std::string>> names;  // objtype -> names by id
  std::string model_name;

  // (re)point the table struct at the vectors; validates lengths
  void finalize();
};

namespace ox {
struct CompileError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// MJCF text -> compiled model. Throws XmlError (parse) or CompileError.
ox_model* compile_mjcf(const std::string& xml);
}  // namespace ox

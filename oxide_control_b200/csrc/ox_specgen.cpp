// ox_specgen: build-time generator of a model-specialised translation unit.
//   ox_specgen <model.xml> <name> > spec_<name>.cu
// It compiles the MJCF with the product's own compiler and prints the unit ox_specsrc.h generates: a model policy type
// whose table lookups are compile-time constants, plus the launchers and the registry entry.
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>

#include "ox_specsrc.h"
#include "ox_xml.h"

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: ox_specgen model.xml name\n"); return 2; }
  std::ifstream f(argv[1]);
  if (!f) { std::fprintf(stderr, "ox_specgen: cannot read %s\n", argv[1]); return 2; }
  std::stringstream ss;
  ss << f.rdbuf();
  ox_model* m = nullptr;
  try {
    m = ox::compile_mjcf(ss.str());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ox_specgen: %s: %s\n", argv[1], e.what());
    return 1;
  }
  const std::string src = ox::spec_source_static(m->t, argv[2], argv[1]);
  std::fwrite(src.data(), 1, src.size(), stdout);
  ox_model_free(m);
  return 0;
}

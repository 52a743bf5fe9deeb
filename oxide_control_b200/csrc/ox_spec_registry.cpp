// Registry of model-specialised kernels (filled by static initialisers of the generated spec_<name>.cu units).
#include <vector>

#include "ox_spec.cuh"

namespace ox {
static std::vector<SpecEntry>& table() {
  static std::vector<SpecEntry> t;
  return t;
}
void register_spec(const SpecEntry& e) { table().push_back(e); }
const SpecEntry* find_spec(uint64_t hash) {
  for (auto& e : table())
    if (e.hash == hash) return &e;
  return nullptr;
}
int spec_count() { return (int)table().size(); }
const SpecEntry* spec_at(int i) { return (i >= 0 && i < (int)table().size()) ? &table()[i] : nullptr; }
}  // namespace ox

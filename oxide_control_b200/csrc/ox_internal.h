// Internal helpers shared by the C-ABI translation units.
#pragma once
#include <string>
namespace ox {
void set_error(const std::string& msg);
}

// last_error.hpp -- the thread-local message behind pvqt_last_error_string(), shared by the
// translation units of libpvqt.so.
#pragma once
#include <string>

namespace pvqt_detail {
void set_last_error(const std::string &message);
}

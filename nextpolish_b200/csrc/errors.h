// errors.h — thread-local last-error string shared by the host and device halves of the ABI.
#pragma once
#include <string>
namespace np {
void set_error(const std::string& e);
const std::string& get_error();
}

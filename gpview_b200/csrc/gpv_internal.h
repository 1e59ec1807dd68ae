// gpview_b200/csrc/gpv_internal.h -- shared between the host (.cpp) and device (.cu) halves of libgpview_b200.so
#pragma once
#include <string>

namespace gpv {
// records the message for gpv_last_error() (thread-local) and returns a non-zero error code
int fail(const std::string& msg);
}

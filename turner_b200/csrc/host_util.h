// Host-only helpers shared by the C ABI and the CLI (no CUDA here).
#pragma once
#include <string>

namespace trn {
extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);
} // namespace trn

// TEST INFRASTRUCTURE ONLY -- stand-in for docopt.cpp@1811022's docopt::value so
// that /root/reference/config.h compiles. oracle/ref_driver.cpp fills TracerConfig
// fields directly; no argument parsing happens here.
#pragma once
#include <string>
namespace docopt {
struct value {
    bool b = false;
    long l = 0;
    std::string s;
    bool has = false;
    bool asBool() const { return b; }
    long asLong() const { return l; }
    const std::string& asString() const { return s; }
    explicit operator bool() const { return has; }
};
} // namespace docopt

// glog stand-in for compiling the reference's header-only depth_vector.hpp in place
// (oracle/Makefile target ref).  TEST INFRASTRUCTURE ONLY.  CHECK_* abort with a message like
// glog does; LOG(...) swallows its stream.
#ifndef EMVS_GLOG_SHIM_H_
#define EMVS_GLOG_SHIM_H_
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <algorithm>
#include <utility>

namespace emvs_glog_shim {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct Fatal {
  const char* what;
  explicit Fatal(const char* w) : what(w) {}
  [[noreturn]] ~Fatal() { std::fprintf(stderr, "Check failed: %s\n", what); std::abort(); }
  template <typename T> Fatal& operator<<(const T&) { return *this; }
};
}  // namespace emvs_glog_shim

#define LOG(severity) ::emvs_glog_shim::NullStream()
#define LOG_FIRST_N(severity, n) ::emvs_glog_shim::NullStream()
#define VLOG(level) ::emvs_glog_shim::NullStream()
#define EMVS_SHIM_CHECK(cond, text) if (cond) {} else ::emvs_glog_shim::Fatal(text)
#define CHECK(c) EMVS_SHIM_CHECK((c), #c)
#define CHECK_GT(a, b) EMVS_SHIM_CHECK((a) > (b), #a " > " #b)
#define CHECK_GE(a, b) EMVS_SHIM_CHECK((a) >= (b), #a " >= " #b)
#define CHECK_LT(a, b) EMVS_SHIM_CHECK((a) < (b), #a " < " #b)
#define CHECK_LE(a, b) EMVS_SHIM_CHECK((a) <= (b), #a " <= " #b)
#define CHECK_EQ(a, b) EMVS_SHIM_CHECK((a) == (b), #a " == " #b)
#endif

// Stand-in for glog's CHECK family (test infrastructure, see ../README.md): a failed check prints the streamed
// message and aborts; DCHECKs are compiled in.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace ref_stub {
struct check_fail {
  std::ostringstream os;
  check_fail(const char* file, int line, const char* what) { os << file << ":" << line << " check failed: " << what << " "; }
  [[noreturn]] ~check_fail() {
    std::cerr << os.str() << std::endl;
    std::abort();
  }
  template <class T>
  check_fail& operator<<(const T& v) {
    os << v;
    return *this;
  }
};
struct voidify {
  void operator&(const check_fail&) {}
};
}  // namespace ref_stub
#define CHECK(c) (c) ? (void)0 : ::ref_stub::voidify() & ::ref_stub::check_fail(__FILE__, __LINE__, #c)
#define REF_STUB_CHECK_OP(a, op, b) CHECK((a)op(b))
#define CHECK_EQ(a, b) REF_STUB_CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) REF_STUB_CHECK_OP(a, !=, b)
#define CHECK_LT(a, b) REF_STUB_CHECK_OP(a, <, b)
#define CHECK_LE(a, b) REF_STUB_CHECK_OP(a, <=, b)
#define CHECK_GT(a, b) REF_STUB_CHECK_OP(a, >, b)
#define CHECK_GE(a, b) REF_STUB_CHECK_OP(a, >=, b)
#define DCHECK(c) CHECK(c)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define LOG(sev) ::std::cerr
#define VLOG(n) if (false) ::std::cerr

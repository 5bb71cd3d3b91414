// Stand-in for glog's CHECK family (test infrastructure, see ../README.md): a failed check prints the streamed
// message and throws std::runtime_error, so that a host process embedding oracle/_ref (pytest, bench.py) reports it
// instead of dying -- unless another exception is already in flight, where it aborts as glog does.  DCHECKs are
// compiled in.
#pragma once
#include <cstdlib>
#include <exception>
#include <stdexcept>
#include <iostream>
#include <sstream>
namespace ref_stub {
struct check_fail {
  std::ostringstream os;
  check_fail(const char* file, int line, const char* what) { os << file << ":" << line << " check failed: " << what << " "; }
  ~check_fail() noexcept(false) {
    std::cerr << os.str() << std::endl;
    if (std::uncaught_exceptions() > 0) std::abort();
    throw std::runtime_error(os.str());
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

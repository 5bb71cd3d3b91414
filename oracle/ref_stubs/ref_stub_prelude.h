// Force-included in front of every reference source compiled into oracle/_ref (test infrastructure): standard
// headers the real Boost headers pull in transitively, and the few Boost free functions used without an include.
#pragma once
#include <unistd.h>
#include <cmath>
#include <cstddef>
#include <functional>
#include <memory>
#include <string>
using std::make_unique;
namespace boost {
template <class T>
inline void hash_combine(std::size_t& seed, const T& v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b97f4a7c15ULL + (seed << 6) + (seed >> 2);
}
}  // namespace boost

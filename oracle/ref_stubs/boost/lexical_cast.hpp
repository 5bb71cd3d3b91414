// Stand-in for boost::lexical_cast (test infrastructure).
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : public std::runtime_error {
  bad_lexical_cast() : std::runtime_error("bad lexical cast") {}
};
template <class T, class S>
T lexical_cast(const S& s) {
  std::stringstream ss;
  ss << s;
  T v;
  if (!(ss >> v) || !(ss >> std::ws).eof()) throw bad_lexical_cast();
  return v;
}
}  // namespace boost

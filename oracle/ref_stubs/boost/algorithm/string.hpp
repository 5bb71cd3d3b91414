// Stand-in for the Boost.StringAlgo pieces the reference's leaf sources use (test infrastructure).
#pragma once
#include <string>
#include <vector>
#include "boost/algorithm/string/predicate.hpp"
namespace boost {
struct is_any_of {
  std::string set;
  explicit is_any_of(const std::string& s) : set(s) {}
  bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
template <class Out, class Pred>
Out& split(Out& out, const std::string& in, Pred pred) {
  out.clear();
  std::string cur;
  for (char c : in) {
    if (pred(c)) { out.push_back(cur); cur.clear(); } else cur += c;
  }
  out.push_back(cur);
  return out;
}
}  // namespace boost

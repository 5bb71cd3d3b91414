// Stand-in for the Boost.StringAlgo predicates the reference uses (test infrastructure).
#pragma once
#include <string>
namespace boost { namespace algorithm {
inline bool starts_with(const std::string& s, const std::string& p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
inline bool ends_with(const std::string& s, const std::string& p) { return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0; }
}
using algorithm::starts_with;
using algorithm::ends_with;
}  // namespace boost

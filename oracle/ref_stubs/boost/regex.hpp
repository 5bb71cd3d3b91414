// Stand-in: boost::regex over std::regex (test infrastructure).
#pragma once
#include <regex>
namespace boost {
using std::regex;
using std::regex_match;
using std::regex_search;
using std::smatch;
}  // namespace boost

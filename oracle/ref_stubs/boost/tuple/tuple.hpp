// Stand-in: boost::tuple over std::tuple (test infrastructure).
#pragma once
#include <tuple>
namespace boost {
template <class... T> using tuple = std::tuple<T...>;
using std::get;
using std::make_tuple;
using std::tie;
}  // namespace boost

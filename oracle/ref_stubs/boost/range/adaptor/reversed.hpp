// Stand-in for boost::adaptors::reverse (test infrastructure).
#pragma once
#include <iterator>
namespace boost { namespace adaptors {
template <class C>
struct reversed_range {
  C& c;
  auto begin() const { return std::rbegin(c); }
  auto end() const { return std::rend(c); }
};
template <class C> reversed_range<C> reverse(C& c) { return reversed_range<C>{c}; }
template <class C> reversed_range<const C> reverse(const C& c) { return reversed_range<const C>{c}; }
}}  // namespace boost::adaptors

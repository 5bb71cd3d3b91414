// Stand-in for boost::iterator_range (test infrastructure).
#pragma once
#include <cstddef>
#include <iterator>
namespace boost {
template <class It>
class iterator_range {
  It m_b, m_e;
 public:
  iterator_range() = default;
  iterator_range(It b, It e) : m_b(b), m_e(e) {}
  It begin() const { return m_b; }
  It end() const { return m_e; }
  bool empty() const { return m_b == m_e; }
  std::size_t size() const { return std::distance(m_b, m_e); }
};
template <class It> iterator_range<It> make_iterator_range(It b, It e) { return iterator_range<It>(b, e); }
}  // namespace boost

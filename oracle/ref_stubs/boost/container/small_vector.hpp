// Stand-in: small_vector without the inline storage (test infrastructure).
#pragma once
#include <cstddef>
#include <vector>
namespace boost { namespace container {
template <class T, std::size_t N>
class small_vector : public std::vector<T> {
 public:
  using std::vector<T>::vector;
};
}}  // namespace boost::container

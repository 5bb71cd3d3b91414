// Stand-in for the Boost.Operators mix-ins the reference's DNA types derive from (test infrastructure).
#pragma once
namespace boost {
struct ref_stub_no_base {};
template <class T, class Base = ref_stub_no_base>
struct totally_ordered : Base {
  friend bool operator>(const T& a, const T& b) { return b < a; }
  friend bool operator<=(const T& a, const T& b) { return !(b < a); }
  friend bool operator>=(const T& a, const T& b) { return !(a < b); }
  friend bool operator!=(const T& a, const T& b) { return !(a == b); }
};
template <class T, class U = T>
struct less_than_comparable {  // T vs U, given T < U and T > U
  friend bool operator<=(const T& a, const U& b) { return !(a > b); }
  friend bool operator>=(const T& a, const U& b) { return !(a < b); }
  friend bool operator>(const U& a, const T& b) { return b < a; }
  friend bool operator<(const U& a, const T& b) { return b > a; }
  friend bool operator<=(const U& a, const T& b) { return !(b < a); }
  friend bool operator>=(const U& a, const T& b) { return !(b > a); }
};
template <class T>
struct less_than_comparable<T, T> {
  friend bool operator>(const T& a, const T& b) { return b < a; }
  friend bool operator<=(const T& a, const T& b) { return !(b < a); }
  friend bool operator>=(const T& a, const T& b) { return !(a < b); }
};
template <class T>
struct equality_comparable {
  friend bool operator!=(const T& a, const T& b) { return !(a == b); }
};
template <class T>
struct addable {
  friend T operator+(T a, const T& b) {
    a += b;
    return a;
  }
};
}  // namespace boost

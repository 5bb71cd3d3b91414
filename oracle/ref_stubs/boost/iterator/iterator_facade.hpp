// Stand-in for boost::iterator_facade (test infrastructure): the iterator operators are written in terms of the
// derived class's dereference / increment / decrement / advance / equal / distance_to members.
#pragma once
#include <cstddef>
#include <iterator>
#include <type_traits>
namespace boost {
class iterator_core_access {
 public:
  template <class I> static typename I::reference dereference(const I& i) { return i.dereference(); }
  template <class I> static void increment(I& i) { i.increment(); }
  template <class I> static void decrement(I& i) { i.decrement(); }
  template <class I, class D> static void advance(I& i, D n) { i.advance(n); }
  template <class I, class J> static bool equal(const I& a, const J& b) { return a.equal(b); }
  template <class I, class J> static std::ptrdiff_t distance_to(const I& a, const J& b) { return a.distance_to(b); }
};
namespace detail_stub {
template <class Ref>
struct arrow_proxy {
  Ref r;
  Ref* operator->() { return &r; }
};
}  // namespace detail_stub
template <class Derived, class Value, class Category, class Reference = Value&, class Difference = std::ptrdiff_t>
class iterator_facade {
  Derived& self() { return *static_cast<Derived*>(this); }
  const Derived& self() const { return *static_cast<const Derived*>(this); }

 public:
  typedef typename std::remove_const<Value>::type value_type;
  typedef Reference reference;
  typedef Value* pointer;
  typedef Difference difference_type;
  typedef Category iterator_category;

  reference operator*() const { return iterator_core_access::dereference(self()); }
  detail_stub::arrow_proxy<reference> operator->() const { return detail_stub::arrow_proxy<reference>{**this}; }
  reference operator[](difference_type n) const { return *(self() + n); }
  Derived& operator++() { iterator_core_access::increment(self()); return self(); }
  Derived operator++(int) { Derived t(self()); ++*this; return t; }
  Derived& operator--() { iterator_core_access::decrement(self()); return self(); }
  Derived operator--(int) { Derived t(self()); --*this; return t; }
  Derived& operator+=(difference_type n) { iterator_core_access::advance(self(), n); return self(); }
  Derived& operator-=(difference_type n) { iterator_core_access::advance(self(), -n); return self(); }
  Derived operator+(difference_type n) const { Derived t(self()); t += n; return t; }
  Derived operator-(difference_type n) const { Derived t(self()); t -= n; return t; }
  friend Derived operator+(difference_type n, const Derived& i) { return i + n; }
};
#define REF_STUB_FACADE_ARGS class D1, class V1, class C1, class R1, class F1, class D2, class V2, class C2, class R2, class F2
#define REF_STUB_FACADE_L iterator_facade<D1, V1, C1, R1, F1>
#define REF_STUB_FACADE_R iterator_facade<D2, V2, C2, R2, F2>
template <REF_STUB_FACADE_ARGS>
bool operator==(const REF_STUB_FACADE_L& a, const REF_STUB_FACADE_R& b) {
  return iterator_core_access::equal(static_cast<const D1&>(a), static_cast<const D2&>(b));
}
template <REF_STUB_FACADE_ARGS>
bool operator!=(const REF_STUB_FACADE_L& a, const REF_STUB_FACADE_R& b) { return !(a == b); }
template <REF_STUB_FACADE_ARGS>
std::ptrdiff_t operator-(const REF_STUB_FACADE_L& a, const REF_STUB_FACADE_R& b) {
  return iterator_core_access::distance_to(static_cast<const D2&>(b), static_cast<const D1&>(a));
}
template <REF_STUB_FACADE_ARGS>
bool operator<(const REF_STUB_FACADE_L& a, const REF_STUB_FACADE_R& b) { return (a - b) < 0; }
template <REF_STUB_FACADE_ARGS>
bool operator>(const REF_STUB_FACADE_L& a, const REF_STUB_FACADE_R& b) { return (a - b) > 0; }
template <REF_STUB_FACADE_ARGS>
bool operator<=(const REF_STUB_FACADE_L& a, const REF_STUB_FACADE_R& b) { return (a - b) <= 0; }
template <REF_STUB_FACADE_ARGS>
bool operator>=(const REF_STUB_FACADE_L& a, const REF_STUB_FACADE_R& b) { return (a - b) >= 0; }
}  // namespace boost

// Stand-in: boost::optional over std::optional (test infrastructure).
#pragma once
#include <functional>
#include <optional>
#include <ostream>
namespace boost {
struct none_t {};
static const none_t none{};
template <class T>
class optional : public std::optional<T> {
 public:
  using std::optional<T>::optional;
  optional() = default;
  optional(none_t) {}
  optional& operator=(none_t) { this->reset(); return *this; }
  template <class U, class = typename std::enable_if<!std::is_same<typename std::decay<U>::type, optional>::value &&
                                                     !std::is_same<typename std::decay<U>::type, none_t>::value>::type>
  optional& operator=(U&& u) {
    std::optional<T>::operator=(std::forward<U>(u));
    return *this;
  }
  T& get() { return **this; }
  const T& get() const { return **this; }
  bool is_initialized() const { return this->has_value(); }
};
template <class T>
std::ostream& operator<<(std::ostream& os, const optional<T>& o) {
  if (o) return os << ' ' << *o;
  return os << "--";
}
}  // namespace boost

// Stand-in for boost::format (test infrastructure): positional / printf directives are replaced in order by the
// streamed arguments; enough for the error messages of the reference's leaf sources.
#pragma once
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
  std::string m_fmt;
  std::vector<std::string> m_args;

 public:
  explicit format(const std::string& f) : m_fmt(f) {}
  explicit format(const char* f) : m_fmt(f) {}
  template <class T>
  format& operator%(const T& v) {
    std::ostringstream os;
    os << v;
    m_args.push_back(os.str());
    return *this;
  }
  std::string str() const {
    std::string out;
    size_t next = 0;
    for (size_t i = 0; i < m_fmt.size(); ++i) {
      if (m_fmt[i] != '%') { out += m_fmt[i]; continue; }
      if (i + 1 < m_fmt.size() && m_fmt[i + 1] == '%') { out += '%'; ++i; continue; }
      size_t j = i + 1;
      while (j < m_fmt.size() && !isalpha((unsigned char)m_fmt[j]) && m_fmt[j] != '%') ++j;
      if (j < m_fmt.size() && m_fmt[j] == '%') {  // %N%
        size_t idx = (size_t)atoi(m_fmt.c_str() + i + 1);
        if (idx >= 1 && idx <= m_args.size()) out += m_args[idx - 1];
      } else {
        while (j < m_fmt.size() && (m_fmt[j] == 'l' || m_fmt[j] == 'h' || m_fmt[j] == 'z')) ++j;
        if (next < m_args.size()) out += m_args[next++];
      }
      i = j;
    }
    return out;
  }
};
inline std::string str(const format& f) { return f.str(); }
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
}  // namespace boost

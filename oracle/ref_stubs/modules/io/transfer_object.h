// Shadows modules/io/transfer_object.h (test infrastructure).  The reference serialises its value types through
// json_spirit / msgpack, neither of which is in this image.  Here TRANSFER_OBJECT declares a member template that
// visits the fields with a small JSON writer / reader (below), which is all the spiral-file bookkeeping members
// (part_info.json, packed_vector.json, bitcount.json ...) need; the JSON text is bookkeeping, not payload, and is
// not meant to equal json_spirit's formatting.
#pragma once
#include <cstdlib>
#include <map>
#include <set>
#include <stdexcept>
#include <stdint.h>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>
#include <boost/container/small_vector.hpp>
#include <boost/tuple/tuple.hpp>
#include "modules/io/io.h"
#include "modules/io/utils.h"

template <class T> struct transfer_info;
class deserialization_error : public io_exception {
 public:
  deserialization_error(const std::string& s) : io_exception(s) {}
};

namespace ref_stub_json {

struct value {
  enum kind_t { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
  std::string text;  // BOOL: "true"/"false", NUM: the literal, STR: unescaped
  std::vector<value> arr;
  std::vector<std::pair<std::string, value>> obj;
  const value* find(const std::string& k) const {
    for (const auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
};

class parser {
 public:
  explicit parser(const std::string& s) : m_p(s.data()), m_e(s.data() + s.size()) {}
  value parse() {
    value v = any();
    ws();
    return v;
  }

 private:
  const char *m_p, *m_e;
  [[noreturn]] void fail(const char* what) { throw deserialization_error(std::string("json: ") + what); }
  void ws() {
    while (m_p < m_e && (*m_p == ' ' || *m_p == '\n' || *m_p == '\t' || *m_p == '\r')) ++m_p;
  }
  std::string str() {
    std::string out;
    ++m_p;
    while (m_p < m_e && *m_p != '"') {
      if (*m_p == '\\' && m_p + 1 < m_e) {
        ++m_p;
        switch (*m_p) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          default: out += *m_p;
        }
      } else {
        out += *m_p;
      }
      ++m_p;
    }
    if (m_p == m_e) fail("unterminated string");
    ++m_p;
    return out;
  }
  value any() {
    ws();
    if (m_p == m_e) fail("unexpected end");
    value v;
    if (*m_p == '{') {
      v.kind = value::OBJ;
      ++m_p;
      ws();
      if (m_p < m_e && *m_p == '}') { ++m_p; return v; }
      for (;;) {
        ws();
        if (m_p == m_e || *m_p != '"') fail("expected key");
        std::string k = str();
        ws();
        if (m_p == m_e || *m_p != ':') fail("expected ':'");
        ++m_p;
        v.obj.emplace_back(k, any());
        ws();
        if (m_p < m_e && *m_p == ',') { ++m_p; continue; }
        if (m_p < m_e && *m_p == '}') { ++m_p; return v; }
        fail("expected ',' or '}'");
      }
    }
    if (*m_p == '[') {
      v.kind = value::ARR;
      ++m_p;
      ws();
      if (m_p < m_e && *m_p == ']') { ++m_p; return v; }
      for (;;) {
        v.arr.push_back(any());
        ws();
        if (m_p < m_e && *m_p == ',') { ++m_p; continue; }
        if (m_p < m_e && *m_p == ']') { ++m_p; return v; }
        fail("expected ',' or ']'");
      }
    }
    if (*m_p == '"') {
      v.kind = value::STR;
      v.text = str();
      return v;
    }
    const char* s = m_p;
    while (m_p < m_e && *m_p != ',' && *m_p != '}' && *m_p != ']' && *m_p != ' ' && *m_p != '\n') ++m_p;
    v.text.assign(s, m_p);
    if (v.text == "null") v.kind = value::NUL;
    else if (v.text == "true" || v.text == "false") v.kind = value::BOOL;
    else v.kind = value::NUM;
    return v;
  }
};

struct writer_ctx;
struct reader_ctx;

template <class T, class = void> struct has_transfer : std::false_type {};
template <class T>
struct has_transfer<T, decltype(std::declval<T&>().ref_stub_transfer(std::declval<writer_ctx&>()), void())> : std::true_type {};

inline void write_string(std::string& out, const std::string& s) {
  out += '"';
  for (char c : s) {
    if (c == '"' || c == '\\') { out += '\\'; out += c; }
    else if (c == '\n') out += "\\n";
    else if (c == '\t') out += "\\t";
    else if (c == '\r') out += "\\r";
    else out += c;
  }
  out += '"';
}

template <class T> void write_value(std::string& out, const T& v);
template <class T> void write_value(std::string& out, const std::vector<T>& v);
template <class V> void write_value(std::string& out, const std::map<std::string, V>& m);
template <class T> void read_value(const value& j, T& v);
template <class T> void read_value(const value& j, std::vector<T>& v);
template <class V> void read_value(const value& j, std::map<std::string, V>& m);

struct writer_ctx {
  std::string& out;
  bool first = true;
  bool is_serialize() const { return true; }
  bool is_human_readable() const { return true; }
  size_t get_version() const { return 0; }
  template <class T>
  void field(const char* name, T& v) {
    if (!first) out += ',';
    first = false;
    write_string(out, name);
    out += ':';
    write_value(out, v);
  }
};
struct reader_ctx {
  const value& obj;
  bool is_serialize() const { return false; }
  bool is_human_readable() const { return true; }
  size_t get_version() const { return 0; }
  template <class T>
  void field(const char* name, T& v) {
    const value* j = obj.find(name);
    if (j && j->kind != value::NUL) read_value(*j, v);
  }
};

template <class T>
void write_value(std::string& out, const T& v) {
  if constexpr (std::is_same<T, bool>::value) {
    out += v ? "true" : "false";
  } else if constexpr (std::is_enum<T>::value) {
    out += std::to_string((long long)v);
  } else if constexpr (std::is_arithmetic<T>::value) {
    out += std::to_string(v);
  } else if constexpr (std::is_same<T, std::string>::value) {
    write_string(out, v);
  } else if constexpr (has_transfer<T>::value) {
    out += '{';
    writer_ctx ctx{out};
    const_cast<T&>(v).ref_stub_transfer(ctx);
    out += '}';
  } else {
    static_assert(sizeof(T) == 0, "oracle/ref_stubs: this field type is not supported by the JSON stand-in");
  }
}
template <class T>
void write_value(std::string& out, const std::vector<T>& v) {
  out += '[';
  for (size_t i = 0; i < v.size(); ++i) {
    if (i) out += ',';
    write_value(out, v[i]);
  }
  out += ']';
}
template <class V>
void write_value(std::string& out, const std::map<std::string, V>& m) {
  out += '{';
  bool first = true;
  for (const auto& kv : m) {
    if (!first) out += ',';
    first = false;
    write_string(out, kv.first);
    out += ':';
    write_value(out, kv.second);
  }
  out += '}';
}

template <class T>
void read_value(const value& j, T& v) {
  if constexpr (std::is_same<T, bool>::value) {
    v = (j.text == "true");
  } else if constexpr (std::is_enum<T>::value) {
    v = T(std::strtoll(j.text.c_str(), nullptr, 10));
  } else if constexpr (std::is_floating_point<T>::value) {
    v = T(std::strtod(j.text.c_str(), nullptr));
  } else if constexpr (std::is_signed<T>::value && std::is_integral<T>::value) {
    v = T(std::strtoll(j.text.c_str(), nullptr, 10));
  } else if constexpr (std::is_integral<T>::value) {
    v = T(std::strtoull(j.text.c_str(), nullptr, 10));
  } else if constexpr (std::is_same<T, std::string>::value) {
    v = j.text;
  } else if constexpr (has_transfer<T>::value) {
    reader_ctx ctx{j};
    v.ref_stub_transfer(ctx);
  } else {
    static_assert(sizeof(T) == 0, "oracle/ref_stubs: this field type is not supported by the JSON stand-in");
  }
}
template <class T>
void read_value(const value& j, std::vector<T>& v) {
  v.clear();
  for (const auto& e : j.arr) {
    v.emplace_back();
    read_value(e, v.back());
  }
}
template <class V>
void read_value(const value& j, std::map<std::string, V>& m) {
  m.clear();
  for (const auto& kv : j.obj) read_value(kv.second, m[kv.first]);
}

}  // namespace ref_stub_json

#define TRANSFER_OBJECT template <class ref_stub_ctx_> void ref_stub_transfer(ref_stub_ctx_& _ctx)
#define VERSION(v) (void)_ctx
#define FIELD(name, ...) _ctx.field(#name, name)
#define FIELD_SPECIAL(...) (void)0
#define OBSOLETE_FIELD(...) (void)0
#define IS_SERIALIZE (_ctx.is_serialize())
#define IS_DESERIALIZE (!_ctx.is_serialize())
#define IS_HUMAN_READABLE (_ctx.is_human_readable())
#define GET_VERSION(v) (_ctx.get_version())
#define BASE_TYPE(native_type, transfer_type)
#define SET_TYPE_ID(type, id)
#define TF_STRICT 1
#define TF_ALLOW_NULL 2
#define TF_NO_DEFAULT 4

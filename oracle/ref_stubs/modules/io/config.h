// Shadows modules/io/config.h (test infrastructure).  The reference keeps its settings as json_spirit values read
// from a JSON file; neither exists in oracle/_ref.  The sources compiled there only ever ask for a few path settings
// through CONF_S(...), so the stand-in is a process-wide table of strings that oracle/ref_shim.cpp fills in
// ("temp_root", "resources_root", "path_bulkdata") before anything runs.
#pragma once
#include <functional>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <type_traits>
#include "modules/io/json_transfer.h"
#include "modules/io/utils.h"

#define CONF_T(type, param) Config::instance().get<type>(#param)
#define CONF_S(param) CONF_T(std::string, param)
#define CONF_CS(param) CONF_S(param).c_str()
#define CONF(param) Proxy(#param)

struct unknown_key_exception : io_exception {
  explicit unknown_key_exception(const std::string& key) : io_exception("unknown config key: " + key) {}
};

namespace ref_stub_config {
inline std::mutex& mu() { static std::mutex m; return m; }
inline std::map<std::string, std::string>& table() { static std::map<std::string, std::string> t; return t; }
inline bool lookup(const std::string& key, std::string* out) {
  std::lock_guard<std::mutex> l(mu());
  auto it = table().find(key);
  if (it == table().end()) return false;
  *out = it->second;
  return true;
}
template <class T>
T parse(const std::string& text) {
  if constexpr (std::is_same<T, std::string>::value) {
    return text;
  } else {
    T v{};
    std::istringstream(text) >> v;
    return v;
  }
}
}  // namespace ref_stub_config

struct Config {
  static Config& instance() { static Config the_one; return the_one; }
  template <class T>
  static void set(const std::string& key, const T& v) {
    std::ostringstream text;
    text << v;
    std::lock_guard<std::mutex> l(ref_stub_config::mu());
    ref_stub_config::table()[key] = text.str();
  }
  template <class T>
  T get(const std::string& key) {
    std::string text;
    if (!ref_stub_config::lookup(key, &text)) throw unknown_key_exception(key);
    return ref_stub_config::parse<T>(text);
  }
  template <class T>
  T get(const std::string& key, const T& fallback) {
    std::string text;
    return ref_stub_config::lookup(key, &text) ? ref_stub_config::parse<T>(text) : fallback;
  }
};

struct Proxy {
  std::string key;
  Proxy(const std::string& k) : key(k) {}
  template <class T>
  operator T() { return Config::instance().get<T>(key); }
};

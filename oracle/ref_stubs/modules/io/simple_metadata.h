// Shadows modules/io/simple_metadata.h (test infrastructure): a key/value sink for run statistics whose values are
// json_spirit values in the reference; the importers compiled into oracle/_ref only ever push numbers into it.
#pragma once
#include <string>
#include "modules/io/json_transfer.h"
class simple_metadata {
 public:
  virtual ~simple_metadata() = default;
  virtual void set_simple_text(const std::string& key, const std::string& json_text) { (void)key; (void)json_text; }
  template <class V>
  void set_simple(const std::string& key, const V& value) { set_simple_text(key, json_serialize(value)); }
};
inline simple_metadata& discard_simple_metadata() {
  static simple_metadata sink;
  return sink;
}

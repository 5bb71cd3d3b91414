// Shadows modules/io/json_transfer.h (test infrastructure): json_serialize / json_deserialize over the small JSON
// writer / reader of the transfer_object.h stand-in (no json_spirit in this image).
#pragma once
#include <string>
#include "modules/io/log.h"
#include "modules/io/transfer_object.h"
template <class T>
std::string json_serialize(const T& obj, bool = false) {
  std::string out;
  ref_stub_json::write_value(out, obj);
  return out;
}
template <class T>
void json_deserialize(T& obj, const std::string& text) {
  ref_stub_json::value j = ref_stub_json::parser(text).parse();
  ref_stub_json::read_value(j, obj);
}
template <class T>
T inline_json_deserialize(const std::string& text) {
  T obj;
  json_deserialize(obj, text);
  return obj;
}

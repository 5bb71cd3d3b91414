// Shadows modules/mapred/manifest.h (test infrastructure): the map-reduce file catalogue is outside the hot path;
// the sources compiled into oracle/_ref only name these types in signatures and members they never exercise.
#pragma once
#include <string>
#include "base/base.h"
#include "modules/io/keyvalue.h"
#include "modules/io/transfer_object.h"
class path {
 public:
  path() = default;
  explicit path(const std::string& s) : m_s(s) {}
  std::string bare_path() const { return m_s; }
 private:
  std::string m_s;
};
class manifest {
 public:
  TRANSFER_OBJECT { VERSION(0); }
  size_t get_num_records() const { return 0; }
  size_t get_size() const { return 0; }
};

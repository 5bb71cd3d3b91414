// Shadows modules/mapred/task.h (test infrastructure): the map-reduce task runtime is outside the hot path; the
// value types of bio_mapred/read_correction.h only derive from / name these.
#pragma once
#include <string>
#include "modules/mapred/manifest.h"
typedef std::string subtask_id;
template <class Derived>
class task_impl {
 public:
  virtual ~task_impl() = default;
};

// Shadows modules/mapred/manifest_parallel.h (test infrastructure): the records of the manifest stand-in are in
// memory (one "file"), so every record is handed to the functor in order with file_info_id 0 and its record number
// (serially: the functor keeps per-file counters, and there is one file).
#pragma once
#include <utility>
#include <vector>
#include "modules/io/parallel.h"
#include "modules/io/progress.h"
#include "modules/mapred/manifest.h"
template <typename Function, typename KeyType, typename ValueType>
inline Function manifest_parallelize(manifest the_manifest, Function f,
                                     progress_handler_t progress = null_progress_handler) {
  typedef std::vector<std::pair<KeyType, ValueType>> records_t;
  const records_t* records = static_cast<const records_t*>(the_manifest.ref_stub_records());
  if (records) {
    for (size_t i = 0; i < records->size(); ++i) f((*records)[i].first, (*records)[i].second, 0, i);
  }
  progress(1.0);
  return f;
}

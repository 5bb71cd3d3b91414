// Shadows modules/mapred/manifest.h (test infrastructure): the map-reduce file catalogue is outside the hot path.
// The sources compiled into oracle/_ref name these types in signatures and members; the one place that reads
// records through a manifest (make_readmap::import_reads_from) gets them from memory: a manifest here is a typed
// list of (key, value) records that oracle/ref_shim.cpp fills in, visited by the manifest_parallelize stand-in.
#pragma once
#include <memory>
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
  size_t get_num_records() const { return m_num_records; }
  size_t count_file_infos() const { return 1; }
  size_t get_size() const { return 0; }
  // in-memory records: a std::vector<std::pair<Key, Value>> behind a type-erased pointer
  template <class Records>
  void ref_stub_set_records(std::shared_ptr<Records> r) {
    m_num_records = r->size();
    m_records = r;
  }
  const void* ref_stub_records() const { return m_records.get(); }
 private:
  std::shared_ptr<const void> m_records;
  size_t m_num_records = 0;
};
class manifest_reader : public readable {
 public:
  explicit manifest_reader(const manifest&) {}
  size_t read(char*, size_t) override { return 0; }
};

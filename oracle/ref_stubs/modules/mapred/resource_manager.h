// Shadows modules/mapred/resource_manager.h (test infrastructure).  In the reference this class moves big tables
// between scratch files and the map-reduce blob store.  oracle/_ref only needs the scratch half: kmer_set asks it
// for two writable mappings while it is built.  They are backed by unlinked files under CONF_S(resources_root), as
// the reference's are by named ones (mapred/resource_manager.cpp:52-68); the blob-store half reports that it is absent.
#pragma once
#include <atomic>
#include <stdexcept>
#include <string>
#include "modules/io/config.h"
#include "modules/io/mmap_buffer.h"
#include "modules/io/progress.h"
#include "modules/mapred/manifest.h"

namespace ref_stub_resources {
inline std::string next_scratch_name() {
  static std::atomic<unsigned long> serial{0};
  return CONF_S(resources_root) + "/scratch-" + std::to_string(::getpid()) + "-" + std::to_string(serial++);
}
[[noreturn]] inline void no_blob_store(const char* what) {
  throw std::logic_error(std::string("oracle/_ref has no blob store: resource_manager::") + what);
}
}  // namespace ref_stub_resources

struct resource_manager {
  explicit resource_manager(bool /*direct*/ = false) {}

  void create_resource(mmap_buffer& mapping, size_t bytes) {
    const std::string name = ref_stub_resources::next_scratch_name();
    mapping.open(name, bytes);
    ::unlink(name.c_str());  // the mapping keeps the pages; nothing is left behind
  }

  template <class... A> void write_resource(A&&...) { ref_stub_resources::no_blob_store("write_resource"); }
  template <class... A> void read_resource(A&&...) { ref_stub_resources::no_blob_store("read_resource"); }
};

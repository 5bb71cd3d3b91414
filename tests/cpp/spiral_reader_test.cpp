// spiral_reader_test.cpp -- reads a seqset (or any) spiral file through the facade's spiral_file_reader /
// seqset_file and prints what it found as JSON: the member list, and for a seqset the decoded entry
// sizes and prev bits as FNV-1a digests (tests/test_spiral_reader.py compares them with Python's view).
//   spiral_reader_test members <file>      spiral_reader_test seqset <file>
#include <cstdio>
#include <string>

#include "bgx_build_seqset.hpp"

static uint64_t fnv(const void* d, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* p = static_cast<const unsigned char*>(d);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  try {
    const std::string mode = argv[1];
    if (mode == "members") {
      bgx_bs::spiral_file_reader r(argv[2]);
      printf("[");
      bool first = true;
      for (const auto& m : r.members()) {
        const std::string d = m.size <= (1u << 20) ? r.read(m.name) : std::string();   // big members: framing only
        printf("%s{\"name\":\"%s\",\"size\":%llu,\"offset\":%llu,\"fnv\":\"%016llx\"}", first ? "" : ",", m.name.c_str(),
               (unsigned long long)m.size, (unsigned long long)m.offset, (unsigned long long)fnv(d.data(), d.size()));
        first = false;
      }
      printf("]\n");
    } else {
      bgx_bs::seqset_file f(argv[2]);
      const bgx_seqset_part p = f.part();
      printf("{\"n\":%llu,\"uuid\":\"%s\",\"max_read_len\":%u,\"sizes\":\"%016llx\"", (unsigned long long)f.size(), f.uuid().c_str(),
             f.max_read_len(), (unsigned long long)fnv(p.sizes, f.size() * 2));
      for (int b = 0; b < 4; ++b)
        printf(",\"prev%d\":\"%016llx\"", b, (unsigned long long)fnv(p.prev_bits[b], (f.size() + 63) / 64 * 8));
      printf("}\n");
    }
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}

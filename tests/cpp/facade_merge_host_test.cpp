// facade_merge_host_test.cpp -- bgx_bs::seqset_merger's host side against the mock C ABI (tests/cpp/mock_bgx.cpp):
//   facade_merge_host_test <seqset a> <seqset b> <old readmap> <out dir>
// writes <out>/0.mergemap, <out>/1.mergemap, <out>/migrated.readmap and prints a JSON summary.
#include <cstdio>

#include "bgx_build_seqset.hpp"

int main(int argc, char** argv) {
  if (argc != 5) return 2;
  try {
    bgx_bs::session s;
    bgx_bs::seqset_file a(argv[1]), b(argv[2]);
    bgx_bs::seqset_merger m(s, {&a, &b});
    bool threw = false;
    try { m.total_merged_entries(); } catch (const bgx_bs::io_exception&) { threw = true; }   // build() first
    m.build();
    const std::string out = argv[4];
    m.write_mergemap(0, out + "/0.mergemap", "merged-uuid");
    m.write_mergemap(1, out + "/1.mergemap", "merged-uuid");
    m.fast_migrate(0, argv[3], out + "/migrated.readmap", "merged-uuid");
    const auto mm = m.fill_mergemap(1);
    const auto flat = m.flat_entries(0, 3, 4);
    bool no_such = false;
    try { m.fill_mergemap(7); } catch (const bgx_bs::io_exception& e) { no_such = std::string(e.what()).find("no such input") != std::string::npos; }
    printf("{\"need_build\":%s,\"total\":%zu,\"n_bits\":%llu,\"n_set\":%llu,\"flat\":[\"%s\",\"%s\",\"%s\",\"%s\"],\"no_such\":%s}\n", threw ? "true" : "false",
           m.total_merged_entries(), (unsigned long long)mm.n_bits, (unsigned long long)mm.n_set, flat[0].c_str(), flat[1].c_str(), flat[2].c_str(),
           flat[3].c_str(), no_such ? "true" : "false");
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}

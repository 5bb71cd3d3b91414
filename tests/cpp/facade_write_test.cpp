// facade_write_test.cpp -- bgx_bs::builder::make_seqset (the facade's seqset spiral-file writer) over whatever the C
// ABI it is linked with serves: with tests/cpp/mock_tables_bgx.cpp, tables prepared by the test.  usage: <out path> <uuid>  |  readmap <out path> <seqset uuid> <is_paired> <max_read_len>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "bgx_build_seqset.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    bgx_bs::count_kmer_options ko;
    bgx_bs::session s(ko);
    if (std::string(argv[1]) == "readmap") {
      if (argc < 6) return 2;
      auto t = bgx_bs::make_readmap::do_make(argv[2], s, argv[3], atoi(argv[4]) != 0, (unsigned)atoi(argv[5]));
      printf("wrote %llu rows\n", (unsigned long long)t.n_rows);
      return 0;
    }
    bgx_bs::builder b(s);
    bgx_bs::seqset_tables t = b.make_seqset(argv[1], bgx_bs::null_progress_handler, argv[2]);
    printf("wrote %llu entries\n", (unsigned long long)t.num_entries);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}

// bgx_bs::multi_session: ONE process drives N GPUs (one host thread per device) through the facade;
// the sharded build must give the tables of a single-GPU build over the same reads.
// usage: multi_session_test <reads.txt (one read per line)> <n_gpus>
#include <fstream>
#include <iostream>

#include "bgx_build_seqset.hpp"

namespace bs = bgx_bs;

static void run_stages(bs::session& s, const std::vector<std::string>& reads, size_t lo, size_t hi, bs::seqset_tables* out) {
  bs::kmer_counter kc(s);
  {
    bs::kmer_counter::prob_pass_processor p(kc);
    for (size_t i = lo; i < hi; ++i) p.add(reads[i]);
    p.flush_all();
  }
  std::unique_ptr<bs::kmer_set> ks = bs::run_kmerize_subtask(&kc);
  bs::correct_reads cr(s, *ks, bs::read_correction_params());
  cr.correct_all();
  bs::builder b(s);
  b.build_chunks();
  *out = b.tables();
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  std::vector<std::string> reads;
  {
    std::ifstream in(argv[1]);
    std::string line;
    while (std::getline(in, line))
      if (!line.empty()) reads.push_back(line);
  }
  const int n = atoi(argv[2]);
  try {
    bs::seqset_tables whole;
    {
      bs::session s;
      run_stages(s, reads, 0, reads.size(), &whole);
    }
    bs::multi_session ms(n);
    std::vector<bs::seqset_tables> part((size_t)n);
    ms.parallel([&](int r, bs::session& s) {
      run_stages(s, reads, reads.size() * r / n, reads.size() * (r + 1) / n, &part[(size_t)r]);
    });
    uint64_t total = 0;
    bool ok = true;
    std::vector<uint64_t> bits[4];
    for (int r = 0; r < n; ++r) {
      total += part[(size_t)r].num_entries;
      for (int b = 0; b < 5; ++b) ok = ok && part[(size_t)r].fixed[b] == whole.fixed[b];
      for (int b = 0; b < 4; ++b) bits[b].insert(bits[b].end(), part[(size_t)r].prev_bits[b].p, part[(size_t)r].prev_bits[b].p + part[(size_t)r].prev_bits[b].n);
    }
    ok = ok && total == whole.num_entries;
    for (int b = 0; b < 4; ++b) {
      ok = ok && bits[b].size() == whole.prev_bits[b].n;
      for (size_t i = 0; ok && i < bits[b].size(); ++i) ok = bits[b][i] == whole.prev_bits[b].p[i];
    }
    std::cout << "{\"gpus\":" << n << ",\"entries\":" << total << ",\"expected\":" << whole.num_entries << ",\"ok\":" << (ok ? "true" : "false") << "}\n";
    return ok ? 0 : 1;
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  }
}

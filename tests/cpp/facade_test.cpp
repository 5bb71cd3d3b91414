// facade_test.cpp -- exercises include/bgx_build_seqset.hpp the way SEQSETMain drives the reference
// stages, on the reference's own known answers:
//   * bs/builder_test.cpp:52-122 (seqset_for_reads): entry counts 129 / 152 / 91 / 91 / 89 / 99
//   * a full create flow (import -> kmerization -> read_correction -> make_seqset) on reads given
//     on the command line as a text file (one read per line), writing <out>/seqset
// Prints one JSON line per case; tests/test_facade.py checks them and compares the written
// seqset members with the oracle.  Build: see tests/test_facade.py.
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "bgx_build_seqset.hpp"

namespace bs = bgx_bs;

// modules/bio_base/dna_testutil.cpp:14-33: each char -> 'C' + 8 bits LSB-first (T=1, A=0) + 'C'
static std::string tseq(const std::string& s) {
  std::string out;
  for (unsigned char ch : s) {
    out += 'C';
    for (int i = 0; i < 8; ++i) out += (ch & (1 << i)) ? 'T' : 'A';
    out += 'C';
  }
  return out;
}
static std::string rc(const std::string& s) {
  std::string o(s.rbegin(), s.rend());
  for (char& c : o) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
  return o;
}

int main(int argc, char** argv) {
  try {
    struct Case { std::vector<std::string> reads; size_t size; };
    std::vector<Case> cases = {
        {{tseq("abcdefg")}, 129},
        {{tseq("abcd"), tseq("cdef"), rc(tseq("efgh"))}, 152},
        {{tseq("ab"), tseq("bc"), tseq("cd"), tseq("be")}, 91},
        {{tseq("AB"), tseq("BC"), tseq("CD"), tseq("BE")}, 91},
        {{tseq("abc"), tseq("cde")}, 89},
        {{tseq("abc"), tseq("efg")}, 99},
    };
    int bad = 0;
    for (size_t i = 0; i < cases.size(); ++i) {
      bs::seqset_tables t = bs::seqset_for_reads(cases[i].reads);
      bool ok = t.num_entries == cases[i].size && t.fixed[4] == t.num_entries;
      bad += !ok;
      std::cout << "{\"case\":\"builder_test_" << i << "\",\"entries\":" << t.num_entries << ",\"expected\":" << cases[i].size
                << ",\"ok\":" << (ok ? "true" : "false") << "}" << std::endl;
    }
    if (argc >= 3) {
      // the create flow of SEQSETMain::run with the reference's stage objects
      std::ifstream in(argv[1]);
      std::string line;
      bs::count_kmer_options ko;            // --kmer-size 30 --min-kmer-count 5
      bs::read_correction_params rp;        // --trim-after-portion 0.7 --max-corrections 8 --min-good-run 2
      bs::session s(ko, rp);
      bs::kmer_counter counter(s);
      counter.start_prob_pass();
      size_t n_in = 0;
      {
        bs::kmer_counter::prob_pass_processor p(counter);   // read_importer_state::process (:119-151)
        while (std::getline(in, line))
          if (!line.empty()) { p.add(line); ++n_in; }
      }
      counter.close_prob_pass();
      std::unique_ptr<bs::kmer_set> ks = bs::run_kmerize_subtask(&counter);          // :692
      bs::correct_reads cr(s, *ks, rp);                                                // :835-912
      cr.add_initial_repo();
      cr.correct_all();
      size_t kept = 0, bases = 0;
      bs::corrected_read out;
      for (size_t i = 0; i < cr.size(); ++i)
        if (cr.correct(i, out)) { ++kept; bases += out.corrected.size(); }
      bs::expander expand(s, false);                                                   // :921-931
      size_t r1 = expand.sort_and_dedup("", "initial", "init_sorted", "", 0, 0);
      expand.expand("init_sorted", "init_expanded", 7, 255);
      expand.sort_and_dedup("init_sorted", "init_expanded", "pass2_sorted", "pass2_expanded", 1, 6);
      size_t fin = expand.sort_and_dedup("pass2_sorted", "pass2_expanded", "complete", "", 0, 0);
      bs::builder b(s);
      bs::seqset_tables t = b.make_seqset(std::string(argv[2]) + "/seqset");           // :944-947
      // make_readmap::do_make (biograph_create.cpp:818-831), unpaired
      bs::make_readmap::tables rm = bs::make_readmap::do_make(std::string(argv[2]) + "/readmap", s, "test-uuid", false, 35);
      std::cout << "{\"case\":\"readmap\",\"rows\":" << rm.n_rows << ",\"entries\":" << rm.n_entries << "}" << std::endl;
      std::cout << "{\"case\":\"create\",\"reads\":" << n_in << ",\"kmers\":" << ks->size() << ",\"corrected_reads\":" << kept
                << ",\"corrected_bases\":" << bases << ",\"round1\":" << r1 << ",\"entries\":" << fin
                << ",\"written_entries\":" << t.num_entries << "}" << std::endl;
    }
    return bad ? 1 : 0;
  } catch (const std::exception& e) {
    std::cerr << "facade_test: " << e.what() << std::endl;
    return 2;
  }
}

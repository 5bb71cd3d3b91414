// merge_core_test.cpp -- runs the per-element bodies of the seqset-merge kernels
// (biograph_b200/csrc/merge_core.cuh) SERIALLY ON THE CPU, in the order merge.cu / seqset.cu launch them,
// so that their logic is checked against the merge oracle in the GPU-less dev container
// (tests/test_merge_core_cpu.py).  Test infrastructure: the product only ever calls these functions from
// CUDA kernels.  The device-wide steps between them (scan, radix sort, compaction) are stood in for by
// std:: algorithms with the same contract.
//
//   merge_core_test <in.bin> <out.bin>
// in : u64 n_parts, u64 nsplits; per part: u64 n, u16 sizes[n] (padded to 8 bytes), u64 prev[4][ceil(n/64)],
//      u64 old_bits[ceil(n/64)] (a bit vector over the part's entries to migrate)
// out: u64 n_merged, u16 sizes[] (padded), u16 shared[] (padded), u64 prev[4][words], u64 missing,
//      per part: u64 mergemap[words]; per part: u64 migrated[words];
//      per part: flat entries as ASCII, '\n' after each
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../biograph_b200/csrc/merge_core.cuh"

using namespace bgx;
using namespace bgx::mergecore;

static std::vector<unsigned char> slurp(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  std::vector<unsigned char> d;
  unsigned char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
  fclose(f);
  return d;
}

struct Part {
  uint64_t n;
  std::vector<uint16_t> sizes;
  std::vector<uint64_t> prev[4], old_bits;
};

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: merge_core_test in.bin out.bin\n"); return 2; }
  std::vector<unsigned char> in = slurp(argv[1]);
  size_t off = 0;
  auto u64 = [&]() { uint64_t v; memcpy(&v, &in[off], 8); off += 8; return v; };
  const uint64_t n_parts = u64(), nsplits = u64();
  std::vector<Part> parts(n_parts);
  uint64_t N = 0;
  for (Part& p : parts) {
    p.n = u64();
    p.sizes.resize(p.n);
    memcpy(p.sizes.data(), &in[off], p.n * 2);
    off += (p.n * 2 + 7) / 8 * 8;
    const uint64_t words = (p.n + 63) / 64;
    for (int b = 0; b < 4; ++b) {
      p.prev[b].resize(words);
      memcpy(p.prev[b].data(), &in[off], words * 8);
      off += words * 8;
    }
    p.old_bits.resize(words);
    memcpy(p.old_bits.data(), &in[off], words * 8);
    off += words * 8;
    N += p.n;
  }

  // ---- stage_merge_seqsets steps 1-2 (merge.cu) ---------------------------------------------------------
  PartTable pt;
  pt.n = (int)n_parts;
  pt.word_base[0] = 0;
  std::vector<std::vector<uint32_t>> woff(n_parts);
  for (uint64_t p = 0; p < n_parts; ++p) {
    woff[p].resize(parts[p].n);
    uint32_t acc = 0;
    for (uint64_t i = 0; i < parts[p].n; ++i) {  // entry_words_kernel + exclusive scan
      woff[p][i] = acc;
      acc += entry_words(parts[p].sizes[i]);
    }
    pt.word_base[p + 1] = pt.word_base[p] + acc;
  }
  const uint64_t total_words = pt.word_base[n_parts];
  std::vector<uint64_t> store(total_words + 2, 0xdeadbeefdeadbeefULL);  // emit_entry must write every word it owns
  store[total_words] = store[total_words + 1] = 0;
  std::vector<uint64_t> keys(N), locs(N);
  std::vector<std::vector<uint64_t>> flat_loc(n_parts);
  uint64_t rec_base = 0;
  for (uint64_t p = 0; p < n_parts; ++p) {
    const Part& P = parts[p];
    const uint64_t n = P.n, words = (n + 63) / 64;
    uint64_t fixed[5] = {0, 0, 0, 0, 0};
    for (int b = 0; b < 4; ++b) {
      uint64_t c = 0;
      for (uint64_t w = 0; w < words; ++w) c += popc64(masked_word(P.prev[b].data(), w, n));
      fixed[b + 1] = fixed[b] + c;
    }
    if (fixed[4] != n) { fprintf(stderr, "Invalid seqset: prev bit totals != entries\n"); return 3; }
    std::vector<uint32_t> next(n), next_alt(n);
    for (int b = 0; b < 4; ++b) {  // select_table(bits_b, n, next, fixed[b])
      uint32_t excl = 0;
      for (uint64_t w = 0; w < words; ++w) {
        scatter_set_bits(P.prev[b].data(), w, n, excl, next.data(), fixed[b]);
        excl += popc64(masked_word(P.prev[b].data(), w, n));
      }
    }
    std::vector<uint64_t> wa(n), wb(n);
    for (uint64_t i = 0; i < n; ++i) double_init(i, fixed[1], fixed[2], fixed[3], wa.data());
    uint64_t* w_in = wa.data(); uint64_t* w_out = wb.data();
    uint32_t* j_in = next.data(); uint32_t* j_out = next_alt.data();
    for (int have = 1; have < 32; have <<= 1) {
      for (uint64_t i = 0; i < n; ++i) double_step(w_in, j_in, have, i, w_out, j_out);
      std::swap(w_in, w_out);
      std::swap(j_in, j_out);
    }
    for (uint64_t i = 0; i < n; ++i)
      emit_entry(w_in, j_in, P.sizes.data(), woff[p].data(), i, pt.word_base[p], rec_base, store.data(), keys.data(), locs.data());
    flat_loc[p].assign(locs.begin() + rec_base, locs.begin() + rec_base + n);
    rec_base += n;
  }
  for (uint64_t w = 0; w < total_words; ++w)
    if (store[w] == 0xdeadbeefdeadbeefULL) { fprintf(stderr, "store word %llu never written\n", (unsigned long long)w); return 3; }

  // ---- build_seqset_from_records (seqset.cu), stood in for by std:: ---------------------------------------------
  const uint64_t* st = store.data();
  std::vector<uint32_t> order(N);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    if (keys[a] != keys[b]) return keys[a] < keys[b];   // the radix key decides first, as on the device
    int lcp;
    return compare_seq(st, loc_addr(locs[a]), (int)loc_len(locs[a]), loc_addr(locs[b]), (int)loc_len(locs[b]), &lcp) < 0;
  });
  std::vector<uint64_t> skeys(N), slocs(N);
  for (uint64_t j = 0; j < N; ++j) { skeys[j] = keys[order[j]]; slocs[j] = locs[order[j]]; }
  std::vector<uint32_t> keep(N), pos(N);
  for (uint64_t j = 0; j < N; ++j)  // dedup_flag_kernel: dropped when a prefix of / equal to the successor
    keep[j] = !(j + 1 < N && prefixes(st, loc_addr(slocs[j]), (int)loc_len(slocs[j]), slocs[j + 1]));
  uint32_t n_kept = 0;
  for (uint64_t j = 0; j < N; ++j) { pos[j] = n_kept; n_kept += keep[j]; }
  const uint64_t mm_words = ((uint64_t)n_kept + 63) / 64;
  std::vector<unsigned long long> mm(mm_words * n_parts, 0);
  for (uint64_t j = 0; j < N; ++j) mergemap_mark(slocs.data(), pos.data(), (uint32_t)j, pt, mm.data(), mm_words);
  std::vector<uint64_t> ekeys(n_kept), elocs(n_kept);
  for (uint64_t j = 0; j < N; ++j)
    if (keep[j]) { ekeys[pos[j]] = skeys[j]; elocs[pos[j]] = slocs[j]; }
  const uint32_t n = n_kept;
  std::vector<uint16_t> sizes(n), shared(n);
  const uint64_t pw = ((uint64_t)n + 63) / 64;
  std::vector<unsigned long long> prev(4 * pw, 0);
  uint64_t missing = 0;
  for (uint32_t i = 0; i < n; ++i) {
    sizes[i] = (uint16_t)loc_len(elocs[i]);
    int lcp = 0;
    if (i) compare_seq(st, loc_addr(elocs[i - 1]), (int)loc_len(elocs[i - 1]), loc_addr(elocs[i]), (int)loc_len(elocs[i]), &lcp);
    shared[i] = (uint16_t)lcp;
    const uint32_t t = merge_prev_target(st, elocs.data(), n, i, 0, n, nsplits);   // merge_prev_kernel
    if (t == kNone) { missing = 1; continue; }
    or_bit(prev.data() + (uint64_t)(ekeys[i] >> 62) * pw, t);
  }

  // ---- migrate_bits (merge.cu) -------------------------------------------------------------------------------------
  std::vector<std::vector<unsigned long long>> migrated(n_parts);
  for (uint64_t p = 0; p < n_parts; ++p) {
    std::vector<uint32_t> sel(parts[p].n);
    uint32_t excl = 0;
    const uint64_t* bits = reinterpret_cast<const uint64_t*>(mm.data() + p * mm_words);
    for (uint64_t w = 0; w < mm_words; ++w) {
      const int c = popc64(masked_word(bits, w, n));
      if ((uint64_t)excl + c > parts[p].n) { fprintf(stderr, "mergemap of part %llu has too many bits\n", (unsigned long long)p); return 3; }
      scatter_set_bits(bits, w, n, excl, sel.data(), 0);
      excl += c;
    }
    if (excl != parts[p].n) { fprintf(stderr, "mergemap bit total != entries of the input\n"); return 3; }
    migrated[p].assign(pw, 0);
    for (uint64_t w = 0; w < (parts[p].n + 63) / 64; ++w) migrate_word(parts[p].old_bits.data(), w, parts[p].n, sel.data(), migrated[p].data());
  }

  // ---- output ----------------------------------------------------------------------------------------------------------
  FILE* f = fopen(argv[2], "wb");
  if (!f) { perror(argv[2]); return 2; }
  auto w64 = [&](uint64_t v) { fwrite(&v, 8, 1, f); };
  auto pad16 = [&](const std::vector<uint16_t>& v) {
    fwrite(v.data(), 2, v.size(), f);
    const uint64_t z = 0;
    fwrite(&z, 1, ((v.size() * 2 + 7) / 8 * 8) - v.size() * 2, f);
  };
  w64(n);
  pad16(sizes);
  pad16(shared);
  fwrite(prev.data(), 8, prev.size(), f);
  w64(missing);
  fwrite(mm.data(), 8, mm.size(), f);
  for (uint64_t p = 0; p < n_parts; ++p) fwrite(migrated[p].data(), 8, migrated[p].size(), f);
  for (uint64_t p = 0; p < n_parts; ++p)
    for (uint64_t l : flat_loc[p]) {   // flat_ascii_kernel
      std::string s;
      for (uint32_t j = 0; j < loc_len(l); ++j) {
        const uint64_t a = loc_addr(l) + j;
        s.push_back("ACGT"[(store[a >> 5] >> (62 - 2 * (a & 31))) & 3]);
      }
      s.push_back('\n');
      fwrite(s.data(), 1, s.size(), f);
    }
  fclose(f);
  return 0;
}

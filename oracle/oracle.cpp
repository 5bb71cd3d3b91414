// oracle.cpp -- CPU restatement of BioGraph's seqset-construction path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under biograph_b200/ may include, link or call this
// file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs use it, and only as the checker / the timed CPU baseline.
//
// Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the
// reference's own golden fixture (golden/e_coli_10000snp.fq -> golden/e_coli_10000snp.bg/seqset,
// every payload member byte-for-byte), the builder_test / expand_test known answers and the
// fast_read_correct_test analytic cases (fixtures committed under tests/golden/); and
// tests/test_ref_vs_oracle.py checks it against the reference's OWN classes (oracle/_ref/libref.so,
// compiled from the sources under /root/reference, see ref_shim.cpp) on the same inputs: counts,
// flags, solid set, corrected reads, every seqset table and encoded member.
//
// Each function cites the reference file:line (relative to the reference checkout) it follows.
// "bs/" abbreviates modules/build_seqset/.
//
// Conventions: reads are ASCII over {A,C,G,T,N}, concatenated, with offs[n+1] giving the
// start of each read.  Base codes A=0 C=1 G=2 T=3 (modules/bio_base/dna_base.h:38-57).
// k-mers are uint64 with the first base in the high bits of the low 2k bits
// (modules/bio_base/kmer.h:30-38).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

inline int base_code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;  // 'N'
  }
}
const char kBaseChar[4] = {'A', 'C', 'G', 'T'};

// modules/bio_base/dna_sequence.cpp:330-351 (rev_comp) -- reverse the 2-bit groups of the
// low 2k bits and complement each base (complement of code c is 3-c).
inline uint64_t rev_comp_kmer(uint64_t kmer, int k) {
  uint64_t x = ~kmer;
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = __builtin_bswap64(x);
  return x >> (64 - 2 * k);
}

// modules/bio_base/dna_sequence.cpp:378-385 (canonicalize): canonical = min(kmer, rc);
// flipped = rc < kmer.
inline uint64_t canonicalize(uint64_t kmer, int k, bool* flipped) {
  uint64_t rc = rev_comp_kmer(kmer, k);
  if (rc < kmer) { *flipped = true; return rc; }
  *flipped = false;
  return kmer;
}

inline uint64_t kmer_mask(int k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1); }

void set_threads(int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
}

// ---------------------------------------------------------------------------------------
// k-mer enumeration: bs/kmer_counter.h:297-326 (pass_processor::add).  'N' resets the window
// and clears is_first; fwd_flag = first k-mer of the read, rev_flag = k-mer ending at the last
// base.  Calls f(kmer, is_first, is_last) for each k-mer instance.
template <typename F>
inline void for_each_kmer(const char* s, int64_t len, int k, F&& f) {
  bool is_first = true;
  int left = k;
  uint64_t kmer = 0;
  const uint64_t mask = kmer_mask(k);
  for (int64_t i = 0; i < len; ++i) {
    int c = base_code(s[i]);
    if (c < 0) { left = k; is_first = false; continue; }
    kmer = (kmer << 2) | (uint64_t)c;
    if (left) --left;
    if (!left) {
      f(kmer & mask, is_first, i == len - 1);
      is_first = false;
    }
  }
}

struct KmerSet {  // modules/bio_mapred/kmer_set.h: sorted canonical k-mers; index = rank
  const uint64_t* kmers;
  const uint8_t* flags;  // bit0 fwd_starts_read, bit1 rev_starts_read
  int64_t n;
  int k;
  // kmer_set's lookup table (modules/bio_mapred/kmer_set.cpp:401-481): lookup[p] = first index whose
  // leading prefix_bits bits are >= p; optional (nullptr: plain binary search over the whole set)
  const uint32_t* lookup = nullptr;
  int prefix_bits = 0;
  // kmer_set::find_table_index (modules/bio_mapred/kmer_set.cpp:296-360): prefix lookup, then binary
  // search of the tails in that range.
  int64_t find(uint64_t canon) const {
    const uint64_t *lo = kmers, *hi = kmers + n;
    if (lookup) {
      uint64_t pfx = canon >> (2 * k - prefix_bits);
      lo = kmers + lookup[pfx];
      hi = kmers + lookup[pfx + 1];
    }
    const uint64_t* p = std::lower_bound(lo, hi, canon);
    if (p == hi || *p != canon) return -1;
    return p - kmers;
  }
};

// the lookup table of a sorted set (empty when the set is too small to bother)
inline std::vector<uint32_t> build_kmer_lookup(const uint64_t* kmers, int64_t n, int k, int* prefix_bits) {
  int pb = 0;
  while (pb < 2 * k && pb < 26 && (1LL << pb) < n) ++pb;
  *prefix_bits = pb;
  std::vector<uint32_t> lk;
  if (pb == 0 || n >= (1LL << 32)) { *prefix_bits = 0; return lk; }
  lk.assign((size_t(1) << pb) + 1, 0);
  const int sh = 2 * k - pb;
  int64_t i = 0;
  for (uint64_t p = 0; p <= (uint64_t(1) << pb); ++p) {
    while (i < n && (kmers[i] >> sh) < p) ++i;
    lk[p] = (uint32_t)i;
  }
  return lk;
}

struct FrcKmer { bool flipped; int64_t index; };

struct FrcOut {
  std::string corrected;  // ASCII
  unsigned corrections = 0;
  std::vector<FrcKmer> kmers;
};

struct FrcParams {
  unsigned max_corrections, min_good_run;
  int k;
  const KmerSet* ks;
  // bs/correct_reads.cpp:163-171 (kmer_lookup_f): canonicalize then find_table_index.
  bool lookup(uint64_t kmer, FrcKmer* ki) const {
    bool fl;
    uint64_t canon = canonicalize(kmer, k, &fl);
    int64_t idx = ks->find(canon);
    if (idx < 0) return false;
    ki->flipped = fl; ki->index = idx;
    return true;
  }
};

inline uint64_t shift_in(uint64_t kmer, int k, int b) { return ((kmer << 2) | (uint64_t)b) & kmer_mask(k); }

// modules/bio_base/fast_read_correct.cpp:16-90 (correct_internal).  `kmer` is the k-mer of the
// bases up to but not including in[0].  Extends while the next k-mer is solid; on a miss (or
// 'N') requires min_run_here good bases so far and budget > 0, skips the bad base, and tries
// the four substitutions recursively; the longest continuation wins and a later base wins only
// if strictly longer.
void correct_internal(FrcOut* res, const char* in, int64_t n, uint64_t kmer, const FrcParams& p,
                      unsigned min_run_here, unsigned budget, bool require_run_at_end) {
  int64_t it = 0;
  FrcKmer ki;
  if (in[0] != 'N') {
    uint64_t nk = shift_in(kmer, p.k, base_code(in[it]));
    while (p.lookup(nk, &ki)) {
      res->corrected.push_back(in[it]);
      res->kmers.push_back(ki);
      ++it;
      if (it == n) return;
      kmer = nk;
      if (in[it] == 'N') break;
      nk = shift_in(kmer, p.k, base_code(in[it]));
    }
  }
  if (res->corrected.size() < min_run_here) return;
  if (budget == 0) return;
  ++it;  // skip the offending base
  FrcOut tries[4];
  unsigned best_size = 0;
  int best_b = -1;
  for (int b = 0; b < 4; ++b) {
    uint64_t tk = shift_in(kmer, p.k, b);
    if (!p.lookup(tk, &ki)) continue;
    tries[b].kmers.push_back(ki);
    if (it != n) {
      correct_internal(&tries[b], in + it, n - it, tk, p, p.min_good_run, budget - 1,
                       require_run_at_end);
    }
    if (require_run_at_end && tries[b].corrected.size() < p.min_good_run) continue;
    if (tries[b].corrected.size() >= best_size) {
      best_size = (unsigned)tries[b].corrected.size() + 1;
      best_b = b;
    }
  }
  if (best_size) {
    res->corrected.push_back(kBaseChar[best_b]);
    res->corrections += 1 + tries[best_b].corrections;
    res->corrected += tries[best_b].corrected;
    res->kmers.insert(res->kmers.end(), tries[best_b].kmers.begin(), tries[best_b].kmers.end());
  }
}

std::string revcomp_ascii(const std::string& s) {  // keeps 'N'
  std::string r(s.rbegin(), s.rend());
  for (char& c : r) {
    int b = base_code(c);
    if (b >= 0) c = kBaseChar[3 - b];
  }
  return r;
}

// modules/bio_base/fast_read_correct.cpp:94-182 (fast_read_correct).
FrcOut fast_read_correct(const char* in, int64_t n, const FrcParams& p) {
  FrcOut result;
  if (n < p.k) return FrcOut{};
  unsigned budget = p.max_corrections;
  int64_t it = 0;
  uint64_t kmer = 0;
  int left = p.k;
  FrcKmer ki;
  // scan right (N resets the window) to the first position whose trailing k-mer is solid
  while (left || !p.lookup(kmer, &ki)) {
    if (it == n) return FrcOut{};
    if (in[it] == 'N') { ++it; left = p.k; continue; }
    kmer = shift_in(kmer, p.k, base_code(in[it]));
    ++it;
    if (left) --left;
  }
  FrcOut right;
  if (it == p.k) {
    result.corrected.assign(in, p.k);
    if (it == n) { result.kmers.push_back(ki); return result; }
    right.kmers.push_back(ki);
  } else {
    int64_t kmer_start = it - p.k;
    std::string left_str = revcomp_ascii(std::string(in, kmer_start));
    FrcOut lc;
    correct_internal(&lc, left_str.data(), (int64_t)left_str.size(), rev_comp_kmer(kmer, p.k), p, 0,
                     budget, false);
    if (lc.corrected.size() != left_str.size()) return FrcOut{};  // left side must fully correct
    result.corrected = revcomp_ascii(lc.corrected);
    result.corrected.append(in + kmer_start, p.k);
    for (auto& x : lc.kmers) x.flipped = !x.flipped;
    result.kmers.assign(lc.kmers.rbegin(), lc.kmers.rend());
    result.kmers.push_back(ki);
    budget -= lc.corrections;
    result.corrections += lc.corrections;
  }
  if (it != n) correct_internal(&right, in + it, n - it, kmer, p, 0, budget, true);
  result.corrected += right.corrected;
  result.corrections += right.corrections;
  result.kmers.insert(result.kmers.end(), right.kmers.begin(), right.kmers.end());
  return result;
}

// ---------------------------------------------------------------------------------------
// Sequence store for the seqset stage: every corrected read and its reverse complement as
// base codes, one byte per base.  A suffix is (offset, len) into the store.
struct Store {
  std::vector<uint8_t> codes;
  std::vector<int64_t> fwd_off, rc_off;
  std::vector<int32_t> len;
};

void build_store(const char* seq, const int64_t* offs, int64_t n, Store* st) {
  int64_t total = offs[n] - offs[0];
  st->codes.resize(2 * total + 8);
  st->fwd_off.resize(n); st->rc_off.resize(n); st->len.resize(n);
  int64_t w = 0;
  for (int64_t r = 0; r < n; ++r) {
    int64_t L = offs[r + 1] - offs[r];
    const char* s = seq + offs[r];
    st->len[r] = (int32_t)L;
    st->fwd_off[r] = w;
    for (int64_t i = 0; i < L; ++i) st->codes[w + i] = (uint8_t)base_code(s[i]);
    w += L;
    st->rc_off[r] = w;
    for (int64_t i = 0; i < L; ++i) st->codes[w + i] = (uint8_t)(3 - base_code(s[L - 1 - i]));
    w += L;
  }
}

struct Rec {      // restated bs/repo_seq.h:85-183 entry_data: inline head + locator
  uint64_t head;  // first 32 bases, MSB-first, zero padded
  int64_t off;    // offset into Store::codes
  int32_t len;
};

inline uint64_t head_of(const uint8_t* p, int len) {
  uint64_t h = 0;
  int m = len < 32 ? len : 32;
  for (int i = 0; i < m; ++i) h |= (uint64_t)p[i] << (62 - 2 * i);
  return h;
}

// Ordering: bs/repo_seq.cpp:660-684 + modules/bio_base/dna_sequence.cpp:528-566:
// lexicographic, A<C<G<T, a proper prefix sorts first.
// returns <0, 0, >0 ; *lcp (optional) = shared prefix length
inline int compare_seq(const uint8_t* a, int la, const uint8_t* b, int lb, int* lcp) {
  int m = la < lb ? la : lb;
  int i = 0;
  while (i < m && a[i] == b[i]) ++i;
  if (lcp) *lcp = i;
  if (i < m) return (int)a[i] - (int)b[i];
  return la - lb;
}

struct RecLess {
  const uint8_t* codes;
  bool operator()(const Rec& x, const Rec& y) const {
    if (x.head != y.head) return x.head < y.head;
    if (x.len <= 32 || y.len <= 32) return x.len < y.len;  // equal padded heads: the shorter is a prefix
    return compare_seq(codes + x.off + 32, x.len - 32, codes + y.off + 32, y.len - 32, nullptr) < 0;
  }
};

inline bool is_prefix_or_equal(const uint8_t* codes, const Rec& a, const Rec& b) {
  if (a.len > b.len) return false;
  return memcmp(codes + a.off, codes + b.off, a.len) == 0;
}

void parallel_sort(std::vector<Rec>& v, const uint8_t* codes) {
  RecLess less{codes};
#ifdef _OPENMP
  int T = omp_get_max_threads();
  size_t n = v.size();
  if (T > 1 && n > 100000) {
    // bs/expand.cpp:199-282 scatters into prefix sections and std::sorts each section in the
    // pool; restated as: sort T contiguous chunks in parallel, then pairwise parallel merges.
    std::vector<size_t> cut(T + 1);
    for (int t = 0; t <= T; ++t) cut[t] = n * t / T;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < T; ++t) std::sort(v.begin() + cut[t], v.begin() + cut[t + 1], less);
    for (int w = 1; w < T; w *= 2) {
#pragma omp parallel for schedule(dynamic, 1)
      for (int t = 0; t < T; t += 2 * w) {
        int mid = std::min(t + w, T), hi = std::min(t + 2 * w, T);
        if (mid < hi) std::inplace_merge(v.begin() + cut[t], v.begin() + cut[mid], v.begin() + cut[hi], less);
      }
    }
    return;
  }
#endif
  std::sort(v.begin(), v.end(), less);
}

// Drop every record that is a prefix of / equal to its successor
// (bs/expand.cpp:19-48 skip_dups; bs/expand_test.cpp:57-75 states the rule).
void dedup_sorted(std::vector<Rec>& v, const uint8_t* codes) {
  size_t n = v.size(), w = 0;
  for (size_t i = 0; i < n; ++i) {
    if (i + 1 < n && is_prefix_or_equal(codes, v[i], v[i + 1])) continue;
    v[w++] = v[i];
  }
  v.resize(w);
}

inline Rec make_rec(const uint8_t* codes, int64_t off, int len) {
  return Rec{head_of(codes + off, len), off, len};
}

// first index i in [lo,hi) with !(E[i] < x) in the sequence order.
inline size_t lower_bound_seq(const std::vector<Rec>& E, size_t lo, size_t hi, const uint8_t* codes, const Rec& x) {
  RecLess less{codes};
  return std::lower_bound(E.begin() + lo, E.begin() + hi, x, less) - E.begin();
}

// For every entry e = b.x of sorted, prefix-free E, find the first entry having x as a prefix
// by a streaming merge over the contiguous range of entries starting with b -- the restated
// form of the pushed/popped iterator merge in bs/expand.cpp:581-661 and bs/builder.cpp:72-112.
// f(e_index, covered, hit_index)
template <typename F>
void popped_merge(const std::vector<Rec>& E, const uint8_t* codes, F&& f) {
  size_t n = E.size();
  if (!n) return;
  size_t chunk = 1 << 14;
  size_t nchunks = (n + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 4)
  for (size_t c = 0; c < nchunks; ++c) {
    size_t lo = c * chunk, hi = std::min(n, lo + chunk);
    size_t j = 0;
    int cur_b = -1;
    for (size_t i = lo; i < hi; ++i) {
      const Rec& e = E[i];
      int b = codes[e.off];
      if (e.len == 1) { f(i, true, (size_t)0); continue; }  // empty pop: prefix of everything
      Rec x = make_rec(codes, e.off + 1, e.len - 1);
      if (b != cur_b || i == lo) { j = lower_bound_seq(E, 0, n, codes, x); cur_b = b; }
      RecLess less{codes};
      while (j < n && less(E[j], x)) ++j;
      bool cov = j < n && is_prefix_or_equal(codes, x, E[j]);
      f(i, cov, j);
    }
  }
}

struct SeqsetTables {
  std::vector<uint16_t> sizes, shared;
  std::vector<uint64_t> prev[4];
  uint64_t fixed[5];
  bool missing_expansion = false;
};

// bs/builder.cpp:8-164 (build_chunks) + :207-263 (make_seqset) + modules/bio_base/seqset.cpp:113-129.
void build_tables(const std::vector<Rec>& E, const uint8_t* codes, SeqsetTables* t) {
  size_t n = E.size();
  t->sizes.resize(n); t->shared.resize(n);
  size_t words = (n + 63) / 64;
  for (int b = 0; b < 4; ++b) t->prev[b].assign(words, 0);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    t->sizes[i] = (uint16_t)E[i].len;
    int lcp = 0;
    if (i) compare_seq(codes + E[i - 1].off, E[i - 1].len, codes + E[i].off, E[i].len, &lcp);
    t->shared[i] = (uint16_t)lcp;
  }
  std::atomic<bool> missing{false};
  std::vector<uint64_t>* prev = t->prev;
  popped_merge(E, codes, [&](size_t i, bool cov, size_t j) {
    if (!cov) { missing = true; return; }  // LOG(FATAL) "Missing expansion?" bs/builder.cpp:96
    int b = codes[E[i].off];
    __atomic_fetch_or(&prev[b][j >> 6], 1ULL << (j & 63), __ATOMIC_RELAXED);
  });
  t->missing_expansion = missing;
  uint64_t off = 0;
  for (int b = 0; b < 4; ++b) {
    t->fixed[b] = off;
    uint64_t c = 0;
    for (uint64_t w : t->prev[b]) c += __builtin_popcountll(w);
    off += c;
  }
  t->fixed[4] = off;
}

int64_t export_tables(const std::vector<Rec>& E, const SeqsetTables& t, uint16_t** sizes, uint16_t** shared,
                      uint64_t** prev, uint64_t* fixed, int64_t** locs) {
  size_t n = E.size(), words = (n + 63) / 64;
  *sizes = (uint16_t*)malloc(std::max<size_t>(1, n) * 2);
  *shared = (uint16_t*)malloc(std::max<size_t>(1, n) * 2);
  *prev = (uint64_t*)malloc(std::max<size_t>(1, words) * 8 * 4);
  memcpy(*sizes, t.sizes.data(), n * 2);
  memcpy(*shared, t.shared.data(), n * 2);
  for (int b = 0; b < 4; ++b) memcpy(*prev + b * words, t.prev[b].data(), words * 8);
  memcpy(fixed, t.fixed, 40);
  if (locs) {
    *locs = (int64_t*)malloc(std::max<size_t>(1, n) * 8);
    for (size_t i = 0; i < n; ++i) (*locs)[i] = E[i].off;
  }
  return t.missing_expansion ? -1 : (int64_t)n;
}

// bs/part_repo.cpp:481-510 (write_with_expansions): x, then every stride-th further suffix,
// `count` records in total, never an empty one.
inline size_t write_with_expansions(std::vector<Rec>& out, const uint8_t* codes, int64_t off, int len,
                                    unsigned stride, unsigned count) {
  size_t w = 1;
  out.push_back(make_rec(codes, off, len));
  --count;
  unsigned until = stride - 1;
  while (len > 1 && count) {
    ++off; --len;
    if (until) { --until; continue; }
    --count; until = stride - 1;
    out.push_back(make_rec(codes, off, len));
    ++w;
  }
  return w;
}

}  // namespace

extern "C" {

void orc_free(void* p) { free(p); }

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

uint64_t orc_rev_comp(uint64_t kmer, int k) { return rev_comp_kmer(kmer, k); }

// Exact canonical k-mer counts with flags, the reference's way: every thread walks its share of the
// reads and increments a shared open-addressing table with compare-and-swap
// (kmer_count_table::increment, bs/kmer_count_table.h:54-103: claim the slot with CAS, OR the
// starts-read flags -- swapped when the instance is flipped --, CAS-increment fwd_count or
// rev_count; exact_pass_processor::flush_part, bs/kmer_counter.cpp:629-697).  The reference's uint8
// counters + uint32 overflow table (:680-685) are one uint32 here (same value).  Its 256 hash
// partitions and multiple exact passes only bound memory; one table does the same work.
//   prefilter_min == 0: every distinct k-mer is counted (the parity oracle).
//   prefilter_min  > 0: the reference's stage 1 first (prob_pass_processor::flush_part, :579-617):
//                       2-bit saturating counters at canon * 11304120250909662091 mod size; stage 2
//                       then only counts k-mers whose cell reached min(3, prefilter_min)
//                       (:243-247, :296-302, :668).  Result-invisible for the solid set;
//                       this is the form the CPU baseline times.
// Output: the counted canonical k-mers ascending.  flags bit0 = fwd_starts_read, bit1 = rev_starts_read.
namespace {
struct CountSlot {
  std::atomic<uint64_t> key;   // canonical k-mer | flags (bits 63/62); ~0 = unused
  std::atomic<uint32_t> fwd, rev;
};
constexpr uint64_t kUnused = ~0ULL, kKmerBits = ~0ULL >> 2, kFwdBit = 1ULL << 63, kRevBit = 1ULL << 62;
inline uint64_t table_hash(uint64_t kmer) { return kmer * 15674341118187572551ULL; }  // kmer_count_table::hash_kmer
inline uint64_t pt_hash(uint64_t kmer) { return kmer * 11304120250909662091ULL; }      // bs/kmer_counter.cpp pt_hash_kmer
}  // namespace

int64_t orc_count_kmers2(const char* seq, const int64_t* offs, int64_t n, int k, int threads, int prefilter_min,
                         uint64_t** out_kmers, uint32_t** out_fwd, uint32_t** out_rev, uint8_t** out_flags) {
  set_threads(threads);
  if (k < 1 || k > 31) return -1;
  int64_t total = 0;
  for (int64_t r = 0; r < n; ++r) { int64_t L = offs[r + 1] - offs[r]; if (L >= k) total += L - k + 1; }
  // ---- stage 1 (optional): probabilistic 2-bit counters ------------------------------------------------
  std::vector<std::atomic<uint8_t>> prob;
  uint64_t prob_size = 0;
  unsigned prob_need = 0;
  if (prefilter_min > 0) {
    prob_need = (unsigned)std::min(3, prefilter_min);
    prob_size = std::max<uint64_t>(512 * 1024, (uint64_t)total);   // one cell per instance (reference: memory budget)
    prob = std::vector<std::atomic<uint8_t>>(prob_size);
    for (auto& c : prob) c.store(0, std::memory_order_relaxed);
#pragma omp parallel for schedule(dynamic, 512)
    for (int64_t r = 0; r < n; ++r) {
      for_each_kmer(seq + offs[r], offs[r + 1] - offs[r], k, [&](uint64_t kmer, bool, bool) {
        bool fl;
        uint64_t canon = canonicalize(kmer, k, &fl);
        std::atomic<uint8_t>& c = prob[pt_hash(canon) % prob_size];
        uint8_t v = c.load(std::memory_order_relaxed);
        while (v < 3 && !c.compare_exchange_weak(v, (uint8_t)(v + 1), std::memory_order_relaxed)) {}
      });
    }
  }
  auto passes = [&](uint64_t canon) {
    return !prob_need || prob[pt_hash(canon) % prob_size].load(std::memory_order_relaxed) >= prob_need;
  };
  // ---- distinct estimate (linear counting over the 1/16 of hash space with 4 low zero bits) -------------
  uint64_t est;
  {
    uint64_t bits = 1;
    while (bits < (uint64_t)std::max<int64_t>(1 << 20, total / 8)) bits <<= 1;
    std::vector<std::atomic<uint64_t>> bm(bits / 64);
    for (auto& w : bm) w.store(0, std::memory_order_relaxed);
#pragma omp parallel for schedule(dynamic, 512)
    for (int64_t r = 0; r < n; ++r) {
      for_each_kmer(seq + offs[r], offs[r + 1] - offs[r], k, [&](uint64_t kmer, bool, bool) {
        bool fl;
        uint64_t canon = canonicalize(kmer, k, &fl);
        uint64_t h = table_hash(canon);
        h ^= h >> 29;
        if ((h & 15) == 0 && passes(canon)) {
          uint64_t b = (h >> 4) & (bits - 1);
          bm[b >> 6].fetch_or(1ULL << (b & 63), std::memory_order_relaxed);
        }
      });
    }
    uint64_t ones = 0;
    for (auto& w : bm) ones += __builtin_popcountll(w.load(std::memory_order_relaxed));
    double zf = std::max(1.0 / (double)bits, 1.0 - (double)ones / (double)bits);
    est = (uint64_t)(16.0 * -(double)bits * std::log(zf));
  }
  uint64_t slots = 1024;
  while (slots < est + est / 2 + 4096) slots <<= 1;
  for (;;) {
    std::vector<CountSlot> table(slots);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)slots; ++i) {
      table[i].key.store(kUnused, std::memory_order_relaxed);
      table[i].fwd.store(0, std::memory_order_relaxed);
      table[i].rev.store(0, std::memory_order_relaxed);
    }
    std::atomic<bool> full{false};
    const uint64_t mask = slots - 1;
    int slot_shift = 64;
    for (uint64_t t = slots; t > 1; t >>= 1) --slot_shift;   // top bits of the multiplicative hash
#pragma omp parallel for schedule(dynamic, 512)
    for (int64_t r = 0; r < n; ++r) {
      if (full.load(std::memory_order_relaxed)) continue;
      for_each_kmer(seq + offs[r], offs[r + 1] - offs[r], k, [&](uint64_t kmer, bool first, bool last) {
        bool fl;
        uint64_t canon = canonicalize(kmer, k, &fl);
        if (!passes(canon)) return;
        uint64_t pos = table_hash(canon) >> slot_shift;
        uint64_t probes = 0;
        for (;;) {
          uint64_t cur = table[pos].key.load(std::memory_order_relaxed);
          if (cur == kUnused) {
            if (table[pos].key.compare_exchange_strong(cur, canon, std::memory_order_relaxed)) cur = canon;
          }
          if ((cur & kKmerBits) == canon) break;
          pos = (pos + 1) & mask;
          if (++probes > mask) { full.store(true); return; }   // "Kmer table (...) too small"
        }
        bool ff = first, rf = last;
        if (fl) std::swap(ff, rf);
        uint64_t nf = (ff ? kFwdBit : 0) | (rf ? kRevBit : 0);
        if (nf) table[pos].key.fetch_or(nf, std::memory_order_relaxed);
        std::atomic<uint32_t>& c = fl ? table[pos].rev : table[pos].fwd;
        uint32_t v = c.load(std::memory_order_relaxed);
        while (v != 0xFFFFFFFFu && !c.compare_exchange_weak(v, v + 1, std::memory_order_relaxed)) {}
      });
    }
    if (full.load()) { slots <<= 1; continue; }
    // ---- gather, order by k-mer (256 buckets on the leading 4 bases, sorted in parallel) -----------------
    const int shift = 2 * k - 8;
    std::vector<uint64_t> bucket_n(257, 0);
    for (uint64_t i = 0; i < slots; ++i) {
      uint64_t key = table[i].key.load(std::memory_order_relaxed);
      if (key != kUnused) ++bucket_n[((key & kKmerBits) >> shift) + 1];
    }
    for (int b = 0; b < 256; ++b) bucket_n[b + 1] += bucket_n[b];
    const uint64_t distinct = bucket_n[256];
    std::vector<uint64_t> idx(distinct), cur(bucket_n.begin(), bucket_n.end() - 1);
    for (uint64_t i = 0; i < slots; ++i) {
      uint64_t key = table[i].key.load(std::memory_order_relaxed);
      if (key != kUnused) idx[cur[(key & kKmerBits) >> shift]++] = i;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < 256; ++b)
      std::sort(idx.begin() + bucket_n[b], idx.begin() + bucket_n[b + 1], [&](uint64_t x, uint64_t y) {
        return (table[x].key.load(std::memory_order_relaxed) & kKmerBits) < (table[y].key.load(std::memory_order_relaxed) & kKmerBits);
      });
    *out_kmers = (uint64_t*)malloc(std::max<size_t>(1, distinct) * 8);
    *out_fwd = (uint32_t*)malloc(std::max<size_t>(1, distinct) * 4);
    *out_rev = (uint32_t*)malloc(std::max<size_t>(1, distinct) * 4);
    *out_flags = (uint8_t*)malloc(std::max<size_t>(1, distinct));
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < (int64_t)distinct; ++w) {
      const CountSlot& e = table[idx[w]];
      uint64_t key = e.key.load(std::memory_order_relaxed);
      (*out_kmers)[w] = key & kKmerBits;
      (*out_fwd)[w] = e.fwd.load(std::memory_order_relaxed);
      (*out_rev)[w] = e.rev.load(std::memory_order_relaxed);
      (*out_flags)[w] = (uint8_t)(((key & kFwdBit) ? 1 : 0) | ((key & kRevBit) ? 2 : 0));
    }
    return (int64_t)distinct;
  }
}

int64_t orc_count_kmers(const char* seq, const int64_t* offs, int64_t n, int k, int threads,
                        uint64_t** out_kmers, uint32_t** out_fwd, uint32_t** out_rev, uint8_t** out_flags) {
  return orc_count_kmers2(seq, offs, n, k, threads, 0, out_kmers, out_fwd, out_rev, out_flags);
}

// Single-read correction (parity hook for modules/bio_base/fast_read_correct_test.cpp).
// Returns corrected length; out must hold `len` chars.
int orc_fast_read_correct(const char* read, int len, const uint64_t* solid, int64_t n_solid, int k,
                          int max_corrections, int min_good_run, char* out, int* corrections) {
  KmerSet ks{solid, nullptr, n_solid, k};
  FrcParams p{(unsigned)max_corrections, (unsigned)min_good_run, k, &ks};
  FrcOut o = fast_read_correct(read, len, p);
  memcpy(out, o.corrected.data(), o.corrected.size());
  if (corrections) *corrections = (int)o.corrections;
  return (int)o.corrected.size();
}

// bs/correct_reads.cpp:154-231 (correct_reads::correct) over a batch.  A read is dropped when
// len < k (:155) or corrected.size() < unsigned(trim_after_portion * len) (:174-178; the CLI
// parses the flag as float and widens it to double, biograph_create.cpp:489-490,731).
// next_fwd / next_rev (:195-210): 1 + number of k-mers until one (other than the first) whose
// flag says a read starts there, for the read and for its reverse complement.
// out_offs has n+1 entries; dropped reads have zero length.  Returns the number kept.
int64_t orc_correct_reads(const char* seq, const int64_t* offs, int64_t n, const uint64_t* solid,
                          const uint8_t* flags, int64_t n_solid, int k, int max_corrections,
                          int min_good_run, float trim_after_portion, int threads, char* out_seq,
                          int64_t* out_offs, uint8_t* kept, int32_t* corrections, int32_t* next_fwd,
                          int32_t* next_rev) {
  set_threads(threads);
  KmerSet ks{solid, flags, n_solid, k};
  int pb = 0;
  std::vector<uint32_t> lookup = build_kmer_lookup(solid, n_solid, k, &pb);
  if (!lookup.empty()) { ks.lookup = lookup.data(); ks.prefix_bits = pb; }
  FrcParams p{(unsigned)max_corrections, (unsigned)min_good_run, k, &ks};
  const double portion = (double)trim_after_portion;
  std::vector<int32_t> out_len(n, 0);
  // first pass: correct into the slot of the input read (corrected.size() <= len always)
  std::vector<char> tmp(offs[n] - offs[0] + 1);
  const int64_t base0 = offs[0];
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < n; ++r) {
    int64_t L = offs[r + 1] - offs[r];
    kept[r] = 0;
    if (corrections) corrections[r] = 0;
    if (next_fwd) next_fwd[r] = 0;
    if (next_rev) next_rev[r] = 0;
    if (L < k) continue;
    FrcOut o = fast_read_correct(seq + offs[r], L, p);
    unsigned needed = (unsigned)(portion * (double)L);
    if (o.corrected.size() < needed) continue;
    if (o.corrected.empty()) continue;  // needed==0 and nothing corrected: CHECK_GT(next_fwd_read,0) would abort
    kept[r] = 1;
    out_len[r] = (int32_t)o.corrected.size();
    memcpy(tmp.data() + (offs[r] - base0), o.corrected.data(), o.corrected.size());
    if (corrections) corrections[r] = (int32_t)o.corrections;
    if (flags) {
      // kmer_starts_read (bs/correct_reads.cpp:308-311): flipped ? rev_starts_read : fwd_starts_read
      auto starts = [&](const FrcKmer& km, bool toggle) {
        bool fl = km.flipped ^ toggle;
        return (flags[km.index] & (fl ? 2 : 1)) != 0;
      };
      int nf = 0;
      for (size_t i = 0; i < o.kmers.size(); ++i) { if (nf > 0 && starts(o.kmers[i], false)) break; ++nf; }
      int nr = 0;
      for (size_t i = o.kmers.size(); i-- > 0;) { if (nr > 0 && starts(o.kmers[i], true)) break; ++nr; }
      if (next_fwd) next_fwd[r] = nf;
      if (next_rev) next_rev[r] = nr;
    }
  }
  int64_t w = 0, nk = 0;
  for (int64_t r = 0; r < n; ++r) {
    out_offs[r] = w;
    if (kept[r]) { memcpy(out_seq + w, tmp.data() + (offs[r] - base0), out_len[r]); w += out_len[r]; ++nk; }
  }
  out_offs[n] = w;
  return nk;
}

// Closed-form seqset (bs/expand_test.cpp:57-75): the sorted set of all
// suffixes of all reads and of their reverse complements, minus every sequence that is a
// prefix of / equal to another.  Reads must be pure ACGT (corrected reads are).
// Returns the number of entries (or -1 on an internal closure violation, which cannot happen
// here).  locs (optional): offset of each entry in the fwd/rc store, for debugging.
int64_t orc_seqset_closed_form(const char* seq, const int64_t* offs, int64_t n, int threads,
                               uint16_t** sizes, uint16_t** shared, uint64_t** prev, uint64_t* fixed) {
  set_threads(threads);
  Store st;
  build_store(seq, offs, n, &st);
  const uint8_t* codes = st.codes.data();
  std::vector<Rec> v;
  size_t total = 0;
  for (int64_t r = 0; r < n; ++r) total += 2 * (size_t)st.len[r];
  v.reserve(total);
  for (int64_t r = 0; r < n; ++r) {
    int L = st.len[r];
    for (int i = 0; i < L; ++i) v.push_back(make_rec(codes, st.fwd_off[r] + i, L - i));
    for (int i = 0; i < L; ++i) v.push_back(make_rec(codes, st.rc_off[r] + i, L - i));
  }
  parallel_sort(v, codes);
  dedup_sorted(v, codes);
  SeqsetTables t;
  build_tables(v, codes, &t);
  return export_tables(v, t, sizes, shared, prev, fixed, nullptr);
}

// Staged seqset = the reference's own algorithm (the CPU baseline):
//   seeds        bs/correct_reads.cpp:215-226 -> bs/part_repo.cpp:53-126: the first next_fwd[r]
//                suffixes of read r and the first next_rev[r] suffixes of its reverse complement
//   round 1      sort_and_dedup("initial" -> "init_sorted")                       biograph_create.cpp:925
//   expand       every popped entry not covered by an entry is written with stride 7, count 255   :926
//   round 2      sort_and_dedup(init_sorted + init_expanded -> pass2_sorted), every surviving NEW
//                entry writes pop_front with stride 1, count 6                    :927
//   round 3      sort_and_dedup(pass2_sorted + pass2_expanded -> complete)        :929
//   builder      bs/builder.cpp
// stats[0..5] = seeds, after round 1, expanded (7/255), after round 2, expanded (1/6), final.
int64_t orc_seqset_staged(const char* seq, const int64_t* offs, int64_t n, const int32_t* next_fwd,
                          const int32_t* next_rev, int threads, uint16_t** sizes, uint16_t** shared,
                          uint64_t** prev, uint64_t* fixed, int64_t* stats) {
  set_threads(threads);
  Store st;
  build_store(seq, offs, n, &st);
  const uint8_t* codes = st.codes.data();
  std::vector<Rec> sorted;
  for (int64_t r = 0; r < n; ++r) {
    int L = st.len[r];
    int nf = next_fwd ? next_fwd[r] : L, nr = next_rev ? next_rev[r] : L;
    for (int i = 0; i < nf && i < L; ++i) sorted.push_back(make_rec(codes, st.fwd_off[r] + i, L - i));
    for (int i = 0; i < nr && i < L; ++i) sorted.push_back(make_rec(codes, st.rc_off[r] + i, L - i));
  }
  if (stats) stats[0] = (int64_t)sorted.size();
  parallel_sort(sorted, codes);
  dedup_sorted(sorted, codes);
  if (stats) stats[1] = (int64_t)sorted.size();

  // expand(stride 7, count 255): bs/expand.cpp:581-661
  std::vector<Rec> expanded;
  {
    int T = orc_max_threads();
    std::vector<std::vector<Rec>> per(T);
    popped_merge(sorted, codes, [&](size_t i, bool cov, size_t) {
      if (cov) return;
#ifdef _OPENMP
      int t = omp_get_thread_num();
#else
      int t = 0;
#endif
      write_with_expansions(per[t], codes, sorted[i].off + 1, sorted[i].len - 1, 7, 255);
    });
    for (auto& p : per) expanded.insert(expanded.end(), p.begin(), p.end());
  }
  if (stats) stats[2] = (int64_t)expanded.size();

  // round 2: sort new, merge with sorted, drop prefixes; surviving new entries expand 1/6.
  // bs/expand.cpp:285-380 (dedup_and_output): a new record survives iff no other record (old
  // or new) has it as a prefix/equal; an old record equal to a new one keeps the old
  // (EQUAL -> "Duplicate; ignore"), so equal runs are ordered new-first below.
  parallel_sort(expanded, codes);
  std::vector<Rec> merged(sorted.size() + expanded.size());
  std::vector<uint8_t> is_new(merged.size());
  {
    RecLess less{codes};
    size_t a = 0, b = 0, w = 0;
    while (a < sorted.size() || b < expanded.size()) {
      bool take_new;
      if (a == sorted.size()) take_new = true;
      else if (b == expanded.size()) take_new = false;
      else take_new = !less(sorted[a], expanded[b]);  // equal -> the new record first, so it is the one dropped
      if (take_new) { merged[w] = expanded[b++]; is_new[w++] = 1; }
      else { merged[w] = sorted[a++]; is_new[w++] = 0; }
    }
  }
  std::vector<Rec> pass2;
  std::vector<Rec> expanded2;
  {
    size_t m = merged.size();
    for (size_t i = 0; i < m; ++i) {
      if (i + 1 < m && is_prefix_or_equal(codes, merged[i], merged[i + 1])) continue;
      pass2.push_back(merged[i]);
      if (is_new[i] && merged[i].len > 1)
        write_with_expansions(expanded2, codes, merged[i].off + 1, merged[i].len - 1, 1, 6);
    }
  }
  if (stats) { stats[3] = (int64_t)pass2.size(); stats[4] = (int64_t)expanded2.size(); }
  // round 3
  pass2.insert(pass2.end(), expanded2.begin(), expanded2.end());
  parallel_sort(pass2, codes);
  dedup_sorted(pass2, codes);
  if (stats) stats[5] = (int64_t)pass2.size();
  SeqsetTables t;
  build_tables(pass2, codes, &t);
  return export_tables(pass2, t, sizes, shared, prev, fixed, nullptr);
}

// modules/io/bitcount.cpp:84-123 (finalize).  accum has ceil((nbits+1)/512) words, subaccum
// ceil(nbits/512).  Returns the total number of set bits.
uint64_t orc_bitcount_finalize(const uint64_t* bits, uint64_t nbits, uint64_t* subaccum, uint64_t* accum) {
  if (nbits == 0) { accum[0] = 0; return 0; }
  uint64_t words = (nbits + 63) / 64;
  uint64_t sub = 0, total = 0;
  for (uint64_t i = 0; i < words; ++i) {
    if (i % 8 == 0) {
      accum[i / 8] = total;
      if (i) subaccum[i / 8 - 1] = sub;
      sub = 0;
    }
    sub <<= 8;
    uint64_t c = (uint64_t)__builtin_popcountll(bits[i]);
    sub |= c;
    total += c;
  }
  uint64_t left = words % 8;
  while (left) { sub <<= 8; left = (left + 1) % 8; }
  subaccum[(nbits + 511) / 512 - 1] = sub;
  if (nbits % 512 == 0) accum[nbits / 512] = total;
  return total;
}

// modules/io/packed_varbit_vector.cpp:174-188 (bits_for_value / elements_for_values) and
// :80-139 (varbit_set): values packed LSB-first, little-endian, bits_per_value = bit_length(max).
// out must hold orc_varbit_words(n, max_value) uint64 words.  Returns bits_per_value.
uint64_t orc_varbit_words(uint64_t n, uint64_t max_value) {
  unsigned bits = 0;
  while (max_value) { ++bits; max_value >>= 1; }
  return (n * bits + 63) / 64;
}
int orc_varbit_pack(const uint16_t* vals, uint64_t n, uint64_t max_value, uint64_t* out) {
  unsigned bits = 0;
  for (uint64_t m = max_value; m; m >>= 1) ++bits;
  uint64_t words = (n * bits + 63) / 64;
  memset(out, 0, words * 8);
  if (!bits) return 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v = vals[i], bit = i * bits;
    out[bit >> 6] |= v << (bit & 63);
    if ((bit & 63) + bits > 64) out[(bit >> 6) + 1] |= v >> (64 - (bit & 63));
  }
  return (int)bits;
}

}  // extern "C"

// merge_core.cuh -- the per-element bodies of the seqset-merge kernels (merge.cu, and the merge-mode
// prev-bit kernel in seqset.cu), written as host+device functions of an element index so that the very
// same code is also run serially by a g++-built harness (tests/cpp/merge_core_test.cpp) against the CPU
// restatement of the reference's merge -- the dev container has no GPU, the logic is checked there and the
// launches on the B200.
//
// Replaces (reference, CPU): seqset_flat_builder::build + seqset_flat::get (modules/bio_base/
// seqset_flat.cpp:40-290, seqset_flat.h:117-140), make_mergemap::count_range / fill_mergemap
// (modules/bio_base/make_mergemap.cpp:188-259), seqset_merger::merge_range (modules/bio_base/
// seqset_merger.cpp:109-197) and make_readmap::create_from_fast_migrate (modules/bio_mapred/
// make_readmap.cpp:459-520).
//
// Flattening on a GPU.  Entry i of a seqset has first base b(i) = #{c : i >= fixed[c]} - 1 and
// pops to next(i) = the (i - fixed[b])-th set bit of prev_b (seqset.cpp:249-254,709-720), so base j of
// entry i is b(next^j(i)).  The reference traces these chains one base at a time per entry
// (seqset_flat_builder::trace); here the chains are DOUBLED: round t turns (first 2^t bases, next^(2^t))
// of every entry into (first 2^(t+1) bases, next^(2^(t+1))) with two gathers, so five rounds give every
// entry its first 32 bases as one word plus the pointer 32 bases on, and an entry of L bases is
// ceil(L/32) hops.  The result is the layout the seqset stage already sorts: a store of 2-bit words
// (every entry on a word boundary, zero padded) and one (key, loc) record per entry.
#pragma once

#include <cstdint>

#include "common.cuh"

#if defined(__CUDACC__)
#define BGX_HD __host__ __device__ __forceinline__
#else
#define BGX_HD inline
#endif

namespace bgx {
namespace mergecore {

constexpr uint32_t kNone = 0xffffffffu;
constexpr int kMaxParts = 64;  // inputs of one merge (the reference's tests go to 15)

struct PartTable {
  int n;
  uint64_t word_base[kMaxParts + 1];  // first store word of each part; [n] = total words
};

BGX_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
BGX_HD int clz64(uint64_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}
BGX_HD int ctz64(uint64_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
BGX_HD void or_bit(unsigned long long* bits, uint64_t pos) {
#if defined(__CUDA_ARCH__)
  atomicOr(&bits[pos >> 6], 1ULL << (pos & 63));
#else
  bits[pos >> 6] |= 1ULL << (pos & 63);
#endif
}

// ---- bit vectors ------------------------------------------------------------------------------------
// word w of an nbits-long bitcount `bits` array with everything at or after nbits cleared
BGX_HD uint64_t masked_word(const uint64_t* bits, uint64_t w, uint64_t nbits) {
  uint64_t v = bits[w];
  const uint64_t lo = w * 64;
  if (lo + 64 > nbits) v = nbits > lo ? v & (~0ULL >> (64 - (nbits - lo))) : 0;
  return v;
}

// positions of the set bits of word w, in order, to out[out_base + excl ...] (excl = set bits before word w):
// the whole vector scattered this way is its select table (bitcount::find_count for every count)
BGX_HD void scatter_set_bits(const uint64_t* bits, uint64_t w, uint64_t nbits, uint32_t excl, uint32_t* out,
                             uint64_t out_base) {
  uint64_t v = masked_word(bits, w, nbits);
  uint64_t o = out_base + excl;
  while (v) {
    out[o++] = (uint32_t)(w * 64 + ctz64(v));
    v &= v - 1;
  }
}

// ---- flattening --------------------------------------------------------------------------------------
// first base of entry i (seqset::entry_get_base, seqset.cpp:249-254); f = fixed[1..3]
BGX_HD uint64_t first_base(uint64_t i, uint64_t f1, uint64_t f2, uint64_t f3) {
  return (uint64_t)(i >= f1) + (uint64_t)(i >= f2) + (uint64_t)(i >= f3);
}

// round 0 of the doubling: one base per entry in the top two bits
BGX_HD void double_init(uint64_t i, uint64_t f1, uint64_t f2, uint64_t f3, uint64_t* w0) {
  w0[i] = first_base(i, f1, f2, f3) << 62;
}

// one doubling round: have = bases held per entry before the round (1, 2, 4, 8, 16)
BGX_HD void double_step(const uint64_t* w_in, const uint32_t* j_in, int have, uint64_t i, uint64_t* w_out,
                        uint32_t* j_out) {
  const uint32_t j = j_in[i];
  w_out[i] = w_in[i] | (w_in[j] >> (2 * have));
  j_out[i] = j_in[j];
}

// words of an entry of `size` bases
BGX_HD uint32_t entry_words(uint32_t size) { return (size + 31) >> 5; }

// entry i of a part -> its store words (zero padded) and its (key, loc) record.
//   w32 / j32 : first 32 bases and the entry 32 pops on, per entry (after five doubling rounds)
//   woff      : exclusive scan of entry_words over the part's entries
//   word_base : first store word of the part; rec_base: first record of the part
BGX_HD void emit_entry(const uint64_t* w32, const uint32_t* j32, const uint16_t* sizes, const uint32_t* woff,
                       uint64_t i, uint64_t word_base, uint64_t rec_base, uint64_t* store, uint64_t* keys,
                       uint64_t* locs) {
  const uint32_t size = sizes[i];
  const uint32_t nw = entry_words(size);
  const uint64_t w0 = word_base + woff[i];
  uint32_t cur = (uint32_t)i;
  uint64_t key = 0;
  for (uint32_t w = 0; w < nw; ++w) {
    uint64_t word = w32[cur];
    const int rem = (int)size - 32 * (int)w;
    if (rem < 32) word &= top_bases_mask(rem);
    store[w0 + w] = word;
    if (w == 0) key = word;
    cur = j32[cur];
  }
  keys[rec_base + i] = key;
  locs[rec_base + i] = make_loc(w0 * 32, size);
}

// ---- sequence comparisons on the flat store ----------------------------------------------------------------
// 32 bases from base address a (the store has one readable pad word at its end)
BGX_HD uint64_t window(const uint64_t* store, uint64_t a) {
  const uint64_t q = a >> 5;
  const unsigned s = (unsigned)(a & 31) * 2;
  const uint64_t hi = store[q];
  if (s == 0) return hi;
  return (hi << s) | (store[q + 1] >> (64 - s));
}

// lexicographic order of the sequences (aa, na) and (ab, nb), a proper prefix first
// (dna_sequence.cpp:528-566): <0, 0, >0; *lcp = shared prefix length
BGX_HD int compare_seq(const uint64_t* store, uint64_t aa, int na, uint64_t ab, int nb, int* lcp) {
  const int m = na < nb ? na : nb;
  int d = 0;
  while (d < m) {
    const int c = m - d < 32 ? m - d : 32;
    const uint64_t msk = top_bases_mask(c);
    const uint64_t wa = window(store, aa + d) & msk, wb = window(store, ab + d) & msk;
    if (wa != wb) {
      *lcp = d + (clz64(wa ^ wb) >> 1);
      return wa < wb ? -1 : 1;
    }
    d += c;
  }
  *lcp = m;
  return na - nb;
}

// (xa, xn) equal to / a prefix of entry e ?
BGX_HD bool prefixes(const uint64_t* store, uint64_t xa, int xn, uint64_t el) {
  if (xn > (int)loc_len(el)) return false;
  int lcp;
  compare_seq(store, xa, xn, loc_addr(el), (int)loc_len(el), &lcp);
  return lcp >= xn;
}

// ---- seqset_merger's prev bits ----------------------------------------------------------------------------
// start of the generate_chunks(0, n, nsplits) chunk that holds entry e (modules/io/parallel.cpp:60-83:
// chunk c = [n*c/nsplits, n*(c+1)/nsplits), empty ones skipped)
BGX_HD uint64_t chunk_start_of(uint64_t e, uint64_t n, uint64_t nsplits) {
  const uint64_t c = ((e + 1) * nsplits - 1) / n;
  return n * c / nsplits;
}

// Where seqset_merger puts the prev bit of entry i = b.x among n sorted prefix-free entries: with
// [lo, hi) the entries x is a prefix of, on the first entry of [lo, hi) that lies in the chunk of entry
// hi - 1 (merge_range: get_base_iterator of a chunk's limit backs up over the candidate whose tail
// prefixes the limit entry, seqset_merger.cpp:80-107, so only the chunk of the LAST prefixed entry sees
// it, and sets the bit on the first overlap it meets, :139-158).  nsplits = 1 is builder::build_chunks'
// rule (first entry of the range, bs/builder.cpp:85-107).  The search for lo runs over [lo0, hi0)
// (the caller may narrow it with a prefix index).  kNone: no entry has x as a prefix (an input was not
// closed under pop_front -- "Missing expansion?").
BGX_HD uint32_t merge_prev_target(const uint64_t* store, const uint64_t* locs, uint32_t n, uint32_t i, uint32_t lo0,
                                  uint32_t hi0, uint64_t nsplits) {
  const uint64_t l = locs[i];
  const uint64_t xa = loc_addr(l) + 1;
  const int xn = (int)loc_len(l) - 1;
  uint32_t lo, last;
  if (xn == 0) {  // the empty sequence prefixes everything
    lo = 0;
    last = n - 1;
  } else {
    uint32_t a = lo0, b = hi0;
    while (a < b) {  // first entry not less than x
      const uint32_t mid = a + ((b - a) >> 1);
      int lcp;
      if (compare_seq(store, loc_addr(locs[mid]), (int)loc_len(locs[mid]), xa, xn, &lcp) < 0) a = mid + 1; else b = mid;
    }
    lo = a;
    if (lo >= n || !prefixes(store, xa, xn, locs[lo])) return kNone;
    // last entry x prefixes: gallop, then bisect (the range is one entry unless x is short)
    uint32_t good = lo, step = 1, bad = n;
    while (good + step < n) {
      if (prefixes(store, xa, xn, locs[good + step])) { good += step; step <<= 1; }
      else { bad = good + step; break; }
    }
    while (good + 1 < bad) {
      const uint32_t mid = good + ((bad - good) >> 1);
      if (prefixes(store, xa, xn, locs[mid])) good = mid; else bad = mid;
    }
    last = good;
  }
  const uint64_t cs = chunk_start_of(last, n, nsplits);
  return cs > lo ? (uint32_t)cs : lo;
}

// ---- mergemap ------------------------------------------------------------------------------------------------
// part that owns the store word `w`
BGX_HD int part_of_word(const PartTable& pt, uint64_t w) {
  int p = 0;
  while (p + 1 < pt.n && w >= pt.word_base[p + 1]) ++p;
  return p;
}

// sorted (pre-dedup) record j of the union: its part's mergemap gets the bit of the merged entry the
// record was folded into = the first surviving record at or after it (make_mergemap.cpp:205-228)
BGX_HD void mergemap_mark(const uint64_t* sorted_locs, const uint32_t* pos, uint32_t j, const PartTable& pt,
                          unsigned long long* mm_bits, uint64_t mm_words) {
  const int p = part_of_word(pt, loc_addr(sorted_locs[j]) >> 5);
  or_bit(mm_bits + (uint64_t)p * mm_words, pos[j]);
}

// ---- readmap migration ---------------------------------------------------------------------------------------
// word w of a part-indexed bit vector (sparse_multi's source_to_mid) re-targeted to merged entry ids:
// bit e -> bit sel[e], sel = select table of the part's mergemap (fast_migrate, make_readmap.cpp:472-486)
BGX_HD void migrate_word(const uint64_t* old_bits, uint64_t w, uint64_t n_old, const uint32_t* sel,
                         unsigned long long* new_bits) {
  uint64_t v = masked_word(old_bits, w, n_old);
  while (v) {
    or_bit(new_bits, sel[w * 64 + ctz64(v)]);
    v &= v - 1;
  }
}

}  // namespace mergecore
}  // namespace bgx

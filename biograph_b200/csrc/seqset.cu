// seqset.cu -- suffix seeds -> sort -> prefix-dedup -> closure -> seqset tables, on the GPU.
//
// Replaces (reference, CPU + temp files): part_repo::write seeds (bs/part_repo.cpp:53-126),
// expander::sort_and_dedup / expand (bs/expand.cpp:418-661, driver biograph_create.cpp:921-931),
// builder::build_chunks / make_seqset (bs/builder.cpp:8-263), bitcount::finalize
// (modules/io/bitcount.cpp:84-123) and seqset::finalize (modules/bio_base/seqset.cpp:113-129).
//
// A suffix record is 16 bytes held as two arrays: key = first 32 bases (MSB-first, zero padded,
// so integer order == sequence order with a proper prefix first when lengths break ties) and
// loc = (base address in the corrected store << 16) | length.  The store keeps every kept read
// AND its reverse complement, so every suffix is a plain forward slice (no rc logic in any
// comparison).
//
// Rounds (result-identical to the reference's seed / stride-7 / stride-1 scheme because the
// final set is the closure):
//   1. seeds: first next_fwd suffixes of each read, first next_rev of its reverse complement
//   2. sort (LSD radix on the top key bits, then tie groups resolved by full comparison) + dedup
//   3. closure walk: for every entry whose pop_front is not covered, emit pop_front and the
//      following suffixes up to the first covered one (one round reaches the closure)
//   4. sort + dedup of (round-2 survivors + walk output)
//   5. tables: sizes, shared (LCP with predecessor), prev bits (lower_bound of pop_front),
//      bitcount accum/subaccum, fixed.
#include <algorithm>
#include <vector>

#include "ctx.h"
#include "merge_core.cuh"

namespace bgx {
namespace {

// records one GPU shard may hold at any time (32-bit indices; the refinement keys keep 31 bits for a group id)
constexpr uint64_t kMaxShardRecords = 1ull << 31;
constexpr int kSmallGroup = 32;   // tie groups up to this size are sorted by one thread
constexpr int kMediumGroup = 512; // ... up to this size by one warp (rank sort); larger ones are refined chunk by chunk
constexpr int kChunkBases = 12;   // bases resolved per refinement round for big tie groups

// ---- comparisons ----------------------------------------------------------------------------
// compare suffixes la, lb from base depth d on (all bases before d known equal).
// <0, 0, >0 ; a proper prefix sorts first (bs/repo_seq.cpp:660-684, dna_sequence.cpp:528-566).
__device__ __forceinline__ int compare_from(const uint64_t* __restrict__ store, uint64_t la, uint64_t lb, int d,
                                            int* lcp) {
  const uint64_t aa = loc_addr(la), ab = loc_addr(lb);
  const int na = (int)loc_len(la), nb = (int)loc_len(lb);
  const int m = min(na, nb);
  while (d < m) {
    int c = min(32, m - d);
    uint64_t msk = top_bases_mask(c);
    uint64_t wa = load_window(store, aa + d) & msk;
    uint64_t wb = load_window(store, ab + d) & msk;
    if (wa != wb) {
      if (lcp) *lcp = d + (__clzll(wa ^ wb) >> 1);
      return wa < wb ? -1 : 1;
    }
    d += c;
  }
  if (lcp) *lcp = m;
  return na - nb;
}

__device__ __forceinline__ bool rec_less(const uint64_t* __restrict__ store, uint64_t ka, uint64_t la, uint64_t kb,
                                         uint64_t lb) {
  if (ka != kb) return ka < kb;
  int na = (int)loc_len(la), nb = (int)loc_len(lb);
  if (na <= 32 || nb <= 32) return na < nb;  // equal padded keys: the shorter one is a prefix
  return compare_from(store, la, lb, 32, nullptr) < 0;
}

// three-way order of two records (<0, 0, >0); 0 = the same sequence
__device__ __forceinline__ int rec_cmp(const uint64_t* __restrict__ store, uint64_t ka, uint64_t la, uint64_t kb,
                                       uint64_t lb) {
  if (ka != kb) return ka < kb ? -1 : 1;
  int na = (int)loc_len(la), nb = (int)loc_len(lb);
  if (na <= 32 || nb <= 32) return na - nb;
  return compare_from(store, la, lb, 32, nullptr);
}

// a prefix of / equal to b ?
__device__ __forceinline__ bool prefix_or_equal(const uint64_t* __restrict__ store, uint64_t ka, uint64_t la,
                                                uint64_t kb, uint64_t lb) {
  int na = (int)loc_len(la), nb = (int)loc_len(lb);
  if (na > nb) return false;
  if ((ka ^ kb) & top_bases_mask(min(na, 32))) return false;
  if (na <= 32) return true;
  int lcp;
  compare_from(store, la, lb, 32, &lcp);
  return lcp >= na;
}

// first index in [lo,hi) whose record is not less than x (hi if none)
__device__ __forceinline__ uint32_t lower_bound_rec(const uint64_t* __restrict__ store,
                                                    const uint64_t* __restrict__ keys,
                                                    const uint64_t* __restrict__ locs, uint32_t lo, uint32_t hi,
                                                    uint64_t xk, uint64_t xl) {
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (rec_less(store, keys[mid], locs[mid], xk, xl)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Prefix-bucket index over a sorted record array: start[b] = first record whose top key bits are
// >= b (start[2^bits] = n).  A lower_bound then only searches the few records of the query's own
// bucket: one L2-resident table read + a search inside one or two DRAM sectors of keys, instead of
// log2(n) dependent probes over the whole array (the reference's merge-joins stream both sides;
// a GPU thread per query wants the range narrowed instead).
struct BucketIndex {
  const uint32_t* start;
  int shift;  // 64 - bits
};

// Thread i closes the buckets between the one of record i-1 and the one of record i.  A long run of
// empty buckets -- everything below / above the key range a rank of a sharded build owns is one --
// is left to gap_fill_kernel, which fills it with a whole grid instead of one thread.
constexpr int kGapInline = 64, kGapList = 1024;
struct Gap {
  unsigned long long b0, b1;  // buckets [b0, b1] get the value v
  uint32_t v;
};
__global__ void bucket_index_kernel(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t n_buckets,
                                    uint32_t* __restrict__ start, Gap* __restrict__ gaps, unsigned int* __restrict__ n_gaps) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;  // thread n closes the table
  const int64_t b_prev = i == 0 ? -1 : (int64_t)(keys[i - 1] >> shift);
  const int64_t b_cur = i == n ? (int64_t)n_buckets : (int64_t)(keys[i] >> shift);
  if (b_cur - b_prev > kGapInline) {
    const unsigned g = atomicAdd(n_gaps, 1u);
    if (g < kGapList) {
      gaps[g] = Gap{(unsigned long long)(b_prev + 1), (unsigned long long)b_cur, i};
      return;
    }
  }
  for (int64_t b = b_prev + 1; b <= b_cur; ++b) start[b] = i;
}

__global__ void gap_fill_kernel(const Gap* __restrict__ gaps, const unsigned int* __restrict__ n_gaps, uint32_t* __restrict__ start) {
  const unsigned ng = min(*n_gaps, (unsigned)kGapList);
  for (unsigned g = 0; g < ng; ++g) {
    const Gap gp = gaps[g];
    for (unsigned long long b = gp.b0 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= gp.b1;
         b += (unsigned long long)gridDim.x * blockDim.x)
      start[b] = gp.v;
  }
}

__device__ __forceinline__ uint32_t lower_bound_idx(const uint64_t* __restrict__ store,
                                                    const uint64_t* __restrict__ keys,
                                                    const uint64_t* __restrict__ locs, const BucketIndex& bi,
                                                    uint64_t xk, uint64_t xl) {
  const uint64_t b = xk >> bi.shift;
  return lower_bound_rec(store, keys, locs, bi.start[b], bi.start[b + 1], xk, xl);
}

// ---- seeds ------------------------------------------------------------------------------------
__global__ void seed_count_kernel(const uint16_t* __restrict__ clen, const uint16_t* __restrict__ nf,
                                  const uint16_t* __restrict__ nr, uint32_t n_reads, uint32_t* __restrict__ cnt) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_reads) cnt[r] = clen[r] ? (uint32_t)nf[r] + nr[r] : 0u;
}

__global__ void seed_emit_kernel(const uint64_t* __restrict__ store, uint64_t fwd_word_base, uint64_t rc_word_base,
                                 const uint32_t* __restrict__ word_off, const uint16_t* __restrict__ clen,
                                 const uint16_t* __restrict__ nf, const uint16_t* __restrict__ nr,
                                 const uint32_t* __restrict__ seed_off, uint32_t n_reads, uint64_t* __restrict__ keys,
                                 uint64_t* __restrict__ locs) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  int L = clen[r];
  if (!L) return;
  uint32_t o = seed_off[r];
  uint64_t fa = (fwd_word_base + word_off[r]) * 32, ra = (rc_word_base + word_off[r]) * 32;
  int f = nf[r], v = nr[r];
  // every seed but the last of its strand pops to the next seed (kLocPopSeed)
  for (int i = 0; i < f; ++i, ++o) {
    keys[o] = suffix_key(store, fa + i, L - i);
    locs[o] = make_loc(fa + i, L - i) | (i + 1 < f ? kLocPopSeed : 0ULL);
  }
  for (int i = 0; i < v; ++i, ++o) {
    keys[o] = suffix_key(store, ra + i, L - i);
    locs[o] = make_loc(ra + i, L - i) | (i + 1 < v ? kLocPopSeed : 0ULL);
  }
}

// ---- tie groups after the radix sort on the top `sbits` bits -------------------------------------
// One thread per record.  Every record measures its run of equal top bits (at most small_limit+1
// to either side), so no thread ever walks a long run: members of runs longer than small_limit
// flag themselves for the refinement path; the leader of a shorter run insertion-sorts it in
// place with the full comparator.
__global__ void __launch_bounds__(128) tie_small_kernel(const uint64_t* __restrict__ store,
                                                        uint64_t* __restrict__ keys, uint64_t* __restrict__ locs,
                                                        uint32_t n, int sbits, int small_limit,
                                                        uint32_t* __restrict__ big_flag,
                                                        unsigned long long* __restrict__ n_big) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int sh = 64 - sbits;  // sbits in [8,64]
  bool big = false;
  uint32_t left = 0, right = 0;
  if (i < n) {
    uint64_t top = keys[i] >> sh;
    while (left <= (uint32_t)small_limit && i > left && (keys[i - left - 1] >> sh) == top) ++left;
    while (right <= (uint32_t)small_limit && i + right + 1 < n && (keys[i + right + 1] >> sh) == top) ++right;
    big = left + right + 1 > (uint32_t)small_limit;
    if (big) big_flag[i] = 1;
  }
  unsigned bm = __ballot_sync(0xffffffffu, big);
  if (lane_id() == 0 && bm) atomicAdd(n_big, (unsigned long long)__popc(bm));
  if (i >= n || big || left != 0 || right == 0) return;
  const uint32_t g = right + 1;
  for (uint32_t a = 1; a < g; ++a) {
    uint64_t k = keys[i + a], l = locs[i + a];
    int j = (int)a - 1;
    while (j >= 0 && rec_less(store, k, l, keys[i + j], locs[i + j])) {
      keys[i + j + 1] = keys[i + j];
      locs[i + j + 1] = locs[i + j];
      --j;
    }
    keys[i + j + 1] = k;
    locs[i + j + 1] = l;
  }
}

// ---- refinement of big tie groups -------------------------------------------------------------------
// members are listed by ascending position (midx) with a non-decreasing dense group id (mgid).
// compound key = gid << 33 | chunk(12 bases at depth D) << 9 | min(len, D + 12)
__global__ void refine_init_kernel(const uint32_t* __restrict__ big_flag, const uint32_t* __restrict__ big_pos,
                                   const uint64_t* __restrict__ keys, uint32_t n, int sbits,
                                   uint32_t* __restrict__ midx, uint32_t* __restrict__ head) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !big_flag[i]) return;
  uint32_t j = big_pos[i];
  midx[j] = i;
  const int sh = 64 - sbits;
  bool is_head = (i == 0) || !big_flag[i - 1] || (keys[i - 1] >> sh) != (keys[i] >> sh);
  head[j] = is_head ? 1u : 0u;
}

// Medium tie groups: one warp per member, only the warp of a group's head works.  It measures
// the group (members are contiguous in the array), and if it has at most `limit` records sorts it
// by rank: every record counts the records that order before it (ties by position, so the result
// is stable), is written to its rank in the alt buffers and copied back.  Larger groups are
// left for the refinement rounds: tied = 1 on all their members.
__global__ void __launch_bounds__(128) tie_medium_kernel(const uint64_t* __restrict__ store, uint64_t* __restrict__ keys,
                                                         uint64_t* __restrict__ locs, uint64_t* __restrict__ keys_alt,
                                                         uint64_t* __restrict__ locs_alt,
                                                         const uint32_t* __restrict__ midx,
                                                         const uint32_t* __restrict__ head, uint32_t m, int limit,
                                                         uint32_t* __restrict__ tied, uint32_t* __restrict__ new_head) {
  const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= m || !head[j]) return;
  const unsigned lane = lane_id();
  // group size, measured up to limit + 1
  uint32_t g = 1;
  for (;;) {
    const uint32_t q = j + g + lane;
    const bool stop = q >= m || head[q] != 0;
    const unsigned sm = __ballot_sync(0xffffffffu, stop);
    if (sm) { g += __ffs(sm) - 1; break; }
    g += 32;
    if (g > (uint32_t)limit) break;
  }
  if (g > (uint32_t)limit) {
    // the whole group (however long) goes to the refinement path: every member marks itself
    // through its head's warp in strides, the head flag stays
    uint32_t q = j;
    for (;;) {
      const uint32_t t = q + lane;
      const bool in = t < m && (t == j || head[t] == 0);
      const unsigned im = __ballot_sync(0xffffffffu, in);
      // members are the leading run of lanes that are still inside the group
      const unsigned run = im == 0xffffffffu ? 32 : (unsigned)__ffs(~im) - 1;
      if (lane < run) { tied[t] = 1u; new_head[t] = t == j ? 1u : 0u; }
      if (run < 32) break;
      q += 32;
    }
    return;
  }
  const uint32_t pos0 = midx[j];
  for (uint32_t e = lane; e < g; e += 32) {
    const uint64_t ke = keys[pos0 + e], le = locs[pos0 + e];
    uint32_t rank = 0;
    for (uint32_t f = 0; f < g; ++f) {
      if (f == e) continue;
      const int cmp = rec_cmp(store, keys[pos0 + f], locs[pos0 + f], ke, le);
      rank += (cmp < 0 || (cmp == 0 && f < e)) ? 1u : 0u;
    }
    keys_alt[pos0 + rank] = ke;
    locs_alt[pos0 + rank] = le;
    tied[j + e] = 0u;
    new_head[j + e] = 0u;
  }
  __syncwarp();
  for (uint32_t e = lane; e < g; e += 32) {
    keys[pos0 + e] = keys_alt[pos0 + e];
    locs[pos0 + e] = locs_alt[pos0 + e];
  }
}

__global__ void refine_key_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ locs,
                                  const uint32_t* __restrict__ midx, const uint32_t* __restrict__ gid_incl /*head scan*/,
                                  const uint32_t* __restrict__ head, uint32_t m, int D, uint64_t* __restrict__ ckey,
                                  uint64_t* __restrict__ cloc) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  uint64_t l = locs[midx[j]];
  int len = (int)loc_len(l);
  uint64_t gid = gid_incl[j] + head[j] - 1;  // exclusive scan + own head - 1 => dense id of this record's group
  int rem = len - D;
  uint64_t chunk = 0;
  if (rem > 0) chunk = (load_window(store, loc_addr(l) + D) & top_bases_mask(min(rem, kChunkBases))) >> (64 - 2 * kChunkBases);
  uint64_t lf = (uint64_t)min(len, D + kChunkBases);
  ckey[j] = (gid << 33) | (chunk << 9) | lf;
  cloc[j] = l;
}

// after sorting (ckey, cloc): write records back to their positions and find the members that are
// still tied (equal compound key and both continue past D + chunk)
__global__ void refine_writeback_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ ckey,
                                        const uint64_t* __restrict__ cloc, const uint32_t* __restrict__ midx,
                                        uint32_t m, int D, uint64_t* __restrict__ keys, uint64_t* __restrict__ locs,
                                        uint32_t* __restrict__ tied, uint32_t* __restrict__ new_head) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  uint64_t l = cloc[j];
  uint32_t pos = midx[j];
  locs[pos] = l;
  keys[pos] = suffix_key(store, loc_addr(l), (int)loc_len(l));
  uint64_t ck = ckey[j];
  bool cont = (int)(ck & 0x1ff) == D + kChunkBases && (int)loc_len(l) > D + kChunkBases;
  bool same_prev = j > 0 && ckey[j - 1] == ck;
  bool same_next = j + 1 < m && ckey[j + 1] == ck;
  // records with equal compound keys that all end exactly at D+chunk are identical sequences;
  // a record ending exactly at D+chunk sorts before longer ones with the same chunk only if the
  // length field differs, which it does not -- so order them by full length in the next round too
  bool t = (same_prev || same_next) && (cont || (int)(ck & 0x1ff) == D + kChunkBases);
  tied[j] = t ? 1u : 0u;
  new_head[j] = (t && !same_prev) ? 1u : 0u;
}

__global__ void refine_compact_kernel(const uint32_t* __restrict__ tied, const uint32_t* __restrict__ tied_pos,
                                      const uint32_t* __restrict__ midx, const uint32_t* __restrict__ new_head,
                                      uint32_t m, uint32_t* __restrict__ midx2, uint32_t* __restrict__ head2) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m || !tied[j]) return;
  uint32_t o = tied_pos[j];
  midx2[o] = midx[j];
  head2[o] = new_head[j];
}

// ---- dedup ----------------------------------------------------------------------------------------
// `next` (optional) is the record that follows this rank's last one in the global order: the
// first record of the next non-empty rank (skip_dups' next_entry, bs/expand.cpp:37-47).
struct NextRec {
  int has;
  uint64_t key, loc;
};
__global__ void dedup_flag_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ keys,
                                  const uint64_t* __restrict__ locs, uint32_t n, NextRec next,
                                  uint32_t* __restrict__ keep) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool drop;
  if (i + 1 < n) drop = prefix_or_equal(store, keys[i], locs[i], keys[i + 1], locs[i + 1]);
  else drop = next.has && prefix_or_equal(store, keys[i], locs[i], next.key, next.loc);
  keep[i] = drop ? 0u : 1u;
}

// Flag, scan and compaction stay three steps on purpose: the flag step is bound by the random
// reads of the corrected store that decide "prefix of / equal to" for records whose 32-base keys
// agree (a third of the seeds at 30x), and wants one record per thread at full occupancy.  A
// single-pass version (tile-local ranks + decoupled look-back across tiles) was measured slower
// on B200 in every tile shape (chr20 30x, 155 M records: 7.9-12.5 ms against 6.4 ms): every tile
// is as slow as its slowest probe chain and holds up the look-back of the tiles behind it.
__global__ void compact_pairs_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ locs,
                                     const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, uint32_t n,
                                     uint64_t* __restrict__ okeys, uint64_t* __restrict__ olocs,
                                     const uint32_t* __restrict__ org_in, uint32_t* __restrict__ org_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  okeys[pos[i]] = keys[i];
  olocs[pos[i]] = locs[i];
  if (org_out != nullptr) org_out[pos[i]] = org_in != nullptr ? org_in[i] : i;
}

// ---- closure walk ------------------------------------------------------------------------------------
__device__ __forceinline__ bool covered(const uint64_t* __restrict__ store, const uint64_t* __restrict__ keys,
                                        const uint64_t* __restrict__ locs, uint32_t n, const BucketIndex& bi,
                                        uint64_t addr, int len, uint32_t* where) {
  if (len == 0) { if (where) *where = 0; return true; }  // the empty sequence is a prefix of everything
  uint64_t xk = suffix_key(store, addr, len), xl = make_loc(addr, len);
  uint32_t lb = lower_bound_idx(store, keys, locs, bi, xk, xl);
  if (where) *where = lb;
  if (lb >= n) return false;
  const uint64_t el = locs[lb];
  if ((el & ~kLocPopSeed) == xl) return true;  // the very same suffix of the same read (the common case): nothing to compare
  return prefix_or_equal(store, xk, xl, keys[lb], el);
}

// phase 1: one thread per entry; entries whose pop_front is not covered start a chain
// where1[i] = the first entry that has pop_front(entry i) as a prefix (kNoWhere when none does): the
// tables step re-bases it through the merge and dedup maps instead of searching again.
constexpr uint32_t kNoWhere = 0xffffffffu;
__global__ void __launch_bounds__(128) walk_phase1_kernel(const uint64_t* __restrict__ store,
                                                          const uint64_t* __restrict__ keys,
                                                          const uint64_t* __restrict__ locs, uint32_t n,
                                                          BucketIndex bi, uint32_t* __restrict__ chains,
                                                          unsigned long long* __restrict__ n_chains,
                                                          uint32_t* __restrict__ where1) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool start = false;
  if (i < n) {
    uint64_t l = locs[i];
    uint32_t where;
    start = !covered(store, keys, locs, n, bi, loc_addr(l) + 1, (int)loc_len(l) - 1, &where);
    where1[i] = start ? kNoWhere : where;
  }
  unsigned mask = __ballot_sync(0xffffffffu, start);
  if (!mask) return;
  unsigned lane = lane_id();
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(n_chains, (unsigned long long)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (start) chains[base + __popc(mask & ((1u << lane) - 1))] = i;
}

// phase 2: one warp per chain; lanes test pops j = 1 + lane + 32*iter in parallel; chain_stop =
// first covered j (or len).  Emits e[j:] for 1 <= j < chain_stop.
__global__ void __launch_bounds__(128) walk_phase2_kernel(const uint64_t* __restrict__ store,
                                                          const uint64_t* __restrict__ keys,
                                                          const uint64_t* __restrict__ locs, uint32_t n,
                                                          BucketIndex bi, const uint32_t* __restrict__ chains,
                                                          uint32_t n_chains, uint32_t* __restrict__ chain_stop) {
  uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n_chains) return;
  unsigned lane = lane_id();
  uint64_t l = locs[chains[c]];
  uint64_t addr = loc_addr(l);
  int len = (int)loc_len(l);
  int stop = len;
  for (int j0 = 2; j0 < len; j0 += 32) {  // j = 1 is known uncovered
    int j = j0 + (int)lane;
    bool cov = j < len && covered(store, keys, locs, n, bi, addr + j, len - j, nullptr);
    unsigned mask = __ballot_sync(0xffffffffu, cov);
    if (mask) { stop = j0 + __ffs(mask) - 1; break; }
  }
  if (lane == 0) chain_stop[c] = (uint32_t)(stop - 1);  // number of emitted suffixes
}

__global__ void walk_emit_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ locs,
                                 const uint32_t* __restrict__ chains, const uint32_t* __restrict__ chain_cnt,
                                 const uint32_t* __restrict__ chain_off, uint32_t n_chains, uint64_t* __restrict__ okeys,
                                 uint64_t* __restrict__ olocs) {
  uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n_chains) return;
  unsigned lane = lane_id();
  uint64_t l = locs[chains[c]];
  uint64_t addr = loc_addr(l);
  int len = (int)loc_len(l);
  uint32_t cnt = chain_cnt[c], off = chain_off[c];
  for (uint32_t t = lane; t < cnt; t += 32) {
    int j = 1 + (int)t;
    okeys[off + t] = suffix_key(store, addr + j, len - j);
    olocs[off + t] = make_loc(addr + j, len - j);
  }
}

// ---- read -> entry lookup (the lookup half of make_readmap) -----------------------------
// For every kept read: the id of the first entry that has the corrected read (resp. its reverse
// complement) as a prefix -- seqset::find_existing_unique as make_readmap calls it
// (modules/bio_mapred/make_readmap.cpp:137-167, modules/bio_base/seqset.cpp:173-188).  Every kept
// read is a seed, so the entry exists; a miss sets *missing ("... was not found in seqset").
__global__ void __launch_bounds__(128) lookup_reads_kernel(const uint64_t* __restrict__ store,
                                                           const uint64_t* __restrict__ keys,
                                                           const uint64_t* __restrict__ locs, uint32_t n, BucketIndex bi,
                                                           const uint32_t* __restrict__ word_off,
                                                           const uint16_t* __restrict__ clen, uint64_t rc_word_base,
                                                           uint32_t n_reads, unsigned long long* __restrict__ fwd_entry,
                                                           unsigned long long* __restrict__ rc_entry, int* __restrict__ missing) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const int L = clen[r];
  unsigned long long f = ~0ULL, v = ~0ULL;
  if (L) {
    uint32_t where;
    if (covered(store, keys, locs, n, bi, (uint64_t)word_off[r] * 32, L, &where)) f = where; else *missing = 1;
    if (covered(store, keys, locs, n, bi, (rc_word_base + word_off[r]) * 32, L, &where)) v = where; else *missing = 1;
  }
  fwd_entry[r] = f;
  rc_entry[r] = v;
}

// ---- readmap tables for unpaired reads (make_readmap::create_from_reads) --------------------
// Row key = entry_id << 12 | type << 10 | read_length: the order of make_readmap.h:187-205 for
// unpaired reads (mate_read_length is 0 throughout, and rows that tie on the key are identical:
// equal entry and length mean equal sequence, hence an equal loop entry).
constexpr unsigned long long kNoLoopEntry = (1ULL << 37) - 1;  // make_readmap::k_no_loop_entry

// Rows of one record (make_readmap.cpp:168-188).  Unpaired: a record is a read.  Paired: reads 2i and
// 2i+1 are mates; a pair with one read dropped is a single read, and the read with the smaller
// sequence is the LOOP_START (:170-175) -- sequence order is (entry id, length) order, because the
// entry of a read is the first entry it is a prefix of.
//   prim = entry << 12 | type << 10 | length      sec = mate_length << 37 | loop entry
// (prim, sec) ascending is the row order of make_readmap.h:187-205.
__global__ void readmap_rows_kernel(const unsigned long long* __restrict__ fwd_entry,
                                    const unsigned long long* __restrict__ rc_entry,
                                    const uint16_t* __restrict__ clen, const uint32_t* __restrict__ pos,
                                    uint32_t n_records, int paired, uint64_t* __restrict__ prim,
                                    uint64_t* __restrict__ sec) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_records) return;
  uint32_t a = paired ? 2 * i : i, b = paired ? 2 * i + 1 : i;
  unsigned long long la = clen[a], lb = paired ? clen[b] : 0ULL;
  if (!la && !lb) return;
  if (!la || (lb && (fwd_entry[a] > fwd_entry[b] || (fwd_entry[a] == fwd_entry[b] && la > lb)))) {
    const uint32_t t = a; a = b; b = t;
    const unsigned long long tl = la; la = lb; lb = tl;
  }
  const uint32_t o = pos[i];
  const unsigned long long e = fwd_entry[a], re = rc_entry[a];
  prim[o] = (e << 12) | (0ULL << 10) | la;              // LOOP_START -> its reverse complement
  sec[o] = re;
  prim[o + 1] = (re << 12) | (1ULL << 10) | la;         // RC -> the mate, if any
  if (!lb) {
    sec[o + 1] = kNoLoopEntry;
    return;
  }
  const unsigned long long me = fwd_entry[b], mre = rc_entry[b];
  sec[o + 1] = (lb << 37) | me;
  prim[o + 2] = (me << 12) | (2ULL << 10) | lb;         // MATE -> its reverse complement
  sec[o + 2] = mre;
  prim[o + 3] = (mre << 12) | (3ULL << 10) | lb;        // MATE_RC -> back to the LOOP_START (claim pass)
  sec[o + 3] = kNoLoopEntry;
}

// rows per record: 0, 2 or 4
__global__ void readmap_count_kernel(const uint16_t* __restrict__ clen, uint32_t n_records, int paired,
                                     uint32_t* __restrict__ cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_records) return;
  const int ka = clen[paired ? 2 * i : i] != 0, kb = paired ? clen[2 * i + 1] != 0 : 0;
  cnt[i] = 2u * (uint32_t)(ka + kb);
}

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* __restrict__ a, uint32_t n, uint64_t x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One thread per sorted row: sparse_multi bits (modules/io/sparse_multi.cpp:90-113), read_lengths,
// is_forward, and the start of every mate loop.  The sequential claim pass of make_readmap.cpp:302-360
// hands the rows of a run of identical rows out in the order in which the LOOP_START rows come by:
//   * the j-th of a run of identical LOOP_START rows takes the j-th RC row of (loop entry, length);
//     a single read's RC row points back, a pair's RC row becomes a claimer of its mate's MATE run
//   * readmap_claim_kernel: the claimers of a MATE run, sorted by LOOP_START row, take its rows in
//     order; the MATE_RC rows go to the same claimers in the same order and close the loops.
__global__ void __launch_bounds__(256) readmap_fill_kernel(const uint64_t* __restrict__ prim,
                                                           const uint64_t* __restrict__ sec, uint32_t m,
                                                           unsigned long long* __restrict__ src_bits,
                                                           uint32_t* __restrict__ dst_bits32,
                                                           uint32_t* __restrict__ fwd_bits32,
                                                           uint16_t* __restrict__ read_lengths,
                                                           unsigned long long* __restrict__ ptr,
                                                           uint64_t* __restrict__ claim_key, uint64_t* __restrict__ claim_val,
                                                           unsigned int* __restrict__ n_claims, int* __restrict__ missing) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool opens = false, is_fwd = false;
  if (i < m) {
    const uint64_t key = prim[i];
    const uint64_t entry = key >> 12;
    const unsigned type = (unsigned)(key >> 10) & 3u;
    const uint64_t len = key & 1023u;
    opens = i == 0 || (prim[i - 1] >> 12) != entry;
    is_fwd = type == 0 || type == 2;
    read_lengths[i] = (uint16_t)len;
    if (opens) atomicOr(&src_bits[entry >> 6], 1ULL << (entry & 63));
    if (type == 0) {
      const uint32_t rank = i - lower_bound_u64(prim, m, key);
      const uint64_t want = ((sec[i] & kNoLoopEntry) << 12) | (1ULL << 10) | len;
      const uint32_t rc = lower_bound_u64(prim, m, want) + rank;
      if (rc >= m || prim[rc] != want) {
        *missing = 1;
      } else {
        ptr[i] = rc;
        const uint64_t rsec = sec[rc];
        if ((rsec & kNoLoopEntry) == kNoLoopEntry) {
          ptr[rc] = i;  // no mate: just point back to the original
        } else {
          const uint64_t mwant = ((rsec & kNoLoopEntry) << 12) | (2ULL << 10) | (rsec >> 37);
          const uint32_t first_mate = lower_bound_u64(prim, m, mwant);
          const unsigned c = atomicAdd(n_claims, 1u);
          claim_key[c] = ((uint64_t)first_mate << 32) | i;
          claim_val[c] = rc;
        }
      }
    }
  }
  const unsigned d = __ballot_sync(0xffffffffu, opens), f = __ballot_sync(0xffffffffu, is_fwd);
  if (lane_id() == 0 && (i >> 5) < ((m + 31) >> 5)) {
    dst_bits32[i >> 5] = d;
    fwd_bits32[i >> 5] = f;
  }
}

__global__ void readmap_claim_kernel(const uint64_t* __restrict__ prim, const uint64_t* __restrict__ sec, uint32_t m,
                                     const uint64_t* __restrict__ claim_key, const uint64_t* __restrict__ claim_val,
                                     uint32_t n_claims, unsigned long long* __restrict__ ptr, int* __restrict__ missing) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_claims) return;
  const uint64_t ck = claim_key[j];
  const uint32_t first_mate = (uint32_t)(ck >> 32), ls = (uint32_t)ck, rc = (uint32_t)claim_val[j];
  const uint32_t rank = j - lower_bound_u64(claim_key, n_claims, (uint64_t)first_mate << 32);
  const uint64_t rsec = sec[rc];
  const uint64_t mlen = rsec >> 37;
  const uint32_t mate = first_mate + rank;
  const uint64_t mwant = ((rsec & kNoLoopEntry) << 12) | (2ULL << 10) | mlen;
  if (mate >= m || prim[mate] != mwant) { *missing = 1; return; }
  const uint64_t rwant = ((sec[mate] & kNoLoopEntry) << 12) | (3ULL << 10) | mlen;
  const uint32_t mrc = lower_bound_u64(prim, m, rwant) + rank;
  if (mrc >= m || prim[mrc] != rwant) { *missing = 1; return; }
  ptr[rc] = mate;
  ptr[mate] = mrc;
  ptr[mrc] = ls;  // save loop back to the beginning
}

// ---- merge of the (few) new records into the sorted survivors ----------------------------------------
// rank[j] = number of old records that sort before new record j; marks[r] counts the new records
// inserted in front of old record r.
__global__ void merge_rank_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ okeys,
                                  const uint64_t* __restrict__ olocs, uint32_t n_old, BucketIndex bi,
                                  const uint64_t* __restrict__ nkeys, const uint64_t* __restrict__ nlocs,
                                  uint32_t n_new, uint32_t* __restrict__ rank, uint32_t* __restrict__ marks) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_new) return;
  uint32_t r = lower_bound_idx(store, okeys, olocs, bi, nkeys[j], nlocs[j]);
  rank[j] = r;
  atomicAdd(&marks[r], 1u);
}

// old record i moves behind the new records ranked <= i: shift = exclusive_scan(marks)[i + 1]
__global__ void merge_scatter_old_kernel(const uint64_t* __restrict__ okeys, const uint64_t* __restrict__ olocs,
                                         uint32_t n_old, const uint32_t* __restrict__ marks_excl,
                                         uint64_t* __restrict__ out_keys, uint64_t* __restrict__ out_locs,
                                         uint32_t* __restrict__ org /*optional: merged position -> old index*/) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_old) return;
  uint32_t d = i + marks_excl[i + 1];
  out_keys[d] = okeys[i];
  out_locs[d] = olocs[i];
  if (org != nullptr) org[d] = i;
}

__global__ void merge_scatter_new_kernel(const uint64_t* __restrict__ nkeys, const uint64_t* __restrict__ nlocs,
                                         uint32_t n_new, const uint32_t* __restrict__ rank,
                                         uint64_t* __restrict__ out_keys, uint64_t* __restrict__ out_locs,
                                         uint32_t* __restrict__ org /*optional: kNoWhere marks a new record*/) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_new) return;
  uint32_t d = rank[j] + j;
  out_keys[d] = nkeys[j];
  out_locs[d] = nlocs[j];
  if (org != nullptr) org[d] = kNoWhere;
}

// ---- tables ---------------------------------------------------------------------------------------------
// The prev-bit target of entry i is the first entry that has pop_front(entry i) as a prefix.  The
// closure walk already found it for every round-1 entry (where1, an index into the round-1 array);
// `Reuse` carries that answer into the final array instead of searching again:
//   round-1 index w -> merged index w + (new records merged in before old w); the new records that
//   sort between the popped sequence x and old w were merged in right in front of old w and all have
//   x as a prefix too (anything between x and a sequence starting with x starts with x), so among them
//   the target is the first one not less than x; dropped records (dedup) are prefixes of their
//   successors, so the target moves on to the next kept one: pos[.].
// Entries that came out of the walk itself, and round-1 entries whose pop_front had no cover then,
// are searched as before.  With no reuse data (org == nullptr) every entry is searched.
struct Reuse {
  const uint32_t* org;         // final index -> round-1 index, kNoWhere for a walk record
  const uint32_t* where1;      // round-1 index -> round-1 index of the cover, kNoWhere if none
  const uint32_t* marks_excl;  // [r + 1] = new records merged in before old r (inclusive of those right in front); nullptr: no merge
  const uint32_t* pos;         // merged index -> final index of the first kept record at or after it; nullptr: no dedup
  const uint64_t* mkeys;       // the merged (pre-dedup) records
  const uint64_t* mlocs;
};
__global__ void __launch_bounds__(128) tables_kernel(const uint64_t* __restrict__ store,
                                                     const uint64_t* __restrict__ keys,
                                                     const uint64_t* __restrict__ locs, uint32_t n, BucketIndex bi,
                                                     Reuse ru, uint16_t* __restrict__ sizes, uint16_t* __restrict__ shared,
                                                     unsigned long long* __restrict__ prev_bits, uint64_t prev_words,
                                                     unsigned int* __restrict__ max_len, int* __restrict__ missing) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned len = 0;
  if (i < n) {
    uint64_t k = keys[i], l = locs[i];
    len = loc_len(l);
    sizes[i] = (uint16_t)len;
    // shared = LCP with the previous entry (bs/builder.cpp:72-80)
    int lcp = 0;
    if (i > 0) {
      uint64_t pk = keys[i - 1], pl = locs[i - 1];
      int m = min((int)len, (int)loc_len(pl));
      uint64_t x = (k ^ pk) & top_bases_mask(min(m, 32));
      if (x) lcp = __clzll(x) >> 1;
      else if (m <= 32) lcp = m;
      else compare_from(store, pl, l, 32, &lcp);
    }
    shared[i] = (uint16_t)lcp;
    // prev bit: first entry having pop_front(e) as a prefix (bs/builder.cpp:85-107)
    uint32_t where = kNoWhere;
    if (ru.org != nullptr && len > 1) {
      const uint32_t o = ru.org[i];
      const uint32_t w = o == kNoWhere ? kNoWhere : ru.where1[o];
      if (w != kNoWhere) {
        uint32_t wm = w;
        if (ru.marks_excl != nullptr) {
          const uint32_t before = ru.marks_excl[w + 1], run = before - ru.marks_excl[w];
          wm = w + before;
          if (run) {
            // new records right in front of the old cover: the first one not less than x wins
            const uint64_t xa = loc_addr(l) + 1;
            const uint64_t xk = suffix_key(store, xa, (int)len - 1), xl = make_loc(xa, len - 1);
            wm = lower_bound_rec(store, ru.mkeys, ru.mlocs, wm - run, wm, xk, xl);
          }
        }
        where = ru.pos != nullptr ? ru.pos[wm] : wm;
      }
    }
    bool cov = where != kNoWhere;
    if (!cov) cov = covered(store, keys, locs, n, bi, loc_addr(l) + 1, (int)len - 1, &where);
    if (!cov) {
      *missing = 1;  // LOG(FATAL) << "Missing expansion?" (bs/builder.cpp:96)
    } else {
      unsigned b = (unsigned)(k >> 62);
      atomicOr(&prev_bits[(uint64_t)b * prev_words + (where >> 6)], 1ULL << (where & 63));
    }
  }
  unsigned mx = __reduce_max_sync(0xffffffffu, len);
  if (lane_id() == 0 && mx) atomicMax(max_len, mx);
}

// merge mode: the prev bit of entry i where seqset_merger::merge_range puts it (merge_core.cuh,
// modules/bio_base/seqset_merger.cpp:109-197 over generate_chunks(0, n, nsplits)); thread per entry
__global__ void __launch_bounds__(128) merge_prev_kernel(const uint64_t* __restrict__ store,
                                                         const uint64_t* __restrict__ keys,
                                                         const uint64_t* __restrict__ locs, uint32_t n, BucketIndex bi,
                                                         uint64_t nsplits, unsigned long long* __restrict__ prev_bits,
                                                         uint64_t prev_words, int* __restrict__ missing) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t l = locs[i];
  const int xn = (int)loc_len(l) - 1;
  uint32_t lo0 = 0, hi0 = n;
  if (xn > 0) {  // narrow the search to the bucket of x's leading bases
    const uint64_t b = suffix_key(store, loc_addr(l) + 1, xn) >> bi.shift;
    lo0 = bi.start[b];
    hi0 = bi.start[b + 1];
  }
  const uint32_t t = mergecore::merge_prev_target(store, locs, n, i, lo0, hi0, nsplits);
  if (t == mergecore::kNone) { *missing = 1; return; }
  mergecore::or_bit(prev_bits + (uint64_t)(keys[i] >> 62) * prev_words, t);
}

// bitcount::finalize (modules/io/bitcount.cpp:84-123): per 512-bit group popcounts
__global__ void bitcount_groups_kernel(const unsigned long long* __restrict__ bits, uint64_t words, uint64_t groups,
                                       uint32_t* __restrict__ group_pop, unsigned long long* __restrict__ subaccum) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  unsigned long long sub = 0;
  uint32_t tot = 0;
  for (int j = 0; j < 8; ++j) {
    uint64_t wi = g * 8 + j;
    sub <<= 8;  // missing words of the last group keep shifting: left-justified (bitcount.cpp:108-113)
    if (wi < words) {
      unsigned c = __popcll(bits[wi]);
      sub |= c;
      tot += c;
    }
  }
  group_pop[g] = tot;
  subaccum[g] = sub;
}

__global__ void bitcount_accum_kernel(const uint32_t* __restrict__ group_excl, const uint32_t* __restrict__ total,
                                      uint64_t groups, uint64_t acc_words, unsigned long long* __restrict__ accum) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= acc_words) return;
  accum[g] = g < groups ? group_excl[g] : *total;  // extra slot only when nbits % 512 == 0
}

__global__ void entries_ascii_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ locs,
                                     const uint64_t* __restrict__ offs, uint64_t first, uint64_t count,
                                     char* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint64_t l = locs[first + i];
  uint64_t a = loc_addr(l);
  int len = (int)loc_len(l);
  char* o = out + offs[i];
  for (int j = 0; j < len; ++j) {
    uint64_t p = a + j;
    o[j] = "ACGT"[(store[p >> 5] >> (62 - 2 * (p & 31))) & 3];
  }
}

__global__ void entry_offs_kernel(const uint64_t* __restrict__ locs, uint64_t first, uint64_t count,
                                  uint32_t* __restrict__ lens) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) lens[i] = loc_len(locs[first + i]);
}

// ==== multi-GPU: routing by splitters, sentinel-aware lookups ========================================
// Rank d owns the records x with splitter[d-1] <= x < splitter[d] in the sequence order.  Before
// anything is sorted the splitters are synthetic bucket boundaries (key = first 8 bases, length
// 0: such a record never reads the store); once every rank holds a sorted range they are the
// ranks' first entries.
constexpr int kMaxRanks = 64;
struct Splitters {
  int n;  // nranks - 1
  uint64_t key[kMaxRanks - 1], loc[kMaxRanks - 1];
};

__device__ __forceinline__ int route_dest(const uint64_t* __restrict__ store, const Splitters& sp, uint64_t k, uint64_t l) {
  int d = 0;
  while (d < sp.n && sp.loc[d] != ~0ULL /* +inf: an empty rank */ && !rec_less(store, k, l, sp.key[d], sp.loc[d])) ++d;
  return d;
}

// histogram of the first 8 bases (65536 buckets): the input of the balanced split
// (the reference pre-sizes its sort sections the same way, bs/part_counts.h:17-32)
__global__ void bucket_hist_kernel(const uint64_t* __restrict__ keys, uint32_t n, unsigned long long* __restrict__ hist) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[keys[i] >> 48], 1ULL);
}

__global__ void __launch_bounds__(256) route_count_kernel(const uint64_t* __restrict__ store,
                                                          const uint64_t* __restrict__ keys,
                                                          const uint64_t* __restrict__ locs, uint32_t n, Splitters sp,
                                                          uint8_t* __restrict__ dest,
                                                          unsigned long long* __restrict__ counts) {
  __shared__ unsigned int sc[kMaxRanks];
  if (threadIdx.x < kMaxRanks) sc[threadIdx.x] = 0;
  __syncthreads();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    // bits 14..15 of a routed loc may carry a tag (prev-bit queries): not part of the length
    int d = route_dest(store, sp, keys[i], locs[i] & ~0xC000ULL);
    dest[i] = (uint8_t)d;
    atomicAdd(&sc[d], 1u);
  }
  __syncthreads();
  if (threadIdx.x <= (unsigned)sp.n && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
}

// per-destination output arrays: the local staging buffer at the destination's offset, or -- the
// default -- the destination rank's receive buffer itself (peer mapping), at this rank's offset in it
struct RouteOut {
  uint64_t* keys[kMaxRanks];
  uint64_t* locs[kMaxRanks];
};
__global__ void __launch_bounds__(256) route_scatter_kernel(const uint64_t* __restrict__ keys,
                                                            const uint64_t* __restrict__ locs,
                                                            const uint8_t* __restrict__ dest, uint32_t n, int nranks,
                                                            unsigned long long* __restrict__ cursors /*zero*/,
                                                            RouteOut out) {
  __shared__ unsigned int sc[kMaxRanks];
  __shared__ unsigned long long sbase[kMaxRanks];
  if (threadIdx.x < kMaxRanks) sc[threadIdx.x] = 0;
  __syncthreads();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned r = 0;
  int d = 0;
  if (i < n) {
    d = dest[i];
    r = atomicAdd(&sc[d], 1u);
  }
  __syncthreads();
  if (threadIdx.x < (unsigned)nranks && sc[threadIdx.x])
    sbase[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
  __syncthreads();
  if (i < n) {
    unsigned long long o = sbase[d] + r;
    out.keys[d][o] = keys[i];
    out.locs[d][o] = locs[i];
  }
}

// covered() against this rank's sorted range plus the record that follows it globally
__device__ __forceinline__ bool covered_next(const uint64_t* __restrict__ store, const uint64_t* __restrict__ keys,
                                             const uint64_t* __restrict__ locs, uint32_t n, const BucketIndex& bi,
                                             const NextRec& next, uint64_t xk, uint64_t xl, uint32_t* where) {
  uint32_t lb = lower_bound_idx(store, keys, locs, bi, xk, xl);
  *where = lb;
  if (lb < n) {
    const uint64_t el = locs[lb];
    return (el & ~kLocPopSeed) == xl || prefix_or_equal(store, xk, xl, keys[lb], el);
  }
  return next.has && prefix_or_equal(store, xk, xl, next.key, next.loc);
}

// queries = pop_front of every local entry (entries of length 1 pop to the empty sequence, which
// every entry covers: skipped).  tag_base: write the entry's first base into loc bits 14..15.
__global__ void pop_queries_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ keys,
                                   const uint64_t* __restrict__ locs, uint32_t n, int tag_base,
                                   uint64_t* __restrict__ qkeys, uint64_t* __restrict__ qlocs,
                                   unsigned long long* __restrict__ n_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool has = false;
  uint64_t qk = 0, ql = 0;
  if (i < n) {
    uint64_t l = locs[i];
    int len = (int)loc_len(l);
    // closure walk (no tag): an entry whose pop_front was a seed itself needs no answer
    if (len > 1 && (tag_base || !(l & kLocPopSeed))) {
      has = true;
      qk = suffix_key(store, loc_addr(l) + 1, len - 1);
      ql = make_loc(loc_addr(l) + 1, len - 1);
      if (tag_base) ql |= (keys[i] >> 62) << 14;
    }
  }
  unsigned mask = __ballot_sync(0xffffffffu, has);
  if (!mask) return;
  unsigned lane = lane_id();
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(n_out, (unsigned long long)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (has) {
    unsigned long long o = base + __popc(mask & ((1u << lane) - 1));
    qkeys[o] = qk;
    qlocs[o] = ql;
  }
}

// flag[i] = 1 if record i is NOT covered by the local range (+ next); cnt[i] = how many records
// it will emit (all its suffixes when emit_suffixes, else itself)
__global__ void __launch_bounds__(128) uncovered_kernel(const uint64_t* __restrict__ store,
                                                        const uint64_t* __restrict__ keys,
                                                        const uint64_t* __restrict__ locs, uint32_t n, BucketIndex bi,
                                                        NextRec next, const uint64_t* __restrict__ xkeys,
                                                        const uint64_t* __restrict__ xlocs, uint32_t m,
                                                        int emit_suffixes, uint32_t* __restrict__ cnt,
                                                        uint32_t* __restrict__ list,
                                                        unsigned long long* __restrict__ n_list) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  bool unc = false;
  if (j < m) {
    uint32_t where;
    unc = !covered_next(store, keys, locs, n, bi, next, xkeys[j], xlocs[j], &where);
    cnt[j] = unc ? (emit_suffixes ? loc_len(xlocs[j]) : 1u) : 0u;
  }
  // the (few) uncovered ones go on a list so the emit kernel only visits them
  unsigned mask = __ballot_sync(0xffffffffu, unc);
  if (!mask) return;
  unsigned lane = lane_id();
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(n_list, (unsigned long long)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (unc) list[base + __popc(mask & ((1u << lane) - 1))] = j;
}

// emit record j itself (and, when emit_suffixes, every further suffix of it) at off[j]
__global__ void emit_uncovered_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ xkeys,
                                      const uint64_t* __restrict__ xlocs, const uint32_t* __restrict__ cnt,
                                      const uint32_t* __restrict__ off, const uint32_t* __restrict__ list,
                                      uint32_t n_list, uint64_t* __restrict__ okeys, uint64_t* __restrict__ olocs) {
  uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_list) return;
  uint32_t j = list[w];
  uint32_t c = cnt[j];
  uint64_t addr = loc_addr(xlocs[j]);
  int len = (int)loc_len(xlocs[j]);
  uint32_t o = off[j];
  for (uint32_t t = lane_id(); t < c; t += 32) {
    okeys[o + t] = suffix_key(store, addr + t, len - (int)t);
    olocs[o + t] = make_loc(addr + t, len - (int)t);
  }
}

// sizes / shared / max length of this rank's range; `prev` = the entry before the range globally
__global__ void __launch_bounds__(128) tables_local_kernel(const uint64_t* __restrict__ store,
                                                           const uint64_t* __restrict__ keys,
                                                           const uint64_t* __restrict__ locs, uint32_t n, NextRec prev,
                                                           uint16_t* __restrict__ sizes, uint16_t* __restrict__ shared,
                                                           unsigned int* __restrict__ max_len) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned len = 0;
  if (i < n) {
    uint64_t k = keys[i], l = locs[i];
    len = loc_len(l);
    sizes[i] = (uint16_t)len;
    int lcp = 0;
    if (i > 0 || prev.has) {
      uint64_t pk = i > 0 ? keys[i - 1] : prev.key, pl = i > 0 ? locs[i - 1] : prev.loc;
      int m = min((int)len, (int)loc_len(pl));
      uint64_t x = (k ^ pk) & top_bases_mask(min(m, 32));
      if (x) lcp = __clzll(x) >> 1;
      else if (m <= 32) lcp = m;
      else compare_from(store, pl, l, 32, &lcp);
    }
    shared[i] = (uint16_t)lcp;
  }
  unsigned mx = __reduce_max_sync(0xffffffffu, len);
  if (lane_id() == 0 && mx) atomicMax(max_len, mx);
}

// routed prev-bit queries: x = pop_front(e) tagged with e's first base.  Sets prev[b][where] for
// the first local entry having x as a prefix; where == n means the next rank's entry 0 (carry).
__global__ void __launch_bounds__(128) prev_apply_kernel(const uint64_t* __restrict__ store,
                                                         const uint64_t* __restrict__ keys,
                                                         const uint64_t* __restrict__ locs, uint32_t n, BucketIndex bi,
                                                         NextRec next, const uint64_t* __restrict__ xkeys,
                                                         const uint64_t* __restrict__ xlocs, uint32_t m,
                                                         unsigned long long* __restrict__ prev_bits, uint64_t prev_words,
                                                         int* __restrict__ carry /*[4]*/, int* __restrict__ missing) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  uint64_t xl = xlocs[j];
  unsigned b = (unsigned)((xl >> 14) & 3);
  xl &= ~0xC000ULL;
  uint32_t where;
  if (!covered_next(store, keys, locs, n, bi, next, xkeys[j], xl, &where)) {
    *missing = 1;  // LOG(FATAL) << "Missing expansion?" (bs/builder.cpp:96)
  } else if (where < n) {
    atomicOr(&prev_bits[(uint64_t)b * prev_words + (where >> 6)], 1ULL << (where & 63));
  } else {
    carry[b] = 1;
  }
}

// entries of length 1 pop to the empty sequence: it is a prefix of the globally first entry
__global__ void single_base_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ locs, uint32_t n,
                                   int* __restrict__ flags /*[4]*/) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && loc_len(locs[i]) == 1) flags[keys[i] >> 62] = 1;
}

__global__ void set_bit0_kernel(unsigned long long* __restrict__ prev_bits, uint64_t prev_words, int b) {
  atomicOr(&prev_bits[(uint64_t)b * prev_words], 1ULL);
}

// accum with the ones of earlier ranks added (bitcount::finalize over the global bit vector)
__global__ void bitcount_accum_offset_kernel(const uint32_t* __restrict__ group_excl, const uint32_t* __restrict__ total,
                                             uint64_t groups, uint64_t acc_words, unsigned long long before,
                                             unsigned long long* __restrict__ accum) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= acc_words) return;
  accum[g] = before + (g < groups ? group_excl[g] : *total);
}

// packed_varbit_vector elements (modules/io/packed_varbit_vector.cpp:80-139): value i occupies bits
// [i*b, (i+1)*b) of the little-endian word stream.  One thread per output word.
__global__ void varbit_pack_kernel(const uint16_t* __restrict__ vals, uint64_t n, unsigned b, uint64_t n_words,
                                   unsigned long long* __restrict__ out) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const uint64_t bit0 = w * 64;
  unsigned long long acc = 0;
  for (uint64_t i = bit0 / b; i < n && i * b < bit0 + 64; ++i) {
    const unsigned long long v = vals[i];
    const int64_t sh = (int64_t)(i * b) - (int64_t)bit0;
    acc |= sh >= 0 ? (v << sh) : (v >> (-sh));
  }
  out[w] = acc;
}

__global__ void iota_u32_kernel(uint32_t* __restrict__ out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)std::max<uint64_t>(1, (n + block - 1) / block); }

uint32_t read_u32(const uint32_t* d, cudaStream_t s) {
  uint32_t h;
  BGX_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  return h;
}
unsigned long long read_u64(const unsigned long long* d, cudaStream_t s) {
  unsigned long long h;
  BGX_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  return h;
}

// bucket index over n sorted records: about 8 records per bucket, at most 2^24 buckets
BucketIndex build_bucket_index(Context* c, const uint64_t* keys, uint32_t n, DevBuf<uint32_t>& buf) {
  cudaStream_t s = c->stream;
  int bits = 0;
  while ((2ull << bits) <= (uint64_t)std::max<uint32_t>(n, 1)) ++bits;  // floor(log2(n))
  bits = std::max(4, std::min(24, bits - 3));
  if (const char* e = getenv("BGX_INDEX_BITS")) bits = std::max(1, std::min(28, atoi(e)));  // experiment hook
  const uint32_t nb = 1u << bits;
  buf.alloc((size_t)nb + 1, s);
  DevBuf<Gap> gaps(kGapList, s);
  DevBuf<unsigned int> n_gaps(1, s);
  BGX_CUDA(cudaMemsetAsync(n_gaps.p, 0, sizeof(unsigned int), s));
  KLAUNCH(bucket_index_kernel)<<<grid_for((uint64_t)n + 1, 256), 256, 0, s>>>(keys, n, 64 - bits, nb, buf.p, gaps.p, n_gaps.p);
  KLAUNCH(gap_fill_kernel)<<<kNumSMs * 4, 256, 0, s>>>(gaps.p, n_gaps.p, buf.p);
  BGX_CUDA(cudaGetLastError());
  return BucketIndex{buf.p, 64 - bits};
}

// Sort n records completely.  keys/locs and the alt buffers all hold >= n elements; the sorted
// data ends in (keys, locs).
void sort_records(Context* c, DevBuf<uint64_t>& keys, DevBuf<uint64_t>& locs, DevBuf<uint64_t>& keys_alt,
                  DevBuf<uint64_t>& locs_alt, uint32_t n, const std::string& tag) {
  cudaStream_t s = c->stream;
  if (n == 0) return;
  // key bits the radix passes sort on.  Auto: 2 bases more than log4(n), rounded up to whole passes
  // -- a few percent of the records then share their top bits with a non-duplicate (duplicates
  // and repeats tie at any width; the tie kernels resolve both), and every 8 bits less is one
  // pass over all records less.  Measured, radix + ties: E. coli 100x 2.29 ms at 48 bits, 2.06 at
  // 32; chr20 30x 16.1 ms at 48, 14.6 at 40, 13.1 at 32.
  int sbits = c->opt.sort_key_bits;
  if (sbits == 0) {
    int bases = 2;
    while (bases < 32 && (1ull << (2 * (bases - 2))) < (uint64_t)n) ++bases;
    sbits = std::max(24, std::min(64, (2 * bases + 7) / 8 * 8));
  }
  if (const char* e = getenv("BGX_SORT_BITS")) sbits = std::max(16, std::min(64, atoi(e) / 8 * 8));  // experiment hook
  int passes = 0;
  {
    ScopedStage st(c, "sort_radix");
    bool alt = radix_sort_pairs(keys.p, locs.p, keys_alt.p, locs_alt.p, n, 64 - sbits, 64, s, &passes);
    if (alt) { std::swap(keys, keys_alt); std::swap(locs, locs_alt); }
    st.stop();
  }
  c->add_stat("sort_radix_passes", passes);
  c->add_stat("sort_radix_records", (double)n * passes);
  c->add_stat("alg_bytes_sort_radix", radix_pass_alg_bytes(n) * passes);
  c->add_stat("useful_bytes_sort", 14.0 * n);

  ScopedStage st(c, "sort_ties");
  DevBuf<uint32_t> big_flag(n, s);
  DevBuf<unsigned long long> n_big(1, s);
  BGX_CUDA(cudaMemsetAsync(big_flag.p, 0, (size_t)n * 4, s));
  BGX_CUDA(cudaMemsetAsync(n_big.p, 0, 8, s));
  int small_limit = kSmallGroup;
  if (const char* e = getenv("BGX_SMALL_GROUP")) small_limit = std::max(1, atoi(e));  // test hook: force the refinement path
  KLAUNCH(tie_small_kernel)<<<grid_for(n, 128), 128, 0, s>>>(c->seq_store(), keys.p, locs.p, n, sbits, small_limit, big_flag.p,
                                                   n_big.p);
  BGX_CUDA(cudaGetLastError());
  uint32_t m = (uint32_t)read_u64(n_big.p, s);
  c->add_stat("tie_big_records_" + tag, m);
  if (m) {
    // refinement rounds over the members of big groups
    DevBuf<uint32_t> big_pos(n, s);
    exclusive_scan_u32(big_flag.p, big_pos.p, n, nullptr, s);
    DevBuf<uint32_t> midx(m, s), head(m, s);
    KLAUNCH(refine_init_kernel)<<<grid_for(n, 256), 256, 0, s>>>(big_flag.p, big_pos.p, keys.p, n, sbits, midx.p, head.p);
    {
      // medium groups are finished by one warp each; what is left goes through the refinement rounds
      int medium_limit = kMediumGroup;
      if (const char* e = getenv("BGX_MEDIUM_GROUP")) medium_limit = std::max(0, atoi(e));  // test hook (0: refinement only)
      DevBuf<uint32_t> tied(m, s), nhead(m, s), tpos(m, s), tot(1, s);
      KLAUNCH(tie_medium_kernel)<<<grid_for((uint64_t)m * 32, 128), 128, 0, s>>>(c->seq_store(), keys.p, locs.p, keys_alt.p, locs_alt.p,
                                                                        midx.p, head.p, m, medium_limit, tied.p, nhead.p);
      exclusive_scan_u32(tied.p, tpos.p, m, tot.p, s);
      BGX_CUDA(cudaGetLastError());
      uint32_t m2 = read_u32(tot.p, s);
      if (m2 != m) {
        DevBuf<uint32_t> midx2(std::max<uint32_t>(m2, 1), s), head2(std::max<uint32_t>(m2, 1), s);
        if (m2) KLAUNCH(refine_compact_kernel)<<<grid_for(m, 256), 256, 0, s>>>(tied.p, tpos.p, midx.p, nhead.p, m, midx2.p, head2.p);
        midx = std::move(midx2);
        head = std::move(head2);
      }
      c->add_stat("tie_refined_records_" + tag, m2);
      m = m2;
    }
    int D = sbits / 2;
    int rounds = 0;
    while (m) {
      DevBuf<uint32_t> gid(m, s);
      exclusive_scan_u32(head.p, gid.p, m, nullptr, s);
      DevBuf<uint64_t> ck(m, s), cl(m, s), ck2(m, s), cl2(m, s);
      KLAUNCH(refine_key_kernel)<<<grid_for(m, 256), 256, 0, s>>>(c->seq_store(), locs.p, midx.p, gid.p, head.p, m, D, ck.p, cl.p);
      bool alt = radix_sort_pairs(ck.p, cl.p, ck2.p, cl2.p, m, 0, 64, s);
      DevBuf<uint32_t> tied(m, s), nhead(m, s), tpos(m, s), tot(1, s);
      KLAUNCH(refine_writeback_kernel)<<<grid_for(m, 256), 256, 0, s>>>(c->seq_store(), alt ? ck2.p : ck.p, alt ? cl2.p : cl.p, midx.p,
                                                              m, D, keys.p, locs.p, tied.p, nhead.p);
      exclusive_scan_u32(tied.p, tpos.p, m, tot.p, s);
      BGX_CUDA(cudaGetLastError());
      uint32_t m2 = read_u32(tot.p, s);
      DevBuf<uint32_t> midx2(std::max<uint32_t>(m2, 1), s), head2(std::max<uint32_t>(m2, 1), s);
      if (m2)
        KLAUNCH(refine_compact_kernel)<<<grid_for(m, 256), 256, 0, s>>>(tied.p, tpos.p, midx.p, nhead.p, m, midx2.p, head2.p);
      midx = std::move(midx2);
      head = std::move(head2);
      m = m2;
      D += kChunkBases;
      ++rounds;
      BGX_CHECK(rounds < 64, "tie refinement did not converge");
    }
    c->add_stat("tie_refine_rounds_" + tag, rounds);
  }
  st.stop();
}

// drop every record that is a prefix of / equal to its successor; returns the survivor count and
// leaves them in (keys, locs) (buffers are swapped with the alt ones).
uint32_t dedup_records(Context* c, DevBuf<uint64_t>& keys, DevBuf<uint64_t>& locs, DevBuf<uint64_t>& keys_alt,
                       DevBuf<uint64_t>& locs_alt, uint32_t n, NextRec next = NextRec{0, 0, 0},
                       DevBuf<uint32_t>* org = nullptr /*in/out: origin of every record*/,
                       DevBuf<uint32_t>* pos_out = nullptr /*out: record -> index of the first survivor at or after it*/) {
  cudaStream_t s = c->stream;
  if (n == 0) return 0;
  ScopedStage st(c, "dedup");
  DevBuf<uint32_t> keep(n, s), pos(n, s), tot(1, s), org2;
  if (org != nullptr) org2.alloc(n, s);
  KLAUNCH(dedup_flag_kernel)<<<grid_for(n, 256), 256, 0, s>>>(c->seq_store(), keys.p, locs.p, n, next, keep.p);
  exclusive_scan_u32(keep.p, pos.p, n, tot.p, s);
  KLAUNCH(compact_pairs_kernel)<<<grid_for(n, 256), 256, 0, s>>>(keys.p, locs.p, keep.p, pos.p, n, keys_alt.p, locs_alt.p,
                                                         org != nullptr ? org->p : nullptr, org != nullptr ? org2.p : nullptr);
  BGX_CUDA(cudaGetLastError());
  uint32_t m = read_u32(tot.p, s);
  std::swap(keys, keys_alt);
  std::swap(locs, locs_alt);
  if (org != nullptr) *org = std::move(org2);
  if (pos_out != nullptr) *pos_out = std::move(pos);
  c->add_stat("alg_bytes_dedup", 16.0 * n + 16.0 * m);
  st.stop();
  return m;
}

// merge n_new sorted records into the n1 sorted records of (keys, locs) by rank; the result
// (n1 + n_new records) ends in (keys, locs); the alt buffers are grown to hold as many.
void merge_new_records(Context* c, DevBuf<uint64_t>& keys, DevBuf<uint64_t>& locs, DevBuf<uint64_t>& keys_alt,
                       DevBuf<uint64_t>& locs_alt, uint32_t n1, const BucketIndex& bi, const DevBuf<uint64_t>& nkeys,
                       const DevBuf<uint64_t>& nlocs, uint32_t n_new, DevBuf<uint32_t>* marks_out = nullptr,
                       DevBuf<uint32_t>* org_out = nullptr) {
  cudaStream_t s = c->stream;
  const uint32_t nm = n1 + n_new;
  ScopedStage st(c, "merge");
  if ((size_t)nm > keys_alt.n) {
    keys_alt.alloc((size_t)nm + 1024, s);
    locs_alt.alloc((size_t)nm + 1024, s);
  }
  DevBuf<uint32_t> rank(n_new, s), marks((size_t)n1 + 2, s);
  BGX_CUDA(cudaMemsetAsync(marks.p, 0, ((size_t)n1 + 2) * 4, s));
  KLAUNCH(merge_rank_kernel)<<<grid_for(n_new, 128), 128, 0, s>>>(c->seq_store(), keys.p, locs.p, n1, bi, nkeys.p, nlocs.p, n_new,
                                                         rank.p, marks.p);
  exclusive_scan_u32(marks.p, marks.p, (size_t)n1 + 2, nullptr, s);
  if (org_out != nullptr) org_out->alloc(nm, s);
  uint32_t* org = org_out != nullptr ? org_out->p : nullptr;
  if (n1) KLAUNCH(merge_scatter_old_kernel)<<<grid_for(n1, 256), 256, 0, s>>>(keys.p, locs.p, n1, marks.p, keys_alt.p, locs_alt.p, org);
  KLAUNCH(merge_scatter_new_kernel)<<<grid_for(n_new, 256), 256, 0, s>>>(nkeys.p, nlocs.p, n_new, rank.p, keys_alt.p, locs_alt.p, org);
  BGX_CUDA(cudaGetLastError());
  if (marks_out != nullptr) *marks_out = std::move(marks);
  std::swap(keys, keys_alt);
  std::swap(locs, locs_alt);
  if ((size_t)nm > keys_alt.n) {  // dedup writes into the alt buffers
    keys_alt.alloc((size_t)nm + 1024, s);
    locs_alt.alloc((size_t)nm + 1024, s);
  }
  c->add_stat("alg_bytes_merge", 32.0 * nm + 8.0 * n1);
  st.stop();
}

}  // namespace

void stage_build_seqset_dist(Context* c);

void stage_build_seqset(Context* c) {
  BGX_CHECK(c->corrected, "bgx_build_seqset: call bgx_correct first");
  // Large inputs: the count table and the solid set (69 GB each per GPU at GRCh38 30x on 8 GPUs) are
  // dead weight from here on; give them back so the corrected store and the records fit.
  // Small inputs keep them, so bgx_export_kmers still works after the build.
  if (c->table.bytes() + c->solid.bytes() > c->total_mem / 4) {
    c->table.release();
    c->solid.release();
    c->counted = false;
    c->set_stat("kmer_tables_released", 1);
  }
  if (c->dist.nranks > 1) return stage_build_seqset_dist(c);
  cudaStream_t s = c->stream;
  ScopedStage st_all(c, "seqset_total");
  const uint32_t n_reads = (uint32_t)c->n_reads;
  BGX_CHECK(c->n_seeds < kMaxShardRecords, "too many seed records for one GPU shard");

  // 1. seeds
  uint32_t n = (uint32_t)c->n_seeds;
  size_t cap = (size_t)n + n / 2 + 1024;
  DevBuf<uint64_t> keys(cap, s), locs(cap, s), keys_alt(cap, s), locs_alt(cap, s);
  {
    ScopedStage st(c, "seed_emit");
    DevBuf<uint32_t> cnt(n_reads, s), off(n_reads, s);
    KLAUNCH(seed_count_kernel)<<<grid_for(n_reads, 256), 256, 0, s>>>(c->clen.p, c->next_fwd.p, c->next_rev.p, n_reads, cnt.p);
    exclusive_scan_u32(cnt.p, off.p, n_reads, nullptr, s);
    KLAUNCH(seed_emit_kernel)<<<grid_for(n_reads, 128), 128, 0, s>>>(c->store.p, 0, c->n_words, c->word_off.p, c->clen.p,
                                                            c->next_fwd.p, c->next_rev.p, off.p, n_reads, keys.p, locs.p);
    BGX_CUDA(cudaGetLastError());
    st.stop();
  }

  build_seqset_from_records(c, keys, locs, keys_alt, locs_alt, n, nullptr);
  st_all.stop();
}

// Steps 2-5 on n (key, loc) records over c->store: what `biograph create` runs after seeding, and what a
// seqset merge runs on one record per input entry (merge.cu; `mh` non-null).
void build_seqset_from_records(Context* c, DevBuf<uint64_t>& keys, DevBuf<uint64_t>& locs, DevBuf<uint64_t>& keys_alt,
                               DevBuf<uint64_t>& locs_alt, uint32_t n, const MergeHooks* mh) {
  cudaStream_t s = c->stream;
  // 2. sort + dedup
  sort_records(c, keys, locs, keys_alt, locs_alt, n, "r1");
  DevBuf<uint32_t> pos1;
  uint32_t n1 = dedup_records(c, keys, locs, keys_alt, locs_alt, n, NextRec{0, 0, 0}, nullptr, mh != nullptr ? &pos1 : nullptr);
  c->set_stat("entries_round1", n1);
  // merge: the alt buffers still hold the sorted records before the dedup, pos1 where each one went
  if (mh != nullptr && n) mh->after_dedup(locs_alt.p, n, pos1.p, n1);
  pos1.release();

  // 3. closure walk: the new records are emitted into their own buffers
  uint32_t n_new = 0;
  DevBuf<uint64_t> nkeys, nlocs;
  DevBuf<uint32_t> index_buf;
  DevBuf<uint32_t> where1, marks_excl, org, pos2;   // the walk's answers and the maps that re-base them (tables_kernel)
  BucketIndex bi1;
  {
    ScopedStage st(c, "walk");
    bi1 = build_bucket_index(c, keys.p, n1, index_buf);
    DevBuf<uint32_t> chains(std::max<uint32_t>(n1, 1), s);
    DevBuf<unsigned long long> n_chains_d(1, s);
    BGX_CUDA(cudaMemsetAsync(n_chains_d.p, 0, 8, s));
    where1.alloc(std::max<uint32_t>(n1, 1), s);
    KLAUNCH(walk_phase1_kernel)<<<grid_for(n1, 128), 128, 0, s>>>(c->store.p, keys.p, locs.p, n1, bi1, chains.p, n_chains_d.p,
                                                          where1.p);
    BGX_CUDA(cudaGetLastError());
    uint32_t n_chains = (uint32_t)read_u64(n_chains_d.p, s);
    c->set_stat("walk_chains", n_chains);
    if (n_chains) {
      DevBuf<uint32_t> ccnt(n_chains, s), coff(n_chains, s), tot(1, s);
      KLAUNCH(walk_phase2_kernel)<<<grid_for((uint64_t)n_chains * 32, 128), 128, 0, s>>>(c->store.p, keys.p, locs.p, n1, bi1, chains.p,
                                                                                n_chains, ccnt.p);
      exclusive_scan_u32(ccnt.p, coff.p, n_chains, tot.p, s);
      BGX_CUDA(cudaGetLastError());
      n_new = read_u32(tot.p, s);
      BGX_CHECK((uint64_t)n1 + n_new < kMaxShardRecords, "too many records for one GPU shard");
      nkeys.alloc(std::max<uint32_t>(n_new, 1), s);
      nlocs.alloc(std::max<uint32_t>(n_new, 1), s);
      KLAUNCH(walk_emit_kernel)<<<grid_for((uint64_t)n_chains * 32, 128), 128, 0, s>>>(c->store.p, locs.p, chains.p, ccnt.p, coff.p,
                                                                              n_chains, nkeys.p, nlocs.p);
      BGX_CUDA(cudaGetLastError());
    }
    st.stop();
  }
  c->set_stat("walk_new_records", n_new);
  // the union of seqsets is closed under pop_front (the cover of a popped entry inside its own part is a
  // prefix of / equal to an entry of the union), so a merge has nothing to add
  BGX_CHECK(mh == nullptr || n_new == 0, "bgx_merge_seqsets: an input is not closed under pop_front (not a seqset)");

  // 4. sort the new records alone, merge them into the sorted survivors by rank, dedup
  uint32_t n2 = n1;
  if (n_new) {
    {
      DevBuf<uint64_t> nkeys_alt(n_new, s), nlocs_alt(n_new, s);
      sort_records(c, nkeys, nlocs, nkeys_alt, nlocs_alt, n_new, "r2");
    }
    const uint32_t nm = n1 + n_new;
    merge_new_records(c, keys, locs, keys_alt, locs_alt, n1, bi1, nkeys, nlocs, n_new, &marks_excl, &org);
    n2 = dedup_records(c, keys, locs, keys_alt, locs_alt, nm, NextRec{0, 0, 0}, &org, &pos2);
  }
  c->n_entries = n2;
  c->n_entries_global = n2;
  c->first_entry_global = 0;
  c->set_stat("entries", n2);

  // 5. tables
  {
    ScopedStage st(c, "tables");
    const uint64_t nb = n2;
    c->prev_words = (nb + 63) / 64;
    c->sub_words = (nb + 511) / 512;
    c->acc_words = (nb + 1 + 511) / 512;
    c->sizes.alloc(std::max<uint64_t>(nb, 1), s);
    c->shared.alloc(std::max<uint64_t>(nb, 1), s);
    c->prev_bits.alloc(std::max<uint64_t>(4 * c->prev_words, 1), s);
    c->prev_sub.alloc(std::max<uint64_t>(4 * c->sub_words, 1), s);
    c->prev_acc.alloc(std::max<uint64_t>(4 * c->acc_words, 1), s);
    BGX_CUDA(cudaMemsetAsync(c->prev_bits.p, 0, std::max<uint64_t>(4 * c->prev_words, 1) * 8, s));
    BGX_CUDA(cudaMemsetAsync(c->prev_acc.p, 0, std::max<uint64_t>(4 * c->acc_words, 1) * 8, s));
    DevBuf<unsigned int> max_len(1, s);
    DevBuf<int> missing(1, s);
    BGX_CUDA(cudaMemsetAsync(max_len.p, 0, 4, s));
    BGX_CUDA(cudaMemsetAsync(missing.p, 0, 4, s));
    if (nb) {
      const BucketIndex bi2 = n2 == n1 && n_new == 0 ? bi1 : build_bucket_index(c, keys.p, n2, index_buf);
      Reuse ru{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
      static const bool reuse = [] { const char* e = getenv("BGX_TABLES_REUSE"); return e ? atoi(e) != 0 : true; }();  // A/B hook
      if (reuse && n_new == 0) {
        // nothing was merged or dropped after the walk: final index == round-1 index
        DevBuf<uint32_t>& ident = org;
        ident.alloc(n2, s);
        KLAUNCH(iota_u32_kernel)<<<grid_for(n2, 256), 256, 0, s>>>(ident.p, n2);
        ru = Reuse{ident.p, where1.p, nullptr, nullptr, nullptr, nullptr};
      } else if (reuse) {
        ru = Reuse{org.p, where1.p, marks_excl.p, pos2.p, keys_alt.p, locs_alt.p};   // the alt buffers still hold the merged records
      }
      KLAUNCH(tables_kernel)<<<grid_for(nb, 128), 128, 0, s>>>(c->store.p, keys.p, locs.p, n2, bi2, ru, c->sizes.p, c->shared.p,
                                                      reinterpret_cast<unsigned long long*>(c->prev_bits.p),
                                                      c->prev_words, max_len.p, missing.p);
      if (mh != nullptr && mh->parallel_splits != 1) {
        // seqset_merger's placement of the prev bits (chunk rule, merge_core.cuh) instead of the builder's
        BGX_CUDA(cudaMemsetAsync(c->prev_bits.p, 0, std::max<uint64_t>(4 * c->prev_words, 1) * 8, s));
        KLAUNCH(merge_prev_kernel)<<<grid_for(nb, 128), 128, 0, s>>>(c->store.p, keys.p, locs.p, n2, bi2, mh->parallel_splits,
                                                            reinterpret_cast<unsigned long long*>(c->prev_bits.p),
                                                            c->prev_words, missing.p);
      }
      DevBuf<uint32_t> gpop(c->sub_words, s), gex(c->sub_words, s), tot(1, s);
      uint64_t off = 0;
      for (int b = 0; b < 4; ++b) {
        KLAUNCH(bitcount_groups_kernel)<<<grid_for(c->sub_words, 256), 256, 0, s>>>(
            reinterpret_cast<unsigned long long*>(c->prev_bits.p) + b * c->prev_words, c->prev_words, c->sub_words, gpop.p,
            reinterpret_cast<unsigned long long*>(c->prev_sub.p) + b * c->sub_words);
        exclusive_scan_u32(gpop.p, gex.p, c->sub_words, tot.p, s);
        KLAUNCH(bitcount_accum_kernel)<<<grid_for(c->acc_words, 256), 256, 0, s>>>(
            gex.p, tot.p, c->sub_words, c->acc_words, reinterpret_cast<unsigned long long*>(c->prev_acc.p) + b * c->acc_words);
        BGX_CUDA(cudaGetLastError());
        c->fixed[b] = off;
        off += read_u32(tot.p, s);
      }
      c->fixed[4] = off;
    } else {
      for (int b = 0; b < 5; ++b) c->fixed[b] = 0;
    }
    unsigned int h_max;
    int h_missing;
    BGX_CUDA(cudaMemcpyAsync(&h_max, max_len.p, 4, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaMemcpyAsync(&h_missing, missing.p, 4, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    c->max_entry_len = h_max;
    BGX_CHECK(!h_missing, "Missing expansion?");  // bs/builder.cpp:96
    // seqset::finalize (seqset.cpp:123-126)
    BGX_CHECK(c->fixed[4] == nb, "Invalid seqset in finalize: prev bit totals != entries");
    st.stop();
  }
  c->ent_key = std::move(keys);
  c->ent_loc = std::move(locs);
  c->built = true;
}

// ==== multi-GPU seqset build =============================================================================
namespace {

struct Routed {
  DevBuf<uint64_t> keys, locs;
  uint32_t n = 0;
};

// Send every record to the rank that owns it (all-to-all over NVLink); returns what this rank got.
Routed route_records(Context* c, const uint64_t* keys, const uint64_t* locs, uint32_t n, const Splitters& sp) {
  cudaStream_t s = c->stream;
  const int N = c->dist.nranks, R = c->dist.rank;
  ScopedStage st(c, "route");
  DevBuf<uint8_t> dest(std::max<uint32_t>(n, 1), s);
  DevBuf<unsigned long long> counts(kMaxRanks, s);
  BGX_CUDA(cudaMemsetAsync(counts.p, 0, kMaxRanks * 8, s));
  ScopedStage st_k(c, "route_kernels");
  if (n) KLAUNCH(route_count_kernel)<<<grid_for(n, 256), 256, 0, s>>>(c->seq_store(), keys, locs, n, sp, dest.p, counts.p);
  std::vector<uint64_t> send_cnt(N), send_off(N), all((size_t)N * N), recv_cnt(N), recv_off(N);
  {
    unsigned long long h[kMaxRanks];
    BGX_CUDA(cudaMemcpyAsync(h, counts.p, N * 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    uint64_t o = 0;
    for (int r = 0; r < N; ++r) { send_cnt[r] = h[r]; send_off[r] = o; o += h[r]; }
  }
  st_k.stop();
  ScopedStage st_x(c, "route_exchange");
  dist_allgather_host_u64(c, send_cnt.data(), N, all.data());
  uint64_t m = 0;
  for (int src = 0; src < N; ++src) { recv_cnt[src] = all[(size_t)src * N + R]; recv_off[src] = m; m += recv_cnt[src]; }
  BGX_CHECK(m < kMaxShardRecords, "too many records for one GPU shard");
  Routed out;
  out.n = (uint32_t)m;
  out.keys.alloc(m + 1024, s);  // slack: the sort / dedup ping-pong buffers are sized alike
  out.locs.alloc(m + 1024, s);
  BGX_CUDA(cudaMemsetAsync(counts.p, 0, kMaxRanks * 8, s));   // the scatter's cursors
  RouteOut ro;
  const Exchange mode = dist_exchange_mode();
  if (mode == Exchange::STORE) {
    // the scatter stores every record straight into its owner's receive buffer over NVLink: this rank's
    // records start where the ranks before it end (the count matrix is known to everyone)
    void* local[2] = {out.keys.p, out.locs.p};
    void* peer[2 * kMaxRanks];
    dist_map_peers_n(c, local, 2, peer);   // also: every owner's buffers are allocated
    for (int d = 0; d < N; ++d) {
      uint64_t off = 0;
      for (int src = 0; src < R; ++src) off += all[(size_t)src * N + d];
      ro.keys[d] = static_cast<uint64_t*>(peer[d]) + off;
      ro.locs[d] = static_cast<uint64_t*>(peer[N + d]) + off;
    }
    if (n) KLAUNCH(route_scatter_kernel)<<<grid_for(n, 256), 256, 0, s>>>(keys, locs, dest.p, n, N, counts.p, ro);
    BGX_CUDA(cudaGetLastError());
    dist_barrier(c);   // every rank's stores have landed
  } else {
    // scatter into local send arrays grouped by destination, then move the groups
    DevBuf<uint64_t> skeys(std::max<uint32_t>(n, 1), s), slocs(std::max<uint32_t>(n, 1), s);
    for (int d = 0; d < N; ++d) { ro.keys[d] = skeys.p + send_off[d]; ro.locs[d] = slocs.p + send_off[d]; }
    if (n) KLAUNCH(route_scatter_kernel)<<<grid_for(n, 256), 256, 0, s>>>(keys, locs, dest.p, n, N, counts.p, ro);
    BGX_CUDA(cudaGetLastError());
    if (mode == Exchange::COPY) {
      // one copy-engine transfer per peer and array, into the owner's receive arrays where the ranks
      // before this one end
      void* local[2] = {out.keys.p, out.locs.p};
      void* peer[2 * kMaxRanks];
      dist_map_peers_n(c, local, 2, peer);   // also: every owner's buffers are allocated
      std::vector<PeerCopy> copies;
      for (int d = 0; d < N; ++d) {
        uint64_t off = 0;
        for (int src = 0; src < R; ++src) off += all[(size_t)src * N + d];
        PeerCopy a, b;
        a.dst = static_cast<uint64_t*>(peer[d]) + off;      a.src = skeys.p + send_off[d];
        b.dst = static_cast<uint64_t*>(peer[N + d]) + off;  b.src = slocs.p + send_off[d];
        a.bytes = b.bytes = send_cnt[d] * 8;
        a.peer = b.peer = d;
        copies.push_back(a);
        copies.push_back(b);
      }
      dist_peer_copies(c, copies);
      dist_barrier(c);   // every rank's records have arrived (and the send arrays may go)
    } else {
      dist_alltoallv(c, skeys.p, send_off.data(), send_cnt.data(), out.keys.p, recv_off.data(), recv_cnt.data(), 8);
      dist_alltoallv(c, slocs.p, send_off.data(), send_cnt.data(), out.locs.p, recv_off.data(), recv_cnt.data(), 8);
      BGX_CUDA(cudaStreamSynchronize(s));  // the send buffers die with this scope
    }
  }
  st_x.stop();
  c->add_stat("route_records_out", (double)n - (double)send_cnt[R]);
  c->add_stat("route_bytes_out", 16.0 * ((double)n - (double)send_cnt[R]));
  st.stop();
  return out;
}

struct RankEnds {  // first and last record of every rank's sorted range
  std::vector<uint64_t> v;  // [rank][5] = has, first key, first loc, last key, last loc
  bool has(int r) const { return v[(size_t)r * 5] != 0; }
  NextRec first(int r) const { return NextRec{1, v[(size_t)r * 5 + 1], v[(size_t)r * 5 + 2]}; }
  NextRec last(int r) const { return NextRec{1, v[(size_t)r * 5 + 3], v[(size_t)r * 5 + 4]}; }
};

RankEnds exchange_ends(Context* c, const DevBuf<uint64_t>& keys, const DevBuf<uint64_t>& locs, uint32_t n) {
  cudaStream_t s = c->stream;
  uint64_t mine[5] = {n ? 1ull : 0ull, 0, 0, 0, 0};
  if (n) {
    BGX_CUDA(cudaMemcpyAsync(&mine[1], keys.p, 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaMemcpyAsync(&mine[2], locs.p, 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaMemcpyAsync(&mine[3], keys.p + (n - 1), 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaMemcpyAsync(&mine[4], locs.p + (n - 1), 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
  }
  RankEnds e;
  e.v.resize((size_t)c->dist.nranks * 5);
  dist_allgather_host_u64(c, mine, 5, e.v.data());
  return e;
}

// the record that follows this rank's range in the global order / the one that precedes it
NextRec next_of(const RankEnds& e, int R, int N) {
  for (int r = R + 1; r < N; ++r)
    if (e.has(r)) return e.first(r);
  return NextRec{0, 0, 0};
}
NextRec prev_of(const RankEnds& e, int R) {
  for (int r = R - 1; r >= 0; --r)
    if (e.has(r)) return e.last(r);
  return NextRec{0, 0, 0};
}

}  // namespace

void stage_build_seqset_dist(Context* c) {
  BGX_CHECK(c->corrected, "bgx_build_seqset: call bgx_correct first");
  cudaStream_t s = c->stream;
  const int N = c->dist.nranks, R = c->dist.rank;
  ScopedStage st_all(c, "seqset_total");
  const uint32_t n_reads = (uint32_t)c->n_reads;
  BGX_CHECK(c->n_seeds < kMaxShardRecords, "too many seed records for one GPU shard");

  // 0. replicate the corrected stores: comparisons past the 32-base key read the sequence, and a
  //    record may be compared on any rank.  Every rank's store sits in an equal-sized slice (the
  //    largest rank's size), so ONE in-place ncclAllGather moves everything (NVSwitch multicast /
  //    ring at NVLink rate) instead of N-1 send/recv pairs.
  {
    ScopedStage st(c, "store_allgather");
    uint64_t mine = 2 * c->n_words + 1;
    std::vector<uint64_t> words(N);
    dist_allgather_host_u64(c, &mine, 1, words.data());
    uint64_t slice = 0;
    for (int r = 0; r < N; ++r) slice = std::max(slice, words[r]);
    slice = (slice + 1) & ~1ull;  // 16-byte multiple
    const uint64_t tot = slice * N;
    BGX_CHECK(tot * 32 < (1ull << 47), "corrected store too large for 48-bit suffix locators");
    c->gstore.alloc(tot + 1, s);
    uint64_t* my_slice = c->gstore.p + (uint64_t)R * slice;
    BGX_CUDA(cudaMemcpyAsync(my_slice, c->store.p, mine * 8, cudaMemcpyDeviceToDevice, s));
    if (slice > mine) BGX_CUDA(cudaMemsetAsync(my_slice + mine, 0, (slice - mine) * 8, s));
    BGX_CUDA(cudaMemsetAsync(c->gstore.p + tot, 0, 8, s));
    dist_allgather_bytes(c, my_slice, c->gstore.p, slice * 8);
    c->gstore_word_base = (uint64_t)R * slice;
    c->add_stat("store_allgather_bytes", 8.0 * (double)tot);
    st.stop();
  }
  const uint64_t* store = c->seq_store();

  // 1. seeds of this rank's reads, addressed in the replicated store
  uint32_t n = (uint32_t)c->n_seeds;
  DevBuf<uint64_t> keys(std::max<uint32_t>(n, 1), s), locs(std::max<uint32_t>(n, 1), s), keys_alt, locs_alt;
  {
    ScopedStage st(c, "seed_emit");
    DevBuf<uint32_t> cnt(std::max<uint32_t>(n_reads, 1), s), off(std::max<uint32_t>(n_reads, 1), s);
    if (n_reads) {
      KLAUNCH(seed_count_kernel)<<<grid_for(n_reads, 256), 256, 0, s>>>(c->clen.p, c->next_fwd.p, c->next_rev.p, n_reads, cnt.p);
      exclusive_scan_u32(cnt.p, off.p, n_reads, nullptr, s);
      KLAUNCH(seed_emit_kernel)<<<grid_for(n_reads, 128), 128, 0, s>>>(store, c->gstore_word_base, c->gstore_word_base + c->n_words,
                                                              c->word_off.p, c->clen.p, c->next_fwd.p, c->next_rev.p, off.p,
                                                              n_reads, keys.p, locs.p);
    }
    BGX_CUDA(cudaGetLastError());
    st.stop();
  }

  // 2. balanced split of the prefix space (histogram of the first 8 bases over all ranks) and
  //    routing of every seed to the owner of its prefix bucket
  Splitters sp;
  sp.n = N - 1;
  {
    ScopedStage st(c, "split");
    const size_t NB = 65536;
    DevBuf<unsigned long long> hist(NB, s);
    BGX_CUDA(cudaMemsetAsync(hist.p, 0, NB * 8, s));
    if (n) KLAUNCH(bucket_hist_kernel)<<<grid_for(n, 256), 256, 0, s>>>(keys.p, n, hist.p);
    dist_allreduce_sum_u64(c, hist.p, NB);
    std::vector<unsigned long long> h(NB);
    BGX_CUDA(cudaMemcpyAsync(h.data(), hist.p, NB * 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    uint64_t total = 0;
    for (size_t b = 0; b < NB; ++b) total += h[b];
    uint64_t cum = 0;
    size_t b = 0;
    for (int r = 1; r < N; ++r) {
      const uint64_t target = total / N * r;
      while (b < NB && cum + h[b] <= target) cum += h[b++];  // bucket b is the first one of rank r
      sp.key[r - 1] = b >= NB ? ~0ULL : ((uint64_t)b << 48);
      sp.loc[r - 1] = b >= NB ? ~0ULL : make_loc(0, 0);
    }
    st.stop();
  }
  {
    Routed r1 = route_records(c, keys.p, locs.p, n, sp);
    keys = std::move(r1.keys);
    locs = std::move(r1.locs);
    n = r1.n;
  }
  c->set_stat("seeds_owned", n);
  keys_alt.alloc((size_t)n + 1024, s);
  locs_alt.alloc((size_t)n + 1024, s);
  if (keys.n < (size_t)n + 1024) {  // sort/dedup ping-pong between equally sized buffers
    DevBuf<uint64_t> k2((size_t)n + 1024, s), l2((size_t)n + 1024, s);
    if (n) {
      BGX_CUDA(cudaMemcpyAsync(k2.p, keys.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
      BGX_CUDA(cudaMemcpyAsync(l2.p, locs.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    }
    keys = std::move(k2);
    locs = std::move(l2);
  }

  // 3. sort + dedup of the owned range; the last record is checked against the next rank's first
  sort_records(c, keys, locs, keys_alt, locs_alt, n, "r1");
  RankEnds ends = exchange_ends(c, keys, locs, n);
  uint32_t n1 = dedup_records(c, keys, locs, keys_alt, locs_alt, n, next_of(ends, R, N));
  c->set_stat("entries_round1", n1);

  // 4. closure: route pop_front(e) of every entry to the owner of its prefix; an uncovered one and
  //    all of its suffixes become candidates, routed to THEIR owners, who keep the uncovered ones.
  //    One round reaches the fixed point: every suffix of anything new was itself a candidate.
  uint32_t n_new = 0;
  DevBuf<uint64_t> nkeys, nlocs;
  DevBuf<uint32_t> index_buf;
  BucketIndex bi1;
  {
    ScopedStage st(c, "walk");
    ends = exchange_ends(c, keys, locs, n1);
    NextRec next = next_of(ends, R, N);
    bi1 = build_bucket_index(c, keys.p, n1, index_buf);
    DevBuf<uint64_t> qk(std::max<uint32_t>(n1, 1), s), ql(std::max<uint32_t>(n1, 1), s);
    DevBuf<unsigned long long> nq_d(1, s);
    BGX_CUDA(cudaMemsetAsync(nq_d.p, 0, 8, s));
    ScopedStage st_q(c, "walk_queries");
    if (n1) KLAUNCH(pop_queries_kernel)<<<grid_for(n1, 256), 256, 0, s>>>(store, keys.p, locs.p, n1, 0, qk.p, ql.p, nq_d.p);
    uint32_t nq = (uint32_t)read_u64(nq_d.p, s);
    st_q.stop();
    Routed q = route_records(c, qk.p, ql.p, nq, sp);
    qk.release();
    ql.release();
    // uncovered queries -> themselves + all their suffixes
    ScopedStage st_u(c, "walk_uncovered");
    DevBuf<uint32_t> cnt(std::max<uint32_t>(q.n, 1), s), off(std::max<uint32_t>(q.n, 1), s), list(std::max<uint32_t>(q.n, 1), s), tot(1, s);
    DevBuf<unsigned long long> nl_d(1, s);
    BGX_CUDA(cudaMemsetAsync(nl_d.p, 0, 8, s));
    if (q.n) KLAUNCH(uncovered_kernel)<<<grid_for(q.n, 128), 128, 0, s>>>(store, keys.p, locs.p, n1, bi1, next, q.keys.p, q.locs.p, q.n,
                                                                  1, cnt.p, list.p, nl_d.p);
    exclusive_scan_u32(cnt.p, off.p, q.n, tot.p, s);
    uint32_t n_cand = read_u32(tot.p, s);
    uint32_t n_list = (uint32_t)read_u64(nl_d.p, s);
    st_u.stop();
    c->set_stat("walk_chains", n_list);
    c->set_stat("walk_candidates", n_cand);
    ScopedStage st_c(c, "walk_candidates");
    DevBuf<uint64_t> ck(std::max<uint32_t>(n_cand, 1), s), cl(std::max<uint32_t>(n_cand, 1), s);
    if (n_list) KLAUNCH(emit_uncovered_kernel)<<<grid_for((uint64_t)n_list * 32, 128), 128, 0, s>>>(store, q.keys.p, q.locs.p, cnt.p,
                                                                                         off.p, list.p, n_list, ck.p, cl.p);
    BGX_CUDA(cudaGetLastError());
    Routed cd = route_records(c, ck.p, cl.p, n_cand, sp);
    // owners keep the candidates nothing covers yet
    DevBuf<uint32_t> cnt2(std::max<uint32_t>(cd.n, 1), s), off2(std::max<uint32_t>(cd.n, 1), s), list2(std::max<uint32_t>(cd.n, 1), s);
    BGX_CUDA(cudaMemsetAsync(nl_d.p, 0, 8, s));
    if (cd.n) KLAUNCH(uncovered_kernel)<<<grid_for(cd.n, 128), 128, 0, s>>>(store, keys.p, locs.p, n1, bi1, next, cd.keys.p, cd.locs.p,
                                                                    cd.n, 0, cnt2.p, list2.p, nl_d.p);
    exclusive_scan_u32(cnt2.p, off2.p, cd.n, tot.p, s);
    n_new = read_u32(tot.p, s);
    nkeys.alloc(std::max<uint32_t>(n_new, 1), s);
    nlocs.alloc(std::max<uint32_t>(n_new, 1), s);
    if (n_new) KLAUNCH(emit_uncovered_kernel)<<<grid_for((uint64_t)n_new * 32, 128), 128, 0, s>>>(store, cd.keys.p, cd.locs.p, cnt2.p,
                                                                                        off2.p, list2.p, n_new, nkeys.p, nlocs.p);
    BGX_CUDA(cudaGetLastError());
    st_c.stop();
    st.stop();
  }
  c->set_stat("walk_new_records", n_new);
  BGX_CHECK((uint64_t)n1 + n_new < kMaxShardRecords, "too many records for one GPU shard");
  uint32_t n2 = n1;
  {
    // every rank takes part in the exchanges below even with nothing new of its own
    if (n_new) {
      DevBuf<uint64_t> nkeys_alt(n_new, s), nlocs_alt(n_new, s);
      sort_records(c, nkeys, nlocs, nkeys_alt, nlocs_alt, n_new, "r2");
      merge_new_records(c, keys, locs, keys_alt, locs_alt, n1, bi1, nkeys, nlocs, n_new);
    }
    ends = exchange_ends(c, keys, locs, n1 + n_new);
    n2 = dedup_records(c, keys, locs, keys_alt, locs_alt, n1 + n_new, next_of(ends, R, N));
  }

  // 5. rebalance: rank r takes the global entries [r*chunk, (r+1)*chunk), chunk a multiple of 512,
  //    so every rank's bit vectors start on a bitcount group boundary and concatenate as they are
  uint64_t Nt = 0, first_global = 0;
  {
    ScopedStage st(c, "rebalance");
    uint64_t mine = n2;
    std::vector<uint64_t> cnts(N), goff(N + 1);
    dist_allgather_host_u64(c, &mine, 1, cnts.data());
    goff[0] = 0;
    for (int r = 0; r < N; ++r) goff[r + 1] = goff[r] + cnts[r];
    Nt = goff[N];
    const uint64_t chunk = std::max<uint64_t>(512, ((Nt + N - 1) / N + 511) / 512 * 512);
    auto lo_of = [&](int r) { return std::min<uint64_t>((uint64_t)r * chunk, Nt); };
    std::vector<uint64_t> send_off(N), send_cnt(N), recv_off(N), recv_cnt(N);
    uint64_t m = 0;
    for (int d = 0; d < N; ++d) {
      uint64_t a = std::max<uint64_t>(goff[R], lo_of(d)), b = std::min<uint64_t>(goff[R + 1], lo_of(d + 1));
      send_cnt[d] = b > a ? b - a : 0;
      send_off[d] = b > a ? a - goff[R] : 0;
    }
    for (int src = 0; src < N; ++src) {
      uint64_t a = std::max<uint64_t>(goff[src], lo_of(R)), b = std::min<uint64_t>(goff[src + 1], lo_of(R + 1));
      recv_cnt[src] = b > a ? b - a : 0;
      recv_off[src] = m;
      m += recv_cnt[src];
    }
    // the ping-pong partners and the walk's records are not needed any more: give them back first
    // (GRCh38 30x on 8 GPUs: 12 GB per rank, the difference between fitting and not)
    keys_alt.release();
    locs_alt.release();
    nkeys.release();
    nlocs.release();
    DevBuf<uint64_t> k2(std::max<uint64_t>(m, 1), s), l2(std::max<uint64_t>(m, 1), s);
    dist_alltoallv(c, keys.p, send_off.data(), send_cnt.data(), k2.p, recv_off.data(), recv_cnt.data(), 8);
    dist_alltoallv(c, locs.p, send_off.data(), send_cnt.data(), l2.p, recv_off.data(), recv_cnt.data(), 8);
    BGX_CUDA(cudaStreamSynchronize(s));
    keys = std::move(k2);
    locs = std::move(l2);
    keys_alt.release();
    locs_alt.release();
    n2 = (uint32_t)m;
    first_global = lo_of(R);
    st.stop();
  }
  c->n_entries = n2;
  c->n_entries_global = Nt;
  c->first_entry_global = first_global;
  c->set_stat("entries", n2);
  c->set_stat("entries_global", (double)Nt);

  // 6. tables of the owned range
  {
    ScopedStage st(c, "tables");
    ends = exchange_ends(c, keys, locs, n2);
    const NextRec next = next_of(ends, R, N), prev = prev_of(ends, R);
    int last_owner = 0;
    for (int r = 0; r < N; ++r)
      if (ends.has(r)) last_owner = r;
    Splitters sp2;
    sp2.n = N - 1;
    for (int r = 1; r < N; ++r) {
      sp2.key[r - 1] = ends.has(r) ? ends.first(r).key : ~0ULL;
      sp2.loc[r - 1] = ends.has(r) ? ends.first(r).loc : ~0ULL;
    }
    const uint64_t nb = n2;
    c->prev_words = (nb + 63) / 64;
    c->sub_words = (nb + 511) / 512;
    c->acc_words = R == last_owner ? (Nt + 1 + 511) / 512 - first_global / 512 : c->sub_words;
    c->sizes.alloc(std::max<uint64_t>(nb, 1), s);
    c->shared.alloc(std::max<uint64_t>(nb, 1), s);
    c->prev_bits.alloc(std::max<uint64_t>(4 * c->prev_words, 1), s);
    c->prev_sub.alloc(std::max<uint64_t>(4 * c->sub_words, 1), s);
    c->prev_acc.alloc(std::max<uint64_t>(4 * c->acc_words, 1), s);
    BGX_CUDA(cudaMemsetAsync(c->prev_bits.p, 0, std::max<uint64_t>(4 * c->prev_words, 1) * 8, s));
    BGX_CUDA(cudaMemsetAsync(c->prev_acc.p, 0, std::max<uint64_t>(4 * c->acc_words, 1) * 8, s));
    unsigned long long* bits = reinterpret_cast<unsigned long long*>(c->prev_bits.p);
    DevBuf<unsigned int> max_len(1, s);
    DevBuf<int> flags(9, s);  // [0..3] carry to the next rank's entry 0, [4..7] single-base entries, [8] missing
    BGX_CUDA(cudaMemsetAsync(max_len.p, 0, 4, s));
    BGX_CUDA(cudaMemsetAsync(flags.p, 0, 9 * 4, s));
    {
      ScopedStage st_l(c, "tables_local");
      if (nb) KLAUNCH(tables_local_kernel)<<<grid_for(nb, 128), 128, 0, s>>>(store, keys.p, locs.p, n2, prev, c->sizes.p, c->shared.p,
                                                                     max_len.p);
      st_l.stop();
    }
    // prev bits: pop_front(e) tagged with e's first base, routed to the rank whose range holds
    // the first entry it is a prefix of (bs/builder.cpp:85-107)
    {
      DevBuf<uint64_t> qk(std::max<uint32_t>(n2, 1), s), ql(std::max<uint32_t>(n2, 1), s);
      DevBuf<unsigned long long> nq_d(1, s);
      BGX_CUDA(cudaMemsetAsync(nq_d.p, 0, 8, s));
      ScopedStage st_q(c, "tables_queries");
      if (n2) {
        KLAUNCH(pop_queries_kernel)<<<grid_for(n2, 256), 256, 0, s>>>(store, keys.p, locs.p, n2, 1, qk.p, ql.p, nq_d.p);
        KLAUNCH(single_base_kernel)<<<grid_for(n2, 256), 256, 0, s>>>(keys.p, locs.p, n2, flags.p + 4);
      }
      uint32_t nq = (uint32_t)read_u64(nq_d.p, s);
      st_q.stop();
      Routed q = route_records(c, qk.p, ql.p, nq, sp2);
      ScopedStage st_p(c, "tables_prev_apply");
      const BucketIndex bi2 = build_bucket_index(c, keys.p, n2, index_buf);
      if (q.n) KLAUNCH(prev_apply_kernel)<<<grid_for(q.n, 128), 128, 0, s>>>(store, keys.p, locs.p, n2, bi2, next, q.keys.p, q.locs.p, q.n,
                                                                     bits, c->prev_words, flags.p, flags.p + 8);
      BGX_CUDA(cudaGetLastError());
      st_p.stop();
    }
    // carries, single-base entries, missing flag, max length: tiny all-gather
    {
      int hf[9];
      unsigned int h_max;
      BGX_CUDA(cudaMemcpyAsync(hf, flags.p, sizeof(hf), cudaMemcpyDeviceToHost, s));
      BGX_CUDA(cudaMemcpyAsync(&h_max, max_len.p, 4, cudaMemcpyDeviceToHost, s));
      BGX_CUDA(cudaStreamSynchronize(s));
      uint64_t mine[10];
      for (int i = 0; i < 9; ++i) mine[i] = (uint64_t)hf[i];
      mine[9] = h_max;
      std::vector<uint64_t> all((size_t)N * 10);
      dist_allgather_host_u64(c, mine, 10, all.data());
      bool missing = false;
      uint64_t mx = 0;
      int first_owner = -1;
      for (int r = 0; r < N; ++r) {
        missing = missing || all[(size_t)r * 10 + 8];
        mx = std::max(mx, all[(size_t)r * 10 + 9]);
        if (first_owner < 0 && ends.has(r)) first_owner = r;
      }
      BGX_CHECK(!missing, "Missing expansion?");  // bs/builder.cpp:96
      c->max_entry_len = (uint32_t)mx;
      if (nb) {
        for (int b = 0; b < 4; ++b) {
          bool set0 = false;
          for (int r = R - 1; r >= 0; --r) {  // carries from the ranks right before this one
            if (all[(size_t)r * 10 + b]) set0 = true;
            if (ends.has(r)) break;
          }
          if (R == first_owner)
            for (int r = 0; r < N; ++r) set0 = set0 || all[(size_t)r * 10 + 4 + b];
          if (set0) KLAUNCH(set_bit0_kernel)<<<1, 1, 0, s>>>(bits, c->prev_words, b);
        }
      }
    }
    // bitcount index over the global bit vector (modules/io/bitcount.cpp:84-123)
    {
      uint64_t pops[4] = {0, 0, 0, 0};
      DevBuf<uint32_t> gpop(std::max<uint64_t>(c->sub_words, 1), s), gex(std::max<uint64_t>(4 * c->sub_words, 1), s), tot(4, s);
      BGX_CUDA(cudaMemsetAsync(tot.p, 0, 16, s));
      for (int b = 0; b < 4 && nb; ++b) {
        KLAUNCH(bitcount_groups_kernel)<<<grid_for(c->sub_words, 256), 256, 0, s>>>(
            bits + b * c->prev_words, c->prev_words, c->sub_words, gpop.p,
            reinterpret_cast<unsigned long long*>(c->prev_sub.p) + b * c->sub_words);
        exclusive_scan_u32(gpop.p, gex.p + b * c->sub_words, c->sub_words, tot.p + b, s);
      }
      uint32_t h_tot[4];
      BGX_CUDA(cudaMemcpyAsync(h_tot, tot.p, 16, cudaMemcpyDeviceToHost, s));
      BGX_CUDA(cudaStreamSynchronize(s));
      for (int b = 0; b < 4; ++b) pops[b] = h_tot[b];
      std::vector<uint64_t> all((size_t)N * 4);
      dist_allgather_host_u64(c, pops, 4, all.data());
      uint64_t off = 0;
      for (int b = 0; b < 4; ++b) {
        uint64_t before = 0, total = 0;
        for (int r = 0; r < N; ++r) {
          if (r < R) before += all[(size_t)r * 4 + b];
          total += all[(size_t)r * 4 + b];
        }
        if (c->acc_words)
          KLAUNCH(bitcount_accum_offset_kernel)<<<grid_for(c->acc_words, 256), 256, 0, s>>>(
              gex.p + b * c->sub_words, tot.p + b, c->sub_words, c->acc_words, before,
              reinterpret_cast<unsigned long long*>(c->prev_acc.p) + b * c->acc_words);
        c->fixed[b] = off;
        off += total;
      }
      c->fixed[4] = off;
      BGX_CUDA(cudaGetLastError());
      BGX_CUDA(cudaStreamSynchronize(s));
      // seqset::finalize (seqset.cpp:123-126)
      BGX_CHECK(c->fixed[4] == Nt, "Invalid seqset in finalize: prev bit totals != entries");
    }
    st.stop();
  }
  c->ent_key = std::move(keys);
  c->ent_loc = std::move(locs);
  c->built = true;
  st_all.stop();
}

void export_varbit(Context* c, int which, uint64_t** words, uint64_t* n_words, uint32_t* bits, uint64_t* max_value) {
  BGX_CHECK(c->built, "bgx_export_varbit: call bgx_build_seqset first");
  BGX_CHECK(which == 0 || which == 1, "bgx_export_varbit: which must be 0 (entry_sizes) or 1 (shared)");
  cudaStream_t s = c->stream;
  // seqset ctor (modules/bio_base/seqset.cpp:27-33): max_value = max_entry_len, resp. max_entry_len - 1
  const uint64_t mv = which == 0 ? c->max_entry_len : (c->max_entry_len ? c->max_entry_len - 1 : 0);
  unsigned b = 0;  // bits_for_value (packed_varbit_vector.cpp:174-182): bit length of max_value (0 for 0)
  while ((mv >> b) != 0) ++b;
  const uint64_t n = c->n_entries, nw = (n * b + 63) / 64;
  DevBuf<unsigned long long> d(std::max<uint64_t>(nw, 1), s);
  if (nw)
    KLAUNCH(varbit_pack_kernel)<<<grid_for(nw, 256), 256, 0, s>>>(which == 0 ? c->sizes.p : c->shared.p, n, b, nw, d.p);
  BGX_CUDA(cudaGetLastError());
  uint64_t* h = (uint64_t*)host_alloc(std::max<uint64_t>(nw, 1) * 8);
  if (nw) BGX_CUDA(cudaMemcpyAsync(h, d.p, nw * 8, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  *words = h;
  *n_words = nw;
  *bits = b;
  *max_value = mv;
}

void lookup_reads(Context* c, uint64_t* n_reads, uint64_t** fwd_entry, uint64_t** rc_entry) {
  BGX_CHECK(c->built, "bgx_lookup_reads: call bgx_build_seqset first");
  BGX_CHECK(c->corrected, "bgx_lookup_reads: the context holds a merged seqset, not a build from reads");
  BGX_CHECK(c->dist.nranks == 1, "bgx_lookup_reads: single-GPU builds only (a sharded build would route the reads like the "
                                 "pop_front queries; not built yet)");
  cudaStream_t s = c->stream;
  const uint64_t n = c->n_reads;
  *n_reads = n;
  *fwd_entry = (uint64_t*)host_alloc(std::max<uint64_t>(n, 1) * 8);
  *rc_entry = (uint64_t*)host_alloc(std::max<uint64_t>(n, 1) * 8);
  if (n == 0) return;
  ScopedStage st(c, "lookup_reads");
  DevBuf<unsigned long long> d_f(n, s), d_v(n, s);
  DevBuf<int> missing(1, s);
  BGX_CUDA(cudaMemsetAsync(missing.p, 0, sizeof(int), s));
  DevBuf<uint32_t> index_buf;
  const BucketIndex bi = build_bucket_index(c, c->ent_key.p, (uint32_t)c->n_entries, index_buf);
  KLAUNCH(lookup_reads_kernel)<<<grid_for(n, 128), 128, 0, s>>>(c->store.p, c->ent_key.p, c->ent_loc.p, (uint32_t)c->n_entries, bi,
                                                        c->word_off.p, c->clen.p, c->n_words, (uint32_t)n, d_f.p, d_v.p,
                                                        missing.p);
  BGX_CUDA(cudaGetLastError());
  int h_missing = 0;
  BGX_CUDA(cudaMemcpyAsync(*fwd_entry, d_f.p, n * 8, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaMemcpyAsync(*rc_entry, d_v.p, n * 8, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaMemcpyAsync(&h_missing, missing.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  st.stop();
  BGX_CHECK(!h_missing, "a corrected read was not found in seqset");  // make_readmap.cpp:150-154
}

static int read_flag_i(const int* d, cudaStream_t s) {
  int h = 0;
  BGX_CUDA(cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  return h;
}

// bitcount::finalize of one bit vector of nbits bits (modules/io/bitcount.cpp:84-123) into host arrays
void bitcount_to_host(Context* c, const unsigned long long* bits, uint64_t nbits, uint64_t* out[3], uint64_t* total) {
  cudaStream_t s = c->stream;
  const uint64_t words = (nbits + 63) / 64, sub_words = (nbits + 511) / 512, acc_words = (nbits + 1 + 511) / 512;
  DevBuf<uint32_t> gpop(std::max<uint64_t>(sub_words, 1), s), gex(std::max<uint64_t>(sub_words, 1), s), tot(1, s);
  DevBuf<unsigned long long> sub(std::max<uint64_t>(sub_words, 1), s), acc(std::max<uint64_t>(acc_words, 1), s);
  BGX_CUDA(cudaMemsetAsync(tot.p, 0, 4, s));
  BGX_CUDA(cudaMemsetAsync(acc.p, 0, std::max<uint64_t>(acc_words, 1) * 8, s));
  if (sub_words) {
    KLAUNCH(bitcount_groups_kernel)<<<grid_for(sub_words, 256), 256, 0, s>>>(bits, words, sub_words, gpop.p, sub.p);
    exclusive_scan_u32(gpop.p, gex.p, sub_words, tot.p, s);
  }
  if (acc_words) KLAUNCH(bitcount_accum_kernel)<<<grid_for(acc_words, 256), 256, 0, s>>>(gex.p, tot.p, sub_words, acc_words, acc.p);
  BGX_CUDA(cudaGetLastError());
  out[0] = (uint64_t*)host_alloc(std::max<uint64_t>(words, 1) * 8);
  out[1] = (uint64_t*)host_alloc(std::max<uint64_t>(sub_words, 1) * 8);
  out[2] = (uint64_t*)host_alloc(std::max<uint64_t>(acc_words, 1) * 8);
  if (words) BGX_CUDA(cudaMemcpyAsync(out[0], bits, words * 8, cudaMemcpyDeviceToHost, s));
  if (sub_words) BGX_CUDA(cudaMemcpyAsync(out[1], sub.p, sub_words * 8, cudaMemcpyDeviceToHost, s));
  if (acc_words) BGX_CUDA(cudaMemcpyAsync(out[2], acc.p, acc_words * 8, cudaMemcpyDeviceToHost, s));
  uint32_t h_tot = 0;
  BGX_CUDA(cudaMemcpyAsync(&h_tot, tot.p, 4, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  if (total != nullptr) *total = h_tot;
}

void build_readmap(Context* c, int paired, uint64_t* n_rows, uint16_t** read_lengths, uint64_t** mate_loop_ptr,
                   uint64_t** is_forward, uint64_t* read_ids_source[3], uint64_t* read_ids_dest[3]) {
  BGX_CHECK(c->built, "bgx_build_readmap: call bgx_build_seqset first");
  BGX_CHECK(c->corrected, "bgx_build_readmap: the context holds a merged seqset, not a build from reads");
  BGX_CHECK(c->dist.nranks == 1, "bgx_build_readmap: single-GPU builds only");
  BGX_CHECK(c->n_entries < kNoLoopEntry, "Entry id too long to fit in mate loop table entry");  // make_readmap.h:77
  BGX_CHECK(!paired || c->n_reads % 2 == 0, "bgx_build_readmap: paired input needs an even number of reads (mates are reads 2i, 2i+1)");
  cudaStream_t s = c->stream;
  ScopedStage st(c, "readmap");
  const uint64_t n = c->n_reads;
  const uint64_t n_rec = paired ? n / 2 : n;
  const uint32_t n_ent = (uint32_t)c->n_entries;
  // 1. the entry of every read and of its reverse complement (find_existing_unique, make_readmap.cpp:137-167)
  DevBuf<unsigned long long> d_f(std::max<uint64_t>(n, 1), s), d_v(std::max<uint64_t>(n, 1), s);
  DevBuf<int> missing(1, s);
  BGX_CUDA(cudaMemsetAsync(missing.p, 0, sizeof(int), s));
  DevBuf<uint32_t> index_buf;
  DevBuf<uint32_t> cnt(std::max<uint64_t>(n_rec, 1), s), pos(std::max<uint64_t>(n_rec, 1), s), tot(1, s);
  BGX_CUDA(cudaMemsetAsync(tot.p, 0, 4, s));
  if (n) {
    const BucketIndex bi = build_bucket_index(c, c->ent_key.p, n_ent, index_buf);
    KLAUNCH(lookup_reads_kernel)<<<grid_for(n, 128), 128, 0, s>>>(c->store.p, c->ent_key.p, c->ent_loc.p, n_ent, bi, c->word_off.p,
                                                          c->clen.p, c->n_words, (uint32_t)n, d_f.p, d_v.p, missing.p);
    KLAUNCH(readmap_count_kernel)<<<grid_for(n_rec, 256), 256, 0, s>>>(c->clen.p, (uint32_t)n_rec, paired, cnt.p);
    exclusive_scan_u32(cnt.p, pos.p, n_rec, tot.p, s);
    BGX_CUDA(cudaGetLastError());
  }
  const uint32_t m = read_u32(tot.p, s);
  BGX_CHECK(!read_flag_i(missing.p, s), "a corrected read was not found in seqset");  // make_readmap.cpp:150-154
  *n_rows = m;
  // 2. rows, sorted by (prim, sec): LSD, the secondary key first
  DevBuf<uint64_t> k0((size_t)m + 1, s), v0((size_t)m + 1, s), k1((size_t)m + 1, s), v1((size_t)m + 1, s);
  const uint64_t *prim = k0.p, *sec = v0.p;
  if (m) {
    KLAUNCH(readmap_rows_kernel)<<<grid_for(n_rec, 256), 256, 0, s>>>(d_f.p, d_v.p, c->clen.p, pos.p, (uint32_t)n_rec, paired, k0.p,
                                                              v0.p);
    BGX_CUDA(cudaGetLastError());
    // pass A: key = sec (10 + 37 bits), value = prim
    uint64_t *sk = v0.p, *sv = k0.p, *sk_alt = v1.p, *sv_alt = k1.p;
    if (radix_sort_pairs(sk, sv, sk_alt, sv_alt, m, 0, 48, s)) { std::swap(sk, sk_alt); std::swap(sv, sv_alt); }
    // pass B: key = prim (37 + 2 + 10 bits), value = sec; stable, so ties keep the order of pass A
    if (radix_sort_pairs(sv, sk, sv_alt, sk_alt, m, 0, 56, s)) { std::swap(sk, sk_alt); std::swap(sv, sv_alt); }
    prim = sv;
    sec = sk;
  }
  // 3. tables
  const uint64_t src_words = ((uint64_t)n_ent + 63) / 64, row_words = ((uint64_t)m + 63) / 64;
  DevBuf<unsigned long long> src_bits(std::max<uint64_t>(src_words, 1), s), dst_bits(std::max<uint64_t>(row_words, 1), s),
      fwd_bits(std::max<uint64_t>(row_words, 1), s), ptr(std::max<uint32_t>(m, 1), s);
  DevBuf<uint16_t> lens(std::max<uint32_t>(m, 1), s);
  BGX_CUDA(cudaMemsetAsync(src_bits.p, 0, std::max<uint64_t>(src_words, 1) * 8, s));
  BGX_CUDA(cudaMemsetAsync(dst_bits.p, 0, std::max<uint64_t>(row_words, 1) * 8, s));
  BGX_CUDA(cudaMemsetAsync(fwd_bits.p, 0, std::max<uint64_t>(row_words, 1) * 8, s));
  if (m) {
    const uint32_t max_claims = m / 4 + 1;  // one per pair
    DevBuf<uint64_t> ck0(max_claims, s), cv0(max_claims, s), ck1(max_claims, s), cv1(max_claims, s);
    DevBuf<unsigned int> n_claims_d(1, s);
    BGX_CUDA(cudaMemsetAsync(n_claims_d.p, 0, 4, s));
    KLAUNCH(readmap_fill_kernel)<<<grid_for(m, 256), 256, 0, s>>>(prim, sec, m, src_bits.p, reinterpret_cast<uint32_t*>(dst_bits.p),
                                                          reinterpret_cast<uint32_t*>(fwd_bits.p), lens.p, ptr.p, ck0.p, cv0.p,
                                                          n_claims_d.p, missing.p);
    BGX_CUDA(cudaGetLastError());
    const uint32_t n_claims = read_u32(n_claims_d.p, s);
    BGX_CHECK(n_claims < max_claims, "internal: more mate claims than pairs");
    if (n_claims) {
      const uint64_t *ck = ck0.p, *cv = cv0.p;
      if (radix_sort_pairs(ck0.p, cv0.p, ck1.p, cv1.p, n_claims, 0, 64, s)) { ck = ck1.p; cv = cv1.p; }
      KLAUNCH(readmap_claim_kernel)<<<grid_for(n_claims, 256), 256, 0, s>>>(prim, sec, m, ck, cv, n_claims, ptr.p, missing.p);
      BGX_CUDA(cudaGetLastError());
    }
    c->set_stat("readmap_pairs", n_claims);
  }
  BGX_CHECK(!read_flag_i(missing.p, s), "internal: a mate loop row has no row to claim");
  // 4. out
  bitcount_to_host(c, src_bits.p, n_ent, read_ids_source);
  bitcount_to_host(c, dst_bits.p, m, read_ids_dest);
  *read_lengths = (uint16_t*)host_alloc(std::max<uint32_t>(m, 1) * 2);
  *mate_loop_ptr = (uint64_t*)host_alloc(std::max<uint32_t>(m, 1) * 8);
  *is_forward = (uint64_t*)host_alloc(std::max<uint64_t>(row_words, 1) * 8);
  if (m) {
    BGX_CUDA(cudaMemcpyAsync(*read_lengths, lens.p, (size_t)m * 2, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaMemcpyAsync(*mate_loop_ptr, ptr.p, (size_t)m * 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaMemcpyAsync(*is_forward, fwd_bits.p, row_words * 8, cudaMemcpyDeviceToHost, s));
  }
  BGX_CUDA(cudaStreamSynchronize(s));
  st.stop();
  c->set_stat("readmap_rows", m);
}

void export_entries_ascii(Context* c, uint64_t first, uint64_t count, char** bases, uint64_t** offs_out) {
  BGX_CHECK(c->built, "bgx_export_entries_ascii: call bgx_build_seqset first");
  BGX_CHECK(first + count <= c->n_entries, "bgx_export_entries_ascii: range out of bounds");
  cudaStream_t s = c->stream;
  std::vector<uint32_t> lens(count);
  DevBuf<uint32_t> d_lens(std::max<uint64_t>(count, 1), s);
  if (count) {
    KLAUNCH(entry_offs_kernel)<<<grid_for(count, 256), 256, 0, s>>>(c->ent_loc.p, first, count, d_lens.p);
    BGX_CUDA(cudaMemcpyAsync(lens.data(), d_lens.p, count * 4, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
  }
  uint64_t* offs = (uint64_t*)host_alloc((count + 1) * 8);
  offs[0] = 0;
  for (uint64_t i = 0; i < count; ++i) offs[i + 1] = offs[i] + lens[i];
  char* out = (char*)host_alloc(std::max<uint64_t>(offs[count], 1));
  if (count) {
    DevBuf<uint64_t> d_offs(count + 1, s);
    DevBuf<char> d_out(std::max<uint64_t>(offs[count], 1), s);
    BGX_CUDA(cudaMemcpyAsync(d_offs.p, offs, (count + 1) * 8, cudaMemcpyHostToDevice, s));
    KLAUNCH(entries_ascii_kernel)<<<grid_for(count, 128), 128, 0, s>>>(c->seq_store(), c->ent_loc.p, d_offs.p, first, count, d_out.p);
    BGX_CUDA(cudaMemcpyAsync(out, d_out.p, offs[count], cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
  }
  *bases = out;
  *offs_out = offs;
}

}  // namespace bgx

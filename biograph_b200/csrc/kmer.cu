// kmer.cu -- k-mer counting over 2-bit packed reads and construction of the solid k-mer set.
//
// Replaces (reference, CPU): prob_pass_processor / exact_pass_processor (bs/kmer_counter.h:275-326,
// bs/kmer_counter.cpp:579-697), kmer_count_table::increment (bs/kmer_count_table.h:54-103),
// kmerizer::run's min-count filter (modules/bio_mapred/kmerize_bf.cpp:290-318) and kmer_set
// (modules/bio_mapred/kmer_set.cpp:522-631, lookup :296-360).
//
// B200 design: two passes, no probabilistic pre-pass, no temp files.
//   pass 1  kmer_partition_kernel: a block pulls a tile of packed reads into shared memory with
//           128-bit coalesced loads; lane l forms the k-mers l, l+32, ... of a read by funnel
//           shifts, canonicalises with brev and hashes; the instances leave bucketed by the top
//           hash bits as coalesced per-partition runs.  A fused linear-counting sample sizes the
//           table.
//   pass 2  kmer_upsert_kernel: the slot of a k-mer is the TOP bits of its hash, so a partition is
//           one contiguous slice of the table that stays in L2 while its tiles are upserted
//           (CAS claim that doubles as the read, RED.add on the counter, RED.or for the flags):
//           16-byte slots, key|flags + both counters in one 32-byte sector.
//   filter  one sweep of the table; the solid k-mers are inserted into the bucketed hash set that
//           correction probes (32-byte buckets of four slots, home bucket = top hash bits).
// Inputs whose instance buffers would not fit next to the table are counted in batches of reads
// (count_batch_reads); multi-GPU, every partition travels to the rank that owns its hash range.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdlib>
#include <vector>

#include "ctx.h"

namespace bgx {
namespace {

// ---- partitioned counting -------------------------------------------------------------------
// A k-mer instance travels between the two kernels as one 8-byte word:
//   bits 0..2k-1 the k-mer as seen in the read (NOT canonical), bit 63 = first k-mer of its read,
//   bit 62 = last k-mer of its read (bs/kmer_counter.h:318-321).  k <= 31, so the fields never meet.
constexpr uint64_t kPkFirst = 1ULL << 63;
constexpr uint64_t kPkLast = 1ULL << 62;
constexpr int kMaxPartBits = 10;            // <= 1024 hash partitions
constexpr int kPartThreads = 256;
constexpr int kPartWarps = kPartThreads / 32;

// The table slot of a k-mer is the TOP bits of its hash, so the top part_bits bits (its partition)
// select a contiguous 1/P slice of the table.
// With 2^rank_bits GPUs the top rank_bits bits of the hash name the owning rank (a contiguous block
// of partitions) and are dropped before indexing that rank's table.
__device__ __forceinline__ uint64_t table_slot(uint64_t h, int log2_slots, int rank_bits) {
  return (h << rank_bits) >> (64 - log2_slots);
}

// Pass 1: extract every k-mer instance, bucket it by hash partition.
//   * the block's reads (kPartWarps*RPW consecutive reads) are pulled into shared memory with
//     128-bit coalesced loads; a warp takes one read at a time, lane l forms the k-mers at
//     positions l, l+32, ... by funnel shifts (semantics of pass_processor::add,
//     bs/kmer_counter.h:297-326: a window containing 'N' is skipped)
//   * rank within (tile, partition) from a shared-memory histogram, tile staged in partition
//     order, one global cursor bump per (block, partition), coalesced per-partition runs out
//   * fused distinct-k-mer estimate: linear counting over the 2^-samp_shift of hash space whose low
//     hash bits are zero (sizes the table; sampling by hash value is unbiased for distinct counts)
// Partition p owns out[part_base[p] .. part_base[p] + cap).  cursors[] keep counting past cap, so
// after an overflowing run they are the exact histogram for the exact re-run.
template <int MAXIT, int RPW>
__global__ void __launch_bounds__(kPartThreads, (MAXIT <= 4 ? 4 : 2))
kmer_partition_kernel(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask,
                      const uint32_t* __restrict__ word_off, const uint16_t* __restrict__ lens, uint32_t n_reads,
                      int k, int part_bits, unsigned long long* __restrict__ cursors,
                      const unsigned long long* __restrict__ part_base, unsigned long long cap,
                      unsigned long long* __restrict__ out, unsigned int* __restrict__ bitmap, uint64_t bit_mask,
                      int samp_shift, int* __restrict__ overflow) {
  constexpr int kTileReads = kPartWarps * RPW;
  constexpr int kTileKmers = kTileReads * MAXIT * 32;
  constexpr int kWordsPerRead = MAXIT + 1;          // MAXIT*32 k-mers of k<=31 bases span <= MAXIT+1 words
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* stage = reinterpret_cast<unsigned long long*>(smem_raw);            // kTileKmers
  uint64_t* rwords = reinterpret_cast<uint64_t*>(stage + kTileKmers);                     // kTileReads*kWordsPerRead + 2
  uint32_t* rmask = reinterpret_cast<uint32_t*>(rwords + kTileReads * kWordsPerRead + 2); // same count
  uint16_t* stage_bin = reinterpret_cast<uint16_t*>(rmask + kTileReads * kWordsPerRead + 2);  // kTileKmers
  const int P = 1 << part_bits;
  unsigned long long* gdst = reinterpret_cast<unsigned long long*>(stage_bin + kTileKmers);   // P
  uint32_t* hist = reinterpret_cast<uint32_t*>(gdst + P);                                     // P
  uint32_t* bin_start = hist + P;                                                             // P
  __shared__ uint32_t tile_word0, tile_nwords;

  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_tiles = (n_reads + kTileReads - 1) / kTileReads;

  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t r0 = tile * kTileReads;
    const uint32_t r1 = min(n_reads, r0 + kTileReads);
    for (int d = tid; d < P; d += kPartThreads) hist[d] = 0;
    if (tid == 0) {
      tile_word0 = word_off[r0] & ~1u;  // 16-byte aligned start
      tile_nwords = word_off[r1] - (word_off[r0] & ~1u);
    }
    __syncthreads();
    {
      // 128-bit coalesced loads of the tile's packed reads (+ N mask) into shared memory
      const uint32_t w0 = tile_word0, nw2 = (tile_nwords + 1) >> 1;
      const uint4* src = reinterpret_cast<const uint4*>(words + w0);
      uint4* dst = reinterpret_cast<uint4*>(rwords);
      for (uint32_t i = tid; i < nw2; i += kPartThreads) dst[i] = ld_stream_u4(src + i);
      if (nmask != nullptr) {
        const uint2* msrc = reinterpret_cast<const uint2*>(nmask + w0);
        uint2* mdst = reinterpret_cast<uint2*>(rmask);
        for (uint32_t i = tid; i < nw2; i += kPartThreads) mdst[i] = msrc[i];
      }
    }
    __syncthreads();

    unsigned long long kw[RPW * MAXIT];
    uint32_t br[RPW * MAXIT];  // bin << 16 | rank ; 0xffffffff = no k-mer
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
      const uint32_t r = r0 + warp * RPW + q;
      int nk = 0;
      uint32_t wb = 0;
      if (r < r1) {
        nk = (int)lens[r] - k + 1;
        wb = word_off[r] - tile_word0;
      }
#pragma unroll
      for (int it = 0; it < MAXIT; ++it) {
        const int p = it * 32 + (int)lane;
        uint32_t code = 0xffffffffu;
        unsigned long long word = 0;
        if (p < nk) {
          const uint64_t hi = rwords[wb + it], lo = rwords[wb + it + 1];
          const unsigned s = lane * 2;
          const uint64_t win = s ? ((hi << s) | (lo >> (64 - s))) : hi;
          bool has_n = false;
          if (nmask != nullptr) {
            const uint32_t mh = rmask[wb + it], ml = rmask[wb + it + 1];
            const uint32_t mwin = lane ? ((mh << lane) | (ml >> (32 - lane))) : mh;
            has_n = (mwin >> (32 - k)) != 0;  // an 'N' inside the window (bs/kmer_counter.h:306-311)
          }
          if (!has_n) {
            const uint64_t kmer = win >> (64 - 2 * k);
            bool flipped;
            const uint64_t canon = canonicalize(kmer, k, flipped);
            const uint64_t h = mix64(canon);
            const uint32_t bin = (uint32_t)(h >> (64 - part_bits));
            const uint32_t rank = atomicAdd(&hist[bin], 1u);
            code = (bin << 16) | rank;
            word = kmer | (p == 0 ? kPkFirst : 0ULL) | (p == nk - 1 ? kPkLast : 0ULL);
            if (bitmap != nullptr && (h & ((1ULL << samp_shift) - 1)) == 0) {
              const uint64_t bit = (h >> samp_shift) & bit_mask;
              atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));  // RED: fire and forget, no read-back stall
            }
          }
        }
        kw[q * MAXIT + it] = word;
        br[q * MAXIT + it] = code;
      }
    }
    __syncthreads();
    // exclusive scan of the tile histogram; one cursor bump per non-empty partition
    {
      const int per = (P + kPartThreads - 1) / kPartThreads;  // 1..4
      uint32_t loc[4];
      uint32_t sum = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = (int)tid * per + j;
        loc[j] = (j < per && d < P) ? hist[d] : 0u;
        sum += loc[j];
      }
      uint32_t tot;
      uint32_t ex = block_excl_scan_u32(sum, &tot);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = (int)tid * per + j;
        if (j < per && d < P) {
          bin_start[d] = ex;
          if (loc[j]) {
            const unsigned long long g = atomicAdd(&cursors[d], (unsigned long long)loc[j]);
            if (g + loc[j] > cap) *overflow = 1;
            gdst[d] = g;
          }
          ex += loc[j];
        }
      }
      if (tid == 0) tile_nwords = tot;  // reuse: number of k-mers staged
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RPW * MAXIT; ++i) {
      if (br[i] != 0xffffffffu) {
        const uint32_t bin = br[i] >> 16, pos = bin_start[bin] + (br[i] & 0xffffu);
        stage[pos] = kw[i];
        stage_bin[pos] = (uint16_t)bin;
      }
    }
    __syncthreads();
    const uint32_t n_staged = tile_nwords;
    for (uint32_t j = tid; j < n_staged; j += kPartThreads) {
      const uint32_t bin = stage_bin[j];
      const unsigned long long o = gdst[bin] + (j - bin_start[bin]);
      if (o < cap) out[part_base[bin] + o] = stage[j];
    }
    __syncthreads();
  }
}

// Pass 2: upsert the partitioned instances into the table.  Block (p, t) handles tile t of
// partition p; blocks are dispatched in index order, so at any moment the resident blocks work
// on one or two partitions = a few MB of table that stay in L2: the CAS / atomicAdd / atomicOr
// traffic (kmer_count_table::increment, bs/kmer_count_table.h:54-103) never leaves L2, and DRAM
// sees the table exactly twice (first touch, final write-back).
constexpr int kUpsItems = 4;
constexpr int kUpsTile = 256 * kUpsItems;
template <bool CAS_FIRST>
__global__ void __launch_bounds__(256, 4) kmer_upsert_kernel(const unsigned long long* const* __restrict__ part_ptr,
                                                             const unsigned long long* __restrict__ part_count,
                                                             uint32_t tiles_per_part, int k,
                                                             CountEntry* __restrict__ table, int log2_slots,
                                                             int rank_bits, int* __restrict__ overflow) {
  const uint32_t p = blockIdx.x / tiles_per_part, t = blockIdx.x % tiles_per_part;
  const unsigned long long cnt = part_count[p];
  const unsigned long long first = (unsigned long long)t * kUpsTile;
  if (first >= cnt) return;
  const unsigned long long* src = part_ptr[p];
  const uint64_t slot_mask = (1ULL << log2_slots) - 1;
  // lock-step phases over the thread's items keep kUpsItems L2 round trips in flight per thread:
  // A) instance words  B) first-probe keys  C) CAS claims of empty slots  D) counters (RED)
  unsigned long long w[kUpsItems], key[kUpsItems], cur[kUpsItems];
  uint32_t slot[kUpsItems];   // offset from the partition-aligned base (fits 32 bits: slots < 2^32 per call)
  bool flip[kUpsItems];
#pragma unroll
  for (int i = 0; i < kUpsItems; ++i) {
    const unsigned long long idx = first + (unsigned long long)i * 256 + threadIdx.x;
    w[i] = idx < cnt ? src[idx] : ~0ULL;   // ~0 = no item (a real word never has all k-mer bits and both flags set for k<=31)
  }
#pragma unroll
  for (int i = 0; i < kUpsItems; ++i) {
    const bool valid = w[i] != ~0ULL;
    bool fl;
    const uint64_t canon = canonicalize(w[i] & kKmerMask, k, fl);
    flip[i] = fl;
    // fwd_flag = first k-mer of the read, rev_flag = last; swapped when flipped
    // (bs/kmer_counter.h:318-321, bs/kmer_count_table.h:82-86)
    const bool is_first = (w[i] & kPkFirst) != 0, is_last = (w[i] & kPkLast) != 0;
    uint64_t f = 0;
    if (fl ? is_last : is_first) f |= kFwdFlag;
    if (fl ? is_first : is_last) f |= kRevFlag;
    key[i] = valid ? (canon | f) : ~0ULL;
    slot[i] = (uint32_t)table_slot(mix64(canon), log2_slots, rank_bits);
  }
  if (CAS_FIRST) {
    // the claim attempt doubles as the read of the slot: one CAS + one RED per instance.  Measured
    // (tools/micro/l2_atomics, 16 MB slice): failing CAS + RED 88 G/s against load + RED 81 G/s,
    // and the separate CAS of the first-touch instances (14 % at 0.5 % error) is gone.
#pragma unroll
    for (int i = 0; i < kUpsItems; ++i)
      cur[i] = key[i] != ~0ULL ? atomicCAS(&table[slot[i]].key, (unsigned long long)kEmptyKey, key[i]) : 0ULL;
  } else {
#pragma unroll
    for (int i = 0; i < kUpsItems; ++i)
      cur[i] = key[i] != ~0ULL ? *reinterpret_cast<volatile unsigned long long*>(&table[slot[i]].key) : 0ULL;
#pragma unroll
    for (int i = 0; i < kUpsItems; ++i)
      if (key[i] != ~0ULL && cur[i] == kEmptyKey)
        cur[i] = atomicCAS(&table[slot[i]].key, (unsigned long long)kEmptyKey, key[i]);  // returns kEmptyKey when claimed
  }
#pragma unroll
  for (int i = 0; i < kUpsItems; ++i) {
    if (key[i] == ~0ULL) continue;
    const uint64_t canon = key[i] & kKmerMask;
    uint64_t f = key[i] & ~kKmerMask;
    uint64_t s = slot[i];
    unsigned long long c = cur[i];
    if (c == kEmptyKey) {
      f = 0;  // claimed by the CAS above: key and flags are in place
    } else if ((c & kKmerMask) == canon) {
      f &= ~c;
    } else {
      // collision: linear probing (rare at load factor <= 2/3)
      bool done = false;
      for (uint64_t probes = 0; probes <= slot_mask; ++probes) {
        s = (s + 1) & slot_mask;
        c = *reinterpret_cast<volatile unsigned long long*>(&table[s].key);
        if (c == kEmptyKey) {
          c = atomicCAS(&table[s].key, (unsigned long long)kEmptyKey, key[i]);
          if (c == kEmptyKey) { f = 0; done = true; break; }
        }
        if ((c & kKmerMask) == canon) { f &= ~c; done = true; break; }
      }
      if (!done) { *overflow = 1; continue; }
    }
    if (f) atomicOr(&table[s].key, (unsigned long long)f);
    // one 32-bit RED on the fwd (low) or rev (high) half of the counter word
    atomicAdd(reinterpret_cast<unsigned int*>(&table[s].cnt) + (flip[i] ? 1 : 0), 1u);
  }
}

// Distinct estimate over partitioned instance words (multi-GPU: run by the owner after the
// exchange, because the fused estimate of pass 1 only saw this rank's own reads).
__global__ void __launch_bounds__(256) kmer_estimate_words_kernel(const unsigned long long* const* __restrict__ part_ptr,
                                                                  const unsigned long long* __restrict__ part_count,
                                                                  uint32_t tiles_per_part, int k,
                                                                  unsigned int* __restrict__ bitmap, uint64_t bit_mask) {
  const uint32_t p = blockIdx.x / tiles_per_part, t = blockIdx.x % tiles_per_part;
  const unsigned long long cnt = part_count[p];
  const unsigned long long* src = part_ptr[p];
#pragma unroll
  for (int i = 0; i < kUpsItems; ++i) {
    const unsigned long long idx = (unsigned long long)t * kUpsTile + (unsigned long long)i * 256 + threadIdx.x;
    if (idx >= cnt) continue;
    bool fl;
    const uint64_t h = mix64(canonicalize(src[idx] & kKmerMask, k, fl));
    if ((h & 15) == 0) {
      const uint64_t bit = (h >> 4) & bit_mask;
      atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
    }
  }
}

// Distinct estimate straight from the packed reads (batched counting: the table must be sized
// before the first batch is upserted, so the estimate cannot ride on pass 1).  A warp per read,
// lane l takes the k-mers l, l+32, ...; same sampling rule as the fused estimate.
__global__ void __launch_bounds__(256) kmer_estimate_reads_kernel(const uint64_t* __restrict__ words,
                                                                  const uint32_t* __restrict__ nmask,
                                                                  const uint32_t* __restrict__ word_off,
                                                                  const uint16_t* __restrict__ lens, uint32_t n_reads,
                                                                  int k, unsigned int* __restrict__ bitmap,
                                                                  uint64_t bit_mask, int samp_shift) {
  const uint32_t r = blockIdx.x * (256 / 32) + (threadIdx.x >> 5);
  if (r >= n_reads) return;
  const unsigned lane = lane_id();
  const int nk = (int)lens[r] - k + 1;
  const uint32_t wb = word_off[r];
  for (int p = (int)lane; p < nk; p += 32) {
    const int w = p >> 5;
    const unsigned sft = lane * 2;
    const uint64_t hi = words[wb + w], lo = words[wb + w + 1];  // the store has one pad word
    const uint64_t win = sft ? ((hi << sft) | (lo >> (64 - sft))) : hi;
    if (nmask != nullptr) {
      const uint32_t mh = nmask[wb + w], ml = nmask[wb + w + 1];
      const uint32_t mwin = lane ? ((mh << lane) | (ml >> (32 - lane))) : mh;
      if ((mwin >> (32 - k)) != 0) continue;
    }
    bool fl;
    const uint64_t h = mix64(canonicalize(win >> (64 - 2 * k), k, fl));
    if ((h & ((1ULL << samp_shift) - 1)) == 0) {
      const uint64_t bit = (h >> samp_shift) & bit_mask;
      atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
    }
  }
}

// out[i] = OR over the N gathered bitmaps (multi-GPU batched counting: union of every rank's sample)
__global__ void bitmap_or_kernel(const unsigned int* __restrict__ gathered, int n_maps, uint64_t n_words,
                                 unsigned int* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  unsigned v = 0;
  for (int m = 0; m < n_maps; ++m) v |= gathered[(uint64_t)m * n_words + i];
  out[i] = v;
}

__global__ void popcount_kernel(const unsigned int* __restrict__ w, uint64_t n, unsigned long long* __restrict__ total) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned c = i < n ? __popc(w[i]) : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ unsigned ws[8];
  if (lane_id() == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int j = 0; j < 8; ++j) t += ws[j];
    if (t) atomicAdd(total, (unsigned long long)t);
  }
}

__global__ void fill_empty_kernel(uint4* __restrict__ t, uint64_t n_slots) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_slots) t[i] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
}

__global__ void fill_u64_kernel(unsigned long long* __restrict__ t, uint64_t n, unsigned long long v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = v;
}

// Sweep: one pass over the table.  Each 256-thread block covers kSweepPerBlock slots, stages its
// passing entries (fwd+rev >= min_count) in shared memory and appends them with ONE global atomic
// per block; distinct/passing totals are block-reduced the same way.
constexpr int kSweepIters = 8;
constexpr int kSweepPerBlock = 256 * kSweepIters;
__global__ void __launch_bounds__(256) table_sweep_kernel(const CountEntry* __restrict__ table, uint64_t n_slots,
                                                          uint32_t min_count,
                                                          unsigned long long* __restrict__ counters /*[0]=distinct,[1]=passing*/,
                                                          unsigned long long* __restrict__ out_key,
                                                          unsigned long long* __restrict__ out_cnt, uint64_t out_cap) {
  __shared__ unsigned long long skey[kSweepPerBlock];
  __shared__ unsigned long long scnt[kSweepPerBlock];
  __shared__ unsigned int n_pass, n_used;
  __shared__ unsigned long long gbase;
  if (threadIdx.x == 0) { n_pass = 0; n_used = 0; }
  __syncthreads();
  const unsigned lane = lane_id();
  uint64_t base = (uint64_t)blockIdx.x * kSweepPerBlock;
  unsigned used_cnt = 0;
#pragma unroll 2
  for (int it = 0; it < kSweepIters; ++it) {
    uint64_t i = base + (uint64_t)it * 256 + threadIdx.x;
    bool used = false, pass = false;
    unsigned long long key = 0, cnt = 0;
    if (i < n_slots) {
      uint4 e = ld_stream_u4(reinterpret_cast<const uint4*>(table) + i);
      key = ((unsigned long long)e.y << 32) | e.x;
      cnt = ((unsigned long long)e.w << 32) | e.z;
      used = key != kEmptyKey;
      uint64_t tot = (cnt & 0xffffffffu) + (cnt >> 32);
      pass = used && tot >= min_count;
    }
    used_cnt += used ? 1u : 0u;
    unsigned pm = __ballot_sync(0xffffffffu, pass);
    if (pm) {
      unsigned wb = 0;
      if (lane == 0) wb = atomicAdd(&n_pass, (unsigned)__popc(pm));
      wb = __shfl_sync(0xffffffffu, wb, 0);
      if (pass) {
        unsigned o = wb + __popc(pm & ((1u << lane) - 1));
        skey[o] = key;
        scnt[o] = cnt;
      }
    }
  }
  used_cnt = __reduce_add_sync(0xffffffffu, used_cnt);
  if (lane == 0 && used_cnt) atomicAdd(&n_used, used_cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (n_used) atomicAdd(&counters[0], (unsigned long long)n_used);
    gbase = n_pass ? atomicAdd(&counters[1], (unsigned long long)n_pass) : 0ULL;
  }
  __syncthreads();
  if (out_key != nullptr) {
    for (unsigned j = threadIdx.x; j < n_pass; j += 256) {
      unsigned long long o = gbase + j;
      if (o < out_cap) {
        out_key[o] = skey[j];
        out_cnt[o] = scnt[j];
      }
    }
  }
}

// insert into the solid set (32-byte buckets of four 8-byte slots, common.cuh).  Keys are distinct,
// so a plain CAS claim suffices; a bucket fills in slot order, a full one sends the key onwards.
// The home bucket is the TOP bits of the same hash whose top bits order the count table, and the
// sweep emits the solid k-mers in table order: consecutive threads insert into consecutive
// buckets, so building the set is a near-sequential write instead of one random DRAM CAS per key.
__global__ void solid_insert_kernel(const unsigned long long* __restrict__ keys, uint64_t n,
                                    unsigned long long* __restrict__ set, uint64_t bucket_mask, int bucket_shift) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long kf = keys[i];
  uint64_t b = mix64(kf & kKmerMask) >> bucket_shift;
  for (;;) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (atomicCAS(&set[4 * b + j], (unsigned long long)kEmptyKey, kf) == kEmptyKey) return;
    b = (b + 1) & bucket_mask;
  }
}

// sort key for export: the canonical k-mer without flag bits; value = index into the compacted list
__global__ void export_prepare_kernel(const unsigned long long* __restrict__ key_flags, uint64_t n,
                                      uint64_t* __restrict__ sort_key, uint64_t* __restrict__ sort_val) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sort_key[i] = key_flags[i] & kKmerMask;
  sort_val[i] = i;
}

__global__ void export_gather_kernel(const uint64_t* __restrict__ sort_key, const uint64_t* __restrict__ sort_val,
                                     const unsigned long long* __restrict__ key_flags,
                                     const unsigned long long* __restrict__ cnt, uint64_t n,
                                     uint64_t* __restrict__ kmers, uint32_t* __restrict__ fwd,
                                     uint32_t* __restrict__ rev, uint8_t* __restrict__ flags) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t j = sort_val[i];
  unsigned long long kf = key_flags[j], c = cnt[j];
  kmers[i] = sort_key[i];
  fwd[i] = (uint32_t)(c & 0xffffffffu);
  rev[i] = (uint32_t)(c >> 32);
  flags[i] = (uint8_t)(((kf & kFwdFlag) ? BGX_FLAG_FWD_STARTS_READ : 0) | ((kf & kRevFlag) ? BGX_FLAG_REV_STARTS_READ : 0));
}

uint64_t pow2_ceil(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace

namespace {

template <int MAXIT, int RPW>
void launch_partition(Context* c, uint64_t r0, uint64_t n_reads, int part_bits, unsigned long long* cursors,
                      const unsigned long long* part_base, unsigned long long cap, unsigned long long* out,
                      unsigned int* bitmap, uint64_t bit_mask, int samp_shift, int* overflow) {
  constexpr int tile_reads = kPartWarps * RPW;
  constexpr int tile_kmers = tile_reads * MAXIT * 32;
  constexpr int wpr = MAXIT + 1;
  const size_t smem = (size_t)tile_kmers * 8 + (size_t)(tile_reads * wpr + 2) * 12 + (size_t)tile_kmers * 2 +
                      ((size_t)16 << part_bits);
  static bool attr_set = false;
  if (!attr_set) {
    const size_t smem_max = (size_t)tile_kmers * 10 + (size_t)(tile_reads * wpr + 2) * 12 + ((size_t)16 << kMaxPartBits);
    BGX_CUDA(cudaFuncSetAttribute(kmer_partition_kernel<MAXIT, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_max));
    attr_set = true;
  }
  int blocks_per_sm = 1;
  BGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kmer_partition_kernel<MAXIT, RPW>,
                                                         kPartThreads, smem));
  blocks_per_sm = std::max(blocks_per_sm, 1);
  // persistent grid: every SM fully occupied, tiles handed out grid-stride
  const unsigned n_tiles = (unsigned)((n_reads + tile_reads - 1) / tile_reads);
  const unsigned grid = std::min<unsigned>(n_tiles, (unsigned)(kNumSMs * blocks_per_sm));
  note_launch();
  // a batch is a range of reads: word offsets are absolute, so only the per-read arrays shift
  kmer_partition_kernel<MAXIT, RPW><<<grid, kPartThreads, smem, c->stream>>>(
      c->words.p, c->has_n ? c->nmask.p : nullptr, c->word_off.p + r0, c->lens.p + r0, (uint32_t)n_reads,
      c->opt.kmer_size, part_bits, cursors, part_base, cap, out, bitmap, bit_mask, samp_shift, overflow);
  BGX_CUDA(cudaGetLastError());
}

void run_partition(Context* c, int maxit, uint64_t r0, uint64_t n_reads, int part_bits, unsigned long long* cursors,
                   const unsigned long long* part_base, unsigned long long cap, unsigned long long* out,
                   unsigned int* bitmap, uint64_t bit_mask, int samp_shift, int* overflow) {
  if (maxit <= 4)
    launch_partition<4, 4>(c, r0, n_reads, part_bits, cursors, part_base, cap, out, bitmap, bit_mask, samp_shift, overflow);
  else
    launch_partition<8, 2>(c, r0, n_reads, part_bits, cursors, part_base, cap, out, bitmap, bit_mask, samp_shift, overflow);
}

int log2_exact(uint64_t x) {
  int l = 0;
  while ((1ULL << l) < x) ++l;
  return l;
}

// the linear-counting sample: bit (h >> shift) & (bits - 1) of the bitmap for hashes with `shift` low zero bits
struct Estimator {
  DevBuf<unsigned int> bitmap;
  DevBuf<unsigned long long> ones;
  uint64_t bits = 0;
  int shift = 4;
  void init(uint64_t bits_, int shift_, cudaStream_t s) {
    bits = bits_;
    shift = shift_;
    bitmap.alloc(bits / 32, s);
    ones.alloc(1, s);
    BGX_CUDA(cudaMemsetAsync(bitmap.p, 0, bits / 8, s));
  }
  // number of distinct k-mers the sample stands for (synchronises the stream)
  uint64_t estimate(cudaStream_t s) {
    unsigned long long h_ones = 0;
    BGX_CUDA(cudaMemsetAsync(ones.p, 0, 8, s));
    KLAUNCH(popcount_kernel)<<<(unsigned)((bits / 32 + 255) / 256), 256, 0, s>>>(bitmap.p, bits / 32, ones.p);
    BGX_CUDA(cudaGetLastError());
    BGX_CUDA(cudaMemcpyAsync(&h_ones, ones.p, 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    const double zero_frac = std::max(1.0 / (double)bits, 1.0 - (double)h_ones / (double)bits);
    return (uint64_t)((double)(1ULL << shift) * -(double)bits * std::log(zero_frac));
  }
};

// pass-1 output for a range of reads
struct Partitioned {
  DevBuf<unsigned long long> pk;
  std::vector<unsigned long long> base, count;  // per partition: first word, words
  unsigned long long cap = 0;                   // stride between partitions unless `exact`
  bool exact = false;                           // re-run with exact offsets (a heavy hitter overfilled a partition)
};

// Pass 1 over reads [r0, r0 + n): K = their k-mer instances.  est (optional) receives the fused sample.
void partition_reads(Context* c, uint64_t r0, uint64_t n, uint64_t K, int part_bits, int maxit, Estimator* est,
                     Partitioned* out) {
  cudaStream_t s = c->stream;
  const int P = 1 << part_bits;
  unsigned long long cap = K / P + K / (8ull * P) + 4096;  // hash partitions are near uniform
  ScopedStage st_alloc(c, "count_setup");
  out->pk.alloc((size_t)cap * P, s);
  DevBuf<unsigned long long> cursors(P, s), part_base(P, s);
  out->base.assign(P, 0);
  out->count.assign(P, 0);
  for (int p = 0; p < P; ++p) out->base[p] = (unsigned long long)p * cap;
  BGX_CUDA(cudaMemcpyAsync(part_base.p, out->base.data(), P * 8, cudaMemcpyHostToDevice, s));
  BGX_CUDA(cudaMemsetAsync(cursors.p, 0, P * 8, s));
  DevBuf<int> overflow(1, s);
  BGX_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
  int h_over = 0;
  st_alloc.stop();
  ScopedStage st(c, "count_partition");
  unsigned int* bitmap = est ? est->bitmap.p : nullptr;
  const uint64_t bit_mask = est ? est->bits - 1 : 0;
  const int shift = est ? est->shift : 4;
  if (n && !c->upload.empty() && r0 == 0 && n == c->n_reads) {
    // the reads are still arriving (bgx_add_reads_packed_async): one launch per upload chunk, each
    // ordered after its own copy only, all appending to the same partitions through the cursors
    uint64_t pos = 0;
    for (Context::UploadChunk& ch : c->upload) {
      if (ch.r0 > pos)
        run_partition(c, maxit, pos, ch.r0 - pos, part_bits, cursors.p, part_base.p, cap, out->pk.p, bitmap, bit_mask, shift, overflow.p);
      BGX_CUDA(cudaStreamWaitEvent(s, ch.ev, 0));
      cudaEventDestroy(ch.ev);
      if (ch.r1 > ch.r0)
        run_partition(c, maxit, ch.r0, ch.r1 - ch.r0, part_bits, cursors.p, part_base.p, cap, out->pk.p, bitmap, bit_mask, shift, overflow.p);
      pos = ch.r1;
    }
    c->upload.clear();
    if (pos < n)
      run_partition(c, maxit, pos, n - pos, part_bits, cursors.p, part_base.p, cap, out->pk.p, bitmap, bit_mask, shift, overflow.p);
  } else if (n) {
    reads_ready(c);
    run_partition(c, maxit, r0, n, part_bits, cursors.p, part_base.p, cap, out->pk.p, bitmap, bit_mask, shift, overflow.p);
  }
  BGX_CUDA(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaMemcpyAsync(out->count.data(), cursors.p, P * 8, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  out->cap = cap;
  out->exact = false;
  if (h_over) {
    // a heavy-hitter k-mer overfilled its partition: the cursors are now the exact histogram,
    // so re-run with exact offsets (the reference has no analogue; its tables are sized up front).
    // The sample bitmap is only OR-ed into, so the second run leaves it as it is.
    uint64_t total = 0;
    cap = 0;
    for (int p = 0; p < P; ++p) {
      out->base[p] = total;
      total += out->count[p];
      cap = std::max<unsigned long long>(cap, out->count[p]);
    }
    out->pk.alloc(std::max<uint64_t>(total, 1), s);
    BGX_CUDA(cudaMemcpyAsync(part_base.p, out->base.data(), P * 8, cudaMemcpyHostToDevice, s));
    BGX_CUDA(cudaMemsetAsync(cursors.p, 0, P * 8, s));
    BGX_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
    run_partition(c, maxit, r0, n, part_bits, cursors.p, part_base.p, cap, out->pk.p, bitmap, bit_mask, shift, overflow.p);
    BGX_CUDA(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    BGX_CHECK(!h_over, "internal: exact partition pass overflowed");
    c->add_stat("count_partition_reruns", 1);
    out->cap = cap;
    out->exact = true;
  }
  st.stop();
}

// the partitions this rank will count: (device address, instance count) each
struct Owned {
  std::vector<unsigned long long> ptr, cnt;
  DevBuf<unsigned long long> rbuf;  // what arrived from the peers
  uint64_t n_inst = 0;
};

// Single GPU: the P partitions where pass 1 left them.  Multi-GPU: every partition goes to its
// owner over NVLink; partition p of source rank s arrives as its own "virtual partition", listed
// partition-major so pass 2 still walks the table slice by slice.
void exchange_partitions(Context* c, const Partitioned& pt, int P, Owned* own) {
  cudaStream_t s = c->stream;
  const int N = c->dist.nranks, R = c->dist.rank;
  own->ptr.clear();
  own->cnt.clear();
  if (N == 1) {
    for (int p = 0; p < P; ++p) {
      own->ptr.push_back((unsigned long long)(uintptr_t)(pt.pk.p + pt.base[p]));
      own->cnt.push_back(pt.count[p]);
    }
  } else {
    ScopedStage st(c, "count_exchange");
    const int Pl = P / N, p0 = R * Pl;
    const unsigned long long cap = pt.cap;
    std::vector<uint64_t> mine(P + 2), all((size_t)N * (P + 2));
    for (int p = 0; p < P; ++p) mine[p] = pt.count[p];
    mine[P] = cap;
    mine[P + 1] = pt.exact ? 1 : 0;  // exact layout: not cap-strided
    dist_allgather_host_u64(c, mine.data(), P + 2, all.data());
    auto cnt_of = [&](int src, int p) { return all[(size_t)src * (P + 2) + p]; };
    bool strided = true;
    for (int r = 0; r < N; ++r) strided = strided && all[(size_t)r * (P + 2) + P + 1] == 0;
    std::vector<P2P> sends, recvs;
    if (strided) {
      // fast path: ONE message per peer -- the cap-strided block of the peer's partitions as it
      // lies in memory (the ~11 % slack travels too; NCCL p2p costs ~17 us per message, measured)
      std::vector<uint64_t> src_off(N, 0);
      uint64_t total = 0;
      for (int src = 0; src < N; ++src) {
        if (src == R) continue;
        src_off[src] = total;
        total += (uint64_t)Pl * all[(size_t)src * (P + 2) + P];
      }
      own->rbuf.alloc(std::max<uint64_t>(total, 1), s);
      for (int d = 0; d < N; ++d) {
        if (d == R) continue;
        P2P x, y;
        x.send = pt.pk.p + (uint64_t)d * Pl * cap;
        x.bytes = (size_t)Pl * cap * 8;
        x.peer = d;
        sends.push_back(x);
        y.recv = own->rbuf.p + src_off[d];
        y.bytes = (size_t)Pl * all[(size_t)d * (P + 2) + P] * 8;
        y.peer = d;
        recvs.push_back(y);
      }
      for (int pl = 0; pl < Pl; ++pl)
        for (int src = 0; src < N; ++src) {
          const uint64_t cap_s = all[(size_t)src * (P + 2) + P];
          const unsigned long long* ptr =
              src == R ? pt.pk.p + (uint64_t)(p0 + pl) * cap : own->rbuf.p + src_off[src] + (uint64_t)pl * cap_s;
          own->ptr.push_back((unsigned long long)(uintptr_t)ptr);
          own->cnt.push_back(cnt_of(src, p0 + pl));
        }
    } else {
      // general path (some rank re-ran pass 1 with exact offsets): one message per partition.
      // A rank that did NOT re-run still holds its partitions cap-strided: base[p] says where.
      uint64_t total = 0;
      std::vector<uint64_t> l_base(Pl);
      for (int pl = 0; pl < Pl; ++pl) {
        l_base[pl] = total;
        for (int src = 0; src < N; ++src) total += cnt_of(src, p0 + pl);
      }
      own->rbuf.alloc(std::max<uint64_t>(total, 1), s);
      for (int p = 0; p < P; ++p) {
        P2P x;
        x.send = pt.pk.p + pt.base[p];
        x.bytes = (size_t)pt.count[p] * 8;
        x.peer = p / Pl;
        sends.push_back(x);
      }
      // recvs from one peer must be posted in the order that peer sends: ascending partition
      for (int src = 0; src < N; ++src)
        for (int pl = 0; pl < Pl; ++pl) {
          uint64_t off = l_base[pl];
          for (int q = 0; q < src; ++q) off += cnt_of(q, p0 + pl);
          P2P x;
          x.recv = own->rbuf.p + off;
          x.bytes = (size_t)cnt_of(src, p0 + pl) * 8;
          x.peer = src;
          recvs.push_back(x);
        }
      for (int pl = 0; pl < Pl; ++pl) {
        uint64_t cnt = 0;
        for (int src = 0; src < N; ++src) cnt += cnt_of(src, p0 + pl);
        own->ptr.push_back((unsigned long long)(uintptr_t)(own->rbuf.p + l_base[pl]));
        own->cnt.push_back(cnt);
      }
    }
    dist_p2p_batch(c, sends, recvs);
    uint64_t out_bytes = 0;
    for (const P2P& x : sends)
      if (x.peer != R) out_bytes += x.bytes;
    c->add_stat("count_exchange_bytes_out", (double)out_bytes);
    st.stop();
  }
  own->n_inst = 0;
  for (unsigned long long v : own->cnt) own->n_inst += v;
}

// device-side view of an Owned list, ready for the tile kernels
struct OwnedDev {
  DevBuf<unsigned long long> ptr, cnt;
  uint32_t V = 0, tiles_per_part = 1;
  const unsigned long long* const* part_ptr() const { return reinterpret_cast<const unsigned long long* const*>(ptr.p); }
};

void upload_owned(Context* c, const Owned& own, OwnedDev* d) {
  cudaStream_t s = c->stream;
  d->V = (uint32_t)own.ptr.size();
  d->ptr.alloc(d->V, s);
  d->cnt.alloc(d->V, s);
  BGX_CUDA(cudaMemcpyAsync(d->ptr.p, own.ptr.data(), d->V * 8, cudaMemcpyHostToDevice, s));
  BGX_CUDA(cudaMemcpyAsync(d->cnt.p, own.cnt.data(), d->V * 8, cudaMemcpyHostToDevice, s));
  uint64_t max_count = 0;
  for (unsigned long long v : own.cnt) max_count = std::max<uint64_t>(max_count, v);
  d->tiles_per_part = (uint32_t)std::max<uint64_t>(1, (max_count + kUpsTile - 1) / kUpsTile);
  BGX_CHECK((uint64_t)d->tiles_per_part * d->V < (1ull << 31), "too many k-mer tiles for one launch");
  // the host vectors must outlive the copies
  BGX_CUDA(cudaStreamSynchronize(s));
}

// table slots for an estimated distinct count: load factor in (1/3, 2/3]
uint64_t slots_for(uint64_t est_distinct) {
  uint64_t slots = pow2_ceil(std::max<uint64_t>(1024, est_distinct + est_distinct / 2));
  if (const char* e = getenv("BGX_TABLE_SLOTS_LOG2")) slots = 1ull << atoi(e);  // experiment hook
  return slots;
}

void alloc_table(Context* c, uint64_t slots) {
  cudaStream_t s = c->stream;
  ScopedStage st_sz(c, "count_table_alloc");
  BGX_CHECK(slots <= (1ull << 32), "k-mer table too large for one GPU shard (slot index is 32-bit)");
  c->table_slots = slots;
  c->table.alloc(slots, s);  // throws "out of device memory" if the table cannot fit
  st_sz.stop();
  ScopedStage st(c, "count_init");
  KLAUNCH(fill_empty_kernel)<<<(unsigned)((slots + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint4*>(c->table.p), slots);
  BGX_CUDA(cudaGetLastError());
  st.stop();
}

void upsert_owned(Context* c, const OwnedDev& od, int rank_bits, int* overflow) {
  ScopedStage st(c, "count_kernel");
  static const bool cas_first = [] { const char* e = getenv("BGX_UPSERT_CAS_FIRST"); return e ? atoi(e) != 0 : true; }();  // experiment hook
  note_launch();
  if (cas_first)
    kmer_upsert_kernel<true><<<od.tiles_per_part * od.V, 256, 0, c->stream>>>(
        od.part_ptr(), od.cnt.p, od.tiles_per_part, c->opt.kmer_size, c->table.p, log2_exact(c->table_slots), rank_bits, overflow);
  else
    kmer_upsert_kernel<false><<<od.tiles_per_part * od.V, 256, 0, c->stream>>>(
      od.part_ptr(), od.cnt.p, od.tiles_per_part, c->opt.kmer_size, c->table.p, log2_exact(c->table_slots), rank_bits, overflow);
  BGX_CUDA(cudaGetLastError());
  st.stop();
}

int read_flag(const int* d, cudaStream_t s) {
  int h = 0;
  BGX_CUDA(cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  return h;
}

// How many batches of reads pass 1 + pass 2 run in.  One batch keeps all K instance words (and,
// multi-GPU, what the peers send) in HBM next to the table; the batched form bounds those buffers
// to ~40 % of the device so that inputs like GRCh38 30x on 8 GPUs (84 + 66 GB of words per GPU
// next to a 69 GB table) still fit.  Every rank must use the same count (each batch is a
// collective exchange).
uint64_t choose_batches(Context* c, uint64_t K_local, uint64_t K_share) {
  const int N = c->dist.nranks;
  uint64_t batches = 1;
  uint64_t batch_reads = c->opt.count_batch_reads > 0 ? (uint64_t)c->opt.count_batch_reads : 0;
  if (const char* e = getenv("BGX_COUNT_BATCH_READS")) batch_reads = strtoull(e, nullptr, 10);  // test hook
  if (batch_reads) {
    batches = std::max<uint64_t>(1, (c->n_reads + batch_reads - 1) / batch_reads);
  } else {
    const double budget = 0.4 * (double)c->total_mem;
    const double need = 9.0 * (double)K_local + (N > 1 ? 9.0 * (double)K_share : 0.0);  // 8 B per word + 1/8 slack
    batches = std::max<uint64_t>(1, (uint64_t)std::ceil(need / budget));
  }
  if (N > 1) {
    std::vector<uint64_t> all(N);
    dist_allgather_host_u64(c, &batches, 1, all.data());
    for (uint64_t v : all) batches = std::max(batches, v);
  }
  return batches;
}

}  // namespace

void stage_count_kmers(Context* c) {
  cudaStream_t s = c->stream;
  const int k = c->opt.kmer_size;
  const int N = c->dist.nranks;
  const int rank_bits = log2_exact((uint64_t)N);
  BGX_CHECK(c->n_reads > 0 || N > 1, "bgx_count_kmers: no reads");
  ScopedStage st_all(c, "count_total");
  const uint64_t K = c->n_kmer_instances;  // this rank's reads
  uint64_t K_all = K;                      // all ranks' reads
  dist_allreduce_sum_host_u64(c, &K_all, 1);
  const uint64_t K_share = K_all / N;      // instances this rank will own (hash-uniform)

  // P partitions so that one partition's slice of the owner's table (~4 B per instance at typical
  // coverage) is at most ~64 MB = half of L2; measured on B200: fewer, larger partitions make
  // pass 1 faster (longer coalesced runs) and pass 2 is insensitive down to 64 MB slices.
  // Rank r owns the contiguous block of partitions [r*P/N, (r+1)*P/N) (same P on every rank).
  int part_bits = 7;
  while (part_bits < kMaxPartBits && (K_all * 4 >> part_bits) > (64ull << 20)) ++part_bits;
  if (const char* e = getenv("BGX_PART_BITS")) part_bits = std::max(1, std::min(kMaxPartBits, atoi(e)));
  part_bits = std::max(part_bits, rank_bits);
  const int P = 1 << part_bits;
  const int maxit = (int)((std::max<int64_t>((int64_t)c->max_len - k + 1, 1) + 31) / 32);
  BGX_CHECK(maxit <= 8, "read longer than 255 bases");
  const uint64_t batches = choose_batches(c, K, K_share);
  c->set_stat("count_batches", (double)batches);
  DevBuf<int> overflow(1, s);
  uint64_t n_inst = 0;   // instances this rank counted
  uint64_t slots = 0;

  if (batches == 1) {
    // ---- everything at once: pass 1 carries the distinct estimate, then the table is sized ---------
    Estimator est;
    est.init(pow2_ceil(std::max<uint64_t>(1 << 20, std::max(K, K_share) / 8)), 4, s);
    Partitioned pt;
    partition_reads(c, 0, c->n_reads, K, part_bits, maxit, &est, &pt);
    Owned own;
    exchange_partitions(c, pt, P, &own);
    OwnedDev od;
    upload_owned(c, own, &od);
    n_inst = own.n_inst;
    if (N > 1) {
      // the owner estimates the distinct count of what it received (pass 1 filled the bitmap with
      // this rank's own reads over the whole hash space: start over)
      BGX_CUDA(cudaMemsetAsync(est.bitmap.p, 0, est.bits / 8, s));
      KLAUNCH(kmer_estimate_words_kernel)<<<od.tiles_per_part * od.V, 256, 0, s>>>(od.part_ptr(), od.cnt.p, od.tiles_per_part, k,
                                                                           est.bitmap.p, est.bits - 1);
      BGX_CUDA(cudaGetLastError());
    }
    uint64_t est_distinct = est.estimate(s);
    est_distinct = std::min<uint64_t>(est_distinct + est_distinct / 16 + 4096, n_inst + 1);
    c->set_stat("kmer_distinct_estimate", (double)est_distinct);
    slots = slots_for(est_distinct);
    for (int tries = 0;; ++tries) {
      alloc_table(c, slots);
      BGX_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
      upsert_owned(c, od, rank_bits, overflow.p);
      if (!read_flag(overflow.p, s)) break;
      // estimate was off (cannot happen for slots > distinct; belt and braces): double and redo.
      // The reference throws io_exception("Kmer table (...) too small") (bs/kmer_count_table.h:75).
      BGX_CHECK(tries < 3, "Kmer table too small");
      slots *= 2;
    }
  } else {
    // ---- batched: estimate from the reads first, then partition / exchange / upsert batch by batch ---
    reads_ready(c);
    Estimator est;
    {
      ScopedStage st(c, "count_estimate");
      // sample so thinly that at most ~2^29 instances are expected in it; the bitmap has twice as many bits
      int shift = 4;
      while ((K_all >> shift) > (1ull << 29)) ++shift;
      est.init(pow2_ceil(std::max<uint64_t>(1 << 20, (K_all >> shift) * 2)), shift, s);
      if (c->n_reads)
        KLAUNCH(kmer_estimate_reads_kernel)<<<(unsigned)((c->n_reads + 7) / 8), 256, 0, s>>>(
            c->words.p, c->has_n ? c->nmask.p : nullptr, c->word_off.p, c->lens.p, (uint32_t)c->n_reads, k, est.bitmap.p,
            est.bits - 1, est.shift);
      BGX_CUDA(cudaGetLastError());
      if (N > 1) {
        // union of every rank's sample; an owner then holds 1/N of the distinct k-mers (hash-uniform)
        DevBuf<unsigned int> gathered((size_t)N * (est.bits / 32), s);
        dist_allgather_bytes(c, est.bitmap.p, gathered.p, est.bits / 8);
        KLAUNCH(bitmap_or_kernel)<<<(unsigned)((est.bits / 32 + 255) / 256), 256, 0, s>>>(gathered.p, N, est.bits / 32, est.bitmap.p);
        BGX_CUDA(cudaGetLastError());
      }
      st.stop();
    }
    uint64_t est_distinct = est.estimate(s) / N;
    est_distinct = est_distinct + est_distinct / 16 + 4096;
    c->set_stat("kmer_distinct_estimate", (double)est_distinct);
    est.bitmap.release();
    slots = slots_for(est_distinct);
    // per-read instance counts are not kept on the host: a batch's K is bounded by its share of the bases
    for (int tries = 0;; ++tries) {
      alloc_table(c, slots);
      BGX_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
      n_inst = 0;
      bool over = false;
      for (uint64_t b = 0; b < batches && !over; ++b) {
        const uint64_t r0 = c->n_reads * b / batches, r1 = c->n_reads * (b + 1) / batches;
        // instances of the batch: at most (max_len - k + 1) per read
        const uint64_t Kb = std::min<uint64_t>(K, (r1 - r0) * (uint64_t)std::max<int64_t>((int64_t)c->max_len - k + 1, 0));
        Partitioned pt;
        partition_reads(c, r0, r1 - r0, Kb, part_bits, maxit, nullptr, &pt);
        Owned own;
        exchange_partitions(c, pt, P, &own);
        OwnedDev od;
        upload_owned(c, own, &od);
        n_inst += own.n_inst;
        upsert_owned(c, od, rank_bits, overflow.p);
        over = read_flag(overflow.p, s) != 0;  // also: the batch's buffers are free to go
      }
      if (N > 1) {  // every rank must take the same branch: the batches are collective
        uint64_t any = over ? 1 : 0;
        dist_allreduce_sum_host_u64(c, &any, 1);
        over = any != 0;
      }
      if (!over) break;
      BGX_CHECK(tries < 3, "Kmer table too small");
      slots *= 2;
    }
  }

  // filter (kmer_passes: fwd+rev >= min_count) and build the solid set: ONE sweep into buffers
  // sized by the bound #solid <= instances / min_count.
  DevBuf<unsigned long long> counters(2, s);
  unsigned long long h_cnt[2];
  {
    ScopedStage st(c, "count_filter");
    uint64_t capf = std::min<uint64_t>(slots, n_inst / (uint64_t)c->opt.min_kmer_count + 1);
    DevBuf<unsigned long long> sk(capf, s), sc(capf, s);
    BGX_CUDA(cudaMemsetAsync(counters.p, 0, 2 * sizeof(unsigned long long), s));
    KLAUNCH(table_sweep_kernel)<<<(unsigned)((slots + kSweepPerBlock - 1) / kSweepPerBlock), 256, 0, s>>>(
        c->table.p, slots, (uint32_t)c->opt.min_kmer_count, counters.p, sk.p, sc.p, capf);
    BGX_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    c->n_distinct = h_cnt[0];
    uint64_t n_solid_local = h_cnt[1];
    BGX_CHECK(n_solid_local <= capf, "internal: solid k-mer bound violated");
    // multi-GPU: every rank needs the whole solid set for correction -> all-gather the owners' lists
    const unsigned long long* all_keys = sk.p;
    DevBuf<unsigned long long> gathered;
    c->n_solid = n_solid_local;
    if (N > 1) {
      std::vector<uint64_t> cnts(N), offs(N), zero(N, 0), mine_cnt(N, n_solid_local);
      dist_allgather_host_u64(c, &n_solid_local, 1, cnts.data());
      uint64_t tot = 0;
      for (int r = 0; r < N; ++r) { offs[r] = tot; tot += cnts[r]; }
      gathered.alloc(std::max<uint64_t>(tot, 1), s);
      dist_alltoallv(c, sk.p, zero.data(), mine_cnt.data(), gathered.p, offs.data(), cnts.data(), 8);
      all_keys = gathered.p;
      c->n_solid = tot;
      c->set_stat("kmer_solid_owned", (double)n_solid_local);
    }
    // "Too many kmers for kmer table!" (kmer_set.cpp:554-556) has no analogue: the set is sized to fit.
    // load factor in (1/4, 1/2]: a lookup is one sector read for all but the few keys of overfull
    // buckets.  Measured on E. coli 100x: 4.35 ms for the probe pass at load factor 0.29 (128 MB)
    // against 5.9 ms at 0.61 (64 MB) -- short probe chains matter more than L2 residency.
    double solid_factor = 2.0;
    if (const char* e = getenv("BGX_SOLID_FACTOR")) solid_factor = std::max(1.05, atof(e));  // experiment hook
    c->solid_slots = pow2_ceil(std::max<uint64_t>(1024, (uint64_t)((double)c->n_solid * solid_factor)));
    c->solid.alloc(c->solid_slots, s);
    KLAUNCH(fill_u64_kernel)<<<(unsigned)((c->solid_slots + 255) / 256), 256, 0, s>>>(c->solid.p, c->solid_slots, kEmptyKey);
    if (c->n_solid)
      KLAUNCH(solid_insert_kernel)<<<(unsigned)((c->n_solid + 255) / 256), 256, 0, s>>>(all_keys, c->n_solid, c->solid.p,
                                                                             c->solid_slots / 4 - 1,
                                                                             64 - log2_exact(c->solid_slots / 4));
    BGX_CUDA(cudaGetLastError());
    st.stop();
  }
  c->counted = true;
  c->corrected = c->built = false;
  st_all.stop();
  c->set_stat("kmer_instances", (double)n_inst);
  c->set_stat("kmer_distinct", (double)c->n_distinct);
  c->set_stat("kmer_solid", (double)c->n_solid);
  c->set_stat("count_table_slots", (double)slots);
  c->set_stat("count_partitions", (double)P);
  // partition pass: reads in, one 8-byte word per instance out; upsert pass: the words back in,
  // the table touched twice (first touch + write-back), 16 B per slot
  c->set_stat("alg_bytes_count_partition", (double)c->n_bases / 4 + 8.0 * (double)K);
  c->set_stat("alg_bytes_count_kernel", 8.0 * (double)n_inst + 32.0 * (double)slots);
  c->set_stat("alg_bytes_count", (double)c->n_bases / 4 + 8.0 * (double)K + 8.0 * (double)n_inst + 48.0 * (double)slots);
}

void export_kmers(Context* c, uint32_t min_count, uint64_t* n_out, uint64_t** kmers, uint32_t** fwd, uint32_t** rev,
                  uint8_t** flags) {
  BGX_CHECK(c->counted && c->table.p,
            "bgx_export_kmers: call bgx_count_kmers first (and before bgx_reset_results; on inputs whose k-mer tables "
            "exceed a quarter of the device memory also before bgx_build_seqset, which releases them)");
  cudaStream_t s = c->stream;
  uint64_t slots = c->table_slots;
  DevBuf<unsigned long long> counters(2, s);
  unsigned long long h_cnt[2];
  BGX_CUDA(cudaMemsetAsync(counters.p, 0, 2 * sizeof(unsigned long long), s));
  const unsigned sweep_grid = (unsigned)((slots + kSweepPerBlock - 1) / kSweepPerBlock);
  KLAUNCH(table_sweep_kernel)<<<sweep_grid, 256, 0, s>>>(c->table.p, slots, min_count, counters.p, nullptr, nullptr, 0);
  BGX_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  uint64_t n = h_cnt[1];
  *n_out = n;
  size_t na = std::max<uint64_t>(n, 1);
  *kmers = (uint64_t*)host_alloc(na * 8);
  *fwd = (uint32_t*)host_alloc(na * 4);
  *rev = (uint32_t*)host_alloc(na * 4);
  *flags = (uint8_t*)host_alloc(na);
  if (n == 0) return;
  DevBuf<unsigned long long> ek(n, s), ec(n, s);
  BGX_CUDA(cudaMemsetAsync(counters.p, 0, 2 * sizeof(unsigned long long), s));
  KLAUNCH(table_sweep_kernel)<<<sweep_grid, 256, 0, s>>>(c->table.p, slots, min_count, counters.p, ek.p, ec.p, n);
  DevBuf<uint64_t> k0(n, s), v0(n, s), k1(n, s), v1(n, s);
  unsigned g = (unsigned)((n + 255) / 256);
  KLAUNCH(export_prepare_kernel)<<<g, 256, 0, s>>>(ek.p, n, k0.p, v0.p);
  int bits = ((2 * c->opt.kmer_size + 7) / 8) * 8;
  bool alt = radix_sort_pairs(k0.p, v0.p, k1.p, v1.p, n, 0, bits, s);
  DevBuf<uint64_t> dk(n, s);
  DevBuf<uint32_t> df(n, s), dr(n, s);
  DevBuf<uint8_t> dfl(n, s);
  KLAUNCH(export_gather_kernel)<<<g, 256, 0, s>>>(alt ? k1.p : k0.p, alt ? v1.p : v0.p, ek.p, ec.p, n, dk.p, df.p, dr.p, dfl.p);
  BGX_CUDA(cudaGetLastError());
  BGX_CUDA(cudaMemcpyAsync(*kmers, dk.p, n * 8, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaMemcpyAsync(*fwd, df.p, n * 4, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaMemcpyAsync(*rev, dr.p, n * 4, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaMemcpyAsync(*flags, dfl.p, n, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
}

}  // namespace bgx

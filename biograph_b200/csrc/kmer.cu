// kmer.cu -- k-mer counting over 2-bit packed reads and construction of the solid k-mer set.
//
// Replaces (reference, CPU): prob_pass_processor / exact_pass_processor (bs/kmer_counter.h:275-326,
// bs/kmer_counter.cpp:579-697), kmer_count_table::increment (bs/kmer_count_table.h:54-103),
// kmerizer::run's min-count filter (modules/bio_mapred/kmerize_bf.cpp:290-318) and kmer_set
// (modules/bio_mapred/kmer_set.cpp:522-631, lookup :296-360).
//
// B200 design: hash-partition twice, then count in shared memory.  No global
// atomics on the counting path, no count table in HBM, no probabilistic pre-pass, no temp files.
//   pass 1  kmer_partition_kernel: a block pulls a tile of packed reads into shared memory with
//           128-bit coalesced loads; lane l forms the k-mers l, l+32, ... of a read by funnel
//           shifts, canonicalises with brev and hashes with khash (a BIJECTION on 2k bits); the
//           instances leave as 8-byte words [remaining hash bits | flipped | rev flag | fwd flag]
//           bucketed by the top hash bits, as coalesced per-partition runs.  A fused linear-counting
//           sample estimates the distinct count.
//   pass 2  kmer_subhist_kernel + kmer_split_kernel: every partition is split by the next hash bits
//           into sub-bins that hold ~2 k distinct k-mers each (exact offsets from a histogram, tiles
//           staged in shared memory, coalesced runs out).
//   pass 3  kmer_count_bins_kernel: one block per sub-bin counts its words in a shared-memory hash
//           table (kmer_count_table::increment semantics: fwd/rev counters, starts-read flags OR-ed,
//           swapped when flipped), then writes every distinct k-mer once -- the hash is inverted to
//           recover the k-mer -- and the k-mers with fwd+rev >= min_count a second time as the solid
//           list.
//   solid   the solid k-mers are inserted into the bucketed hash set that correction probes
//           (32-byte buckets of four slots, home bucket = top hash bits = arrival order).
// Inputs whose instance words would not fit HBM are counted in batches of HASH RANGE (every batch
// re-reads the reads and keeps the k-mers whose top hash bits name the batch); multi-GPU, every
// partition travels to the rank that owns its hash range.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdlib>
#include <string>
#include <vector>

#include "ctx.h"

namespace bgx {

// ---- planning (pure host arithmetic; bgx_debug_count_plan exposes it to the CPU tests) ---------------------
static uint64_t pow2_ceil_u64(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

// Hash-range batches of the counting (a power of two).  One batch keeps all K instance words twice
// (pass-1 output and the split copy; multi-GPU also what the peers send) in HBM; more batches bound
// those buffers to ~45 % of the device so that inputs like GRCh38 30x on 8 GPUs still fit.
// batch_reads (option count_batch_reads / BGX_COUNT_BATCH_READS) asks for at least ceil(reads / that).
uint64_t plan_count_batches(uint64_t K_local, uint64_t K_share, int N, uint64_t total_mem, uint64_t batch_reads,
                            uint64_t n_reads) {
  uint64_t batches = 1;
  if (batch_reads) {
    batches = std::max<uint64_t>(1, (n_reads + batch_reads - 1) / batch_reads);
  } else {
    const double budget = 0.45 * (double)total_mem;
    // 8 B per word + 1/8 slack out of pass 1, what arrives from the peers, the split copy
    const double need = 9.0 * (double)K_local + (N > 1 ? 9.0 : 0.0) * (double)K_share + 8.0 * (double)K_share;
    batches = std::max<uint64_t>(1, (uint64_t)std::ceil(need / budget));
    // ... and so that the distinct k-mers one rank counts per batch fit its sub-bins at a load factor of
    // ~0.65: (128 partitions on one GPU, 64 per rank on several) x 2048 sub-bins x 4096 slots.  The distinct
    // count is not known yet: a fifth of the instances is typical (E. coli 100x 0.15, chr20 30x 0.18 at
    // 0.5 % errors); a worse input overflows a bin and re-runs with larger tables.
    const double cap_distinct = (N == 1 ? 128.0 : 64.0) * 2048.0 * 4096.0 * 0.65;
    batches = std::max<uint64_t>(batches, (uint64_t)std::ceil((double)K_share / 5.0 / cap_distinct));
  }
  return pow2_ceil_u64(batches);
}

// Hash partitions per batch (log2).  One GPU: 128, 256 once a partition would pass 8 M words.  Several
// GPUs: 64 per rank -- pass 1's runs shrink with the TOTAL count (1024 partitions cost it 85 % more time on
// B200) and pass 2 splits finer instead (up to 2048 sub-bins per partition).  Rank r owns the contiguous
// block [r*P/N, (r+1)*P/N).
int plan_part_bits(uint64_t K_share, uint64_t batches, int rank_bits) {
  int part_bits = 7;
  while (part_bits < 8 && ((K_share / batches) >> part_bits) > (8ull << 20)) ++part_bits;
  part_bits = std::max(part_bits, rank_bits + 6);
  return std::max(std::min(part_bits, 10), rank_bits);
}

namespace {

// ---- instance words ---------------------------------------------------------------------------
// A k-mer instance travels between the passes as one 8-byte word:
//   bits 3..  the hash of the canonical k-mer WITHOUT its top `drop` bits (batch + partition bits:
//             they are the same for every word of a partition)
//   bit 2     flipped (the instance was the reverse complement of the canonical k-mer -> rev_count)
//   bit 1     rev_starts_read, bit 0 fwd_starts_read: first / last k-mer of its read, swapped when
//             flipped (bs/kmer_counter.h:318-321, bs/kmer_count_table.h:82-86)
// 2k - drop + 3 <= 62 - 7 + 3 bits, so k = 31 fits too.
constexpr uint64_t kWFwd = 1, kWRev = 2, kWFlip = 4;
constexpr int kMaxPartBits = 10;            // <= 1024 hash partitions per batch
constexpr int kMaxSubBits = 11;             // <= 2048 sub-bins per partition
constexpr int kPartThreads = 256;
constexpr int kPartWarps = kPartThreads / 32;

// Pass 1: extract every k-mer instance of the batch, bucket it by hash partition.
//   * the block's reads (kPartWarps*RPW consecutive reads) are pulled into shared memory with
//     128-bit coalesced loads; a warp takes one read at a time, lane l forms the k-mers at
//     positions l, l+32, ... by funnel shifts (semantics of pass_processor::add,
//     bs/kmer_counter.h:297-326: a window containing 'N' is skipped)
//   * rank within (tile, partition) from a shared-memory histogram, tile staged in partition
//     order, one global cursor bump per (block, partition), coalesced per-partition runs out
//   * fused distinct-k-mer estimate: linear counting over the 2^-samp_shift of hash space whose low
//     hash bits are zero (sampling by hash value is unbiased for distinct counts)
// Partition p owns the cap words at the device address part_addr[p] -- in this GPU's memory or,
// multi-GPU, in the memory of the GPU that owns the partition (mapped with CUDA IPC): the runs then
// leave over NVLink as they are produced, and the k-mer exchange IS this kernel (no send buffer, no
// separate all-to-all).  cursors[] keep counting past cap, so after an overflowing run they are the
// exact histogram for the exact re-run.
struct PartGeom {
  int k;
  int batch_bits;   // the batch keeps the k-mers whose top batch_bits hash bits == batch
  uint32_t batch;
  int part_bits;    // next part_bits bits = partition
};

template <int MAXIT, int RPW>
__global__ void __launch_bounds__(kPartThreads, (MAXIT <= 4 ? 4 : 2))
kmer_partition_kernel(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask,
                      const uint32_t* __restrict__ word_off, const uint16_t* __restrict__ lens, uint32_t n_reads,
                      PartGeom G, unsigned long long* __restrict__ cursors,
                      const unsigned long long* __restrict__ part_addr, unsigned long long cap,
                      unsigned int* __restrict__ bitmap, uint64_t bit_mask, int samp_shift, int* __restrict__ overflow) {
  constexpr int kTileReads = kPartWarps * RPW;
  constexpr int kTileKmers = kTileReads * MAXIT * 32;
  constexpr int kWordsPerRead = MAXIT + 1;          // MAXIT*32 k-mers of k<=31 bases span <= MAXIT+1 words
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* stage = reinterpret_cast<unsigned long long*>(smem_raw);            // kTileKmers
  uint64_t* rwords = reinterpret_cast<uint64_t*>(stage + kTileKmers);                     // kTileReads*kWordsPerRead + 2
  uint32_t* rmask = reinterpret_cast<uint32_t*>(rwords + kTileReads * kWordsPerRead + 2); // same count
  uint16_t* stage_bin = reinterpret_cast<uint16_t*>(rmask + kTileReads * kWordsPerRead + 2);  // kTileKmers
  const int k = G.k, part_bits = G.part_bits;
  const int P = 1 << part_bits;
  unsigned long long* gdst = reinterpret_cast<unsigned long long*>(stage_bin + kTileKmers);   // P: address of staged element 0
  uint32_t* hist = reinterpret_cast<uint32_t*>(gdst + P);                                     // P; later: staged index limit
  uint32_t* bin_start = hist + P;                                                             // P
  __shared__ uint32_t tile_word0, tile_nwords;

  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_tiles = (n_reads + kTileReads - 1) / kTileReads;
  const int low_bits = 2 * k - G.batch_bits - part_bits;   // hash bits a word keeps
  const uint64_t low_mask = (1ULL << low_bits) - 1;

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
            const uint64_t h = khash(canon, k);
            const uint64_t hp = h >> low_bits;   // batch | partition
            if ((uint32_t)(hp >> part_bits) == G.batch) {
              const uint32_t bin = (uint32_t)hp & (uint32_t)(P - 1);
              const uint32_t rank = atomicAdd(&hist[bin], 1u);
              code = (bin << 16) | rank;
              // fwd_flag = first k-mer of the read, rev_flag = last; swapped when flipped
              const bool is_first = p == 0, is_last = p == nk - 1;
              word = ((h & low_mask) << 3) | (flipped ? kWFlip : 0ULL) |
                     ((flipped ? is_last : is_first) ? kWFwd : 0ULL) | ((flipped ? is_first : is_last) ? kWRev : 0ULL);
              if (bitmap != nullptr && (h & ((1ULL << samp_shift) - 1)) == 0) {
                const uint64_t bit = ((h & low_mask) >> samp_shift) & bit_mask;
                atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));  // RED: fire and forget, no read-back stall
              }
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
            // staged element i of this partition (ex <= i < ex + loc) goes to part_addr + 8 * (g + i - ex);
            // the ones past the partition's cap are dropped (the exact re-run places them)
            gdst[d] = part_addr[d] + 8ull * (g - ex);
            hist[d] = g >= cap ? 0u : ex + (uint32_t)min((unsigned long long)loc[j], cap - g);
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
      if (j < hist[bin]) *reinterpret_cast<unsigned long long*>(gdst[bin] + 8ull * j) = stage[j];
    }
    __syncthreads();
  }
}

// ---- pass 2: split every partition into sub-bins by the next hash bits -----------------------------
// A "virtual partition" is a run of words of one partition from one source (single GPU: the
// partition itself; multi-GPU: what one rank sent for it): vptr / vcnt / vpl (its local partition
// index).  Block (v, t) takes tile t of virtual partition v.  sub-bin of a word = bits
// [sub_shift, sub_shift + sub_bits) of its hash field.
constexpr int kSplitThreads = 512;
constexpr int kSplitItems = 16;
constexpr int kSplitTile = kSplitThreads * kSplitItems;  // 8192 words
constexpr int kHistTiles = 4;                            // the histogram kernel takes 4 tiles per block

__global__ void __launch_bounds__(kSplitThreads) kmer_subhist_kernel(const unsigned long long* const* __restrict__ vptr,
                                                                     const unsigned long long* __restrict__ vcnt,
                                                                     const uint32_t* __restrict__ vpl,
                                                                     uint32_t blocks_per_part, int sub_shift, int sub_bits,
                                                                     unsigned long long* __restrict__ hist) {
  __shared__ uint32_t h[1 << kMaxSubBits];   // 8 KB
  const uint32_t v = blockIdx.x / blocks_per_part, t = blockIdx.x % blocks_per_part;
  const unsigned long long cnt = vcnt[v];
  const unsigned long long first = (unsigned long long)t * (kSplitTile * kHistTiles);
  if (first >= cnt) return;
  const int S = 1 << sub_bits;
  for (int d = threadIdx.x; d < S; d += kSplitThreads) h[d] = 0;
  __syncthreads();
  const unsigned long long* src = vptr[v];
  const unsigned long long last = min(cnt, first + (unsigned long long)kSplitTile * kHistTiles);
  const uint32_t smask = (uint32_t)S - 1;
#pragma unroll 4
  for (unsigned long long i = first + threadIdx.x; i < last; i += kSplitThreads)
    atomicAdd(&h[(uint32_t)(src[i] >> (3 + sub_shift)) & smask], 1u);
  __syncthreads();
  unsigned long long* g = hist + (size_t)vpl[v] * S;
  for (int d = threadIdx.x; d < S; d += kSplitThreads)
    if (h[d]) atomicAdd(&g[d], (unsigned long long)h[d]);
}

// bin_off = exclusive scan of the sub-bin histogram (one block; n <= 2^20), total -> *total_out
__global__ void __launch_bounds__(1024) scan_u64_kernel(const unsigned long long* __restrict__ in, uint64_t n,
                                                        unsigned long long* __restrict__ out,
                                                        unsigned long long* __restrict__ total_out,
                                                        unsigned long long* __restrict__ max_out) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned long long mx = 0;
  for (uint64_t base = 0; base < n; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const unsigned long long v = i < n ? in[i] : 0ULL;
    mx = max(mx, v);
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const unsigned long long w = wsum[lane];
      unsigned long long wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= (unsigned)o) wi += u;
      }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const unsigned long long carry = carry_s;
    if (i < n) out[i] = carry + wsum[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wsum[31] + inc;
    __syncthreads();
  }
  // block max of the bin sizes (the longest bin bounds pass 3's tail)
  for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) wsum[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w) mx = max(mx, wsum[w]);
    *total_out = carry_s;
    *max_out = mx;
  }
}

__global__ void __launch_bounds__(kSplitThreads, 2) kmer_split_kernel(const unsigned long long* const* __restrict__ vptr,
                                                                      const unsigned long long* __restrict__ vcnt,
                                                                      const uint32_t* __restrict__ vpl,
                                                                      uint32_t tiles_per_part, int sub_shift, int sub_bits,
                                                                      unsigned long long* __restrict__ cursors,
                                                                      const unsigned long long* __restrict__ bin_off,
                                                                      unsigned long long* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* stage = reinterpret_cast<unsigned long long*>(smem_raw);     // kSplitTile
  const int S = 1 << sub_bits;
  unsigned long long* gdst = reinterpret_cast<unsigned long long*>(stage + kSplitTile);      // S
  uint32_t* hist = reinterpret_cast<uint32_t*>(gdst + S);                          // S
  uint32_t* bin_start = hist + S;                                                  // S
  __shared__ uint32_t n_tile_s;

  const uint32_t v = blockIdx.x / tiles_per_part, t = blockIdx.x % tiles_per_part;
  const unsigned long long cnt = vcnt[v];
  const unsigned long long first = (unsigned long long)t * kSplitTile;
  if (first >= cnt) return;
  const unsigned tid = threadIdx.x;
  for (int d = tid; d < S; d += kSplitThreads) hist[d] = 0;
  __syncthreads();
  const unsigned long long* src = vptr[v] + first;
  const uint32_t n_tile = (uint32_t)min((unsigned long long)kSplitTile, cnt - first);
  const uint32_t smask = (uint32_t)S - 1;
  unsigned long long w[kSplitItems];
  uint32_t rk[kSplitItems / 2];  // two 16-bit ranks per register (rank < 8192); the bin is recomputed from the word
#pragma unroll
  for (int i = 0; i < kSplitItems; ++i) {
    const uint32_t j = i * kSplitThreads + tid;
    w[i] = j < n_tile ? src[j] : 0ULL;
  }
#pragma unroll
  for (int i = 0; i < kSplitItems; ++i) {
    const uint32_t j = i * kSplitThreads + tid;
    uint32_t r = 0;
    if (j < n_tile) r = atomicAdd(&hist[(uint32_t)(w[i] >> (3 + sub_shift)) & smask], 1u);
    if (i & 1) rk[i >> 1] |= r << 16; else rk[i >> 1] = r;
  }
  __syncthreads();
  {
    const int per = (S + kSplitThreads - 1) / kSplitThreads;  // 1..4
    uint32_t loc[4];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = (int)tid * per + j;
      loc[j] = (j < per && d < S) ? hist[d] : 0u;
      sum += loc[j];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan_u32(sum, &tot);
    const size_t gb = (size_t)vpl[v] * S;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = (int)tid * per + j;
      if (j < per && d < S) {
        bin_start[d] = ex;
        // staged element i of this sub-bin goes to out + (bin_off + cursor + i - ex)
        if (loc[j])
          gdst[d] = (unsigned long long)(uintptr_t)out +
                    8ull * (bin_off[gb + d] + atomicAdd(&cursors[gb + d], (unsigned long long)loc[j]) - ex);
        ex += loc[j];
      }
    }
    if (tid == 0) n_tile_s = tot;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSplitItems; ++i) {
    const uint32_t j = i * kSplitThreads + tid;
    if (j < n_tile) {
      const uint32_t bin = (uint32_t)(w[i] >> (3 + sub_shift)) & smask;
      const uint32_t pos = bin_start[bin] + ((rk[i >> 1] >> (16 * (i & 1))) & 0xffffu);
      stage[pos] = w[i];
    }
  }
  __syncthreads();
  const uint32_t n_staged = n_tile_s;
  for (uint32_t j = tid; j < n_staged; j += kSplitThreads) {
    const unsigned long long wj = stage[j];   // the word names its own sub-bin
    *reinterpret_cast<unsigned long long*>(gdst[(uint32_t)(wj >> (3 + sub_shift)) & smask] + 8ull * j) = wj;
  }
}

// ---- pass 3: count one sub-bin per block in a shared-memory hash table ----------------------------------
// Table of C = 2^c_log2 slots: keys[s] = hash field | fwd/rev flag bits (63/62, as
// kmer_count_table.h:21-30), cnt[0][s] = fwd_count, cnt[1][s] = rev_count (uint32: the reference's
// uint8 counters + uint32 overflow table, bs/kmer_counter.cpp:680-685).  Open addressing, linear
// probing; the slot is the hash bits right below the sub-bin bits.  When the bin is done every used
// slot is written once: {canonical k-mer | flags, rev << 32 | fwd} to the distinct list, and the
// k-mer | flags again to the solid list if fwd + rev >= min_count (kmer_passes,
// modules/bio_mapred/kmerize_bf.cpp:290-318).
constexpr int kBinItems = 4;
struct BinGeom {
  int k;
  int low_bits;       // hash bits a word keeps (2k - batch_bits - part_bits)
  int sub_bits;
  int c_log2;
  uint64_t prefix0;   // (batch << part_bits | first partition of this rank): hash bits above the word's, for bin 0
  uint32_t min_count;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) kmer_count_bins_kernel(const unsigned long long* __restrict__ words,
                                                                  const unsigned long long* __restrict__ bin_off,
                                                                  const unsigned long long* __restrict__ bin_cnt,
                                                                  BinGeom G, CountEntry* __restrict__ out_all,
                                                                  unsigned long long cap_all,
                                                                  unsigned long long* __restrict__ out_solid,
                                                                  unsigned long long cap_solid,
                                                                  unsigned long long* __restrict__ counters /*[0] distinct [1] solid*/,
                                                                  int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t C = 1u << G.c_log2;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);  // C
  unsigned int* cnt = reinterpret_cast<unsigned int*>(keys + C);               // 2 * C
  __shared__ unsigned long long base_all, base_solid;
  __shared__ unsigned int blk_used, blk_solid, cur_all, cur_solid;
  const uint32_t bin = blockIdx.x;
  const unsigned long long n = bin_cnt[bin];
  if (n == 0) return;
  const unsigned tid = threadIdx.x, lane = tid & 31;
  const unsigned long long* src = words + bin_off[bin];
  // the first tile is on its way while the table is cleared
  unsigned long long w[kBinItems], wn[kBinItems];
#pragma unroll
  for (int j = 0; j < kBinItems; ++j) {
    const unsigned long long i = (unsigned long long)j * THREADS + tid;
    w[j] = i < n ? src[i] : ~0ULL;
  }
  for (uint32_t s = tid; s < C; s += THREADS) {
    keys[s] = kEmptyKey;
    cnt[s] = 0;
    cnt[C + s] = 0;
  }
  if (tid == 0) { blk_used = 0; blk_solid = 0; cur_all = 0; cur_solid = 0; }
  __syncthreads();
  const int slot_shift = max(G.low_bits - G.sub_bits - G.c_log2, 0);
  const uint32_t cmask = C - 1;
  const uint64_t field_mask = (1ULL << G.low_bits) - 1;
  bool full = false;
  // an instance word never has all 64 bits set (at most 2k - 7 + 3 <= 58 are used): ~0 = no item
  for (unsigned long long i0 = 0; i0 < n; i0 += (unsigned long long)THREADS * kBinItems) {
#pragma unroll
    for (int j = 0; j < kBinItems; ++j) {   // next tile: in flight while this one is counted
      const unsigned long long i = i0 + (unsigned long long)(kBinItems + j) * THREADS + tid;
      wn[j] = i < n ? src[i] : ~0ULL;
    }
#pragma unroll
    for (int j = 0; j < kBinItems; ++j) {
      if (w[j] == ~0ULL) continue;
      const uint64_t hl = w[j] >> 3;
      const uint64_t flags = ((w[j] & kWFwd) << 63) | ((w[j] & kWRev) << 61);   // -> kFwdFlag (bit 63), kRevFlag (bit 62)
      uint32_t s = (uint32_t)(hl >> slot_shift) & cmask;
      bool done = false;
      for (uint32_t probes = 0; probes < C; ++probes) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&keys[s]);
        if (cur == kEmptyKey) {
          cur = atomicCAS(&keys[s], (unsigned long long)kEmptyKey, (unsigned long long)(hl | flags));
          if (cur == kEmptyKey) { done = true; break; }   // claimed: key and flags are in place
        }
        if ((cur & field_mask) == hl) {
          if (flags & ~cur) atomicOr(&keys[s], (unsigned long long)flags);
          done = true;
          break;
        }
        s = (s + 1) & cmask;
      }
      if (!done) { full = true; continue; }
      atomicAdd(&cnt[((w[j] & kWFlip) ? C : 0u) + s], 1u);
    }
#pragma unroll
    for (int j = 0; j < kBinItems; ++j) w[j] = wn[j];
  }
  if (full) *overflow = 1;   // more distinct k-mers than slots: the host re-runs with finer sub-bins
  __syncthreads();
  // ---- write-out: every used slot once.  Count (strided, conflict-free), reserve the block's ranges of
  // the two lists with one global atomic each, then hand out positions warp by warp. -------------------
  uint32_t n_used = 0, n_solid = 0;
  for (uint32_t s = tid; s < C; s += THREADS) {
    if (keys[s] != kEmptyKey) {
      ++n_used;
      if ((unsigned long long)cnt[s] + cnt[C + s] >= G.min_count) ++n_solid;
    }
  }
  n_used = __reduce_add_sync(0xffffffffu, n_used);
  n_solid = __reduce_add_sync(0xffffffffu, n_solid);
  if (lane == 0) {
    if (n_used) atomicAdd(&blk_used, n_used);
    if (n_solid) atomicAdd(&blk_solid, n_solid);
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned long long a = atomicAdd(&counters[0], (unsigned long long)blk_used);
    const unsigned long long b = blk_solid ? atomicAdd(&counters[1], (unsigned long long)blk_solid) : 0ULL;
    base_all = a;
    base_solid = b;
    if ((cap_all && a + blk_used > cap_all) || (blk_solid && b + blk_solid > cap_solid)) *overflow = 2;
  }
  __syncthreads();
  const unsigned long long ba = base_all, bs = base_solid;
  const uint64_t prefix = (G.prefix0 + (bin >> G.sub_bits)) << G.low_bits;
  const unsigned lt = (1u << lane) - 1;
  for (uint32_t s0 = 0; s0 < C; s0 += THREADS) {
    const uint32_t s = s0 + tid;
    const unsigned long long kf = keys[s];
    const bool used = kf != kEmptyKey;
    const unsigned long long f = cnt[s], r = cnt[C + s];
    const bool solid = used && f + r >= G.min_count;
    const unsigned um = __ballot_sync(0xffffffffu, used), sm = __ballot_sync(0xffffffffu, solid);
    if (!um) continue;
    unsigned wa = 0, ws = 0;
    if (lane == 0) {
      wa = atomicAdd(&cur_all, (unsigned)__popc(um));
      if (sm) ws = atomicAdd(&cur_solid, (unsigned)__popc(sm));
    }
    wa = __shfl_sync(0xffffffffu, wa, 0);
    ws = __shfl_sync(0xffffffffu, ws, 0);
    if (used) {
      const uint64_t canon = khash_inv(prefix | (kf & field_mask), G.k);
      const unsigned long long key = canon | (kf & (kFwdFlag | kRevFlag));
      const unsigned long long oa = ba + wa + __popc(um & lt);
      if (oa < cap_all) {
        uint4 e;
        e.x = (unsigned)key; e.y = (unsigned)(key >> 32); e.z = (unsigned)f; e.w = (unsigned)r;
        reinterpret_cast<uint4*>(out_all)[oa] = e;
      }
      if (solid) {
        const unsigned long long os = bs + ws + __popc(sm & lt);
        if (os < cap_solid) out_solid[os] = key;
      }
    }
  }
}

// Distinct estimate over partitioned instance words (multi-GPU: run by the owner after the
// exchange, because the fused estimate of pass 1 only saw this rank's own reads).
__global__ void __launch_bounds__(256) kmer_estimate_words_kernel(const unsigned long long* const* __restrict__ vptr,
                                                                  const unsigned long long* __restrict__ vcnt,
                                                                  uint32_t tiles_per_part, int samp_shift,
                                                                  unsigned int* __restrict__ bitmap, uint64_t bit_mask) {
  const uint32_t v = blockIdx.x / tiles_per_part, t = blockIdx.x % tiles_per_part;
  const unsigned long long cnt = vcnt[v];
  const unsigned long long* src = vptr[v];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const unsigned long long idx = (unsigned long long)t * 4096 + (unsigned long long)i * 256 + threadIdx.x;
    if (idx >= cnt) continue;
    const uint64_t hl = src[idx] >> 3;
    if ((hl & ((1ULL << samp_shift) - 1)) == 0) {
      const uint64_t bit = (hl >> samp_shift) & bit_mask;
      atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
    }
  }
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

__global__ void fill_u64_kernel(unsigned long long* __restrict__ t, uint64_t n, unsigned long long v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = v;
}

// Sweep of the distinct list (export only).  Each 256-thread block covers kSweepPerBlock entries,
// stages the passing ones (fwd+rev >= min_count) in shared memory and appends them with ONE global
// atomic per block.
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
// The home bucket is the TOP bits of the hash that orders the counting bins, and pass 3 emits the
// solid k-mers roughly bin by bin: consecutive threads insert into nearby buckets, so building the
// set is a near-sequential write instead of one random DRAM CAS per key.
__global__ void solid_insert_kernel(const unsigned long long* __restrict__ keys, uint64_t n, int k,
                                    unsigned long long* __restrict__ set, uint64_t bucket_mask, int bucket_shift) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long kf = keys[i];
  if (kf == kEmptyKey) return;   // padding of a gathered list
  uint64_t b = khash(kf & kKmerMask, k) >> bucket_shift;
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

int log2_exact(uint64_t x) {
  int l = 0;
  while ((1ULL << l) < x) ++l;
  return l;
}

template <int MAXIT, int RPW>
void launch_partition(Context* c, uint64_t r0, uint64_t n_reads, const PartGeom& G, unsigned long long* cursors,
                      const unsigned long long* part_addr, unsigned long long cap, unsigned int* bitmap, uint64_t bit_mask,
                      int samp_shift, int* overflow) {
  constexpr int tile_reads = kPartWarps * RPW;
  constexpr int tile_kmers = tile_reads * MAXIT * 32;
  constexpr int wpr = MAXIT + 1;
  const size_t smem = (size_t)tile_kmers * 8 + (size_t)(tile_reads * wpr + 2) * 12 + (size_t)tile_kmers * 2 +
                      ((size_t)16 << G.part_bits);
  // the opt-in is per device: set it on every launch (cheap) rather than once per process
  BGX_CUDA(cudaFuncSetAttribute(kmer_partition_kernel<MAXIT, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks_per_sm = 1;
  BGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kmer_partition_kernel<MAXIT, RPW>,
                                                         kPartThreads, smem));
  blocks_per_sm = std::max(blocks_per_sm, 1);
  // persistent grid: every SM fully occupied, tiles handed out grid-stride
  const unsigned n_tiles = (unsigned)((n_reads + tile_reads - 1) / tile_reads);
  const unsigned grid = std::min<unsigned>(n_tiles, (unsigned)(kNumSMs * blocks_per_sm));
  note_launch();
  // a range of reads: word offsets are absolute, so only the per-read arrays shift
  kmer_partition_kernel<MAXIT, RPW><<<grid, kPartThreads, smem, c->stream>>>(
      c->words.p, c->has_n ? c->nmask.p : nullptr, c->word_off.p + r0, c->lens.p + r0, (uint32_t)n_reads, G, cursors,
      part_addr, cap, bitmap, bit_mask, samp_shift, overflow);
  BGX_CUDA(cudaGetLastError());
}

void run_partition(Context* c, int maxit, uint64_t r0, uint64_t n_reads, const PartGeom& G, unsigned long long* cursors,
                   const unsigned long long* part_addr, unsigned long long cap, unsigned int* bitmap, uint64_t bit_mask,
                   int samp_shift, int* overflow) {
  if (maxit <= 4)
    launch_partition<4, 4>(c, r0, n_reads, G, cursors, part_addr, cap, bitmap, bit_mask, samp_shift, overflow);
  else
    launch_partition<8, 2>(c, r0, n_reads, G, cursors, part_addr, cap, bitmap, bit_mask, samp_shift, overflow);
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

// pass-1 output for one batch (single GPU, or the NCCL form of the exchange)
struct Partitioned {
  DevBuf<unsigned long long> pk;
  std::vector<unsigned long long> base, count;  // per partition: first word, words
  unsigned long long cap = 0;                   // stride between partitions unless `exact`
  bool exact = false;                           // re-run with exact offsets (a heavy hitter overfilled a partition)
};

// One sweep of pass 1 over all reads: partition p goes to the cap words at addr[p].  Returns the
// overflow flag; counts[p] = instances of partition p (exact even when the sweep overflowed).
int sweep_partition(Context* c, const PartGeom& G, int maxit, Estimator* est, const std::vector<unsigned long long>& addr,
                    unsigned long long cap, std::vector<unsigned long long>* counts) {
  cudaStream_t s = c->stream;
  const uint64_t n = c->n_reads;
  const int P = 1 << G.part_bits;
  DevBuf<unsigned long long> cursors(P, s), part_addr(P, s);
  DevBuf<int> overflow(1, s);
  BGX_CUDA(cudaMemcpyAsync(part_addr.p, addr.data(), P * 8, cudaMemcpyHostToDevice, s));
  BGX_CUDA(cudaMemsetAsync(cursors.p, 0, P * 8, s));
  BGX_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
  ScopedStage st(c, "count_partition");
  unsigned int* bitmap = est ? est->bitmap.p : nullptr;
  const uint64_t bit_mask = est ? est->bits - 1 : 0;
  const int shift = est ? est->shift : 4;
  if (n && !c->upload.empty()) {
    // the reads are still arriving (bgx_add_reads_packed_async): one launch per upload chunk, each
    // ordered after its own copy only, all appending to the same partitions through the cursors
    uint64_t pos = 0;
    for (Context::UploadChunk& ch : c->upload) {
      if (ch.r0 > pos) run_partition(c, maxit, pos, ch.r0 - pos, G, cursors.p, part_addr.p, cap, bitmap, bit_mask, shift, overflow.p);
      BGX_CUDA(cudaStreamWaitEvent(s, ch.ev, 0));
      cudaEventDestroy(ch.ev);
      if (ch.r1 > ch.r0) run_partition(c, maxit, ch.r0, ch.r1 - ch.r0, G, cursors.p, part_addr.p, cap, bitmap, bit_mask, shift, overflow.p);
      pos = ch.r1;
    }
    c->upload.clear();
    if (pos < n) run_partition(c, maxit, pos, n - pos, G, cursors.p, part_addr.p, cap, bitmap, bit_mask, shift, overflow.p);
  } else if (n) {
    reads_ready(c);
    run_partition(c, maxit, 0, n, G, cursors.p, part_addr.p, cap, bitmap, bit_mask, shift, overflow.p);
  }
  int h_over = 0;
  counts->assign(P, 0);
  BGX_CUDA(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaMemcpyAsync(counts->data(), cursors.p, P * 8, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  st.stop();
  return h_over;
}

// Pass 1 into a buffer of this GPU: K = the k-mer instances expected in the batch.  est (optional)
// receives the fused sample.
void partition_reads(Context* c, uint64_t K, const PartGeom& G, int maxit, Estimator* est, Partitioned* out) {
  cudaStream_t s = c->stream;
  const int P = 1 << G.part_bits;
  unsigned long long cap = K / P + K / (8ull * P) + 4096;  // hash partitions are near uniform
  out->pk.alloc((size_t)cap * P, s);
  out->base.assign(P, 0);
  std::vector<unsigned long long> addr(P);
  for (int p = 0; p < P; ++p) {
    out->base[p] = (unsigned long long)p * cap;
    addr[p] = (unsigned long long)(uintptr_t)(out->pk.p + out->base[p]);
  }
  int h_over = sweep_partition(c, G, maxit, est, addr, cap, &out->count);
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
    for (int p = 0; p < P; ++p) addr[p] = (unsigned long long)(uintptr_t)(out->pk.p + out->base[p]);
    h_over = sweep_partition(c, G, maxit, est, addr, cap, &out->count);
    BGX_CHECK(!h_over, "internal: exact partition pass overflowed");
    c->add_stat("count_partition_reruns", 1);
    out->cap = cap;
    out->exact = true;
  }
}

// the partitions this rank will count: (device address, instance count, local partition) each
struct Owned {
  std::vector<unsigned long long> ptr, cnt;
  std::vector<uint32_t> pl;
  DevBuf<unsigned long long> rbuf;  // what arrived from the peers
  uint64_t n_inst = 0;
};

// Single GPU: the P partitions where pass 1 left them.  Multi-GPU: every partition goes to its
// owner over NVLink; partition p of source rank s arrives as its own "virtual partition", listed
// partition-major.
void exchange_partitions(Context* c, const Partitioned& pt, int P, Owned* own) {
  cudaStream_t s = c->stream;
  const int N = c->dist.nranks, R = c->dist.rank;
  own->ptr.clear();
  own->cnt.clear();
  own->pl.clear();
  if (N == 1) {
    for (int p = 0; p < P; ++p) {
      own->ptr.push_back((unsigned long long)(uintptr_t)(pt.pk.p + pt.base[p]));
      own->cnt.push_back(pt.count[p]);
      own->pl.push_back((uint32_t)p);
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
          own->pl.push_back((uint32_t)pl);
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
        own->pl.push_back((uint32_t)pl);
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

// Multi-GPU, the default: pass 1 stores every run straight into the memory of the GPU that owns its
// partition (peer mapping over NVLink), so counting and exchange are ONE kernel and nothing is staged
// on the sender.  The owner's buffer has one region of `cap` words per (source rank, local partition);
// K_max = the most instances any rank brings to this batch (the same value on every rank).  If a
// region overflows anywhere (a heavy hitter), every rank re-runs the sweep against exact offsets.
void partition_direct(Context* c, uint64_t K_max, const PartGeom& G, int maxit, Owned* own, Partitioned* pt) {
  cudaStream_t s = c->stream;
  const int N = c->dist.nranks, R = c->dist.rank;
  const int P = 1 << G.part_bits, Pl = P / N;
  unsigned long long cap = K_max / P + K_max / (8ull * P) + 4096;
  void* peer[64];
  std::vector<unsigned long long> addr(P), counts;
  const bool staged = dist_exchange_mode() == Exchange::COPY;   // pass 1 writes local memory, copy engines move the blocks
  {
    ScopedStage st(c, "count_exchange");
    own->rbuf.alloc((size_t)N * Pl * cap, s);
    if (staged) pt->pk.alloc((size_t)P * cap, s);
    dist_map_peers(c, own->rbuf.p, peer);   // also: every owner's buffer is ready to be written
    st.stop();
  }
  for (int p = 0; p < P; ++p)
    addr[p] = staged ? (unsigned long long)(uintptr_t)(pt->pk.p + (size_t)p * cap)
                     : (unsigned long long)(uintptr_t)(static_cast<unsigned long long*>(peer[p / Pl]) + ((size_t)R * Pl + p % Pl) * cap);
  int h_over = sweep_partition(c, G, maxit, nullptr, addr, cap, &counts);
  ScopedStage st(c, "count_exchange");
  std::vector<uint64_t> mine(P + 1), all((size_t)N * (P + 1));
  for (int p = 0; p < P; ++p) mine[p] = counts[p];
  mine[P] = h_over ? 1 : 0;
  dist_allgather_host_u64(c, mine.data(), P + 1, all.data());   // also: every rank's stores have landed
  auto cnt_of = [&](int src, int p) { return all[(size_t)src * (P + 1) + p]; };
  bool any_over = false;
  for (int r = 0; r < N; ++r) any_over = any_over || all[(size_t)r * (P + 1) + P] != 0;
  std::vector<uint64_t> off((size_t)N * Pl, 0);   // [src][pl] -> first word in this owner's buffer
  bool own_in_pk = false;                          // this rank's own block stays where pass 1 wrote it
  if (!any_over) {
    for (int src = 0; src < N; ++src)
      for (int pl = 0; pl < Pl; ++pl) off[(size_t)src * Pl + pl] = ((size_t)src * Pl + pl) * cap;
    if (staged) {
      // the block of partitions a peer owns is contiguous here and lands contiguously there: ONE copy per peer
      std::vector<PeerCopy> copies;
      for (int d = 0; d < N; ++d) {
        if (d == R) continue;
        PeerCopy pc;
        pc.dst = static_cast<unsigned long long*>(peer[d]) + (size_t)R * Pl * cap;
        pc.src = pt->pk.p + (size_t)d * Pl * cap;
        pc.bytes = (size_t)Pl * cap * 8;
        pc.peer = d;
        copies.push_back(pc);
      }
      dist_peer_copies(c, copies);
      dist_barrier(c);   // every rank's blocks have arrived
      own_in_pk = true;
    }
  } else {
    // exact layout, identical arithmetic on every rank: owner d lays its regions out partition-major
    auto layout = [&](int d, std::vector<uint64_t>* o) {
      uint64_t total = 0;
      for (int pl = 0; pl < Pl; ++pl)
        for (int src = 0; src < N; ++src) {
          (*o)[(size_t)src * Pl + pl] = total;
          total += cnt_of(src, d * Pl + pl);
        }
      return total;
    };
    const uint64_t total = layout(R, &off);
    own->rbuf.alloc(std::max<uint64_t>(total, 1), s);
    dist_map_peers(c, own->rbuf.p, peer);
    cap = 0;
    std::vector<uint64_t> od((size_t)N * Pl);
    for (int d = 0; d < N; ++d) {
      layout(d, &od);
      for (int pl = 0; pl < Pl; ++pl) {
        addr[d * Pl + pl] = (unsigned long long)(uintptr_t)(static_cast<unsigned long long*>(peer[d]) + od[(size_t)R * Pl + pl]);
        cap = std::max<unsigned long long>(cap, counts[d * Pl + pl]);
      }
    }
    st.stop();
    h_over = sweep_partition(c, G, maxit, nullptr, addr, cap, &counts);
    BGX_CHECK(!h_over, "internal: exact partition pass overflowed");
    c->add_stat("count_partition_reruns", 1);
    ScopedStage st2(c, "count_exchange");
    uint64_t dummy = 0;
    std::vector<uint64_t> dummies(N);
    dist_allgather_host_u64(c, &dummy, 1, dummies.data());   // every rank's stores have landed
    st2.stop();
  }
  own->ptr.clear();
  own->cnt.clear();
  own->pl.clear();
  uint64_t out_words = 0;
  for (int pl = 0; pl < Pl; ++pl)
    for (int src = 0; src < N; ++src) {
      const unsigned long long* ptr = own_in_pk && src == R ? pt->pk.p + ((size_t)R * Pl + pl) * cap
                                                            : own->rbuf.p + off[(size_t)src * Pl + pl];
      own->ptr.push_back((unsigned long long)(uintptr_t)ptr);
      own->cnt.push_back(cnt_of(src, R * Pl + pl));
      own->pl.push_back((uint32_t)pl);
    }
  for (int p = 0; p < P; ++p)
    if (p / Pl != R) out_words += counts[p];
  c->add_stat("count_exchange_bytes_out", 8.0 * (double)out_words);
  own->n_inst = 0;
  for (unsigned long long v : own->cnt) own->n_inst += v;
  st.stop();
}

// device-side view of an Owned list, ready for the tile kernels
struct OwnedDev {
  DevBuf<unsigned long long> ptr, cnt;
  DevBuf<uint32_t> pl;
  uint32_t V = 0;
  uint64_t max_count = 0;
  const unsigned long long* const* part_ptr() const { return reinterpret_cast<const unsigned long long* const*>(ptr.p); }
  uint32_t blocks(uint64_t tile) const { return (uint32_t)std::max<uint64_t>(1, (max_count + tile - 1) / tile); }
};

void upload_owned(Context* c, const Owned& own, OwnedDev* d) {
  cudaStream_t s = c->stream;
  d->V = (uint32_t)own.ptr.size();
  d->ptr.alloc(d->V, s);
  d->cnt.alloc(d->V, s);
  d->pl.alloc(d->V, s);
  BGX_CUDA(cudaMemcpyAsync(d->ptr.p, own.ptr.data(), d->V * 8, cudaMemcpyHostToDevice, s));
  BGX_CUDA(cudaMemcpyAsync(d->cnt.p, own.cnt.data(), d->V * 8, cudaMemcpyHostToDevice, s));
  BGX_CUDA(cudaMemcpyAsync(d->pl.p, own.pl.data(), d->V * 4, cudaMemcpyHostToDevice, s));
  d->max_count = 0;
  for (unsigned long long v : own.cnt) d->max_count = std::max<uint64_t>(d->max_count, v);
  BGX_CHECK((uint64_t)d->blocks(4096) * d->V < (1ull << 31), "too many k-mer tiles for one launch");
  // the host vectors must outlive the copies
  BGX_CUDA(cudaStreamSynchronize(s));
}

// How many hash-range batches the counting runs in (a power of two).  One batch keeps all K instance
// words twice (pass-1 output and the split copy; multi-GPU also what the peers send) in HBM; more
// batches bound those buffers to ~45 % of the device so that inputs like GRCh38 30x on 8 GPUs
// (3 x 75 GB of words per GPU) still fit.  Every rank must use the same count (each batch is a
// collective exchange).  count_batch_reads (option / BGX_COUNT_BATCH_READS) asks for at least
// ceil(reads / that) batches.
uint64_t choose_batches(Context* c, uint64_t K_local, uint64_t K_share) {
  const int N = c->dist.nranks;
  uint64_t batch_reads = c->opt.count_batch_reads > 0 ? (uint64_t)c->opt.count_batch_reads : 0;
  if (const char* e = getenv("BGX_COUNT_BATCH_READS")) batch_reads = strtoull(e, nullptr, 10);  // test hook
  uint64_t batches = plan_count_batches(K_local, K_share, N, c->total_mem, batch_reads, c->n_reads);
  if (N > 1) {
    std::vector<uint64_t> all(N);
    dist_allgather_host_u64(c, &batches, 1, all.data());
    for (uint64_t v : all) batches = std::max(batches, v);
  }
  BGX_CHECK(batches <= 4096, "k-mer counting would need more than 4096 batches");
  return batches;
}

// results of one batch, appended to the lists of the whole run at the end
struct BatchOut {
  DevBuf<CountEntry> all;
  DevBuf<unsigned long long> solid;
  uint64_t n_all = 0, n_solid = 0;
};

}  // namespace

void stage_count_kmers(Context* c) {
  cudaStream_t s = c->stream;
  const int k = c->opt.kmer_size;
  const int N = c->dist.nranks, R = c->dist.rank;
  const int rank_bits = log2_exact((uint64_t)N);
  BGX_CHECK(c->n_reads > 0 || N > 1, "bgx_count_kmers: no reads");
  ScopedStage st_all(c, "count_total");
  c->table.release();
  c->solid.release();
  const uint64_t K = c->n_kmer_instances;  // this rank's reads
  uint64_t K_all = K;                      // all ranks' reads
  dist_allreduce_sum_host_u64(c, &K_all, 1);
  const uint64_t K_share = K_all / N;      // instances this rank will own (hash-uniform)
  uint64_t K_max = K;                      // the most any rank brings
  if (N > 1) {
    std::vector<uint64_t> ks(N);
    dist_allgather_host_u64(c, &K, 1, ks.data());
    for (uint64_t v : ks) K_max = std::max(K_max, v);
  }
  // BGX_EXCHANGE=nccl: partition locally, then one NCCL send/recv per peer (the round-1 form; A/B hook)
  const bool direct_exchange = dist_exchange_mode() != Exchange::NCCL;
  const int maxit = (int)((std::max<int64_t>((int64_t)c->max_len - k + 1, 1) + 31) / 32);
  BGX_CHECK(maxit <= 8, "read longer than 255 bases");
  const uint64_t batches = choose_batches(c, K, K_share);
  const int batch_bits = log2_exact(batches);
  c->set_stat("count_batches", (double)batches);

  int part_bits = plan_part_bits(K_share, batches, rank_bits);
  if (const char* e = getenv("BGX_PART_BITS")) part_bits = std::max(rank_bits, std::min(kMaxPartBits, atoi(e)));  // experiment hook
  BGX_CHECK(2 * k - batch_bits - part_bits >= 16, "k-mer too short for this many batches / partitions");
  int c_log2 = 12;  // 4096 slots = 64 KB of shared memory per block, three blocks per SM
  if (const char* e = getenv("BGX_BIN_SLOTS_LOG2")) c_log2 = std::max(9, std::min(13, atoi(e)));  // experiment hook

  // The per-k-mer counts of ALL distinct k-mers (bgx_export_kmers below min_count) are kept when the
  // counting is one batch, or small; a big batched count only keeps the solid k-mers.
  const bool keep_all = batches == 1 || 16.0 * (double)K_share < 0.05 * (double)c->total_mem;
  DevBuf<int> overflow(1, s);
  std::vector<BatchOut> outs(batches);
  uint64_t n_inst = 0, total_all = 0, total_solid = 0;
  double alg_part = 0, alg_split = 0, alg_bins = 0;
  int sub_bits_used = 0;

  for (uint64_t b = 0; b < batches; ++b) {
    PartGeom G{k, batch_bits, (uint32_t)b, part_bits};
    int P = 1 << part_bits;
    const int low_bits = 2 * k - batch_bits - part_bits;
    // ---- pass 1 (+ exchange) -----------------------------------------------------------------------
    Estimator est;
    const uint64_t Kb = K / batches + K / (16 * batches) + 1024;  // this rank's instances in the batch (hash-uniform)
    est.init(pow2_ceil(std::max<uint64_t>(1 << 20, std::max(Kb, K_share / batches) / 8)), 4, s);
    Partitioned pt;
    Owned own;
    if (N > 1 && direct_exchange) {
      partition_direct(c, K_max / batches + K_max / (16 * batches) + 1024, G, maxit, &own, &pt);
    } else {
      partition_reads(c, Kb, G, maxit, N == 1 ? &est : nullptr, &pt);
      exchange_partitions(c, pt, P, &own);
    }
    OwnedDev od;
    upload_owned(c, own, &od);
    n_inst += own.n_inst;
    alg_part += (double)c->n_bases / 4 + 8.0 * (double)own.n_inst;
    if (N > 1) {
      // the owner estimates the distinct count of what it received
      KLAUNCH(kmer_estimate_words_kernel)<<<od.blocks(4096) * od.V, 256, 0, s>>>(od.part_ptr(), od.cnt.p, od.blocks(4096),
                                                                          est.shift, est.bitmap.p, est.bits - 1);
      BGX_CUDA(cudaGetLastError());
    }
    uint64_t est_distinct = est.estimate(s);
    est_distinct = std::min<uint64_t>(est_distinct + est_distinct / 16 + 4096, own.n_inst + 1);
    c->add_stat("kmer_distinct_estimate", (double)est_distinct);
    est.bitmap.release();

    // ---- pass 2 + 3, finer sub-bins if a bin overflows its table -------------------------------------
    const int Pl = P / N;
    // sub-bins so that a bin holds ~half a table of distinct k-mers on average
    const uint64_t per_bin = (1ull << c_log2) / 2;
    int sub_bits = 0;
    while (sub_bits < kMaxSubBits && ((uint64_t)Pl << sub_bits) * per_bin < est_distinct) ++sub_bits;
    // the 2048-way split writes 4-word runs (chr20 on 2 GPUs: 15.4 ms against 11.0 at 1024): stay at 1024 while
    // the tables fill to less than two thirds (the count kernel pays ~10 % for the longer probe chains)
    if (sub_bits == kMaxSubBits && ((uint64_t)Pl << (kMaxSubBits - 1)) * (per_bin + per_bin / 3) >= est_distinct) --sub_bits;
    if (const char* e = getenv("BGX_SUB_BITS")) sub_bits = std::max(0, std::min(kMaxSubBits, atoi(e)));  // test hook
    BatchOut& bo = outs[b];
    for (int tries = 0;; ++tries) {
      sub_bits = std::min(sub_bits, low_bits - 1);
      const int S = 1 << sub_bits;
      const size_t n_bins = (size_t)Pl * S;
      const int sub_shift = low_bits - sub_bits;
      DevBuf<unsigned long long> hist(n_bins, s), bin_off(n_bins, s), cursors(n_bins, s), scal(2, s);
      DevBuf<unsigned long long> split(std::max<uint64_t>(own.n_inst, 1), s);
      {
        ScopedStage st(c, "count_split");
        BGX_CUDA(cudaMemsetAsync(hist.p, 0, n_bins * 8, s));
        BGX_CUDA(cudaMemsetAsync(cursors.p, 0, n_bins * 8, s));
        if (own.n_inst) {
          const uint32_t hb = od.blocks((uint64_t)kSplitTile * kHistTiles);
          KLAUNCH(kmer_subhist_kernel)<<<hb * od.V, kSplitThreads, 0, s>>>(od.part_ptr(), od.cnt.p, od.pl.p, hb, sub_shift, sub_bits,
                                                                     hist.p);
        }
        KLAUNCH(scan_u64_kernel)<<<1, 1024, 0, s>>>(hist.p, n_bins, bin_off.p, scal.p, scal.p + 1);
        if (own.n_inst) {
          const size_t smem = (size_t)kSplitTile * 8 + (size_t)S * 16;
          BGX_CUDA(cudaFuncSetAttribute(kmer_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          const uint32_t tb = od.blocks(kSplitTile);
          KLAUNCH(kmer_split_kernel)<<<tb * od.V, kSplitThreads, smem, s>>>(od.part_ptr(), od.cnt.p, od.pl.p, tb, sub_shift, sub_bits,
                                                                      cursors.p, bin_off.p, split.p);
        }
        BGX_CUDA(cudaGetLastError());
        st.stop();
      }
      unsigned long long h_scal[2] = {0, 0};
      BGX_CUDA(cudaMemcpyAsync(h_scal, scal.p, 16, cudaMemcpyDeviceToHost, s));
      BGX_CUDA(cudaStreamSynchronize(s));
      BGX_CHECK(h_scal[0] == own.n_inst, "internal: sub-bin histogram does not add up");
      c->set_stat("count_largest_bin", (double)h_scal[1]);

      ScopedStage st(c, "count_kernel");
      const uint64_t cap_est = est_distinct + est_distinct / 8 + 65536;
      const uint64_t cap_all = keep_all ? cap_est : 0;   // 0: the distinct list is not written
      const uint64_t cap_solid = std::min<uint64_t>(cap_est, own.n_inst / (uint64_t)c->opt.min_kmer_count + 1);
      bo.all.alloc(std::max<uint64_t>(cap_all, 1), s);
      bo.solid.alloc(cap_solid, s);
      DevBuf<unsigned long long> counters(2, s);
      BGX_CUDA(cudaMemsetAsync(counters.p, 0, 16, s));
      BGX_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
      BinGeom BG{k, low_bits, sub_bits, c_log2, ((uint64_t)b << part_bits) | (uint64_t)(R * Pl), (uint32_t)c->opt.min_kmer_count};
      const size_t smem = (size_t)16 << c_log2;
      static const int bin_threads = [] { const char* e = getenv("BGX_BIN_THREADS"); return e ? atoi(e) : 512; }();  // experiment hook
      note_launch();
      if (bin_threads == 256) {
        BGX_CUDA(cudaFuncSetAttribute(kmer_count_bins_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kmer_count_bins_kernel<256><<<(unsigned)n_bins, 256, smem, s>>>(split.p, bin_off.p, hist.p, BG, bo.all.p, cap_all, bo.solid.p,
                                                                      cap_solid, counters.p, overflow.p);
      } else {
        BGX_CUDA(cudaFuncSetAttribute(kmer_count_bins_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kmer_count_bins_kernel<512><<<(unsigned)n_bins, 512, smem, s>>>(split.p, bin_off.p, hist.p, BG, bo.all.p, cap_all, bo.solid.p,
                                                                      cap_solid, counters.p, overflow.p);
      }
      BGX_CUDA(cudaGetLastError());
      unsigned long long h_cnt[2];
      int h_over = 0;
      BGX_CUDA(cudaMemcpyAsync(h_cnt, counters.p, 16, cudaMemcpyDeviceToHost, s));
      BGX_CUDA(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      BGX_CUDA(cudaStreamSynchronize(s));
      st.stop();
      alg_split += 24.0 * (double)own.n_inst;
      alg_bins += 8.0 * (double)own.n_inst + 16.0 * (double)h_cnt[0] + 8.0 * (double)h_cnt[1];
      if (!h_over) {
        bo.n_all = h_cnt[0];
        bo.n_solid = h_cnt[1];
        sub_bits_used = std::max(sub_bits_used, sub_bits);
        if (batches > 1) {
          // the lists were sized by an upper bound: keep exact-sized copies while the other batches run
          DevBuf<unsigned long long> exact(std::max<uint64_t>(bo.n_solid, 1), s);
          if (bo.n_solid) BGX_CUDA(cudaMemcpyAsync(exact.p, bo.solid.p, bo.n_solid * 8, cudaMemcpyDeviceToDevice, s));
          bo.solid = std::move(exact);
          if (keep_all) {
            DevBuf<CountEntry> exact_all(std::max<uint64_t>(bo.n_all, 1), s);
            if (bo.n_all) BGX_CUDA(cudaMemcpyAsync(exact_all.p, bo.all.p, bo.n_all * sizeof(CountEntry), cudaMemcpyDeviceToDevice, s));
            bo.all = std::move(exact_all);
          } else {
            bo.all.release();
          }
        }
        break;
      }
      // a sub-bin had more distinct k-mers than table slots, or the estimate was low: finer bins,
      // then larger tables, larger lists.  The reference throws io_exception("Kmer table (...) too
      // small") when ITS table fills (bs/kmer_count_table.h:75).
      BGX_CHECK(tries < 6, "Kmer table too small");
      c->add_stat("count_bin_reruns", 1);
      if (h_over == 2) est_distinct = std::max<uint64_t>(est_distinct * 2, h_cnt[0] + h_cnt[0] / 8);
      else if (sub_bits < std::min(kMaxSubBits, low_bits - 1)) ++sub_bits;
      else if (c_log2 < 13) ++c_log2;
      else BGX_CHECK(false, "Kmer table too small");
    }
    total_all += bo.n_all;
    total_solid += bo.n_solid;
  }

  // ---- the lists of the whole run: batch 0's buffers as they are, or the batches concatenated ---------
  DevBuf<unsigned long long> solid_list;
  {
    ScopedStage st(c, "count_filter");
    if (batches == 1) {
      c->table = std::move(outs[0].all);
      solid_list = std::move(outs[0].solid);
    } else {
      if (keep_all) c->table.alloc(std::max<uint64_t>(total_all, 1), s);
      solid_list.alloc(std::max<uint64_t>(total_solid, 1), s);
      uint64_t oa = 0, os = 0;
      for (BatchOut& bo : outs) {
        if (keep_all && bo.n_all)
          BGX_CUDA(cudaMemcpyAsync(c->table.p + oa, bo.all.p, bo.n_all * sizeof(CountEntry), cudaMemcpyDeviceToDevice, s));
        if (bo.n_solid) BGX_CUDA(cudaMemcpyAsync(solid_list.p + os, bo.solid.p, bo.n_solid * 8, cudaMemcpyDeviceToDevice, s));
        oa += bo.n_all;
        os += bo.n_solid;
        bo.all.release();
        bo.solid.release();
      }
      c->set_stat("kmer_counts_dropped", keep_all ? 0 : 1);
    }
    c->table_slots = c->table.p ? total_all : 0;
    c->n_distinct = total_all;
    uint64_t n_solid_local = total_solid;
    // multi-GPU: every rank needs the whole solid set for correction -> all-gather the owners' lists
    const unsigned long long* all_keys = solid_list.p;
    DevBuf<unsigned long long> gathered;
    c->n_solid = n_solid_local;
    uint64_t n_insert = n_solid_local;   // list entries handed to the insert kernel (padding included)
    if (N > 1) {
      // equal-sized slices (the largest owner's count, padded with empty keys): ONE in-place ncclAllGather
      std::vector<uint64_t> cnts(N);
      dist_allgather_host_u64(c, &n_solid_local, 1, cnts.data());
      uint64_t tot = 0, slice = 1;
      for (int r = 0; r < N; ++r) { tot += cnts[r]; slice = std::max(slice, cnts[r]); }
      gathered.alloc(slice * N, s);
      unsigned long long* mine = gathered.p + (uint64_t)R * slice;
      if (n_solid_local) BGX_CUDA(cudaMemcpyAsync(mine, solid_list.p, n_solid_local * 8, cudaMemcpyDeviceToDevice, s));
      if (slice > n_solid_local)
        KLAUNCH(fill_u64_kernel)<<<(unsigned)((slice - n_solid_local + 255) / 256), 256, 0, s>>>(mine + n_solid_local, slice - n_solid_local, kEmptyKey);
      dist_allgather_bytes(c, mine, gathered.p, slice * 8);
      all_keys = gathered.p;
      n_insert = slice * N;
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
    if (n_insert)
      KLAUNCH(solid_insert_kernel)<<<(unsigned)((n_insert + 255) / 256), 256, 0, s>>>(all_keys, n_insert, k, c->solid.p,
                                                                             c->solid_slots / 4 - 1,
                                                                             2 * k - log2_exact(c->solid_slots / 4));
    BGX_CUDA(cudaGetLastError());
    BGX_CUDA(cudaStreamSynchronize(s));  // the lists die with this scope
    st.stop();
  }
  c->counted = true;
  c->corrected = c->built = false;
  st_all.stop();
  c->set_stat("kmer_instances", (double)n_inst);
  c->set_stat("kmer_distinct", (double)c->n_distinct);
  c->set_stat("kmer_solid", (double)c->n_solid);
  c->set_stat("count_partitions", (double)(1 << part_bits));
  c->set_stat("count_sub_bins", (double)(1 << sub_bits_used));
  // partition pass: reads in, one 8-byte word per instance out; split: the words in twice (histogram,
  // scatter) and out once; bins: the words in once, every distinct k-mer out once (16 B) + the solid list
  c->set_stat("alg_bytes_count_partition", alg_part);
  c->set_stat("alg_bytes_count_split", alg_split);
  c->set_stat("alg_bytes_count_kernel", alg_bins);
  c->set_stat("alg_bytes_count", alg_part + alg_split + alg_bins);
}

void export_kmers(Context* c, uint32_t min_count, uint64_t* n_out, uint64_t** kmers, uint32_t** fwd, uint32_t** rev,
                  uint8_t** flags) {
  BGX_CHECK(c->counted && c->table.p,
            "bgx_export_kmers: call bgx_count_kmers first (and before bgx_reset_results; on inputs whose k-mer lists "
            "exceed a quarter of the device memory also before bgx_build_seqset, which releases them; inputs counted in "
            "batches that would not fit do not keep per-k-mer counts at all)");
  cudaStream_t s = c->stream;
  uint64_t slots = c->table_slots;
  DevBuf<unsigned long long> counters(2, s);
  unsigned long long h_cnt[2];
  BGX_CUDA(cudaMemsetAsync(counters.p, 0, 2 * sizeof(unsigned long long), s));
  const unsigned sweep_grid = (unsigned)std::max<uint64_t>(1, (slots + kSweepPerBlock - 1) / kSweepPerBlock);
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

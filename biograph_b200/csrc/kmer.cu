// kmer.cu -- k-mer counting over 2-bit packed reads and construction of the solid k-mer set.
//
// Replaces (reference, CPU): prob_pass_processor / exact_pass_processor (bs/kmer_counter.h:275-326,
// bs/kmer_counter.cpp:579-697), kmer_count_table::increment (bs/kmer_count_table.h:54-103),
// kmerizer::run's min-count filter (modules/bio_mapred/kmerize_bf.cpp:290-318) and kmer_set
// (modules/bio_mapred/kmer_set.cpp:522-631, lookup :296-360).
//
// B200 design: one pass (no probabilistic pre-pass, no temp files): a warp takes one read, its
// lanes pull the read's words once (coalesced), every lane forms the k-mers at positions
// lane, lane+32, ... by funnel shifts out of warp-shuffled words, canonicalises with brev, and
// upserts into ONE open-addressing table in HBM (16-byte slots: key|flags + both counters in the
// same 32-byte DRAM sector) with atomicCAS / atomicAdd / atomicOr.  The filter pass sweeps the
// table once and builds the 8-byte-slot solid hash set that correction probes.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "ctx.h"

namespace bgx {
namespace {

__device__ __forceinline__ void table_upsert(CountEntry* __restrict__ table, uint64_t slot_mask, uint64_t canon,
                                             bool flipped, uint64_t flags, int* __restrict__ overflow) {
  uint64_t slot = mix64(canon) & slot_mask;
  for (uint64_t probes = 0;; ++probes) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&table[slot].key);
    if (cur == kEmptyKey) {
      unsigned long long old = atomicCAS(&table[slot].key, (unsigned long long)kEmptyKey,
                                         (unsigned long long)(canon | flags));
      if (old == kEmptyKey) { flags = 0; break; }
      cur = old;
    }
    if ((cur & kKmerMask) == canon) {
      flags &= ~cur;
      break;
    }
    slot = (slot + 1) & slot_mask;
    if (probes > slot_mask) { *overflow = 1; return; }
  }
  if (flags) atomicOr(&table[slot].key, (unsigned long long)flags);
  atomicAdd(&table[slot].cnt, flipped ? (1ULL << 32) : 1ULL);
}

// Enumerate the k-mers of read r cooperatively in one warp: lanes < nw pull the read's words once
// (coalesced), every lane forms the k-mers at positions lane, lane+32, ... by funnel shifts out
// of warp-shuffled words.  f(canon, flipped, flags) is called for every valid k-mer instance
// (semantics of pass_processor::add, bs/kmer_counter.h:297-326).
template <typename F>
__device__ __forceinline__ void warp_read_kmers(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask,
                                                uint32_t base, int L, int k, F&& f) {
  const unsigned lane = lane_id();
  const int nk = L - k + 1;
  const unsigned nw = (unsigned)(L + 31) >> 5;
  uint64_t myw = lane < nw ? words[base + lane] : 0;
  uint32_t mym = (nmask != nullptr && lane < nw) ? nmask[base + lane] : 0;
  const int iters = (nk + 31) >> 5;
  for (int it = 0; it < iters; ++it) {
    uint64_t hi = __shfl_sync(0xffffffffu, myw, it);
    uint64_t lo = __shfl_sync(0xffffffffu, myw, it + 1);
    uint32_t mh = __shfl_sync(0xffffffffu, mym, it);
    uint32_t ml = __shfl_sync(0xffffffffu, mym, it + 1);
    int p = it * 32 + (int)lane;
    if (p >= nk) continue;
    unsigned s = lane * 2;
    uint64_t win = s ? ((hi << s) | (lo >> (64 - s))) : hi;
    uint32_t mwin = lane ? ((mh << lane) | (ml >> (32 - lane))) : mh;
    if (mwin >> (32 - k)) continue;  // an 'N' inside the window (bs/kmer_counter.h:306-311)
    uint64_t kmer = win >> (64 - 2 * k);
    bool flipped;
    uint64_t canon = canonicalize(kmer, k, flipped);
    // fwd_flag = first k-mer of the read, rev_flag = last; swapped when flipped
    // (bs/kmer_counter.h:318-321, bs/kmer_count_table.h:82-86)
    bool first = (p == 0), last = (p == nk - 1);
    uint64_t flags = 0;
    if (flipped ? last : first) flags |= kFwdFlag;
    if (flipped ? first : last) flags |= kRevFlag;
    f(canon, flipped, flags);
  }
}

// warp per read (grid-stride)
__global__ void __launch_bounds__(256) kmer_count_kernel(const uint64_t* __restrict__ words,
                                                         const uint32_t* __restrict__ nmask,
                                                         const uint32_t* __restrict__ word_off,
                                                         const uint16_t* __restrict__ lens, uint32_t n_reads, int k,
                                                         CountEntry* __restrict__ table, uint64_t slot_mask,
                                                         int* __restrict__ overflow) {
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += warps_total) {
    const int L = lens[r];
    if (L < k) continue;
    warp_read_kmers(words, nmask, word_off[r], L, k, [&](uint64_t canon, bool flipped, uint64_t flags) {
      table_upsert(table, slot_mask, canon, flipped, flags, overflow);
    });
  }
}

// Distinct-k-mer estimate for sizing the table: linear counting over the 1/16 of the hash space
// whose low 4 hash bits are zero (sampling by hash value is unbiased for distinct counts).
__global__ void __launch_bounds__(256) kmer_estimate_kernel(const uint64_t* __restrict__ words,
                                                            const uint32_t* __restrict__ nmask,
                                                            const uint32_t* __restrict__ word_off,
                                                            const uint16_t* __restrict__ lens, uint32_t n_reads, int k,
                                                            unsigned int* __restrict__ bitmap, uint64_t bit_mask) {
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += warps_total) {
    const int L = lens[r];
    if (L < k) continue;
    warp_read_kmers(words, nmask, word_off[r], L, k, [&](uint64_t canon, bool, uint64_t) {
      uint64_t h = mix64(canon ^ 0x9e3779b97f4a7c15ULL);
      if ((h & 15) == 0) {
        uint64_t bit = (h >> 4) & bit_mask;
        unsigned int m = 1u << (bit & 31);
        if (!(bitmap[bit >> 5] & m)) atomicOr(&bitmap[bit >> 5], m);
      }
    });
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

// insert into the solid set (8-byte slots).  Keys are distinct, so a plain CAS claim suffices.
__global__ void solid_insert_kernel(const unsigned long long* __restrict__ keys, uint64_t n,
                                    unsigned long long* __restrict__ set, uint64_t slot_mask) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long kf = keys[i];
  uint64_t slot = mix64(kf & kKmerMask) & slot_mask;
  while (atomicCAS(&set[slot], (unsigned long long)kEmptyKey, kf) != kEmptyKey) slot = (slot + 1) & slot_mask;
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

void stage_count_kmers(Context* c) {
  cudaStream_t s = c->stream;
  const int k = c->opt.kmer_size;
  BGX_CHECK(c->n_reads > 0, "bgx_count_kmers: no reads");
  ScopedStage st_all(c, "count_total");

  // Size the table from a distinct-k-mer estimate (one cheap extra pass over the packed reads):
  // load factor in (1/3, 2/3].  An overflow (estimate off) is retried below with a doubled table.
  const unsigned grid_reads = (unsigned)std::min<uint64_t>((c->n_reads * 32 + 255) / 256, (uint64_t)kNumSMs * 8);
  uint64_t est_distinct = 0;
  {
    ScopedStage st(c, "count_estimate");
    uint64_t bits = pow2_ceil(std::max<uint64_t>(1 << 20, c->n_kmer_instances / 8));
    DevBuf<unsigned int> bitmap(bits / 32, s);
    DevBuf<unsigned long long> ones(1, s);
    BGX_CUDA(cudaMemsetAsync(bitmap.p, 0, bits / 8, s));
    BGX_CUDA(cudaMemsetAsync(ones.p, 0, 8, s));
    KLAUNCH(kmer_estimate_kernel)<<<grid_reads, 256, 0, s>>>(c->words.p, c->has_n ? c->nmask.p : nullptr, c->word_off.p,
                                                           c->lens.p, (uint32_t)c->n_reads, k, bitmap.p, bits - 1);
    KLAUNCH(popcount_kernel)<<<(unsigned)((bits / 32 + 255) / 256), 256, 0, s>>>(bitmap.p, bits / 32, ones.p);
    BGX_CUDA(cudaGetLastError());
    unsigned long long h_ones = 0;
    BGX_CUDA(cudaMemcpyAsync(&h_ones, ones.p, 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    double zero_frac = std::max(1.0 / (double)bits, 1.0 - (double)h_ones / (double)bits);
    est_distinct = (uint64_t)(16.0 * -(double)bits * std::log(zero_frac));
    est_distinct = std::min<uint64_t>(est_distinct + est_distinct / 16 + 4096, c->n_kmer_instances + 1);
    st.stop();
  }
  c->set_stat("kmer_distinct_estimate", (double)est_distinct);
  uint64_t slots = pow2_ceil(std::max<uint64_t>(1024, est_distinct + est_distinct / 2));
  if (const char* e = getenv("BGX_TABLE_SLOTS_LOG2")) slots = 1ull << atoi(e);  // experiment hook
  size_t free_b = 0, total_b = 0;
  BGX_CUDA(cudaMemGetInfo(&free_b, &total_b));
  BGX_CHECK(slots * sizeof(CountEntry) < free_b, "not enough device memory for the k-mer table");
  c->table_slots = slots;
  c->table.alloc(slots, s);
  {
    ScopedStage st(c, "count_init");
    KLAUNCH(fill_empty_kernel)<<<(unsigned)((slots + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint4*>(c->table.p), slots);
    st.stop();
  }
  DevBuf<int> overflow(1, s);
  BGX_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
  {
    ScopedStage st(c, "count_kernel");
    // persistent-style grid: 148 SMs x 8 resident 256-thread CTAs
    KLAUNCH(kmer_count_kernel)<<<grid_reads, 256, 0, s>>>(c->words.p, c->has_n ? c->nmask.p : nullptr, c->word_off.p, c->lens.p,
                                             (uint32_t)c->n_reads, k, c->table.p, slots - 1, overflow.p);
    BGX_CUDA(cudaGetLastError());
    st.stop();
  }
  int h_over = 0;
  BGX_CUDA(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  // the reference throws io_exception("Kmer table (...) too small") (bs/kmer_count_table.h:75)
  BGX_CHECK(!h_over, "Kmer table too small");

  // filter (kmer_passes: fwd+rev >= min_count) and build the solid set: ONE sweep into buffers
  // sized by the bound #solid <= K / min_count.
  DevBuf<unsigned long long> counters(2, s);
  unsigned long long h_cnt[2];
  {
    ScopedStage st(c, "count_filter");
    uint64_t cap = std::min<uint64_t>(slots, c->n_kmer_instances / (uint64_t)c->opt.min_kmer_count + 1);
    DevBuf<unsigned long long> sk(cap, s), sc(cap, s);
    BGX_CUDA(cudaMemsetAsync(counters.p, 0, 2 * sizeof(unsigned long long), s));
    KLAUNCH(table_sweep_kernel)<<<(unsigned)((slots + kSweepPerBlock - 1) / kSweepPerBlock), 256, 0, s>>>(
        c->table.p, slots, (uint32_t)c->opt.min_kmer_count, counters.p, sk.p, sc.p, cap);
    BGX_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    c->n_distinct = h_cnt[0];
    c->n_solid = h_cnt[1];
    BGX_CHECK(c->n_solid <= cap, "internal: solid k-mer bound violated");
    // "Too many kmers for kmer table!" (kmer_set.cpp:554-556) has no analogue: the set is sized to fit.
    c->solid_slots = pow2_ceil(std::max<uint64_t>(1024, c->n_solid * 2));
    c->solid.alloc(c->solid_slots, s);
    KLAUNCH(fill_u64_kernel)<<<(unsigned)((c->solid_slots + 255) / 256), 256, 0, s>>>(c->solid.p, c->solid_slots, kEmptyKey);
    if (c->n_solid)
      KLAUNCH(solid_insert_kernel)<<<(unsigned)((c->n_solid + 255) / 256), 256, 0, s>>>(sk.p, c->n_solid, c->solid.p,
                                                                             c->solid_slots - 1);
    BGX_CUDA(cudaGetLastError());
    st.stop();
  }
  c->counted = true;
  c->corrected = c->built = false;
  st_all.stop();
  c->set_stat("kmer_instances", (double)c->n_kmer_instances);
  c->set_stat("kmer_distinct", (double)c->n_distinct);
  c->set_stat("kmer_solid", (double)c->n_solid);
  c->set_stat("count_table_slots", (double)slots);
  // B/4 + K*32 + T*16*2
  c->set_stat("alg_bytes_count", (double)c->n_bases / 4 + 32.0 * (double)c->n_kmer_instances + 32.0 * (double)slots);
}

void export_kmers(Context* c, uint32_t min_count, uint64_t* n_out, uint64_t** kmers, uint32_t** fwd, uint32_t** rev,
                  uint8_t** flags) {
  BGX_CHECK(c->counted && c->table.p, "bgx_export_kmers: call bgx_count_kmers first (and before bgx_reset_results)");
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
  *kmers = (uint64_t*)malloc(na * 8);
  *fwd = (uint32_t*)malloc(na * 4);
  *rev = (uint32_t*)malloc(na * 4);
  *flags = (uint8_t*)malloc(na);
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

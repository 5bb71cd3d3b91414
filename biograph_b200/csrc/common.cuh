// common.cuh -- shared device helpers for the bgx kernels (sm_100a).
//
// Data layout in HBM:
//   base words : uint64, 32 bases per word, first base in the HIGH bits (so integer order ==
//                lexicographic order A<C<G<T).  Every read starts on a word boundary; unused
//                trailing bits are zero.  This is the reference's dna_sequence byte order
//                (modules/bio_base/dna_sequence.h:95-99) read as big-endian 64-bit words.
//   k-mer      : uint64, first base in the high bits of the low 2k bits (modules/bio_base/kmer.h:30-38).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace bgx {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define BGX_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      throw ::bgx::Error(std::string(#expr) + " failed: " + cudaGetErrorString(e__));        \
  } while (0)

#define BGX_CHECK(cond, msg)                                 \
  do {                                                       \
    if (!(cond)) throw ::bgx::Error(std::string(msg));       \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// every kernel launch of this library goes through KLAUNCH so that bgx_launch_count() can report
// how many of OUR kernels ran (bench.py "gpu_launches")
extern unsigned long long g_launches;
inline void note_launch() { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); }
#define KLAUNCH(k) (::bgx::note_launch(), (k))

constexpr uint64_t kEmptyKey = ~0ULL;                 // kmer_count_table::k_unused_entry
constexpr uint64_t kKmerMask = (1ULL << 62) - 1;      // kmer_count_table::k_kmer_mask
constexpr uint64_t kFwdFlag = 1ULL << 63;             // k_fwd_flag  (fwd_starts_read)
constexpr uint64_t kRevFlag = 1ULL << 62;             // k_rev_flag  (rev_starts_read)

// suffix locator: base address in the corrected-base store << 16 | tag bits | length
//   bits 0..12   length (reads are at most 255 bases)
//   bit  13      kLocPopSeed: the suffix one base shorter was emitted as a seed too, so its pop_front is
//                covered by construction (it survives the dedup or is a prefix of what does) -- the
//                sharded closure walk does not have to ask its owner
//   bits 14..15  first base of the popped entry, on routed prev-bit queries only
constexpr int kLocLenBits = 16;
constexpr uint64_t kLocLenMask = (1u << 13) - 1;
constexpr uint64_t kLocPopSeed = 1u << 13;
__host__ __device__ __forceinline__ uint64_t make_loc(uint64_t addr, uint32_t len) { return (addr << kLocLenBits) | len; }
__host__ __device__ __forceinline__ uint64_t loc_addr(uint64_t loc) { return loc >> kLocLenBits; }
__host__ __device__ __forceinline__ uint32_t loc_len(uint64_t loc) { return (uint32_t)(loc & kLocLenMask); }

// mask keeping the top nb bases of a word (nb in [0,32])
__host__ __device__ __forceinline__ uint64_t top_bases_mask(int nb) {
  return nb >= 32 ? ~0ULL : ~(~0ULL >> (2 * nb));
}

__host__ __device__ __forceinline__ uint64_t kmer_low_mask(int k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1); }

#ifdef __CUDACC__

// 32 bases starting at base address a (the array must have one readable pad word at the end).
__device__ __forceinline__ uint64_t load_window(const uint64_t* __restrict__ w, uint64_t a) {
  uint64_t q = a >> 5;
  unsigned s = (unsigned)(a & 31) * 2;
  uint64_t hi = w[q];
  if (s == 0) return hi;
  uint64_t lo = w[q + 1];
  return (hi << s) | (lo >> (64 - s));
}

// first min(len,32) bases of the suffix at addr, zero padded: the radix key of a suffix
__device__ __forceinline__ uint64_t suffix_key(const uint64_t* __restrict__ w, uint64_t addr, int len) {
  return load_window(w, addr) & top_bases_mask(len);
}

// modules/bio_base/dna_sequence.cpp:330-351: reverse complement of a k-mer held in the low 2k bits
__device__ __forceinline__ uint64_t revcomp_kmer(uint64_t x, int k) {
  x = __brevll(~x);
  x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
  return x >> (64 - 2 * k);
}

// modules/bio_base/dna_sequence.cpp:378-385
__device__ __forceinline__ uint64_t canonicalize(uint64_t kmer, int k, bool& flipped) {
  uint64_t rc = revcomp_kmer(kmer, k);
  flipped = rc < kmer;
  return flipped ? rc : kmer;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// Bijective hash of a canonical k-mer on its own 2k-bit domain: fold the high half onto the low
// half (xor-shift by half the width is an involution on 2k bits), multiply by an odd constant mod
// 2^2k (invertible), fold again.  Every input bit reaches the low half through the first fold and
// from there all higher product bits, so the TOP bits -- the ones the counting passes partition by,
// and the home bucket of the solid set -- are well mixed (measured on E. coli k-mers: bin sizes are
// Poisson); the second fold mixes the low half.  Because it is a bijection, a table indexed by the
// top bits only has to remember the remaining bits, and the k-mer comes back with khash_inv().
__host__ __device__ __forceinline__ uint64_t khash(uint64_t x, int k) {
  x ^= x >> k;
  x = (x * 0xff51afd7ed558ccdULL) & kmer_low_mask(k);
  x ^= x >> k;
  return x;
}
__host__ __device__ __forceinline__ uint64_t khash_inv(uint64_t x, int k) {
  x ^= x >> k;
  x = (x * 0x4f74430c22a54005ULL) & kmer_low_mask(k);  // inverse of 0xff51afd7ed558ccd mod 2^64
  x ^= x >> k;
  return x;
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (blockDim.x <= 1024, a multiple of 32);
// returns the exclusive prefix, *total = block sum.  Ends with a barrier, so it can be called
// back to back.
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t tot;
  unsigned lane = lane_id(), warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  uint32_t inc = warp_incl_scan_u32(v);
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nw ? wsum[lane] : 0;
    uint32_t wi = warp_incl_scan_u32(w);
    wsum[lane] = wi - w;
    if (lane == 31) tot = wi;
  }
  __syncthreads();
  uint32_t r = wsum[warp] + inc - v;
  *total = tot;
  __syncthreads();
  return r;
}

// streaming 128-bit loads/stores that do not pollute L1
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---- solid k-mer set: open addressing over 32-byte buckets of four 8-byte slots (key | flags) ----
// One bucket = one DRAM sector = one 256-bit load.  A key lives in the first bucket, walking
// linearly from its home bucket, that had a free slot when it was inserted; buckets fill in slot
// order and nothing is ever deleted, so a lookup stops at the first bucket with an empty slot.
// pf: L2 prefetch-size hint of the load (PTX .L2::64B / ::128B / ::256B; 0 = none given).  A probe wants
// ONE 32-byte sector; what the memory system fetches per miss is measured per setting (profiles/r2m_*).
__device__ __forceinline__ void ld_bucket4(const unsigned long long* __restrict__ set, uint64_t bucket,
                                           unsigned long long (&k)[4], int pf = 0) {
  const unsigned long long* p = set + 4 * bucket;
  if (pf == 64)
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(k[0]), "=l"(k[1]), "=l"(k[2]), "=l"(k[3]) : "l"(p));
  else if (pf == 128)
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(k[0]), "=l"(k[1]), "=l"(k[2]), "=l"(k[3]) : "l"(p));
  else if (pf == 256)
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(k[0]), "=l"(k[1]), "=l"(k[2]), "=l"(k[3]) : "l"(p));
  else if (pf == 1)   // plain coherent load (no .nc)
    asm volatile("ld.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(k[0]), "=l"(k[1]), "=l"(k[2]), "=l"(k[3]) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(k[0]), "=l"(k[1]), "=l"(k[2]), "=l"(k[3]) : "l"(p));
}
// the stored word of `canon` if the bucket holds it, else kEmptyKey; *more = the bucket is full
// and does not hold it (the key may sit in the next bucket)
__device__ __forceinline__ unsigned long long bucket4_match(const unsigned long long (&k)[4], uint64_t canon, bool* more) {
  unsigned long long e = kEmptyKey;
  bool full = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (k[i] == kEmptyKey) full = false;
    else if ((k[i] & kKmerMask) == canon) e = k[i];
  }
  *more = full && e == kEmptyKey;
  return e;
}

#endif  // __CUDACC__

}  // namespace bgx

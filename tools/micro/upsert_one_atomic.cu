// upsert_one_atomic.cu -- micro-benchmark for DESIGN.md section 7 item 1: how much faster would the k-mer
// upsert be if an instance cost ONE atomic with return instead of a CAS (the read) + a RED (the count)?
//
// Model of one hash partition (what kmer_upsert_kernel sees while a table slice is L2 resident):
//   * a slice of 2^log2_slots slots, n instances, a fraction `solid` of them repeats of few keys
//     (each solid key ~70 times, as at 100x coverage), the rest distinct (sequencing errors)
//   * variant A (current): 16-byte slots {key, cnt}: CAS(key, EMPTY -> mine) then RED.add on cnt
//   * variant B (proposed): 8-byte slots [tag:35 | rev:14 | fwd:15]: ATOM.add(+1) with return; the
//     returned word carries the tag: match -> done; empty (tag 0) -> claim with a CAS on the tag
//     bits and keep the increment; other key -> undo with RED.add(-1) and probe the next slot
//   * variant C: B with 16-byte slots (tag word + spare), to separate "one atomic" from "half the
//     bytes per slot"
// Collisions, saturation and the spill table are left out on purpose (a few percent of the work);
// the claim race of B is handled correctly so the counts can be checked (sum of counters == n).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/upsert_one_atomic tools/micro/upsert_one_atomic.cu
// usage: upsert_one_atomic [log2_slots=20] [parts=64] [solid_permille=860]
// Every part is one slice of 2^log2_slots slots with 3 instances per slot (as the real table at 100x:
// ~0.42 distinct error keys and ~0.03 solid keys per slot, the solid ones ~86 times each); the
// parts run back to back, like the partitions of kmer_upsert_kernel.
// (First attempt, r1g: 2^26 instances into ONE 2^20-slot slice overfilled the table and never finished.)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// instance i -> key: solid instances cycle through n_solid keys, the others are unique
__device__ __forceinline__ uint64_t key_of(uint64_t i, uint64_t n_solid, unsigned solid_permille) {
  const uint64_t h = mix64(i * 0x9E3779B97F4A7C15ULL + 1);
  if ((h % 1000) < solid_permille) return 1 + (mix64(h) % n_solid);
  return (1ULL << 40) + i;
}

struct SlotA { unsigned long long key, cnt; };

__global__ void __launch_bounds__(256, 4) upsert_a(SlotA* __restrict__ t, uint64_t mask, uint64_t n, uint64_t n_solid,
                                                   unsigned solid_permille) {
  constexpr int ILP = 4;
  const uint64_t i0 = (uint64_t)blockIdx.x * (256 * ILP) + threadIdx.x;
  unsigned long long key[ILP], cur[ILP];
  uint64_t s[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) {
    const uint64_t i = i0 + j * 256;
    key[j] = i < n ? key_of(i, n_solid, solid_permille) : 0;
    s[j] = mix64(key[j]) & mask;
  }
#pragma unroll
  for (int j = 0; j < ILP; ++j) cur[j] = key[j] ? atomicCAS(&t[s[j]].key, 0ULL, key[j]) : 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) {
    if (!key[j]) continue;
    unsigned long long c = cur[j];
    uint64_t q = s[j];
    while (c != 0 && c != key[j]) {  // linear probing
      q = (q + 1) & mask;
      c = atomicCAS(&t[q].key, 0ULL, key[j]);
    }
    atomicAdd(reinterpret_cast<unsigned int*>(&t[q].cnt), 1u);
  }
}

constexpr int kCntBits = 29;  // 15 + 14
constexpr unsigned long long kCntMask = (1ULL << kCntBits) - 1;

template <int STRIDE>  // slot stride in 8-byte words: 1 (variant B) or 2 (variant C)
__global__ void __launch_bounds__(256, 4) upsert_b(unsigned long long* __restrict__ t, uint64_t mask, uint64_t n,
                                                   uint64_t n_solid, unsigned solid_permille) {
  constexpr int ILP = 4;
  const uint64_t i0 = (uint64_t)blockIdx.x * (256 * ILP) + threadIdx.x;
  unsigned long long tag[ILP], old[ILP];
  uint64_t s[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) {
    const uint64_t i = i0 + j * 256;
    const uint64_t k = i < n ? key_of(i, n_solid, solid_permille) : 0;
    const uint64_t h = mix64(k);
    s[j] = h & mask;
    tag[j] = k ? ((h >> 29) | 1ULL) << kCntBits : 0;  // 35 tag bits, never zero (stands in for the quotient)
  }
#pragma unroll
  for (int j = 0; j < ILP; ++j) old[j] = tag[j] ? atomicAdd(&t[s[j] * STRIDE], 1ULL) : 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) {
    if (!tag[j]) continue;
    uint64_t q = s[j];
    unsigned long long o = old[j];
    for (;;) {
      unsigned long long cur_tag = o & ~kCntMask;
      if (cur_tag == 0) {
        // empty: install the tag; the increment is already in.  Another key may race for the slot.
        unsigned long long seen = o + 1;  // at least our own increment is there
        for (;;) {
          seen = *reinterpret_cast<volatile unsigned long long*>(&t[q * STRIDE]);
          if (seen & ~kCntMask) break;
          if (atomicCAS(&t[q * STRIDE], seen, seen | tag[j]) == seen) { seen |= tag[j]; break; }
        }
        cur_tag = seen & ~kCntMask;
      }
      if (cur_tag == tag[j]) break;
      atomicAdd(&t[q * STRIDE], ~0ULL);  // someone else's slot: undo, move on
      q = (q + 1) & mask;
      o = atomicAdd(&t[q * STRIDE], 1ULL);
    }
  }
}

__global__ void sum_a(const SlotA* t, uint64_t slots, unsigned long long* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < slots && t[i].cnt) atomicAdd(out, t[i].cnt);
}
template <int STRIDE>
__global__ void sum_b(const unsigned long long* t, uint64_t slots, unsigned long long* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < slots && (t[i * STRIDE] & kCntMask)) atomicAdd(out, t[i * STRIDE] & kCntMask);
}

int main(int argc, char** argv) {
  const int log2_slots = argc > 1 ? atoi(argv[1]) : 20;
  const int parts = argc > 2 ? atoi(argv[2]) : 64;
  const unsigned solid_permille = argc > 3 ? (unsigned)atoi(argv[3]) : 860;
  const uint64_t slots = 1ull << log2_slots, mask = slots - 1;
  const uint64_t n = 3 * slots;                                   // instances per part
  const uint64_t n_err = n / 1000 * (1000 - solid_permille);      // distinct error keys per part
  const uint64_t n_solid = n / 1000 * solid_permille / 86;        // solid keys per part, ~86 instances each
  printf("%d parts x (2^%d slots, %llu instances: %u permille solid over %llu keys, ~%llu error keys)\n", parts, log2_slots,
         (unsigned long long)n, solid_permille, (unsigned long long)n_solid, (unsigned long long)n_err);
  fflush(stdout);
  char* buf;
  unsigned long long* out;
  cudaMalloc(&buf, (size_t)parts * slots * 16);
  cudaMalloc(&out, 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const unsigned grid = (unsigned)((n + 1023) / 1024);
  const char* names[] = {"A: CAS + RED, 16-byte slots (current)", "B: one ATOM.add with return, 8-byte slots",
                         "C: one ATOM.add with return, 16-byte stride"};
  for (int v = 0; v < 3; ++v) {
    float best = 1e30f;
    unsigned long long total = 0;
    const size_t slice_bytes = slots * (v == 1 ? 8 : 16);
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemset(buf, 0, (size_t)parts * slots * 16);
      cudaMemset(out, 0, 8);
      cudaEventRecord(a);
      for (int p = 0; p < parts; ++p) {
        char* slice = buf + (size_t)p * slice_bytes;
        if (v == 0) upsert_a<<<grid, 256>>>((SlotA*)slice, mask, n, n_solid, solid_permille);
        if (v == 1) upsert_b<1><<<grid, 256>>>((unsigned long long*)slice, mask, n, n_solid, solid_permille);
        if (v == 2) upsert_b<2><<<grid, 256>>>((unsigned long long*)slice, mask, n, n_solid, solid_permille);
      }
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      best = ms < best ? ms : best;
      const uint64_t all = (uint64_t)parts * slots;
      if (v == 0) sum_a<<<(unsigned)((all + 255) / 256), 256>>>((SlotA*)buf, all, out);
      if (v == 1) sum_b<1><<<(unsigned)((all + 255) / 256), 256>>>((unsigned long long*)buf, all, out);
      if (v == 2) sum_b<2><<<(unsigned)((all + 255) / 256), 256>>>((unsigned long long*)buf, all, out);
      cudaMemcpy(&total, out, 8, cudaMemcpyDeviceToHost);
    }
    const uint64_t tot_n = (uint64_t)parts * n;
    printf("%-48s %8.3f ms  %7.2f G instances/s   counted %llu of %llu%s\n", names[v], best, tot_n / best / 1e6, total,
           (unsigned long long)tot_n, total == tot_n ? "" : "  (MISMATCH)");
    fflush(stdout);
  }
  return 0;
}

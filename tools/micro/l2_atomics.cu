// l2_atomics.cu -- micro-benchmark: random accesses into an L2-resident slice of 16-byte slots (the
// k-mer upsert pattern): how many loads / REDs / CASes per second the chip sustains when every lane
// of a warp touches a different 32-byte sector.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_atomics tools/micro/l2_atomics.cu
// usage: l2_atomics [log2_slots=20] [n=2^28]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

struct Slot { unsigned long long key, cnt; };

// V: 0 = volatile load of key   1 = RED.add.u32 on cnt   2 = load + RED   3 = atomicAdd u64 with return (ATOMG)
//    4 = CAS on key (always fails: key != expected)   5 = load + RED + atomicOr(0) on key
//    6 = RED.add.u64 on cnt   7 = load 16 B (ld.v2.u64) + RED
template <int V, int ILP>
__global__ void __launch_bounds__(256, 4) bench(Slot* __restrict__ t, uint64_t mask, uint64_t n, uint64_t seed,
                                                unsigned long long* __restrict__ out) {
  uint64_t i0 = (uint64_t)blockIdx.x * (256 * ILP) + threadIdx.x;
  uint64_t s[ILP];
  unsigned long long v[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) s[j] = mix64(i0 + j * 256 + seed) & mask;
  unsigned long long acc = 0;
  if (V == 0 || V == 2 || V == 5) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = *reinterpret_cast<volatile unsigned long long*>(&t[s[j]].key);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
  }
  if (V == 7) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      unsigned long long a, b;
      asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(&t[s[j]]));
      v[j] = a + b;
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
  }
  if (V == 1 || V == 2 || V == 5 || V == 7) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) atomicAdd(reinterpret_cast<unsigned int*>(&t[s[j]].cnt) + (acc == 77 ? 1 : 0), 1u);
  }
  if (V == 6) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) atomicAdd(&t[s[j]].cnt, 1ULL);
  }
  if (V == 5) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) atomicOr(&t[s[j]].key, (unsigned long long)(acc == 77));
  }
  if (V == 3) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = atomicAdd(&t[s[j]].cnt, 1ULL);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
  }
  if (V == 4) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = atomicCAS(&t[s[j]].key, ~0ULL, 5ULL);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
  }
  if (V == 8) {  // load and RED independent: the RED does not wait for the load
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = *reinterpret_cast<volatile unsigned long long*>(&t[s[j]].key);
#pragma unroll
    for (int j = 0; j < ILP; ++j) atomicAdd(reinterpret_cast<unsigned int*>(&t[s[j]].cnt), 1u);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
  }
  if (V == 9) {  // relaxed.gpu load, then dependent RED
#pragma unroll
    for (int j = 0; j < ILP; ++j) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v[j]) : "l"(&t[s[j]].key));
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
#pragma unroll
    for (int j = 0; j < ILP; ++j) atomicAdd(reinterpret_cast<unsigned int*>(&t[s[j]].cnt) + (acc == 77 ? 1 : 0), 1u);
  }
  if (V == 10) {  // failing CAS as the load, then dependent RED
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = atomicCAS(&t[s[j]].key, ~0ULL, 5ULL);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
#pragma unroll
    for (int j = 0; j < ILP; ++j) atomicAdd(reinterpret_cast<unsigned int*>(&t[s[j]].cnt) + (acc == 77 ? 1 : 0), 1u);
  }
  if (V == 11) {  // load, then a dependent RED PER ITEM (each RED waits only for its own load)
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = *reinterpret_cast<volatile unsigned long long*>(&t[s[j]].key);
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      atomicAdd(reinterpret_cast<unsigned int*>(&t[s[j]].cnt) + (v[j] == 77 ? 1 : 0), 1u);
    }
  }
  if (V == 12) {  // load of the key, RED on a counter in a DIFFERENT array (separate sector)
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = *reinterpret_cast<volatile unsigned long long*>(&t[s[j]].key);
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
#pragma unroll
    for (int j = 0; j < ILP; ++j)
      atomicAdd(reinterpret_cast<unsigned int*>(&t[(s[j] + (mask >> 1) + 1) & mask].cnt) + (acc == 77 ? 1 : 0), 1u);
  }
  if (acc == 0x1234567) out[0] = acc;
}

__global__ void fill(Slot* t, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { t[i].key = i * 0x9E3779B97F4A7C15ULL | 1; t[i].cnt = 0; }
}

template <int V, int ILP = 4>
void run(const char* name, Slot* t, uint64_t mask, uint64_t n, unsigned long long* out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  unsigned grid = (unsigned)((n + 256 * ILP - 1) / (256 * ILP));
  bench<V, ILP><<<grid, 256>>>(t, mask, n, 1, out);
  cudaEventRecord(a);
  bench<V, ILP><<<grid, 256>>>(t, mask, n, 2, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("%-44s %8.3f ms  %7.2f G items/s\n", name, ms, n / ms / 1e6);
}

int main(int argc, char** argv) {
  int log2_slots = argc > 1 ? atoi(argv[1]) : 20;  // 2^20 * 16 B = 16 MiB: one partition's slice
  uint64_t n = argc > 2 ? strtoull(argv[2], 0, 10) : (1ull << 28);
  uint64_t slots = 1ull << log2_slots;
  Slot* t;
  unsigned long long* out;
  cudaMalloc(&t, slots * sizeof(Slot));
  cudaMalloc(&out, 8);
  fill<<<(unsigned)((slots + 255) / 256), 256>>>(t, slots);
  cudaDeviceSynchronize();
  printf("slice 2^%d slots x 16 B = %.1f MiB, %llu items\n", log2_slots, slots * 16.0 / (1 << 20), (unsigned long long)n);
  run<0>("volatile load key", t, slots - 1, n, out);
  run<1>("RED.add.u32 cnt", t, slots - 1, n, out);
  run<6>("RED.add.u64 cnt", t, slots - 1, n, out);
  run<2>("load key + RED.add.u32", t, slots - 1, n, out);
  run<7>("load 16 B + RED.add.u32", t, slots - 1, n, out);
  run<5>("load + RED.add.u32 + RED.or.u64", t, slots - 1, n, out);
  run<3>("ATOM.add.u64 with return", t, slots - 1, n, out);
  run<4>("CAS.u64 (fails)", t, slots - 1, n, out);
  run<8>("load + independent RED", t, slots - 1, n, out);
  run<9>("ld.relaxed.gpu + RED", t, slots - 1, n, out);
  run<10>("CAS (fails) + RED", t, slots - 1, n, out);
  run<11>("load + RED, per-item dependency", t, slots - 1, n, out);
  run<12>("load + RED on another sector", t, slots - 1, n, out);
  run<2, 1>("load key + RED.add.u32, ILP 1", t, slots - 1, n, out);
  run<2, 2>("load key + RED.add.u32, ILP 2", t, slots - 1, n, out);
  run<2, 8>("load key + RED.add.u32, ILP 8", t, slots - 1, n, out);
  run<11, 8>("load + RED per-item dep, ILP 8", t, slots - 1, n, out);
  run<0, 8>("volatile load key, ILP 8", t, slots - 1, n, out);
  return 0;
}

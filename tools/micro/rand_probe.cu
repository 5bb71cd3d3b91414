// rand_probe.cu -- micro-benchmark: random 8-byte probes into a table far larger than L2, with
// different load instructions and cudaLimitMaxL2FetchGranularity settings.  Prints the time per
// variant; run under `ncu --metrics dram__bytes_read.sum` to see the bytes DRAM really moved.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/rand_probe tools/micro/rand_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

template <int V>
__device__ __forceinline__ unsigned long long load8(const unsigned long long* p) {
  unsigned long long r;
  if (V == 0) r = __ldg(p);
  else if (V == 1) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(p));
  else if (V == 2) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
  else if (V == 3) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(p));
  else if (V == 4) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
  }
  else if (V == 5) asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(r) : "l"(p));
  else if (V == 6) asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(r) : "l"(p));
  else r = *p;
  return r;
}

template <int V, int ILP>
__global__ void __launch_bounds__(256) probe(const unsigned long long* __restrict__ t, uint64_t mask, uint64_t n,
                                             unsigned long long* __restrict__ out) {
  uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * ILP;
  unsigned long long acc = 0;
  unsigned long long v[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) v[j] = i + j < n ? load8<V>(t + (mix64(i + j + 12345) & mask)) : 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) acc += v[j];
  if (acc == 0x1234567) out[0] = acc;
}

__global__ void fill(unsigned long long* t, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = i * 0x9E3779B97F4A7C15ULL;
}

template <int V>
float run(const unsigned long long* t, uint64_t mask, uint64_t n, unsigned long long* out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  constexpr int ILP = 4;
  unsigned grid = (unsigned)((n / ILP + 255) / 256);
  probe<V, ILP><<<grid, 256>>>(t, mask, n, out);
  cudaEventRecord(a);
  probe<V, ILP><<<grid, 256>>>(t, mask, n, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main(int argc, char** argv) {
  int log2_slots = argc > 1 ? atoi(argv[1]) : 27;   // 2^27 * 8 B = 1 GiB
  uint64_t n = argc > 2 ? strtoull(argv[2], 0, 10) : (1ull << 29);
  int gran = argc > 3 ? atoi(argv[3]) : 0;
  if (gran) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("set L2 fetch granularity %d -> %s, now %zu\n", gran, cudaGetErrorString(e), got);
  }
  uint64_t slots = 1ull << log2_slots;
  unsigned long long *t, *out;
  cudaMalloc(&t, slots * 8);
  cudaMalloc(&out, 8);
  fill<<<(unsigned)((slots + 255) / 256), 256>>>(t, slots);
  cudaDeviceSynchronize();
  const char* names[] = {"__ldg", "ld.cg", "ld.nc.L1::no_allocate", "ld.volatile", "ld.nc.no_alloc.L2::cache_hint(evict_first)",
                         "ld.relaxed.gpu", "ld.cs", "plain"};
  float ms[8];
  ms[0] = run<0>(t, slots - 1, n, out); ms[1] = run<1>(t, slots - 1, n, out); ms[2] = run<2>(t, slots - 1, n, out);
  ms[3] = run<3>(t, slots - 1, n, out); ms[4] = run<4>(t, slots - 1, n, out); ms[5] = run<5>(t, slots - 1, n, out);
  ms[6] = run<6>(t, slots - 1, n, out); ms[7] = run<7>(t, slots - 1, n, out);
  for (int v = 0; v < 8; ++v)
    printf("table 2^%d x 8 B, %llu probes, %-32s %8.3f ms  %7.2f Gprobes/s\n", log2_slots, (unsigned long long)n, names[v], ms[v],
           n / ms[v] / 1e6);
  return 0;
}

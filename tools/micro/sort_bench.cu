// sort_bench.cu -- times radix_sort_pairs (prims.cu) on random (key, value) pairs and checks the result
// (sorted on the chosen bits, stable, permutation checksum).  Iteration tool for the onesweep kernel.
// build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr \
//             -I biograph_b200/csrc -o tools/micro/sort_bench tools/micro/sort_bench.cu
// usage: sort_bench [n=15000000] [bits=48] [reps=5] [distinct_log2=0 (0 = full random)]
#include <cstdio>
#include <cstdlib>

#include "../../biograph_b200/csrc/prims.cu"

using namespace bgx;

__global__ void gen_kernel(uint64_t* k, uint64_t* v, uint64_t n, uint64_t seed, int distinct_log2) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t x = mix64(i * 0x9E3779B97F4A7C15ULL + seed);
  if (distinct_log2) x = mix64(x & ((1ULL << distinct_log2) - 1)) ;
  k[i] = x;
  v[i] = i;
}

// bad[0] = order violations, bad[1] = stability violations; sum[0] = sum of values, sum[1] = xor of keys
__global__ void check_kernel(const uint64_t* k, const uint64_t* v, uint64_t n, int begin_bit, unsigned long long* bad,
                             unsigned long long* sum) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(&sum[0], (unsigned long long)v[i]);
  atomicXor(&sum[1], (unsigned long long)k[i]);
  if (i + 1 < n) {
    uint64_t a = k[i] >> begin_bit, b = k[i + 1] >> begin_bit;
    if (a > b) atomicAdd(&bad[0], 1ULL);
    if (a == b && v[i] > v[i + 1]) atomicAdd(&bad[1], 1ULL);
  }
}

int main(int argc, char** argv) {
  uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 15000000ULL;
  int bits = argc > 2 ? atoi(argv[2]) : 48;
  int reps = argc > 3 ? atoi(argv[3]) : 5;
  int distinct_log2 = argc > 4 ? atoi(argv[4]) : 0;
  cudaStream_t s;
  cudaStreamCreate(&s);
  DevBuf<uint64_t> k0(n, s), v0(n, s), k1(n, s), v1(n, s);
  DevBuf<unsigned long long> chk(4, s);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f, tot = 0;
  int passes = 0;
  bool ok = true;
  for (int r = 0; r < reps + 1; ++r) {
    gen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(k0.p, v0.p, n, 12345 + r, distinct_log2);
    cudaEventRecord(a, s);
    bool alt = radix_sort_pairs(k0.p, v0.p, k1.p, v1.p, n, 64 - bits, 64, s, &passes);
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (r) { best = ms < best ? ms : best; tot += ms; }
    cudaMemsetAsync(chk.p, 0, 32, s);
    check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(alt ? k1.p : k0.p, alt ? v1.p : v0.p, n, 64 - bits, chk.p, chk.p + 2);
    unsigned long long h[4];
    cudaMemcpyAsync(h, chk.p, 32, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaError_t e = cudaGetLastError();
    unsigned long long want = (unsigned long long)(n * (n - 1) / 2);
    if (e != cudaSuccess || h[0] || h[1] || h[2] != want) {
      ok = false;
      printf("rep %d: FAILED err=%s order=%llu stability=%llu sum=%llu want=%llu\n", r, cudaGetErrorString(e), h[0], h[1], h[2], want);
    }
  }
  double bytes = (double)passes * 32.0 * (double)n;
  printf("n=%llu bits=%d passes=%d distinct_log2=%d : best %.3f ms mean %.3f ms  -> %.1f GB/s algorithmic (best), %.3f ms/pass  %s\n",
         (unsigned long long)n, bits, passes, distinct_log2, best, tot / reps, bytes / best / 1e6, best / passes, ok ? "OK" : "FAILED");
  return ok ? 0 : 1;
}

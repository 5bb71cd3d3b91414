// prims.cu -- exclusive scan and LSD radix sort kernels (sm_100a, hand-written; no CUB).
#include "prims.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <unordered_map>

namespace bgx {

unsigned long long g_launches = 0;

// ---- device memory arena ---------------------------------------------------------------------
// A size-keyed cache of cudaMalloc blocks.  Every buffer of the path is allocated and released
// in the same order with the same sizes on every run over the same reads, so after the first
// run every request is served from the cache with no driver call (cudaMallocAsync's pool was
// measured to stall for tens of ms when multi-GB blocks were recycled in a different order).
// Reuse is safe because all work of a context is ordered on one stream.
namespace {
struct Block {
  void* p;
  cudaStream_t s;
};
struct Arena {
  std::mutex mu;
  std::multimap<size_t, Block> free_blocks;                       // size -> block
  std::unordered_map<void*, std::pair<size_t, cudaStream_t>> live;  // block -> (size, stream)
  size_t live_bytes = 0, peak_bytes = 0;
  // frees every cached block (of one stream, or of all when s == nullptr)
  void trim_locked(cudaStream_t s, bool all) {
    for (auto it = free_blocks.begin(); it != free_blocks.end();) {
      if (all || it->second.s == s) {
        cudaFree(it->second.p);
        it = free_blocks.erase(it);
      } else {
        ++it;
      }
    }
  }
};
Arena& arena() {
  static Arena a;
  return a;
}
}  // namespace

void* dev_alloc(size_t bytes, cudaStream_t s) {
  Arena& a = arena();
  bytes = (bytes + 511) & ~(size_t)511;
  std::lock_guard<std::mutex> lk(a.mu);
  void* p = nullptr;
  size_t sz = bytes;
  // exact-size match among this stream's cached blocks: sizes repeat run after run, and a
  // near-fit policy lets a smaller request steal the block a later exact request needs
  for (auto it = a.free_blocks.lower_bound(bytes); it != a.free_blocks.end() && it->first == bytes; ++it) {
    if (it->second.s != s) continue;
    p = it->second.p;
    sz = it->first;
    a.free_blocks.erase(it);
    break;
  }
  if (!p) {
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cudaDeviceSynchronize();
      a.trim_locked(nullptr, true);  // give cached blocks back and retry once
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      throw Error("out of device memory allocating " + std::to_string(bytes) + " bytes (" +
                  std::to_string(a.live_bytes) + " live)");
    }
  }
  a.live[p] = {sz, s};
  a.live_bytes += sz;
  a.peak_bytes = std::max(a.peak_bytes, a.live_bytes);
  return p;
}

void dev_free(void* p, cudaStream_t) {
  if (!p) return;
  Arena& a = arena();
  std::lock_guard<std::mutex> lk(a.mu);
  auto it = a.live.find(p);
  if (it == a.live.end()) return;
  a.free_blocks.emplace(it->second.first, Block{p, it->second.second});
  a.live_bytes -= it->second.first;
  a.live.erase(it);
}

void dev_trim(cudaStream_t s) {
  Arena& a = arena();
  cudaStreamSynchronize(s);
  std::lock_guard<std::mutex> lk(a.mu);
  a.trim_locked(s, false);
}

size_t dev_peak_bytes(bool reset) {
  Arena& a = arena();
  std::lock_guard<std::mutex> lk(a.mu);
  size_t v = a.peak_bytes;
  if (reset) a.peak_bytes = a.live_bytes;
  return v;
}

namespace {

// ------------------------------------------------------------------------------------------
// exclusive scan (uint32): reduce per block -> scan block sums -> scan per block with carry-in
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in,
                                                                 uint32_t* __restrict__ block_sums, size_t n) {
  size_t base = (size_t)blockIdx.x * SC_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    size_t idx = base + (size_t)i * SC_THREADS + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  uint32_t tot;
  block_excl_scan_u32(s, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_blocksums_kernel(uint32_t* __restrict__ block_sums, size_t nb,
                                                             uint32_t* __restrict__ total_out) {
  uint32_t carry = 0;
  for (size_t base = 0; base < nb; base += 1024) {
    size_t idx = base + threadIdx.x;
    uint32_t v = idx < nb ? block_sums[idx] : 0;
    uint32_t tot;
    uint32_t ex = block_excl_scan_u32(v, &tot);
    if (idx < nb) block_sums[idx] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SC_THREADS) scan_final_kernel(const uint32_t* __restrict__ in,
                                                                uint32_t* __restrict__ out,
                                                                const uint32_t* __restrict__ block_sums, size_t n) {
  size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_ITEMS;
  uint32_t v[SC_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    size_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan_u32(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    size_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
}

// ------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits.  Per pass:
//   radix_count_kernel  : per-tile digit histogram (keys only, 8 B/elem)
//   exclusive scan of the digit-major (digit, tile) count matrix
//   radix_scatter_kernel: stable in-tile ranking (warp match_any), tile staged in shared
//                         memory in digit order, written out as coalesced per-digit runs.
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 elements
constexpr int RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) radix_count_kernel(const uint64_t* __restrict__ keys,
                                                                 uint32_t* __restrict__ counts, uint32_t n,
                                                                 int shift, uint32_t ntiles) {
  __shared__ uint32_t hist[RS_WARPS][RADIX];
  for (int i = threadIdx.x; i < RS_WARPS * RADIX; i += RS_THREADS) (&hist[0][0])[i] = 0;
  __syncthreads();
  unsigned warp = threadIdx.x >> 5;
  uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int i = 0; i < RS_ITEMS; ++i) {
    uint32_t idx = base + i * RS_THREADS + threadIdx.x;
    if (idx < n) {
      unsigned d = (unsigned)(keys[idx] >> shift) & (RADIX - 1);
      atomicAdd(&hist[warp][d], 1u);
    }
  }
  __syncthreads();
  unsigned d = threadIdx.x;
  uint32_t c = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) c += hist[w][d];
  counts[(size_t)d * ntiles + blockIdx.x] = c;
}

__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint64_t* __restrict__ vals_in,
                     uint64_t* __restrict__ keys_out, uint64_t* __restrict__ vals_out,
                     const uint32_t* __restrict__ offsets, uint32_t n, int shift, uint32_t ntiles) {
  __shared__ uint32_t warp_cnt[RS_WARPS][RADIX];
  __shared__ uint32_t digit_start[RADIX];
  __shared__ uint32_t glob_base[RADIX];
  __shared__ uint8_t stage_digit[RS_TILE];
  extern __shared__ uint64_t stage[];  // RS_TILE words

  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile_base = blockIdx.x * RS_TILE;
  const uint32_t nvalid = min((uint32_t)RS_TILE, n - tile_base);
  const uint32_t warp_base = warp * (RS_ITEMS * 32);

  for (int i = tid; i < RS_WARPS * RADIX; i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;

  uint64_t k[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    uint32_t e = warp_base + i * 32 + lane;
    k[i] = e < nvalid ? keys_in[tile_base + e] : ~0ULL;  // pads sort to the very end of the tile
  }
  __syncthreads();

  uint16_t rank[RS_ITEMS];
  const unsigned lt_mask = (1u << lane) - 1;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    unsigned d = (unsigned)(k[i] >> shift) & (RADIX - 1);
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if ((int)lane == leader) {
      old = warp_cnt[warp][d];
      warp_cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[i] = (uint16_t)(old + __popc(peers & lt_mask));
    __syncwarp();
  }
  __syncthreads();

  {
    unsigned d = tid;  // RS_THREADS == RADIX
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c = warp_cnt[w][d];
      warp_cnt[w][d] = sum;
      sum += c;
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan_u32(sum, &tot);
    digit_start[d] = ex;
    glob_base[d] = offsets[(size_t)d * ntiles + blockIdx.x] - ex;
  }
  __syncthreads();

  uint16_t pos[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    unsigned d = (unsigned)(k[i] >> shift) & (RADIX - 1);
    pos[i] = (uint16_t)(digit_start[d] + warp_cnt[warp][d] + rank[i]);
    stage[pos[i]] = k[i];
    stage_digit[pos[i]] = (uint8_t)d;
  }
  __syncthreads();
  for (uint32_t j = tid; j < nvalid; j += RS_THREADS) {
    keys_out[glob_base[stage_digit[j]] + j] = stage[j];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    uint32_t e = warp_base + i * 32 + lane;
    if (e < nvalid) stage[pos[i]] = vals_in[tile_base + e];
  }
  __syncthreads();
  for (uint32_t j = tid; j < nvalid; j += RS_THREADS) {
    vals_out[glob_base[stage_digit[j]] + j] = stage[j];
  }
}

}  // namespace

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total_out, cudaStream_t s) {
  if (n == 0) {
    if (total_out) BGX_CUDA(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), s));
    return;
  }
  size_t nb = (n + SC_TILE - 1) / SC_TILE;
  DevBuf<uint32_t> block_sums(nb, s);
  KLAUNCH(scan_reduce_kernel)<<<(unsigned)nb, SC_THREADS, 0, s>>>(in, block_sums.p, n);
  KLAUNCH(scan_blocksums_kernel)<<<1, 1024, 0, s>>>(block_sums.p, nb, total_out);
  KLAUNCH(scan_final_kernel)<<<(unsigned)nb, SC_THREADS, 0, s>>>(in, out, block_sums.p, n);
  BGX_CUDA(cudaGetLastError());
}

bool radix_sort_pairs(uint64_t* keys, uint64_t* vals, uint64_t* keys_alt, uint64_t* vals_alt, size_t n,
                      int begin_bit, int end_bit, cudaStream_t s, int* passes_out) {
  BGX_CHECK(n < (1ull << 32), "radix_sort_pairs: n must be < 2^32");
  BGX_CHECK(begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit, "radix_sort_pairs: bad bit range");
  int passes = 0;
  if (n == 0) { if (passes_out) *passes_out = 0; return false; }
  uint32_t ntiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
  DevBuf<uint32_t> counts((size_t)RADIX * ntiles, s);
  static bool attr_set = false;
  if (!attr_set) {
    BGX_CUDA(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  RS_TILE * (int)sizeof(uint64_t)));
    attr_set = true;
  }
  bool in_alt = false;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    const uint64_t* ki = in_alt ? keys_alt : keys;
    const uint64_t* vi = in_alt ? vals_alt : vals;
    uint64_t* ko = in_alt ? keys : keys_alt;
    uint64_t* vo = in_alt ? vals : vals_alt;
    KLAUNCH(radix_count_kernel)<<<ntiles, RS_THREADS, 0, s>>>(ki, counts.p, (uint32_t)n, shift, ntiles);
    exclusive_scan_u32(counts.p, counts.p, (size_t)RADIX * ntiles, nullptr, s);
    KLAUNCH(radix_scatter_kernel)<<<ntiles, RS_THREADS, RS_TILE * sizeof(uint64_t), s>>>(ki, vi, ko, vo, counts.p,
                                                                               (uint32_t)n, shift, ntiles);
    BGX_CUDA(cudaGetLastError());
    in_alt = !in_alt;
    ++passes;
  }
  if (passes_out) *passes_out = passes;
  return in_alt;
}

}  // namespace bgx

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
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      throw Error("out of device memory allocating " + std::to_string(bytes) + " bytes (" +
                  std::to_string(a.live_bytes) + " live in this library, " + std::to_string(free_b) + " of " +
                  std::to_string(total_b) + " free on the device)");
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

// ---- pinned host arena ------------------------------------------------------------------------------
namespace {
struct HostArena {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  std::unordered_map<void*, size_t> live;
  size_t cached_bytes = 0;
};
HostArena& host_arena() {
  static HostArena a;
  return a;
}
constexpr size_t kHostCacheLimit = size_t(8) << 30;  // keep at most 8 GiB of idle pinned memory
}  // namespace

void* host_alloc(size_t bytes) {
  HostArena& a = host_arena();
  bytes = (std::max<size_t>(bytes, 1) + 4095) & ~(size_t)4095;
  std::lock_guard<std::mutex> lk(a.mu);
  auto it = a.free_blocks.find(bytes);
  void* p = nullptr;
  if (it != a.free_blocks.end()) {
    p = it->second;
    a.cached_bytes -= bytes;
    a.free_blocks.erase(it);
  } else {
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      for (auto& kv : a.free_blocks) cudaFreeHost(kv.second);
      a.free_blocks.clear();
      a.cached_bytes = 0;
      if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        throw Error("out of pinned host memory allocating " + std::to_string(bytes) + " bytes");
      }
    }
  }
  a.live[p] = bytes;
  return p;
}

void host_free(void* p) {
  if (!p) return;
  HostArena& a = host_arena();
  std::lock_guard<std::mutex> lk(a.mu);
  auto it = a.live.find(p);
  if (it == a.live.end()) {  // not ours (defensive): plain heap memory
    free(p);
    return;
  }
  size_t bytes = it->second;
  a.live.erase(it);
  if (a.cached_bytes + bytes > kHostCacheLimit) {
    cudaFreeHost(p);
  } else {
    a.free_blocks.emplace(bytes, p);
    a.cached_bytes += bytes;
  }
}

void host_trim() {
  HostArena& a = host_arena();
  std::lock_guard<std::mutex> lk(a.mu);
  for (auto& kv : a.free_blocks) cudaFreeHost(kv.second);
  a.free_blocks.clear();
  a.cached_bytes = 0;
}

size_t dev_live_bytes() {
  Arena& a = arena();
  std::lock_guard<std::mutex> lk(a.mu);
  return a.live_bytes;
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
// LSD radix sort of (uint64 key, uint64 value) pairs, 8-bit digits, "one sweep" per digit:
//   radix_hist_kernel   : ONE read of the keys builds the global digit histograms of every pass
//   radix_bases_kernel  : exclusive scan of each pass's 256 bins -> first output index per digit
//   onesweep_kernel     : per pass, per tile of 4096 records (256 threads x 16 records, 3 CTAs per SM):
//       - the tile's keys and values are pulled into shared memory by TMA bulk copies
//         (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), no registers held
//       - stable in-tile ranking: warp-striped items, match_any per digit, per-warp counters bumped
//         with shared-memory atomics by the group leaders; the 16 items of a thread are independent
//         instruction streams (loads, matches and atomics pipeline instead of forming one chain)
//       - thread d owns digit d: it publishes the tile's count, and fetches the exclusive prefix
//         over earlier tiles by decoupled look-back, several predecessors per round trip; the loads
//         are issued before the in-tile reorder and consumed after it
//       - records are reordered in shared memory and leave as coalesced per-digit runs, so a pass
//         reads and writes every record exactly once (2*N*16 bytes per pass)
// Tiles are handed out by an atomic ticket, so every tile a block can wait on has started.
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 records
constexpr int RADIX = 256;
constexpr int RS_LOOKBACK = 4;           // predecessor tiles fetched per look-back round trip
// tile status words are 64-bit (2 flag bits + count) so that a sort may hold up to 2^32 records
constexpr uint64_t ST_AGG = 1ull << 62;    // tile aggregate available
constexpr uint64_t ST_INCL = 1ull << 63;   // inclusive prefix available
constexpr uint64_t ST_VAL = (1ull << 62) - 1;
static_assert(RS_THREADS == RADIX, "one thread per digit");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared; completion is signalled on the mbarrier as transferred bytes.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t ld_status(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int RH_THREADS = 256;
constexpr int RH_ITEMS = 16;
constexpr int RH_MAXPASS = 8;
__global__ void __launch_bounds__(RH_THREADS) radix_hist_kernel(const uint64_t* __restrict__ keys, uint32_t n,
                                                                int begin_bit, int passes,
                                                                uint32_t* __restrict__ ghist /*[passes][256]*/) {
  __shared__ uint32_t h[RH_MAXPASS][RADIX];
  for (int i = threadIdx.x; i < passes * RADIX; i += RH_THREADS) (&h[0][0])[i] = 0;
  __syncthreads();
  for (uint32_t base = blockIdx.x * (RH_THREADS * RH_ITEMS); base < n; base += gridDim.x * (RH_THREADS * RH_ITEMS)) {
#pragma unroll 4
    for (int i = 0; i < RH_ITEMS; ++i) {
      uint32_t idx = base + i * RH_THREADS + threadIdx.x;
      if (idx < n) {
        uint64_t k = keys[idx] >> begin_bit;
        for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(unsigned)(k >> (8 * p)) & (RADIX - 1)], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * RADIX; i += RH_THREADS) {
    uint32_t v = (&h[0][0])[i];
    if (v) atomicAdd(&ghist[i], v);
  }
}

// one block of 256 threads per pass: ghist[p][d] -> exclusive prefix over d
__global__ void __launch_bounds__(RADIX) radix_bases_kernel(uint32_t* __restrict__ ghist) {
  uint32_t* g = ghist + blockIdx.x * RADIX;
  uint32_t v = g[threadIdx.x], tot;
  uint32_t ex = block_excl_scan_u32(v, &tot);
  g[threadIdx.x] = ex;
}

struct OnesweepSmem {
  uint64_t keys[RS_TILE];    // TMA destination; then the reorder stage of the keys, then of the values
  uint64_t vals[RS_TILE];    // TMA destination
  uint32_t warp_cnt[RS_WARPS][RADIX];
  uint32_t digit_start[RADIX];
  uint32_t glob_base[RADIX];
  uint32_t warp_part[RS_WARPS];
  uint64_t bar_keys, bar_vals;
  uint32_t tile;
};

__global__ void __launch_bounds__(RS_THREADS, 3)
onesweep_kernel(const uint64_t* __restrict__ keys_in, const uint64_t* __restrict__ vals_in,
                uint64_t* __restrict__ keys_out, uint64_t* __restrict__ vals_out,
                const uint32_t* __restrict__ digit_base /*[256] exclusive*/, uint64_t* __restrict__ status /*[ntiles][256]*/,
                uint32_t* __restrict__ ticket, uint32_t n, int shift) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  OnesweepSmem& sm = *reinterpret_cast<OnesweepSmem*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    sm.tile = atomicAdd(ticket, 1u);
    mbar_init(&sm.bar_keys, 1);
    mbar_init(&sm.bar_vals, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    uint4* z = reinterpret_cast<uint4*>(&sm.warp_cnt[0][0]);
#pragma unroll
    for (int i = 0; i < RS_WARPS * RADIX / 4 / RS_THREADS; ++i) z[tid + i * RS_THREADS] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const uint32_t tile = sm.tile;
  const uint32_t tile_base = tile * RS_TILE;
  const uint32_t nvalid = min((uint32_t)RS_TILE, n - tile_base);
  if (tid == 0) {
    const uint32_t bytes = ((nvalid * 8u) + 15u) & ~15u;  // buffers are padded to 512 B by the arena
    mbar_expect_tx(&sm.bar_keys, bytes);
    tma_load_1d(sm.keys, keys_in + tile_base, bytes, &sm.bar_keys);
    mbar_expect_tx(&sm.bar_vals, bytes);
    tma_load_1d(sm.vals, vals_in + tile_base, bytes, &sm.bar_vals);
  }
  mbar_wait(&sm.bar_keys, 0);

  // ---- stable ranking: item (i, lane) of warp w is tile element w*512 + i*32 + lane ------------
  const uint32_t warp_base = warp * (RS_ITEMS * 32);
  const unsigned lt_mask = (1u << lane) - 1;
  uint64_t k[RS_ITEMS];
  uint32_t packed_rank[RS_ITEMS / 2];  // two 16-bit ranks per register; later two 16-bit stage positions
  uint32_t digs[RS_ITEMS / 4];         // four 8-bit digits per register
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) k[i] = sm.keys[warp_base + i * 32 + lane];
  uint32_t* my_cnt = sm.warp_cnt[warp];
#pragma unroll
  for (int h = 0; h < RS_ITEMS; h += 8) {
    unsigned peers[8];
    uint32_t old[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = h + q;
      const uint32_t e = warp_base + i * 32 + lane;
      // pads (beyond nvalid) rank as digit 255 after every real record of the tile
      const unsigned d = e < nvalid ? (unsigned)(k[i] >> shift) & (RADIX - 1) : (RADIX - 1);
      if (i & 3) digs[i >> 2] |= d << (8 * (i & 3)); else digs[i >> 2] = d;
      peers[q] = __match_any_sync(0xffffffffu, d);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = h + q;
      const unsigned d = (digs[i >> 2] >> (8 * (i & 3))) & 0xffu;
      old[q] = 0;
      // the lowest lane of each digit group bumps the warp's counter; groups of one instruction
      // have different digits, successive instructions are ordered by the warp barrier
      if ((peers[q] & lt_mask) == 0) old[q] = atomicAdd(&my_cnt[d], (uint32_t)__popc(peers[q]));
      __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = h + q;
      const uint32_t r = __shfl_sync(0xffffffffu, old[q], __ffs(peers[q]) - 1) + __popc(peers[q] & lt_mask);
      if (i & 1) packed_rank[i >> 1] |= r << 16; else packed_rank[i >> 1] = r;
    }
  }
  __syncthreads();

  // ---- thread d owns digit d: warp offsets, tile count, look-back loads, in-tile digit starts -------
  const unsigned d_own = tid;
  uint32_t sum = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    const uint32_t c = sm.warp_cnt[w][d_own];
    sm.warp_cnt[w][d_own] = sum;
    sum += c;
  }
  // pads were counted as digit 255: remove them from the published count
  const uint32_t real = (d_own == RADIX - 1) ? sum - (RS_TILE - nvalid) : sum;
  uint64_t* const st_mine = status + (size_t)tile * RADIX + d_own;
  st_status(st_mine, (tile == 0 ? ST_INCL : ST_AGG) | (uint64_t)real);
  uint64_t lb[RS_LOOKBACK];
#pragma unroll
  for (int q = 0; q < RS_LOOKBACK; ++q)
    lb[q] = tile > (uint32_t)q ? ld_status(status + (size_t)(tile - 1 - q) * RADIX + d_own) : ST_INCL;
  uint32_t ex;
  {
    const uint32_t inc = warp_incl_scan_u32(sum);
    if (lane == 31) sm.warp_part[warp] = inc;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) before += w < (int)warp ? sm.warp_part[w] : 0u;
    ex = before + inc - sum;
    sm.digit_start[d_own] = ex;
  }
  __syncthreads();

  // ---- keys: registers -> shared (digit order) ---------------------------------------------------------
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const unsigned d = (digs[i >> 2] >> (8 * (i & 3))) & 0xffu;
    const uint32_t r = (packed_rank[i >> 1] >> (16 * (i & 1))) & 0xffffu;
    const uint32_t pos = sm.digit_start[d] + my_cnt[d] + r;
    sm.keys[pos] = k[i];
    if (i & 1) packed_rank[i >> 1] = (packed_rank[i >> 1] & 0xffffu) | (pos << 16);
    else packed_rank[i >> 1] = (packed_rank[i >> 1] & 0xffff0000u) | pos;
  }
  // ---- finish the look-back: exclusive prefix of this digit over the earlier tiles -------------------
  {
    uint32_t excl = 0;
    if (tile != 0) {
      int64_t j = (int64_t)tile - 1;  // tile whose status lb[0] holds
      for (;;) {
        bool done = false;
#pragma unroll
        for (int q = 0; q < RS_LOOKBACK; ++q) {
          if (!done) {
            uint64_t v = lb[q];
            if (j - q >= 0) {
              while (v == 0) v = ld_status(status + (size_t)(j - q) * RADIX + d_own);
              excl += (uint32_t)(v & ST_VAL);
            }
            done = (v & ST_INCL) != 0;
          }
        }
        if (done) break;
        j -= RS_LOOKBACK;
#pragma unroll
        for (int q = 0; q < RS_LOOKBACK; ++q)
          lb[q] = j - q >= 0 ? ld_status(status + (size_t)(j - q) * RADIX + d_own) : ST_INCL;
      }
      st_status(st_mine, ST_INCL | (uint64_t)(excl + real));
    }
    sm.glob_base[d_own] = digit_base[d_own] + excl - ex;
  }
  __syncthreads();

  // ---- keys: shared -> global, coalesced per-digit runs ------------------------------------------------
  // element j of the reordered tile names its own digit; the thread keeps it for the value it writes next
  uint32_t out_digs[RS_ITEMS / 4];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const uint32_t j = tid + i * RS_THREADS;
    unsigned d = 0;
    if (j < nvalid) {
      const uint64_t key = sm.keys[j];
      d = (unsigned)(key >> shift) & (RADIX - 1);
      keys_out[sm.glob_base[d] + j] = key;
    }
    if (i & 3) out_digs[i >> 2] |= d << (8 * (i & 3)); else out_digs[i >> 2] = d;
  }
  // ---- values: same permutation, staged through the (now free) key buffer ------------------------------
  mbar_wait(&sm.bar_vals, 0);
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) k[i] = sm.vals[warp_base + i * 32 + lane];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) sm.keys[(packed_rank[i >> 1] >> (16 * (i & 1))) & 0xffffu] = k[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const uint32_t j = tid + i * RS_THREADS;
    if (j < nvalid) vals_out[sm.glob_base[(out_digs[i >> 2] >> (8 * (i & 3))) & 0xffu] + j] = sm.keys[j];
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
  BGX_CHECK(n < kMaxSortRecords, "radix_sort_pairs: too many records for one sort");
  BGX_CHECK(begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit, "radix_sort_pairs: bad bit range");
  const int passes = (end_bit - begin_bit + 7) / 8;
  if (passes_out) *passes_out = n ? passes : 0;
  if (n == 0 || passes == 0) return false;
  const uint32_t ntiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
  // zeroed scratch: [passes][256] histograms | [passes] tickets, and [passes][ntiles][256] 64-bit status words
  const size_t hist_words = (size_t)passes * RADIX, status_words = (size_t)ntiles * RADIX;
  const size_t ticket_off = hist_words;
  DevBuf<uint32_t> scratch(hist_words + passes, s);
  DevBuf<uint64_t> status(status_words * passes, s);
  BGX_CUDA(cudaMemsetAsync(scratch.p, 0, scratch.n * sizeof(uint32_t), s));
  BGX_CUDA(cudaMemsetAsync(status.p, 0, status.n * sizeof(uint64_t), s));
  // the opt-in is per device: set it on every call (cheap) rather than once per process
  BGX_CUDA(cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OnesweepSmem)));
  const unsigned hist_grid = (unsigned)std::min<size_t>((n + RH_THREADS * RH_ITEMS - 1) / (RH_THREADS * RH_ITEMS), (size_t)kNumSMs * 8);
  KLAUNCH(radix_hist_kernel)<<<hist_grid, RH_THREADS, 0, s>>>(keys, (uint32_t)n, begin_bit, passes, scratch.p);
  KLAUNCH(radix_bases_kernel)<<<passes, RADIX, 0, s>>>(scratch.p);
  bool in_alt = false;
  for (int p = 0; p < passes; ++p) {
    const uint64_t* ki = in_alt ? keys_alt : keys;
    const uint64_t* vi = in_alt ? vals_alt : vals;
    uint64_t* ko = in_alt ? keys : keys_alt;
    uint64_t* vo = in_alt ? vals : vals_alt;
    KLAUNCH(onesweep_kernel)<<<ntiles, RS_THREADS, sizeof(OnesweepSmem), s>>>(
        ki, vi, ko, vo, scratch.p + (size_t)p * RADIX, status.p + status_words * p,
        scratch.p + ticket_off + p, (uint32_t)n, begin_bit + 8 * p);
    in_alt = !in_alt;
  }
  BGX_CUDA(cudaGetLastError());
  return in_alt;
}

}  // namespace bgx

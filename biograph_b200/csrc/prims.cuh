// prims.cuh -- device-wide primitives written for this path (no CUB/Thrust): exclusive scan
// and the LSD radix sort of (uint64 key, uint64 value) pairs used for suffix records and k-mers.
#pragma once

#include <cstddef>
#include <cstdint>

#include "common.cuh"

namespace bgx {

// Device allocation from the library's caching arena (prims.cu).  Blocks are recycled by size
// among allocations of the same stream (one stream per context), so reuse is stream-ordered.
void* dev_alloc(size_t bytes, cudaStream_t s);
void dev_free(void* p, cudaStream_t s);
void dev_trim(cudaStream_t s);         // return the cached blocks of stream s to the driver
size_t dev_peak_bytes(bool reset);     // high-water mark of live bytes
size_t dev_live_bytes();               // bytes currently handed out

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(size_t n_, cudaStream_t s_) { alloc(n_, s_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n_, cudaStream_t s_) {
    release();
    s = s_; n = n_;
    p = static_cast<T*>(dev_alloc((n_ ? n_ : 1) * sizeof(T), s_));
  }
  void release() {
    if (p) dev_free(p, s);
    p = nullptr; n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// Host buffers handed to the caller (every bgx_export_* output): page-locked, so device-to-host
// copies run at PCIe speed, and recycled by size like the device blocks.  bgx_free -> host_free.
void* host_alloc(size_t bytes);
void host_free(void* p);
void host_trim();

// out[i] = sum(in[0..i)), in place allowed.  If total_out != nullptr the grand total is written
// there (device pointer).
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total_out, cudaStream_t s);

// most records one sort / one GPU shard may hold: indices are 32-bit everywhere, with headroom for
// the tile rounding of the kernels
constexpr uint64_t kMaxSortRecords = 0xFFF00000ull;

// Stable LSD radix sort of n (key, value) pairs on key bits [begin_bit, end_bit), 8 bits per
// pass.  Ping-pongs between (keys, vals) and (keys_alt, vals_alt); returns true if the sorted
// data ended up in the *_alt buffers.  n < kMaxSortRecords.
// passes_out (optional) receives the number of passes run.
bool radix_sort_pairs(uint64_t* keys, uint64_t* vals, uint64_t* keys_alt, uint64_t* vals_alt, size_t n,
                      int begin_bit, int end_bit, cudaStream_t s, int* passes_out = nullptr);

// algorithmic HBM bytes of one radix pass over n pairs (read+write every 16-byte record;
// SURVEY 8d "2*N*S") and the extra bytes this implementation moves (separate digit-count read).
inline double radix_pass_alg_bytes(size_t n) { return 2.0 * 16.0 * (double)n; }

}  // namespace bgx

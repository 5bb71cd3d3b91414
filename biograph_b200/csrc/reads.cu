// reads.cu -- ingest of reads into the device-resident 2-bit store.
// Replaces the per-read packing the reference does on the host in
// prob_pass_processor::add / dna_sequence construction (bs/kmer_counter.h:297-326,
// modules/bio_base/dna_sequence.cpp) with one H2D copy + one packing kernel.
#include <vector>

#include "ctx.h"

namespace bgx {
namespace {

// warp per read; lane w packs word w of the read (32 ASCII bases -> one MSB-first uint64 + N mask)
__global__ void __launch_bounds__(256) pack_ascii_kernel(const char* __restrict__ bases,
                                                         const uint64_t* __restrict__ offs, uint64_t n_reads,
                                                         const uint32_t* __restrict__ word_off,
                                                         uint64_t* __restrict__ words, uint32_t* __restrict__ nmask,
                                                         int* __restrict__ any_n) {
  uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_reads) return;
  unsigned lane = lane_id();
  uint64_t b0 = offs[r], len = offs[r + 1] - b0;
  unsigned nw = (unsigned)((len + 31) >> 5);
  bool sawn = false;
  for (unsigned w = lane; w < nw; w += 32) {
    uint64_t word = 0;
    uint32_t m = 0;
    unsigned cnt = (unsigned)min((uint64_t)32, len - (uint64_t)w * 32);
    const char* p = bases + b0 + (uint64_t)w * 32;
    for (unsigned j = 0; j < cnt; ++j) {
      char ch = p[j];
      unsigned code;
      switch (ch) {
        case 'A': case 'a': code = 0; break;
        case 'C': case 'c': code = 1; break;
        case 'G': case 'g': code = 2; break;
        case 'T': case 't': code = 3; break;
        default: code = 0; m |= 1u << (31 - j); break;  // FASTQ alphabet is ACGTN (bio_format/fastq.cpp:82)
      }
      word |= (uint64_t)code << (62 - 2 * j);
    }
    words[word_off[r] + w] = word;
    nmask[word_off[r] + w] = m;
    sawn |= (m != 0);
  }
  if (sawn) *any_n = 1;
}

__global__ void bswap_words_kernel(uint64_t* __restrict__ w, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t x = w[i];
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  w[i] = ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}

__global__ void any_nonzero_kernel(const uint32_t* __restrict__ m, uint64_t n, int* __restrict__ flag) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && m[i]) *flag = 1;
}

template <typename T>
void grow(DevBuf<T>& buf, size_t used, size_t need, cudaStream_t s) {
  if (need <= buf.n) return;
  size_t cap = std::max(need, buf.n + buf.n / 2);
  DevBuf<T> nb(cap, s);
  if (used) BGX_CUDA(cudaMemcpyAsync(nb.p, buf.p, used * sizeof(T), cudaMemcpyDeviceToDevice, s));
  buf = std::move(nb);
}

// common bookkeeping: returns host word offsets (relative to the whole store)
void append_geometry(Context* c, const std::vector<uint16_t>& lens, std::vector<uint32_t>* woff) {
  uint64_t n = lens.size();
  woff->resize(n + 1);
  uint64_t w = c->n_words;
  const int k = c->opt.kmer_size;
  for (uint64_t r = 0; r < n; ++r) {
    (*woff)[r] = (uint32_t)w;
    w += (lens[r] + 31u) >> 5;
    c->n_bases += lens[r];
    if (lens[r] >= k) c->n_kmer_instances += lens[r] - k + 1;
    if (lens[r] > c->max_len) c->max_len = lens[r];
  }
  BGX_CHECK(w < (1ull << 32) - 2, "too many bases for one GPU shard (word offsets are 32-bit)");
  (*woff)[n] = (uint32_t)w;
}

void ensure_capacity(Context* c, uint64_t new_reads, uint64_t new_words) {
  cudaStream_t s = c->stream;
  grow(c->words, c->n_words + (c->n_words ? 1 : 0), c->n_words + new_words + 1, s);
  grow(c->nmask, c->n_words + (c->n_words ? 1 : 0), c->n_words + new_words + 1, s);
  grow(c->word_off, c->n_reads + (c->n_reads ? 1 : 0), c->n_reads + new_reads + 1, s);
  grow(c->lens, c->n_reads, c->n_reads + new_reads, s);
}

void finish_append(Context* c, const std::vector<uint16_t>& lens, const std::vector<uint32_t>& woff, int* d_flag) {
  cudaStream_t s = c->stream;
  uint64_t n = lens.size();
  BGX_CUDA(cudaMemcpyAsync(c->word_off.p + c->n_reads, woff.data(), (n + 1) * sizeof(uint32_t),
                           cudaMemcpyHostToDevice, s));
  BGX_CUDA(cudaMemcpyAsync(c->lens.p + c->n_reads, lens.data(), n * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
  // zero the pad word that load_window may touch
  BGX_CUDA(cudaMemsetAsync(c->words.p + woff[n], 0, sizeof(uint64_t), s));
  BGX_CUDA(cudaMemsetAsync(c->nmask.p + woff[n], 0, sizeof(uint32_t), s));
  int flag = 0;
  BGX_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));  // host vectors (lens/woff) must outlive the copies
  c->has_n = c->has_n || flag != 0;
  c->n_reads += n;
  c->n_words = woff[n];
  c->counted = c->corrected = c->built = false;
}

}  // namespace

void reads_ready(Context* c) {
  for (Context::UploadChunk& ch : c->upload) {
    BGX_CUDA(cudaStreamWaitEvent(c->stream, ch.ev, 0));
    cudaEventDestroy(ch.ev);
  }
  c->upload.clear();
}

void reads_append_ascii(Context* c, const char* bases, const uint64_t* offs, uint64_t n) {
  if (n == 0) return;
  reads_ready(c);
  cudaStream_t s = c->stream;
  std::vector<uint16_t> lens(n);
  for (uint64_t r = 0; r < n; ++r) {
    uint64_t L = offs[r + 1] - offs[r];
    BGX_CHECK(L <= BGX_MAX_READ_LEN, "read longer than 255 bases (the reference needs --allow-long-reads)");
    lens[r] = (uint16_t)L;
  }
  std::vector<uint32_t> woff;
  uint64_t words_before = c->n_words;
  append_geometry(c, lens, &woff);
  ensure_capacity(c, n, woff[n] - words_before);
  uint64_t nbytes = offs[n] - offs[0];
  DevBuf<char> d_bases(nbytes + 1, s);
  DevBuf<uint64_t> d_offs(n + 1, s);
  DevBuf<int> d_flag(1, s);
  BGX_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), s));
  BGX_CUDA(cudaMemcpyAsync(d_bases.p, bases + offs[0], nbytes, cudaMemcpyHostToDevice, s));
  std::vector<uint64_t> rel(n + 1);
  for (uint64_t r = 0; r <= n; ++r) rel[r] = offs[r] - offs[0];
  BGX_CUDA(cudaMemcpyAsync(d_offs.p, rel.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  // word_off for the new reads must be on the device before packing
  BGX_CUDA(cudaMemcpyAsync(c->word_off.p + c->n_reads, woff.data(), (n + 1) * sizeof(uint32_t),
                           cudaMemcpyHostToDevice, s));
  uint64_t threads = n * 32;
  KLAUNCH(pack_ascii_kernel)<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d_bases.p, d_offs.p, n,
                                                                      c->word_off.p + c->n_reads, c->words.p,
                                                                      c->nmask.p, d_flag.p);
  BGX_CUDA(cudaGetLastError());
  c->add_stat("h2d_bytes", (double)nbytes + (double)(n + 1) * 12 + (double)n * 2);
  finish_append(c, lens, woff, d_flag.p);
}

namespace {

// lengths -> words per read + totals, on the device (no host loop over the reads)
// tot: [0] words [1] bases [2] k-mer instances [3] max length [4] reads longer than 255
__global__ void geometry_kernel(const uint16_t* __restrict__ lens, uint64_t n, int k, uint32_t* __restrict__ nwords,
                                unsigned long long* __restrict__ tot) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned L = r < n ? lens[r] : 0;
  unsigned nw = (L + 31) >> 5;
  if (r < n) nwords[r] = nw;
  unsigned kin = L >= (unsigned)k ? L - k + 1 : 0;
  unsigned sw = __reduce_add_sync(0xffffffffu, nw), sb = __reduce_add_sync(0xffffffffu, L),
           sk = __reduce_add_sync(0xffffffffu, kin), mx = __reduce_max_sync(0xffffffffu, L),
           bad = __reduce_add_sync(0xffffffffu, L > BGX_MAX_READ_LEN ? 1u : 0u);
  if (lane_id() == 0) {
    if (sw) atomicAdd(&tot[0], (unsigned long long)sw);
    if (sb) atomicAdd(&tot[1], (unsigned long long)sb);
    if (sk) atomicAdd(&tot[2], (unsigned long long)sk);
    if (mx) atomicMax(&tot[3], (unsigned long long)mx);
    if (bad) atomicAdd(&tot[4], (unsigned long long)bad);
  }
}

__global__ void add_base_kernel(uint32_t* __restrict__ off, uint64_t n, uint32_t base) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off[i] += base;
}

}  // namespace

void reads_append_packed(Context* c, const uint8_t* packed, const uint32_t* n_mask, const uint64_t* word_offs,
                         const uint16_t* lens_in, uint64_t n, bool async) {
  if (n == 0) return;
  reads_ready(c);
  cudaStream_t s = c->stream;
  // the chunked form pays off for big appends only, and the N mask (needed up front: has_n picks
  // the kernel variants) keeps the synchronous path
  constexpr int kChunks = 8;
  async = async && n_mask == nullptr && n >= (1u << 18);
  uint32_t h_bound[kChunks + 1];  // first word of every chunk, relative to this append
  const int k = c->opt.kmer_size;
  // lengths go straight to the device; word offsets, totals and the length check are computed there
  grow(c->word_off, c->n_reads + (c->n_reads ? 1 : 0), c->n_reads + n + 1, s);
  grow(c->lens, c->n_reads, c->n_reads + n, s);
  BGX_CUDA(cudaMemcpyAsync(c->lens.p + c->n_reads, lens_in, n * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
  DevBuf<uint32_t> nwords(n + 1, s);
  DevBuf<unsigned long long> tot(5, s);
  BGX_CUDA(cudaMemsetAsync(tot.p, 0, 5 * 8, s));
  BGX_CUDA(cudaMemsetAsync(nwords.p + n, 0, 4, s));
  KLAUNCH(geometry_kernel)<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->lens.p + c->n_reads, n, k, nwords.p, tot.p);
  exclusive_scan_u32(nwords.p, c->word_off.p + c->n_reads, n + 1, nullptr, s);
  unsigned long long h[5];
  BGX_CUDA(cudaMemcpyAsync(h, tot.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  if (async)
    for (int j = 0; j <= kChunks; ++j)
      BGX_CUDA(cudaMemcpyAsync(&h_bound[j], c->word_off.p + c->n_reads + n * j / kChunks, 4, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  BGX_CHECK(h[4] == 0, "read longer than 255 bases");
  const uint64_t new_words = h[0], words_before = c->n_words;
  const uint64_t src_word0 = word_offs ? word_offs[0] : 0;
  if (word_offs)
    BGX_CHECK(word_offs[n] - word_offs[0] == new_words, "bgx_add_reads_packed: reads must be densely word-packed");
  BGX_CHECK(words_before + new_words < (1ull << 32) - 2, "too many bases for one GPU shard (word offsets are 32-bit)");
  if (words_before)
    KLAUNCH(add_base_kernel)<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>(c->word_off.p + c->n_reads, n + 1, (uint32_t)words_before);
  grow(c->words, words_before + (words_before ? 1 : 0), words_before + new_words + 1, s);
  grow(c->nmask, words_before + (words_before ? 1 : 0), words_before + new_words + 1, s);
  DevBuf<int> d_flag(1, s);
  BGX_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), s));
  if (async) {
    // words travel on the copy stream, chunk by chunk; everything queued so far on the main stream
    // (buffer growth, the N-mask memset) is ordered before the first copy
    if (!c->copy_stream) BGX_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    BGX_CUDA(cudaMemsetAsync(c->nmask.p + words_before, 0, (new_words + 1) * 4, s));
    cudaEvent_t ready;
    BGX_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    BGX_CUDA(cudaEventRecord(ready, s));
    BGX_CUDA(cudaStreamWaitEvent(c->copy_stream, ready, 0));
    cudaEventDestroy(ready);
    cudaStream_t cs = c->copy_stream;
    for (int j = 0; j < kChunks; ++j) {
      const uint64_t w0 = h_bound[j], w1 = h_bound[j + 1];
      if (w1 > w0) {
        BGX_CUDA(cudaMemcpyAsync(c->words.p + words_before + w0, packed + 8 * (src_word0 + w0), (w1 - w0) * 8,
                                 cudaMemcpyHostToDevice, cs));
        KLAUNCH(bswap_words_kernel)<<<(unsigned)((w1 - w0 + 255) / 256), 256, 0, cs>>>(c->words.p + words_before + w0, w1 - w0);
      }
      if (j == kChunks - 1) BGX_CUDA(cudaMemsetAsync(c->words.p + words_before + new_words, 0, sizeof(uint64_t), cs));
      Context::UploadChunk ch;
      ch.r0 = c->n_reads + n * j / kChunks;
      ch.r1 = c->n_reads + n * (j + 1) / kChunks;
      BGX_CUDA(cudaEventCreateWithFlags(&ch.ev, cudaEventDisableTiming));
      BGX_CUDA(cudaEventRecord(ch.ev, cs));
      c->upload.push_back(ch);
    }
    BGX_CUDA(cudaGetLastError());
  } else {
  BGX_CUDA(cudaMemcpyAsync(c->words.p + words_before, packed + 8 * src_word0, new_words * 8, cudaMemcpyHostToDevice, s));
  if (new_words) KLAUNCH(bswap_words_kernel)<<<(unsigned)((new_words + 255) / 256), 256, 0, s>>>(c->words.p + words_before, new_words);
  if (n_mask) {
    BGX_CUDA(cudaMemcpyAsync(c->nmask.p + words_before, n_mask + src_word0, new_words * 4, cudaMemcpyHostToDevice, s));
    if (new_words) KLAUNCH(any_nonzero_kernel)<<<(unsigned)((new_words + 255) / 256), 256, 0, s>>>(c->nmask.p + words_before, new_words, d_flag.p);
  } else {
    BGX_CUDA(cudaMemsetAsync(c->nmask.p + words_before, 0, new_words * 4, s));
  }
  // zero the pad word that load_window may touch
  BGX_CUDA(cudaMemsetAsync(c->words.p + words_before + new_words, 0, sizeof(uint64_t), s));
  BGX_CUDA(cudaMemsetAsync(c->nmask.p + words_before + new_words, 0, sizeof(uint32_t), s));
  BGX_CUDA(cudaGetLastError());
  }
  int flag = 0;
  if (!async) {
    BGX_CUDA(cudaMemcpyAsync(&flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
  }
  c->add_stat("h2d_bytes", (double)new_words * (n_mask ? 12 : 8) + (double)n * 2);
  c->has_n = c->has_n || flag != 0;
  c->n_reads += n;
  c->n_words = words_before + new_words;
  c->n_bases += h[1];
  c->n_kmer_instances += h[2];
  c->max_len = std::max<uint32_t>(c->max_len, (uint32_t)h[3]);
  c->counted = c->corrected = c->built = false;
}

}  // namespace bgx

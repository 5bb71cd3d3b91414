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
                                                         int* __restrict__ any_n,
                                                         const uint16_t* __restrict__ lens16) {
  uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_reads) return;
  unsigned lane = lane_id();
  // lens16 given: offs[r] is where read r starts in the text (FASTQ: the reads are not contiguous)
  uint64_t b0 = offs[r], len = lens16 ? (uint64_t)lens16[r] : offs[r + 1] - b0;
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

// ---- FASTQ text -> reads (semantics of fastq_reader::read, modules/bio_format/fastq.cpp:40-126) --
// Line ends are found in two passes over the text (count per 256-byte chunk, scan, write), then
// one thread per record checks its four lines the way the reference does and hands the position
// and length of the sequence line to the packing kernel.  The first failing line wins
// (line number << 4 | error code, atomicMin).
constexpr int kFqChunk = 256;
enum FqError : unsigned {
  FQ_ID_SHORT = 1, FQ_ID_AT, FQ_SEQ_EMPTY, FQ_SEQ_CHARS, FQ_PLUS_EMPTY, FQ_PLUS, FQ_QUAL_LEN, FQ_TOO_LONG, FQ_BLANK
};

__global__ void fq_count_kernel(const char* __restrict__ text, uint64_t size, uint32_t* __restrict__ cnt) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t b = t * kFqChunk;
  if (b >= size) return;
  const uint64_t e = min(size, b + kFqChunk);
  uint32_t n = 0;
  for (uint64_t i = b; i < e; ++i) n += text[i] == '\n';
  cnt[t] = n;
}

__global__ void fq_positions_kernel(const char* __restrict__ text, uint64_t size, const uint32_t* __restrict__ first,
                                    uint64_t* __restrict__ nl_pos) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t b = t * kFqChunk;
  if (b >= size) return;
  const uint64_t e = min(size, b + kFqChunk);
  uint32_t o = first[t];
  for (uint64_t i = b; i < e; ++i)
    if (text[i] == '\n') nl_pos[o++] = i;
}

__global__ void fq_records_kernel(const char* __restrict__ text, const uint64_t* __restrict__ nl_pos, uint64_t n_records,
                                  uint64_t* __restrict__ seq_start, uint16_t* __restrict__ seq_len,
                                  unsigned long long* __restrict__ first_error) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_records) return;
  uint64_t st[4], ln[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t li = 4 * r + j;
    st[j] = li ? nl_pos[li - 1] + 1 : 0;
    ln[j] = nl_pos[li] - st[j];
  }
  unsigned err = 0;
  int bad_line = 0;
  if (ln[0] == 0) { err = FQ_BLANK; bad_line = 0; }
  else if (ln[0] < 2) { err = FQ_ID_SHORT; bad_line = 0; }
  else if (text[st[0]] != '@') { err = FQ_ID_AT; bad_line = 0; }
  else if (ln[1] == 0) { err = FQ_SEQ_EMPTY; bad_line = 1; }
  else if (ln[1] > BGX_MAX_READ_LEN) { err = FQ_TOO_LONG; bad_line = 1; }
  else {
    for (uint64_t i = 0; i < ln[1]; ++i) {
      const char ch = text[st[1] + i];
      if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T' && ch != 'N') { err = FQ_SEQ_CHARS; bad_line = 1; break; }
    }
    if (!err) {
      if (ln[2] == 0) { err = FQ_PLUS_EMPTY; bad_line = 2; }
      else if (text[st[2]] != '+') { err = FQ_PLUS; bad_line = 2; }
      else if (ln[3] != ln[1]) { err = FQ_QUAL_LEN; bad_line = 3; }
    }
  }
  if (err) atomicMin(first_error, ((unsigned long long)(4 * r + bad_line + 1) << 4) | err);
  seq_start[r] = st[1];
  seq_len[r] = (uint16_t)min(ln[1], (uint64_t)BGX_MAX_READ_LEN);
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
                                                                      c->nmask.p, d_flag.p, nullptr);
  BGX_CUDA(cudaGetLastError());
  c->add_stat("h2d_bytes", (double)nbytes + (double)(n + 1) * 12 + (double)n * 2);
  finish_append(c, lens, woff, d_flag.p);
}

void reads_append_fastq(Context* c, const char* text, uint64_t size, uint64_t* n_added) {
  *n_added = 0;
  if (size == 0) return;
  reads_ready(c);
  cudaStream_t s = c->stream;
  // a last line without its newline: the reference's readline fails on it (fastq.cpp:49-53,70-73,...)
  BGX_CHECK(text[size - 1] == '\n', "Partial line in fastq file (the text must end with a newline)");
  DevBuf<char> d_text(size, s);
  BGX_CUDA(cudaMemcpyAsync(d_text.p, text, size, cudaMemcpyHostToDevice, s));
  const uint64_t n_chunks = (size + kFqChunk - 1) / kFqChunk;
  BGX_CHECK(n_chunks < (1ull << 32), "FASTQ text too large for one call (split it at a record boundary)");
  DevBuf<uint32_t> cnt(n_chunks, s), first(n_chunks, s), tot(1, s);
  KLAUNCH(fq_count_kernel)<<<(unsigned)((n_chunks + 255) / 256), 256, 0, s>>>(d_text.p, size, cnt.p);
  exclusive_scan_u32(cnt.p, first.p, n_chunks, tot.p, s);
  uint32_t n_lines = 0;
  BGX_CUDA(cudaMemcpyAsync(&n_lines, tot.p, 4, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  DevBuf<uint64_t> nl_pos(std::max<uint32_t>(n_lines, 1), s);
  KLAUNCH(fq_positions_kernel)<<<(unsigned)((n_chunks + 255) / 256), 256, 0, s>>>(d_text.p, size, first.p, nl_pos.p);
  BGX_CUDA(cudaGetLastError());
  // blank lines after the last record are skipped like the reference skips them before a record
  // (fastq.cpp:45-58); blank lines between records are not supported here and reported as such
  uint64_t trailing = 0;
  while (trailing + 1 < size && text[size - 2 - trailing] == '\n') ++trailing;
  if (size == 1 || trailing + 1 == size) return;  // nothing but newlines
  const uint64_t lines = (uint64_t)n_lines - trailing;
  const uint64_t n = lines / 4;  // whole records; a cut-off last record is reported after the records before it
  DevBuf<uint64_t> seq_start(std::max<uint64_t>(n, 1), s);
  DevBuf<uint16_t> seq_len(std::max<uint64_t>(n, 1), s);
  DevBuf<unsigned long long> first_error(1, s);
  BGX_CUDA(cudaMemsetAsync(first_error.p, 0xff, 8, s));
  if (n) KLAUNCH(fq_records_kernel)<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_text.p, nl_pos.p, n, seq_start.p, seq_len.p, first_error.p);
  BGX_CUDA(cudaGetLastError());
  unsigned long long h_err = 0;
  std::vector<uint16_t> lens(n);
  BGX_CUDA(cudaMemcpyAsync(&h_err, first_error.p, 8, cudaMemcpyDeviceToHost, s));
  if (n) BGX_CUDA(cudaMemcpyAsync(lens.data(), seq_len.p, n * 2, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  if (h_err != ~0ULL) {
    static const char* msg[] = {"", "Sequence id too short", "Sequence id missing @", "Expecting sequence, found empty line",
                                "Sequence contains unexpected characters", "Expecting +, found empty line",
                                "Expecting + as first char of line", "Quality line not same length as sequence",
                                "read longer than 255 bases (the reference needs --allow-long-reads)",
                                "blank line between records (only trailing blank lines are supported)"};
    throw Error("line " + std::to_string(h_err >> 4) + ": " + msg[h_err & 15]);
  }
  if (lines % 4 != 0) {
    static const char* what[4] = {"", "sequence", "+", "quality"};
    throw Error("line " + std::to_string(lines + 1) + ": End of file while reading " + what[lines % 4] + " line");
  }
  if (n == 0) return;
  std::vector<uint32_t> woff;
  const uint64_t words_before = c->n_words;
  append_geometry(c, lens, &woff);
  ensure_capacity(c, n, woff[n] - words_before);
  DevBuf<int> d_flag(1, s);
  BGX_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), s));
  BGX_CUDA(cudaMemcpyAsync(c->word_off.p + c->n_reads, woff.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  KLAUNCH(pack_ascii_kernel)<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(d_text.p, seq_start.p, n, c->word_off.p + c->n_reads,
                                                                     c->words.p, c->nmask.p, d_flag.p, seq_len.p);
  BGX_CUDA(cudaGetLastError());
  c->add_stat("h2d_bytes", (double)size + (double)(n + 1) * 4);
  *n_added = n;
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
  BGX_CUDA(cudaMemcpyAsync(c->lens.p + c->n_reads, lens_in, n * sizeof(uint16_t), cudaMemcpyDefault, s));
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
                                 cudaMemcpyDefault, cs));
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
  BGX_CUDA(cudaMemcpyAsync(c->words.p + words_before, packed + 8 * src_word0, new_words * 8, cudaMemcpyDefault, s));
  if (new_words) KLAUNCH(bswap_words_kernel)<<<(unsigned)((new_words + 255) / 256), 256, 0, s>>>(c->words.p + words_before, new_words);
  if (n_mask) {
    BGX_CUDA(cudaMemcpyAsync(c->nmask.p + words_before, n_mask + src_word0, new_words * 4, cudaMemcpyDefault, s));
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

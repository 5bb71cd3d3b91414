// api.cu -- the C ABI (include/bgx.h) over the CUDA stages.  No CPU fallback: every entry
// point needs a CUDA device and fails loudly without one.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

#include "ctx.h"

namespace bgx {
void export_entries_ascii(Context* c, uint64_t first, uint64_t count, char** bases, uint64_t** offs_out);

namespace {
thread_local std::string g_last_error;

template <typename F>
int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  } catch (...) {
    g_last_error = "unknown error";
    return 1;
  }
}

__global__ void corrected_ascii_kernel(const uint64_t* __restrict__ store, const uint32_t* __restrict__ word_off,
                                       const uint16_t* __restrict__ clen, const uint64_t* __restrict__ out_off,
                                       uint64_t n_reads, char* __restrict__ out) {
  uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_reads) return;
  int L = clen[r];
  const uint64_t* w = store + word_off[r];
  char* o = out + out_off[r];
  for (int j = lane_id(); j < L; j += 32) o[j] = "ACGT"[(w[j >> 5] >> (62 - 2 * (j & 31))) & 3];
}

__global__ void reads_ascii_kernel(const uint64_t* __restrict__ words, const uint32_t* __restrict__ nmask,
                                   const uint32_t* __restrict__ word_off, const uint16_t* __restrict__ lens,
                                   const uint64_t* __restrict__ out_off, uint64_t n_reads, char* __restrict__ out) {
  uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_reads) return;
  int L = lens[r];
  const uint64_t* w = words + word_off[r];
  const uint32_t* m = nmask ? nmask + word_off[r] : nullptr;
  char* o = out + out_off[r];
  for (int j = lane_id(); j < L; j += 32) {
    const bool is_n = m && ((m[j >> 5] >> (31 - (j & 31))) & 1u);
    o[j] = is_n ? 'N' : "ACGT"[(w[j >> 5] >> (62 - 2 * (j & 31))) & 3];
  }
}

template <typename T>
T* to_host(const T* d, size_t n, cudaStream_t s) {
  T* h = (T*)host_alloc(std::max<size_t>(n, 1) * sizeof(T));
  if (n) BGX_CUDA(cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, s));
  return h;
}

}  // namespace
}  // namespace bgx

using namespace bgx;

struct bgx_ctx {
  Context c;
};

extern "C" {

void bgx_default_options(bgx_options* o) {
  memset(o, 0, sizeof(*o));
  o->kmer_size = 30;
  o->min_kmer_count = 5;
  o->max_corrections = 8;
  o->min_good_run = 2;
  o->trim_after_portion = 0.7f;
  o->device = 0;
  o->sort_key_bits = 0;  // auto
}

const char* bgx_last_error(void) { return g_last_error.c_str(); }
const char* bgx_version(void) { return "bgx 0.1 (sm_100a)"; }

int bgx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int bgx_create(const bgx_options* opts, bgx_ctx** out) {
  *out = nullptr;
  return guard([&] {
    bgx_options o;
    if (opts) o = *opts; else bgx_default_options(&o);
    // bs/kmer_counter.cpp:52-54: k in [16,31] (k-mer + 2 flag bits must fit 64 bits)
    BGX_CHECK(o.kmer_size >= 16 && o.kmer_size <= 31, "kmer_size must be in [16,31]");
    BGX_CHECK(o.min_kmer_count >= 1, "min_kmer_count must be >= 1");
    BGX_CHECK(o.max_corrections >= 0 && o.max_corrections <= 32, "max_corrections must be in [0,32]");  // biograph_create.cpp:486
    BGX_CHECK(o.min_good_run >= 0, "min_good_run must be >= 0");
    BGX_CHECK(o.trim_after_portion >= 0.f && o.trim_after_portion <= 1.f, "trim_after_portion must be in [0,1]");
    BGX_CHECK(o.sort_key_bits == 0 || (o.sort_key_bits >= 16 && o.sort_key_bits <= 64 && o.sort_key_bits % 8 == 0),
              "sort_key_bits must be 0 (auto) or a multiple of 8 in [16,64]");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    BGX_CHECK(e == cudaSuccess && ndev > 0, "no CUDA device: bgx has no CPU fallback");
    BGX_CHECK(o.device >= 0 && o.device < ndev, "bad device ordinal");
    BGX_CUDA(cudaSetDevice(o.device));
    {
      // The hot tables are probed at random: ask L2 to fetch 32-byte sectors instead of
      // promoting misses to 64/128 bytes (hint; BGX_L2_FETCH overrides for experiments).
      size_t gran = 32;
      if (const char* e = getenv("BGX_L2_FETCH")) gran = (size_t)atoi(e);
      if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    }
    bgx_ctx* x = new bgx_ctx();
    x->c.opt = o;
    x->c.device = o.device;
    {
      size_t free_b = 0, total_b = 0;
      BGX_CUDA(cudaMemGetInfo(&free_b, &total_b));
      x->c.total_mem = total_b;
    }
    BGX_CUDA(cudaStreamCreateWithFlags(&x->c.stream, cudaStreamNonBlocking));
    *out = x;
  });
}

void bgx_destroy(bgx_ctx* x) {
  if (!x) return;
  cudaSetDevice(x->c.device);
  cudaStream_t s = x->c.stream;
  if (x->c.copy_stream) {
    cudaStreamSynchronize(x->c.copy_stream);
    for (Context::UploadChunk& ch : x->c.upload) cudaEventDestroy(ch.ev);
    x->c.upload.clear();
    cudaStreamDestroy(x->c.copy_stream);
  }
  cudaStreamSynchronize(s);
  {
    Context& c = x->c;
    // release buffers while the stream is alive
    c.words.release(); c.nmask.release(); c.word_off.release(); c.lens.release();
    c.table.release(); c.solid.release(); c.store.release(); c.gstore.release(); c.clen.release(); c.ncorr.release();
    c.next_fwd.release(); c.next_rev.release(); c.ent_key.release(); c.ent_loc.release();
    c.sizes.release(); c.shared.release(); c.prev_bits.release(); c.prev_sub.release(); c.prev_acc.release();
    merge_release(&c);
  }
  dist_destroy(&x->c);
  dev_trim(s);
  cudaStreamDestroy(s);
  delete x;
}

void bgx_free(void* p) { host_free(p); }

#define CTX_GUARD(...)                          \
  if (!x) { g_last_error = "null context"; return 1; } \
  return guard([&] {                            \
    BGX_CUDA(cudaSetDevice(x->c.device));       \
    Context* c = &x->c;                         \
    (void)c;                                    \
    __VA_ARGS__                                 \
  });

int bgx_add_reads_ascii(bgx_ctx* x, const char* bases, const uint64_t* offs, uint64_t n_reads) {
  CTX_GUARD({ reads_append_ascii(c, bases, offs, n_reads); })
}

int bgx_add_reads_fastq(bgx_ctx* x, const char* text, uint64_t size, uint64_t* n_reads) {
  CTX_GUARD({ reads_append_fastq(c, text, size, n_reads); })
}

int bgx_add_reads_packed(bgx_ctx* x, const uint8_t* packed, const uint32_t* n_mask, const uint64_t* word_offs,
                         const uint16_t* lens, uint64_t n_reads) {
  CTX_GUARD({ reads_append_packed(c, packed, n_mask, word_offs, lens, n_reads); })
}

int bgx_add_reads_packed_async(bgx_ctx* x, const uint8_t* packed, const uint32_t* n_mask, const uint64_t* word_offs,
                               const uint16_t* lens, uint64_t n_reads) {
  CTX_GUARD({ reads_append_packed(c, packed, n_mask, word_offs, lens, n_reads, true); })
}

int bgx_count_kmers(bgx_ctx* x) { CTX_GUARD({ stage_count_kmers(c); }) }

int bgx_export_kmers(bgx_ctx* x, uint32_t min_count, uint64_t* n, uint64_t** kmers, uint32_t** fwd, uint32_t** rev,
                     uint8_t** flags) {
  CTX_GUARD({ export_kmers(c, min_count, n, kmers, fwd, rev, flags); })
}

int bgx_correct(bgx_ctx* x) { CTX_GUARD({ stage_correct(c); }) }

int bgx_seed_uncorrected(bgx_ctx* x) { CTX_GUARD({ stage_seed_uncorrected(c); }) }

int bgx_export_varbit(bgx_ctx* x, int32_t which, uint64_t** words, uint64_t* n_words, uint32_t* bits_per_value,
                      uint64_t* max_value) {
  CTX_GUARD({ export_varbit(c, which, words, n_words, bits_per_value, max_value); })
}

int bgx_export_corrected(bgx_ctx* x, uint64_t* n_reads, uint16_t** lens, char** bases, uint64_t* n_bases,
                         uint8_t** corrections, uint16_t** next_fwd, uint16_t** next_rev) {
  CTX_GUARD({
    BGX_CHECK(c->corrected, "bgx_export_corrected: call bgx_correct first");
    cudaStream_t s = c->stream;
    uint64_t n = c->n_reads;
    if (n_reads) *n_reads = n;
    uint16_t* h_len = to_host(c->clen.p, n, s);
    if (corrections) *corrections = to_host(c->ncorr.p, n, s);
    if (next_fwd) *next_fwd = to_host(c->next_fwd.p, n, s);
    if (next_rev) *next_rev = to_host(c->next_rev.p, n, s);
    BGX_CUDA(cudaStreamSynchronize(s));
    std::vector<uint64_t> off(n + 1);
    off[0] = 0;
    for (uint64_t r = 0; r < n; ++r) off[r + 1] = off[r] + h_len[r];
    if (n_bases) *n_bases = off[n];
    if (bases) {
      char* out = (char*)host_alloc(std::max<uint64_t>(off[n], 1));
      DevBuf<uint64_t> d_off(n + 1, s);
      DevBuf<char> d_out(std::max<uint64_t>(off[n], 1), s);
      BGX_CUDA(cudaMemcpyAsync(d_off.p, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
      if (n) KLAUNCH(corrected_ascii_kernel)<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(c->store.p, c->word_off.p, c->clen.p,
                                                                                  d_off.p, n, d_out.p);
      BGX_CUDA(cudaGetLastError());
      BGX_CUDA(cudaMemcpyAsync(out, d_out.p, off[n], cudaMemcpyDeviceToHost, s));
      BGX_CUDA(cudaStreamSynchronize(s));
      *bases = out;
    }
    if (lens) *lens = h_len; else host_free(h_len);
  })
}

int bgx_export_reads(bgx_ctx* x, uint64_t* n_reads, uint16_t** lens, char** bases, uint64_t* n_bases) {
  CTX_GUARD({
    reads_ready(c);
    cudaStream_t s = c->stream;
    const uint64_t n = c->n_reads;
    if (n_reads) *n_reads = n;
    uint16_t* h_len = to_host(c->lens.p, n, s);
    BGX_CUDA(cudaStreamSynchronize(s));
    std::vector<uint64_t> off(n + 1);
    off[0] = 0;
    for (uint64_t r = 0; r < n; ++r) off[r + 1] = off[r] + h_len[r];
    if (n_bases) *n_bases = off[n];
    if (bases) {
      char* out = (char*)host_alloc(std::max<uint64_t>(off[n], 1));
      DevBuf<uint64_t> d_off(n + 1, s);
      DevBuf<char> d_out(std::max<uint64_t>(off[n], 1), s);
      BGX_CUDA(cudaMemcpyAsync(d_off.p, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
      if (n) KLAUNCH(reads_ascii_kernel)<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(
          c->words.p, c->has_n ? c->nmask.p : nullptr, c->word_off.p, c->lens.p, d_off.p, n, d_out.p);
      BGX_CUDA(cudaGetLastError());
      BGX_CUDA(cudaMemcpyAsync(out, d_out.p, off[n], cudaMemcpyDeviceToHost, s));
      BGX_CUDA(cudaStreamSynchronize(s));
      *bases = out;
    }
    if (lens) *lens = h_len; else host_free(h_len);
  })
}

int bgx_build_seqset(bgx_ctx* x) { CTX_GUARD({ stage_build_seqset(c); }) }

int bgx_export_seqset(bgx_ctx* x, uint64_t* n_entries, uint32_t* max_entry_len, uint16_t** sizes, uint16_t** shared,
                      uint64_t* prev_bits[4], uint64_t* prev_subaccum[4], uint64_t* prev_accum[4], uint64_t fixed[5]) {
  CTX_GUARD({
    BGX_CHECK(c->built, "bgx_export_seqset: call bgx_build_seqset first");
    cudaStream_t s = c->stream;
    uint64_t n = c->n_entries;
    if (n_entries) *n_entries = n;
    if (max_entry_len) *max_entry_len = c->max_entry_len;
    if (sizes) *sizes = to_host(c->sizes.p, n, s);
    if (shared) *shared = to_host(c->shared.p, n, s);
    for (int b = 0; b < 4; ++b) {
      if (prev_bits) prev_bits[b] = to_host(c->prev_bits.p + b * c->prev_words, c->prev_words, s);
      if (prev_subaccum) prev_subaccum[b] = to_host(c->prev_sub.p + b * c->sub_words, c->sub_words, s);
      if (prev_accum) prev_accum[b] = to_host(c->prev_acc.p + b * c->acc_words, c->acc_words, s);
    }
    if (fixed) memcpy(fixed, c->fixed, sizeof(c->fixed));
    BGX_CUDA(cudaStreamSynchronize(s));
    c->add_stat("d2h_bytes", (double)n * 4 + 4.0 * 8 * (c->prev_words + c->sub_words + c->acc_words) + 40);
  })
}

int bgx_lookup_reads(bgx_ctx* x, uint64_t* n_reads, uint64_t** fwd_entry, uint64_t** rc_entry) {
  CTX_GUARD({ lookup_reads(c, n_reads, fwd_entry, rc_entry); })
}

int bgx_build_readmap(bgx_ctx* x, int32_t paired, uint64_t* n_rows, uint16_t** read_lengths, uint64_t** mate_loop_ptr,
                      uint64_t** is_forward, uint64_t* read_ids_source[3], uint64_t* read_ids_dest[3]) {
  CTX_GUARD({ build_readmap(c, paired, n_rows, read_lengths, mate_loop_ptr, is_forward, read_ids_source, read_ids_dest); })
}

int bgx_export_entries_ascii(bgx_ctx* x, uint64_t first, uint64_t count, char** bases, uint64_t** offs) {
  CTX_GUARD({ export_entries_ascii(c, first, count, bases, offs); })
}

int bgx_merge_seqsets(bgx_ctx* x, const bgx_seqset_part* parts, uint32_t n_parts, uint64_t parallel_splits) {
  CTX_GUARD({ stage_merge_seqsets(c, parts, n_parts, parallel_splits); })
}

int bgx_export_mergemap(bgx_ctx* x, uint32_t part, uint64_t* merged_entries[3], uint64_t* n_bits, uint64_t* n_set) {
  CTX_GUARD({ export_mergemap(c, part, merged_entries, n_bits, n_set); })
}

int bgx_migrate_bits(bgx_ctx* x, uint32_t part, const uint64_t* old_bits, uint64_t n_old, uint64_t* migrated[3],
                     uint64_t* n_bits) {
  CTX_GUARD({ migrate_bits(c, part, old_bits, n_old, migrated, n_bits); })
}

int bgx_export_flat_ascii(bgx_ctx* x, uint32_t part, uint64_t first, uint64_t count, char** bases, uint64_t** offs) {
  CTX_GUARD({ export_flat_ascii(c, part, first, count, bases, offs); })
}

int bgx_run(bgx_ctx* x) {
  CTX_GUARD({
    stage_count_kmers(c);
    stage_correct(c);
    stage_build_seqset(c);
  })
}

int bgx_reset_results(bgx_ctx* x) {
  CTX_GUARD({
    c->table.release(); c->solid.release(); c->store.release(); c->gstore.release(); c->clen.release(); c->ncorr.release();
    c->next_fwd.release(); c->next_rev.release(); c->ent_key.release(); c->ent_loc.release();
    c->sizes.release(); c->shared.release(); c->prev_bits.release(); c->prev_sub.release(); c->prev_acc.release();
    merge_release(c);
    c->counted = c->corrected = c->built = false;
    c->stats.clear();
    c->stat_order.clear();
  })
}

int bgx_clear_reads(bgx_ctx* x) {
  CTX_GUARD({
    // copies of an async append still in flight read the caller's host buffer: wait for them here, so
    // that "clear, then free the host buffer" is safe, and order them before the buffers are reused
    if (c->copy_stream && !c->upload.empty()) BGX_CUDA(cudaStreamSynchronize(c->copy_stream));
    reads_ready(c);
    c->words.release(); c->nmask.release(); c->word_off.release(); c->lens.release();
    c->n_reads = c->n_words = c->n_bases = c->n_kmer_instances = 0;
    c->has_n = false;
    c->max_len = 0;
    c->counted = c->corrected = c->built = false;
  })
}

int bgx_stats_json(bgx_ctx* x, char* buf, size_t cap) {
  CTX_GUARD({
    std::ostringstream os;
    os.precision(17);
    os << "{\"n_reads\":" << c->n_reads << ",\"n_bases\":" << c->n_bases << ",\"has_n\":" << (c->has_n ? "true" : "false");
    for (const auto& k : c->stat_order) os << ",\"" << k << "\":" << c->stats[k];
    os << ",\"peak_device_bytes\":" << dev_peak_bytes(false) << ",\"live_device_bytes\":" << dev_live_bytes();
    os << "}";
    std::string sjson = os.str();
    BGX_CHECK(sjson.size() + 1 <= cap, "bgx_stats_json: buffer too small");
    memcpy(buf, sjson.c_str(), sjson.size() + 1);
  })
}

int bgx_timer_start(bgx_ctx* x) {
  CTX_GUARD({
    if (!c->t0) { BGX_CUDA(cudaEventCreate(&c->t0)); BGX_CUDA(cudaEventCreate(&c->t1)); }
    BGX_CUDA(cudaEventRecord(c->t0, c->stream));
  })
}

int bgx_timer_stop(bgx_ctx* x, double* elapsed_ms) {
  CTX_GUARD({
    BGX_CHECK(c->t0 != nullptr, "bgx_timer_stop without bgx_timer_start");
    BGX_CUDA(cudaEventRecord(c->t1, c->stream));
    BGX_CUDA(cudaEventSynchronize(c->t1));
    float ms = 0;
    BGX_CUDA(cudaEventElapsedTime(&ms, c->t0, c->t1));
    *elapsed_ms = ms;
  })
}

int bgx_dist_unique_id(uint8_t id[128]) {
  return guard([&] { dist_get_unique_id(id); });
}

int bgx_dist_init(bgx_ctx* x, int32_t world_size, int32_t rank, const uint8_t id[128]) {
  CTX_GUARD({ dist_init(c, world_size, rank, id); })
}

int bgx_seqset_layout(bgx_ctx* x, uint64_t layout[6]) {
  CTX_GUARD({
    BGX_CHECK(c->built, "bgx_seqset_layout: call bgx_build_seqset first");
    layout[0] = c->n_entries;
    layout[1] = c->n_entries_global;
    layout[2] = c->first_entry_global;
    layout[3] = c->prev_words;
    layout[4] = c->sub_words;
    layout[5] = c->acc_words;
  })
}

int bgx_debug_sort_pairs(bgx_ctx* x, uint64_t* keys, uint64_t* vals, uint64_t n, int begin_bit, int end_bit) {
  CTX_GUARD({
    cudaStream_t s = c->stream;
    DevBuf<uint64_t> k0(n + 1, s), v0(n + 1, s), k1(n + 1, s), v1(n + 1, s);
    BGX_CUDA(cudaMemcpyAsync(k0.p, keys, n * 8, cudaMemcpyHostToDevice, s));
    BGX_CUDA(cudaMemcpyAsync(v0.p, vals, n * 8, cudaMemcpyHostToDevice, s));
    bool alt = radix_sort_pairs(k0.p, v0.p, k1.p, v1.p, n, begin_bit, end_bit, s);
    BGX_CUDA(cudaMemcpyAsync(keys, alt ? k1.p : k0.p, n * 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaMemcpyAsync(vals, alt ? v1.p : v0.p, n * 8, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
  })
}

// host-only hooks for the CPU tests: the k-mer hash and its inverse, and the counting plan
uint64_t bgx_debug_khash(uint64_t x, int32_t k, int32_t inverse) { return inverse ? khash_inv(x, k) : khash(x, k); }
void bgx_debug_count_plan(uint64_t k_local, uint64_t k_share, int32_t n_ranks, uint64_t total_mem, uint64_t batch_reads,
                          uint64_t n_reads, uint64_t* batches, int32_t* part_bits) {
  int rank_bits = 0;
  while ((1 << rank_bits) < n_ranks) ++rank_bits;
  *batches = plan_count_batches(k_local, k_share, n_ranks, total_mem, batch_reads, n_reads);
  *part_bits = plan_part_bits(k_share, *batches, rank_bits);
}

uint64_t bgx_launch_count(void) { return __atomic_load_n(&bgx::g_launches, __ATOMIC_RELAXED); }

}  // extern "C"

// merge.cu -- `biograph merge`'s seqset path on the GPU: flatten every input seqset,
// merge the entry sets, build the merged tables and the per-input mergemaps, migrate readmap bits.
//
// Replaces (reference, CPU; driver modules/biograph/biograph_merge.cpp:199-330):
//   seqset_flat_builder::build     modules/bio_base/seqset_flat.cpp:232-290  (entry sequences)
//   make_mergemap::build / fill    modules/bio_base/make_mergemap.cpp:22-44,188-259
//   seqset_merger::build           modules/bio_base/seqset_merger.cpp:53-78,109-197
//   make_readmap::fast_migrate     modules/bio_mapred/make_readmap.cpp:459-520 (the read_ids part)
//
// The per-element work lives in merge_core.cuh (also run on the CPU by tests/cpp/merge_core_test.cpp);
// this file holds the thin kernels around it and the host sequence.  The merge itself is NOT a k-way
// queue merge (make_mergemap.cpp:188-233 pops one entry at a time): every input entry becomes one
// (key, loc) record over a common 2-bit store and the union goes through the seqset stage's radix sort,
// prefix-dedup and tables kernels (build_seqset_from_records, seqset.cu) -- the merged seqset is the
// prefix-dedup of the sorted union, and the position map of that dedup IS the mergemap.
#include <algorithm>
#include <vector>

#include "ctx.h"
#include "merge_core.cuh"

namespace bgx {
using namespace mergecore;

namespace {

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)std::max<uint64_t>(1, (n + block - 1) / block); }

uint32_t read_u32(const uint32_t* d, cudaStream_t s) {
  uint32_t h;
  BGX_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  return h;
}

__global__ void popc_words_kernel(const uint64_t* __restrict__ bits, uint64_t words, uint64_t nbits,
                                  uint32_t* __restrict__ cnt) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < words) cnt[w] = (uint32_t)popc64(masked_word(bits, w, nbits));
}

__global__ void scatter_bits_kernel(const uint64_t* __restrict__ bits, uint64_t words, uint64_t nbits,
                                    const uint32_t* __restrict__ excl, uint32_t* __restrict__ out, uint64_t out_base) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < words) scatter_set_bits(bits, w, nbits, excl[w], out, out_base);
}

__global__ void double_init_kernel(uint64_t n, uint64_t f1, uint64_t f2, uint64_t f3, uint64_t* __restrict__ w0) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) double_init(i, f1, f2, f3, w0);
}

__global__ void double_step_kernel(const uint64_t* __restrict__ w_in, const uint32_t* __restrict__ j_in, int have,
                                   uint64_t n, uint64_t* __restrict__ w_out, uint32_t* __restrict__ j_out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) double_step(w_in, j_in, have, i, w_out, j_out);
}

// words per entry; flags a size no seqset entry can have
__global__ void entry_words_kernel(const uint16_t* __restrict__ sizes, uint64_t n, uint32_t* __restrict__ cnt,
                                   int* __restrict__ bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t sz = sizes[i];
  if (sz == 0 || sz > BGX_MAX_READ_LEN) *bad = 1;
  cnt[i] = entry_words(sz);
}

__global__ void emit_entries_kernel(const uint64_t* __restrict__ w32, const uint32_t* __restrict__ j32,
                                    const uint16_t* __restrict__ sizes, const uint32_t* __restrict__ woff, uint64_t n,
                                    uint64_t word_base, uint64_t rec_base, uint64_t* __restrict__ store,
                                    uint64_t* __restrict__ keys, uint64_t* __restrict__ locs) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) emit_entry(w32, j32, sizes, woff, i, word_base, rec_base, store, keys, locs);
}

__global__ void mergemap_mark_kernel(const uint64_t* __restrict__ sorted_locs, const uint32_t* __restrict__ pos,
                                     uint32_t n, PartTable pt, unsigned long long* __restrict__ mm_bits,
                                     uint64_t mm_words) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) mergemap_mark(sorted_locs, pos, j, pt, mm_bits, mm_words);
}

__global__ void migrate_kernel(const uint64_t* __restrict__ old_bits, uint64_t words, uint64_t n_old,
                               const uint32_t* __restrict__ sel, unsigned long long* __restrict__ new_bits) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < words) migrate_word(old_bits, w, n_old, sel, new_bits);
}

__global__ void flat_lens_kernel(const uint64_t* __restrict__ locs, uint64_t first, uint64_t count,
                                 uint32_t* __restrict__ lens) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) lens[i] = loc_len(locs[first + i]);
}

__global__ void flat_ascii_kernel(const uint64_t* __restrict__ store, const uint64_t* __restrict__ locs,
                                  const uint64_t* __restrict__ offs, uint64_t first, uint64_t count,
                                  char* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t l = locs[first + i];
  const uint64_t a = loc_addr(l);
  const int len = (int)loc_len(l);
  char* o = out + offs[i];
  for (int j = 0; j < len; ++j) {
    const uint64_t p = a + j;
    o[j] = "ACGT"[(store[p >> 5] >> (62 - 2 * (p & 31))) & 3];
  }
}

// positions of the set bits of a device bit vector, in order, into out[out_base ...]; returns how many
uint32_t select_table(Context* c, const uint64_t* d_bits, uint64_t nbits, uint32_t* d_out, uint64_t out_base) {
  cudaStream_t s = c->stream;
  const uint64_t words = (nbits + 63) / 64;
  if (words == 0) return 0;
  DevBuf<uint32_t> cnt(words, s), tot(1, s);
  KLAUNCH(popc_words_kernel)<<<grid_for(words, 256), 256, 0, s>>>(d_bits, words, nbits, cnt.p);
  exclusive_scan_u32(cnt.p, cnt.p, words, tot.p, s);
  KLAUNCH(scatter_bits_kernel)<<<grid_for(words, 256), 256, 0, s>>>(d_bits, words, nbits, cnt.p, d_out, out_base);
  BGX_CUDA(cudaGetLastError());
  return read_u32(tot.p, s);
}

// set bits of a device bit vector
uint32_t count_bits(Context* c, const uint64_t* d_bits, uint64_t nbits) {
  cudaStream_t s = c->stream;
  const uint64_t words = (nbits + 63) / 64;
  if (words == 0) return 0;
  DevBuf<uint32_t> cnt(words, s), tot(1, s);
  KLAUNCH(popc_words_kernel)<<<grid_for(words, 256), 256, 0, s>>>(d_bits, words, nbits, cnt.p);
  exclusive_scan_u32(cnt.p, cnt.p, words, tot.p, s);
  BGX_CUDA(cudaGetLastError());
  return read_u32(tot.p, s);
}

}  // namespace

void merge_release(Context* c) {
  c->merge_parts.clear();
  c->mergemap.release();
  c->mergemap_words = 0;
}

void stage_merge_seqsets(Context* c, const bgx_seqset_part* parts, uint32_t n_parts, uint64_t parallel_splits) {
  BGX_CHECK(parts != nullptr && n_parts >= 1, "bgx_merge_seqsets: no inputs");
  BGX_CHECK(n_parts <= (uint32_t)kMaxParts, "bgx_merge_seqsets: at most 64 inputs");
  BGX_CHECK(c->dist.nranks == 1, "bgx_merge_seqsets: single-GPU contexts only");
  BGX_CHECK(parallel_splits <= (1ull << 31), "bgx_merge_seqsets: parallel_splits out of range");
  if (parallel_splits == 0) parallel_splits = 100000;  // g_parallel_splits (modules/io/parallel.cpp:13)
  cudaStream_t s = c->stream;
  ScopedStage st_all(c, "merge_total");

  // whatever an earlier build or merge left behind goes first (the reads, if any, stay)
  c->store.release(); c->gstore.release(); c->ent_key.release(); c->ent_loc.release();
  c->sizes.release(); c->shared.release(); c->prev_bits.release(); c->prev_sub.release(); c->prev_acc.release();
  c->built = c->corrected = false;   // the corrected store, if there was one, is gone
  merge_release(c);

  uint64_t N = 0;
  for (uint32_t p = 0; p < n_parts; ++p) {
    BGX_CHECK(parts[p].n_entries >= 1, "bgx_merge_seqsets: an input has no entries");
    BGX_CHECK(parts[p].n_entries < (1ull << 29), "bgx_merge_seqsets: an input has too many entries for one GPU");
    BGX_CHECK(parts[p].sizes != nullptr, "bgx_merge_seqsets: sizes missing");
    for (int b = 0; b < 4; ++b) BGX_CHECK(parts[p].prev_bits[b] != nullptr, "bgx_merge_seqsets: prev bits missing");
    N += parts[p].n_entries;
  }
  BGX_CHECK(N < (1ull << 31), "bgx_merge_seqsets: too many entries in total for one GPU");

  // ---- 1. sizes of every input: words per entry, store offsets ------------------------------------------------
  ScopedStage st_flat(c, "merge_flatten");
  std::vector<DevBuf<uint16_t>> d_sizes(n_parts);
  std::vector<DevBuf<uint32_t>> d_woff(n_parts);
  PartTable pt;
  pt.n = (int)n_parts;
  pt.word_base[0] = 0;
  DevBuf<int> bad(1, s);
  BGX_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
  for (uint32_t p = 0; p < n_parts; ++p) {
    const uint64_t n = parts[p].n_entries;
    d_sizes[p].alloc(n, s);
    d_woff[p].alloc(n, s);
    BGX_CUDA(cudaMemcpyAsync(d_sizes[p].p, parts[p].sizes, n * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
    DevBuf<uint32_t> tot(1, s);
    KLAUNCH(entry_words_kernel)<<<grid_for(n, 256), 256, 0, s>>>(d_sizes[p].p, n, d_woff[p].p, bad.p);
    exclusive_scan_u32(d_woff[p].p, d_woff[p].p, n, tot.p, s);
    BGX_CUDA(cudaGetLastError());
    pt.word_base[p + 1] = pt.word_base[p] + read_u32(tot.p, s);
  }
  {
    int h_bad = 0;
    BGX_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
    BGX_CHECK(!h_bad, "bgx_merge_seqsets: an entry size of 0 or above 255 (not a seqset this library can hold)");
  }
  const uint64_t total_words = pt.word_base[n_parts];
  c->add_stat("h2d_bytes", 2.0 * (double)N);

  // ---- 2. flatten every input into the common store; one record per entry ---------------------------------------
  DevBuf<uint64_t> store(total_words + 2, s);   // + pad words: window() reads one word past a sequence
  BGX_CUDA(cudaMemsetAsync(store.p + total_words, 0, 2 * sizeof(uint64_t), s));
  const size_t cap = (size_t)N + 1024;
  DevBuf<uint64_t> keys(cap, s), locs(cap, s), keys_alt(cap, s), locs_alt(cap, s);
  c->merge_parts.resize(n_parts);
  uint64_t rec_base = 0;
  for (uint32_t p = 0; p < n_parts; ++p) {
    const uint64_t n = parts[p].n_entries;
    const uint64_t words = (n + 63) / 64;
    DevBuf<uint64_t> bits(4 * words, s);
    for (int b = 0; b < 4; ++b)
      BGX_CUDA(cudaMemcpyAsync(bits.p + b * words, parts[p].prev_bits[b], words * 8, cudaMemcpyHostToDevice, s));
    c->add_stat("h2d_bytes", 32.0 * (double)words);
    // next[i] = pop_front of entry i: the set bits of prev_b, in order, are the pops of the entries that
    // start with b, which are entries [fixed[b], fixed[b+1]) (seqset.cpp:113-129,709-720)
    DevBuf<uint32_t> next(n, s), next_alt(n, s);
    uint64_t fixed[5] = {0, 0, 0, 0, 0};
    for (int b = 0; b < 4; ++b) {
      const uint32_t cnt = count_bits(c, bits.p + b * words, n);
      fixed[b + 1] = fixed[b] + cnt;
    }
    BGX_CHECK(fixed[4] == n, "bgx_merge_seqsets: Invalid seqset: prev bit totals != entries");  // seqset.cpp:123-126
    for (int b = 0; b < 4; ++b) select_table(c, bits.p + b * words, n, next.p, fixed[b]);
    bits.release();
    // five doubling rounds: (1 base, next) -> (32 bases, next^32)
    DevBuf<uint64_t> wa(n, s), wb(n, s);
    KLAUNCH(double_init_kernel)<<<grid_for(n, 256), 256, 0, s>>>(n, fixed[1], fixed[2], fixed[3], wa.p);
    uint64_t* w_in = wa.p; uint64_t* w_out = wb.p;
    uint32_t* j_in = next.p; uint32_t* j_out = next_alt.p;
    for (int have = 1; have < 32; have <<= 1) {
      KLAUNCH(double_step_kernel)<<<grid_for(n, 256), 256, 0, s>>>(w_in, j_in, have, n, w_out, j_out);
      std::swap(w_in, w_out);
      std::swap(j_in, j_out);
    }
    // (w_in, j_in) now hold the 32-base words and the 32-pop pointers
    KLAUNCH(emit_entries_kernel)<<<grid_for(n, 128), 128, 0, s>>>(w_in, j_in, d_sizes[p].p, d_woff[p].p, n, pt.word_base[p],
                                                            rec_base, store.p, keys.p, locs.p);
    BGX_CUDA(cudaGetLastError());
    Context::MergePart& mp = c->merge_parts[p];
    mp.n = n;
    mp.flat_loc.alloc(n, s);
    BGX_CUDA(cudaMemcpyAsync(mp.flat_loc.p, locs.p + rec_base, n * 8, cudaMemcpyDeviceToDevice, s));
    BGX_CUDA(cudaStreamSynchronize(s));   // the temporaries of this input go back to the arena in stream order
    d_sizes[p].release();
    d_woff[p].release();
    rec_base += n;
  }
  c->store = std::move(store);
  c->set_stat("merge_inputs", n_parts);
  c->set_stat("merge_input_entries", (double)N);
  c->set_stat("merge_flat_words", (double)total_words);
  st_flat.stop();

  // ---- 3. sort + prefix-dedup + tables; the first dedup's position map gives the mergemaps -----------------------
  MergeHooks mh;
  mh.parallel_splits = parallel_splits;
  mh.after_dedup = [&](const uint64_t* sorted_locs, uint32_t n, const uint32_t* pos, uint32_t n_kept) {
    c->mergemap_words = ((uint64_t)n_kept + 63) / 64;
    c->mergemap.alloc(std::max<uint64_t>(c->mergemap_words * n_parts, 1), s);
    BGX_CUDA(cudaMemsetAsync(c->mergemap.p, 0, std::max<uint64_t>(c->mergemap_words * n_parts, 1) * 8, s));
    KLAUNCH(mergemap_mark_kernel)<<<grid_for(n, 256), 256, 0, s>>>(sorted_locs, pos, n, pt, c->mergemap.p, c->mergemap_words);
    BGX_CUDA(cudaGetLastError());
  };
  build_seqset_from_records(c, keys, locs, keys_alt, locs_alt, (uint32_t)N, &mh);
  c->corrected = false;   // the store holds flat entries, not corrected reads
  c->set_stat("merge_entries", (double)c->n_entries);
  st_all.stop();
}

void export_mergemap(Context* c, uint32_t part, uint64_t* out[3], uint64_t* n_bits, uint64_t* n_set) {
  BGX_CHECK(c->built && !c->merge_parts.empty(), "bgx_export_mergemap: call bgx_merge_seqsets first");
  BGX_CHECK(part < c->merge_parts.size(), "bgx_export_mergemap: no such input");
  uint64_t total = 0;
  bitcount_to_host(c, c->mergemap.p + (uint64_t)part * c->mergemap_words, c->n_entries, out, &total);
  // seqset_merger.cpp:33: the mergemap of an input has as many bits set as the input has entries
  BGX_CHECK(total == c->merge_parts[part].n, "bgx_export_mergemap: mergemap bit total != entries of the input");
  if (n_bits) *n_bits = c->n_entries;
  if (n_set) *n_set = total;
}

void migrate_bits(Context* c, uint32_t part, const uint64_t* old_bits, uint64_t n_old, uint64_t* out[3], uint64_t* n_bits) {
  BGX_CHECK(c->built && !c->merge_parts.empty(), "bgx_migrate_bits: call bgx_merge_seqsets first");
  BGX_CHECK(part < c->merge_parts.size(), "bgx_migrate_bits: no such input");
  BGX_CHECK(n_old == c->merge_parts[part].n, "bgx_migrate_bits: the bit vector is not indexed by the input's entries");
  BGX_CHECK(old_bits != nullptr, "bgx_migrate_bits: bits missing");
  cudaStream_t s = c->stream;
  const uint64_t words = (n_old + 63) / 64, new_words = (c->n_entries + 63) / 64;
  DevBuf<uint64_t> d_old(words, s);
  BGX_CUDA(cudaMemcpyAsync(d_old.p, old_bits, words * 8, cudaMemcpyHostToDevice, s));
  // select table of the input's mergemap: input entry e -> merged entry (bitcount::find_count)
  const uint64_t* mm = reinterpret_cast<const uint64_t*>(c->mergemap.p) + (uint64_t)part * c->mergemap_words;
  BGX_CHECK(count_bits(c, mm, c->n_entries) == n_old, "bgx_migrate_bits: mergemap bit total != entries of the input");
  DevBuf<uint32_t> sel(n_old, s);
  select_table(c, mm, c->n_entries, sel.p, 0);
  DevBuf<unsigned long long> d_new(std::max<uint64_t>(new_words, 1), s);
  BGX_CUDA(cudaMemsetAsync(d_new.p, 0, std::max<uint64_t>(new_words, 1) * 8, s));
  KLAUNCH(migrate_kernel)<<<grid_for(words, 256), 256, 0, s>>>(d_old.p, words, n_old, sel.p, d_new.p);
  BGX_CUDA(cudaGetLastError());
  bitcount_to_host(c, d_new.p, c->n_entries, out);
  if (n_bits) *n_bits = c->n_entries;
}

void export_flat_ascii(Context* c, uint32_t part, uint64_t first, uint64_t count, char** bases, uint64_t** offs_out) {
  BGX_CHECK(c->built && !c->merge_parts.empty(), "bgx_export_flat_ascii: call bgx_merge_seqsets first");
  BGX_CHECK(part < c->merge_parts.size(), "bgx_export_flat_ascii: no such input");
  const Context::MergePart& mp = c->merge_parts[part];
  BGX_CHECK(first + count <= mp.n, "bgx_export_flat_ascii: range out of bounds");
  cudaStream_t s = c->stream;
  std::vector<uint32_t> lens(count);
  if (count) {
    DevBuf<uint32_t> d_lens(count, s);
    KLAUNCH(flat_lens_kernel)<<<grid_for(count, 256), 256, 0, s>>>(mp.flat_loc.p, first, count, d_lens.p);
    BGX_CUDA(cudaMemcpyAsync(lens.data(), d_lens.p, count * 4, cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
  }
  uint64_t* offs = (uint64_t*)host_alloc((count + 1) * 8);
  offs[0] = 0;
  for (uint64_t i = 0; i < count; ++i) offs[i + 1] = offs[i] + lens[i];
  char* out = (char*)host_alloc(std::max<uint64_t>(offs[count], 1));
  if (count) {
    DevBuf<uint64_t> d_offs(count + 1, s);
    DevBuf<char> d_out(std::max<uint64_t>(offs[count], 1), s);
    BGX_CUDA(cudaMemcpyAsync(d_offs.p, offs, (count + 1) * 8, cudaMemcpyHostToDevice, s));
    KLAUNCH(flat_ascii_kernel)<<<grid_for(count, 128), 128, 0, s>>>(c->store.p, mp.flat_loc.p, d_offs.p, first, count, d_out.p);
    BGX_CUDA(cudaMemcpyAsync(out, d_out.p, offs[count], cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
  }
  *bases = out;
  *offs_out = offs;
}

}  // namespace bgx

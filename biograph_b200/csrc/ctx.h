// ctx.h -- the bgx context: device-resident state of one seqset build on one GPU.
#pragma once

#include <chrono>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/bgx.h"
#include "dist.h"
#include "prims.cuh"

namespace bgx {

struct CountEntry {            // 16 B, one per slot: key+flags and both counters share a DRAM sector
  unsigned long long key;      // canonical k-mer | kFwdFlag | kRevFlag ; kEmptyKey when unused
  unsigned long long cnt;      // rev_count << 32 | fwd_count
};

struct StageTimer {
  cudaEvent_t a = nullptr, b = nullptr;
};

struct Context {
  bgx_options opt{};
  int device = 0;
  size_t total_mem = 0;                    // HBM of the device (read once at creation)
  cudaStream_t stream = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;  // bgx_timer_start/stop
  Dist dist;                               // multi-GPU: world size, rank, NCCL communicator

  // ---- reads (device resident) ---------------------------------------------------------
  uint64_t n_reads = 0;
  uint64_t n_words = 0;          // total 64-bit base words (32 bases each)
  uint64_t n_bases = 0;          // sum of read lengths
  uint64_t n_kmer_instances = 0; // sum max(0, len-k+1)
  bool has_n = false;
  uint32_t max_len = 0;
  DevBuf<uint64_t> words;        // n_words + 1 (pad)
  DevBuf<uint32_t> nmask;        // n_words + 1, only if has_n
  DevBuf<uint32_t> word_off;     // n_reads + 1
  DevBuf<uint16_t> lens;         // n_reads
  size_t cap_words = 0, cap_reads = 0;
  // bgx_add_reads_packed_async: the packed words travel on their own stream in chunks of reads;
  // pass 1 of counting starts on a chunk as soon as it has landed (PCIe copy under compute)
  struct UploadChunk {
    uint64_t r0, r1;   // reads [r0, r1)
    cudaEvent_t ev;    // recorded on copy_stream after the chunk's words are in place
  };
  cudaStream_t copy_stream = nullptr;
  std::vector<UploadChunk> upload;  // chunks not yet ordered before the main stream, ascending

  // ---- k-mer stage -----------------------------------------------------------------------
  bool counted = false;
  DevBuf<CountEntry> table;      // open addressing, power-of-two slots
  uint64_t table_slots = 0;
  uint64_t n_distinct = 0, n_solid = 0;
  DevBuf<unsigned long long> solid;  // hash set of solid k-mers (key|flags), power-of-two slots
  uint64_t solid_slots = 0;

  // ---- correction stage --------------------------------------------------------------------
  bool corrected = false;
  DevBuf<uint64_t> store;        // corrected base store: [fwd reads | rc reads] 2*n_words + 1 words
  DevBuf<uint64_t> gstore;       // multi-GPU: every rank's store, concatenated in rank order (replicated)
  uint64_t gstore_word_base = 0; // word offset of this rank's store inside gstore
  const uint64_t* seq_store() const { return gstore.p ? gstore.p : store.p; }
  DevBuf<uint16_t> clen;         // corrected length per read (0 = dropped)
  DevBuf<uint8_t> ncorr;         // substitutions per read
  DevBuf<uint16_t> next_fwd, next_rev;
  uint64_t n_kept = 0, kept_bases = 0, n_seeds = 0;

  // ---- seqset stage ------------------------------------------------------------------------
  bool built = false;
  uint64_t n_entries = 0;
  uint32_t max_entry_len = 0;
  DevBuf<uint64_t> ent_key, ent_loc;   // final sorted entries
  DevBuf<uint16_t> sizes, shared;
  DevBuf<uint64_t> prev_bits;          // 4 * prev_words
  DevBuf<uint64_t> prev_sub, prev_acc; // 4 * sub_words, 4 * acc_words
  uint64_t prev_words = 0, sub_words = 0, acc_words = 0;
  uint64_t fixed[5] = {0, 0, 0, 0, 0};
  uint64_t n_entries_global = 0;       // multi-GPU: entries over all ranks; this rank holds
  uint64_t first_entry_global = 0;     //   [first_entry_global, first_entry_global + n_entries)

  // ---- seqset merge (bgx_merge_seqsets; `biograph merge`) -----------------------------
  // after a merge: store = the flattened entries of every input (seqset_flat), ent_* / tables = the merged
  // seqset, and per input its flat locators (entry i -> store address, length) and its mergemap
  struct MergePart {
    uint64_t n = 0;            // entries of the input
    DevBuf<uint64_t> flat_loc; // n locators into `store`: seqset_flat::get(i)
  };
  std::vector<MergePart> merge_parts;
  DevBuf<unsigned long long> mergemap;   // merge_parts.size() bit vectors of n_entries bits, mergemap_words apart
  uint64_t mergemap_words = 0;

  // ---- stats ---------------------------------------------------------------------------------
  std::map<std::string, double> stats;       // numeric stats (ms, counts, bytes)
  std::vector<std::string> stat_order;
  void set_stat(const std::string& k, double v) {
    if (!stats.count(k)) stat_order.push_back(k);
    stats[k] = v;
  }
  void add_stat(const std::string& k, double v) {
    if (!stats.count(k)) { stat_order.push_back(k); stats[k] = 0; }
    stats[k] += v;
  }
};

// RAII CUDA-event stage timer: records on construction, on stop() synchronises and stores ms.
struct ScopedStage {
  Context* c;
  std::string name;
  cudaEvent_t a, b;
  bool done = false;
  std::chrono::steady_clock::time_point h0;
  ScopedStage(Context* c_, const std::string& n) : c(c_), name(n) {
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, c->stream);
    h0 = std::chrono::steady_clock::now();
  }
  double stop() {
    if (done) return 0;
    done = true;
    cudaEventRecord(b, c->stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    c->add_stat("ms_" + name, ms);
    // host wall clock over the same span: a gap to the device time is host-side stall
    c->add_stat("hostms_" + name,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count());
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms;
  }
  ~ScopedStage() { if (!done) { cudaEventDestroy(a); cudaEventDestroy(b); } }
};

// stage entry points (each in its own .cu)
void reads_append_ascii(Context* c, const char* bases, const uint64_t* offs, uint64_t n);
void reads_append_packed(Context* c, const uint8_t* packed, const uint32_t* n_mask, const uint64_t* word_offs,
                         const uint16_t* lens, uint64_t n, bool async = false);
void reads_append_fastq(Context* c, const char* text, uint64_t size, uint64_t* n_added);
// orders every pending upload chunk before whatever the main stream does next (and forgets them)
void reads_ready(Context* c);
void stage_count_kmers(Context* c);
// planning of the counting (pure host arithmetic, kmer.cu)
uint64_t plan_count_batches(uint64_t K_local, uint64_t K_share, int N, uint64_t total_mem, uint64_t batch_reads,
                            uint64_t n_reads);
int plan_part_bits(uint64_t K_share, uint64_t batches, int rank_bits);
void export_kmers(Context* c, uint32_t min_count, uint64_t* n, uint64_t** kmers, uint32_t** fwd, uint32_t** rev,
                  uint8_t** flags);
void stage_correct(Context* c);
void stage_seed_uncorrected(Context* c);
void export_varbit(Context* c, int which, uint64_t** words, uint64_t* n_words, uint32_t* bits, uint64_t* max_value);
void stage_build_seqset(Context* c);
// sort -> dedup -> closure -> tables on n (key, loc) records over c->store (seqset.cu).  A merge passes
// hooks: after_dedup sees the sorted records before the first dedup and where each one went (the
// mergemap), parallel_splits selects seqset_merger's placement of the prev bits (1 = the builder's).
struct MergeHooks {
  uint64_t parallel_splits = 100000;   // g_parallel_splits, modules/io/parallel.cpp:13
  std::function<void(const uint64_t* sorted_locs, uint32_t n, const uint32_t* pos, uint32_t n_kept)> after_dedup;
};
void build_seqset_from_records(Context* c, DevBuf<uint64_t>& keys, DevBuf<uint64_t>& locs, DevBuf<uint64_t>& keys_alt,
                               DevBuf<uint64_t>& locs_alt, uint32_t n, const MergeHooks* mh);
// bitcount::finalize of a device bit vector into host arrays {bits, subaccum, accum}; *total = set bits
void bitcount_to_host(Context* c, const unsigned long long* bits, uint64_t nbits, uint64_t* out[3], uint64_t* total = nullptr);
// merge.cu
void stage_merge_seqsets(Context* c, const bgx_seqset_part* parts, uint32_t n_parts, uint64_t parallel_splits);
void export_mergemap(Context* c, uint32_t part, uint64_t* out[3], uint64_t* n_bits, uint64_t* n_set);
void migrate_bits(Context* c, uint32_t part, const uint64_t* old_bits, uint64_t n_old, uint64_t* out[3], uint64_t* n_bits);
void export_flat_ascii(Context* c, uint32_t part, uint64_t first, uint64_t count, char** bases, uint64_t** offs);
void merge_release(Context* c);
void lookup_reads(Context* c, uint64_t* n_reads, uint64_t** fwd_entry, uint64_t** rc_entry);
void build_readmap(Context* c, int paired, uint64_t* n_rows, uint16_t** read_lengths, uint64_t** mate_loop_ptr,
                   uint64_t** is_forward, uint64_t* read_ids_source[3], uint64_t* read_ids_dest[3]);

}  // namespace bgx

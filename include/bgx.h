/* bgx.h -- C ABI of the B200-native seqset construction path ("bgx").
 *
 * This is the drop-in boundary for BioGraph's `biograph create` hot path
 * (modules/build_seqset + the k-mer glue in modules/bio_mapred): k-mer counting, k-mer based
 * read correction, and suffix seed/expand + sort/dedup/shared-prefix/prev-bit computation that
 * emits the seqset tables.  Every entry point names the reference interface it replaces
 * (paths relative to the reference checkout; "bs/" = modules/build_seqset/).
 *
 * Conventions
 *   - plain C linkage, plain pointers and sizes; no CUDA or torch types.
 *   - every call returns 0 on success, non-zero on error; bgx_last_error() gives the
 *     thread-local message (the reference throws io_exception / CHECK-aborts instead).
 *   - the caller owns every input buffer; the library owns every output buffer until
 *     bgx_free() (host memory) or bgx_destroy().
 *   - one context drives one GPU (one process per GPU; multi-GPU runs shard above this ABI,
 *     see DESIGN.md).  There is NO CPU fallback: without a CUDA device bgx_create fails.
 *   - base codes A=0 C=1 G=2 T=3 (modules/bio_base/dna_base.h:38-57); k-mers are uint64 with the
 *     first base in the high bits of the low 2k bits (modules/bio_base/kmer.h:30-38).
 */
#ifndef BGX_H_
#define BGX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgx_ctx bgx_ctx;

/* `biograph create` flags that reach the path (modules/biograph/biograph_create.cpp:258-335,
 * validated :483-498) plus the device to run on. */
typedef struct bgx_options {
  int32_t kmer_size;          /* --kmer-size, default 30; 16..31 (bs/kmer_counter.cpp:52-54) */
  int32_t min_kmer_count;     /* --min-kmer-count, default 5 */
  int32_t max_corrections;    /* --max-corrections, default 8; 0..32 as the CLI allows (biograph_create.cpp:486) */
  int32_t min_good_run;       /* --min-good-run, default 2 */
  float trim_after_portion;   /* --trim-after-portion, default 0.7f (parsed as float: :489-490) */
  int32_t device;             /* CUDA device ordinal */
  int32_t sort_key_bits;      /* radix-sorted key bits per round (multiple of 8, 16..64); 0 = auto from the record count */
  int32_t count_batch_reads;  /* k-mer counting: reads per batch; 0 = auto (one batch unless the k-mer
                               * instance buffers would not fit next to the table) */
} bgx_options;

#define BGX_FLAG_FWD_STARTS_READ 1u /* kmer_set::k_fwd_starts_read */
#define BGX_FLAG_REV_STARTS_READ 2u /* kmer_set::k_rev_starts_read */
#define BGX_MAX_READ_LEN 255        /* biograph_create.cpp:134-139 (reads > 255 need --allow-long-reads) */

void bgx_default_options(bgx_options* opts);
const char* bgx_last_error(void);
const char* bgx_version(void);
int bgx_device_count(void);

/* replaces: kmer_counter(count_kmer_options) + part_repo + correct_reads + expander + builder
 * construction in SEQSETMain::run (modules/biograph/biograph_create.cpp:540-779). */
int bgx_create(const bgx_options* opts, bgx_ctx** out);
void bgx_destroy(bgx_ctx* ctx);
void bgx_free(void* host_ptr);

/* replaces: prob_pass_processor::add(string_view) (bs/kmer_counter.h:297-326) called from
 * read_importer_state::process (biograph_create.cpp:119-151).  ASCII over {A,C,G,T,N}; read r is
 * bases[offs[r] .. offs[r+1]).  May be called repeatedly; reads are appended (H2D copy included).
 * Thread-compatible, not thread-safe: serialise calls on one context. */
int bgx_add_reads_ascii(bgx_ctx* ctx, const char* bases, const uint64_t* offs, uint64_t n_reads);

/* Second step beyond the path (read import): a chunk of uncompressed FASTQ text, split
 * and checked on the GPU with the semantics and error texts of fastq_reader::read
 * (modules/bio_format/fastq.cpp:40-126: four lines per record, '@' id line, sequence over ACGTN,
 * '+' line, quality line as long as the sequence), then packed like bgx_add_reads_ascii.  The text
 * must hold whole records and end with a newline; blank lines are accepted after the last record
 * only; CR LF line ends are not handled.  On any error nothing is appended, so a caller can hand the same
 * bytes to a record-by-record parser instead (bgx-create does: blank lines between records, CR LF and an
 * unterminated last line then get fastq_reader's own treatment).  *n_reads = records added.  gzip, BAM/CRAM and
 * read names/pairing stay with the importer. */
int bgx_add_reads_fastq(bgx_ctx* ctx, const char* text, uint64_t size, uint64_t* n_reads);

/* Same, for reads already 2-bit packed (T0 of the benchmark clock).
 *   packed   : dna_sequence byte order (4 bases/byte, first base in the high bits,
 *              modules/bio_base/dna_sequence.h:95-99); read r starts at byte 8*word_offs[r] and
 *              occupies ceil(lens[r]/32) 8-byte words; unused trailing bits must be zero.
 *   n_mask   : optional (NULL = no N anywhere); one uint32 per 8-byte word of `packed`; bit
 *              (31 - j%32) set means base j of the read is 'N' (its 2-bit code must be 0).
 *   word_offs: n_reads+1 entries, word_offs[n_reads] = total words; or NULL for "read 0 starts at
 *              word 0" (reads are always dense, so the offsets follow from lens; only the first
 *              and last entry are looked at).
 * Word offsets, base / k-mer totals and the length check are computed on the device: no host
 * loop over the reads.  Pinned host memory makes the copies asynchronous; pageable works too.
 * packed / n_mask / lens may also be DEVICE pointers (reads produced on the GPU, e.g. by a
 * generator or an upstream decoder): the copies are cudaMemcpyDefault; word_offs is read on the
 * host and must then be NULL. */
int bgx_add_reads_packed(bgx_ctx* ctx, const uint8_t* packed, const uint32_t* n_mask,
                         const uint64_t* word_offs, const uint16_t* lens, uint64_t n_reads);

/* Same arguments and result as bgx_add_reads_packed, but the packed words are copied on a second
 * stream in chunks of reads and the call returns without waiting for them: pass 1 of
 * bgx_count_kmers starts on each chunk as it lands, so the PCIe copy runs under compute.
 * `packed` must stay valid and unchanged until the next bgx_count_kmers / bgx_run / bgx_correct /
 * bgx_clear_reads / bgx_destroy returns.  With an N mask, or for small appends, it is the synchronous call. */
int bgx_add_reads_packed_async(bgx_ctx* ctx, const uint8_t* packed, const uint32_t* n_mask,
                               const uint64_t* word_offs, const uint16_t* lens, uint64_t n_reads);

/* replaces: kmer_counter::close_prob_pass + run_kmerize_subtask (bs/kmer_counter.cpp:234-404,
 * modules/bio_mapred/kmerize_bf.h:81-84): exact canonical k-mer counts with fwd/rev counts and
 * starts-read flags, min-count filter, and the device k-mer set used by correction. */
int bgx_count_kmers(bgx_ctx* ctx);

/* Parity hook + histogram feed (kmer_counter::extract_exact_counts, kmerize_bf.cpp:353-429).
 * All k-mers with fwd+rev >= min_count, ascending; flags use BGX_FLAG_*.  Arrays are
 * bgx_free()'d by the caller. */
int bgx_export_kmers(bgx_ctx* ctx, uint32_t min_count, uint64_t* n, uint64_t** kmers,
                     uint32_t** fwd_counts, uint32_t** rev_counts, uint8_t** flags);

/* The resident reads back as ASCII ('N' where the mask says so), in append order, concatenated;
 * lens[r] = length of read r.  Serves the facade's correct(const unaligned_read&, corrected_read&)
 * (bs/correct_reads.h:14-22), which answers per read by sequence.  Arrays are bgx_free()'d by the caller. */
int bgx_export_reads(bgx_ctx* ctx, uint64_t* n_reads, uint16_t** lens, char** bases, uint64_t* n_bases);

/* replaces: correct_reads::correct over all reads (bs/correct_reads.cpp:154-231) including
 * fast_read_correct (modules/bio_base/fast_read_correct.cpp) and the seed counts. */
int bgx_correct(bgx_ctx* ctx);

/* replaces: the seeding of seqset_for_reads (modules/bio_base/seqset_testutil.cpp:19-59:
 * part_repo::write(read, 1, 1) for every read, no correction) -- "the simplest whole-stage-3
 * operator" (SURVEY 8b).  Treats the resident reads as already corrected: every read and its
 * reverse complement seed exactly one suffix record.  Reads must not contain 'N'.  Replaces
 * bgx_count_kmers + bgx_correct; follow with bgx_build_seqset. */
int bgx_seed_uncorrected(bgx_ctx* ctx);

/* Parity hook + make_readmap input (corrected_reads kv sink, biograph_create.cpp:868-893).
 * lens[r] = corrected length, 0 if the read was dropped; bases = ASCII of kept reads
 * concatenated in input order; corrections[r], next_fwd[r], next_rev[r] as in
 * bs/correct_reads.cpp:195-210.  Any output pointer may be NULL. */
int bgx_export_corrected(bgx_ctx* ctx, uint64_t* n_reads, uint16_t** lens, char** bases,
                         uint64_t* n_bases, uint8_t** corrections, uint16_t** next_fwd,
                         uint16_t** next_rev);

/* replaces: expander::sort_and_dedup / expander::expand rounds + builder::build_chunks
 * (biograph_create.cpp:914-931, bs/expand.cpp, bs/builder.cpp:8-164). */
int bgx_build_seqset(bgx_ctx* ctx);

/* replaces: builder::make_seqset + seqset::finalize (bs/builder.cpp:207-263,
 * modules/bio_base/seqset.cpp:113-129).  sizes/shared are uint16 per entry; prev_bits[b] is the
 * bitcount `bits` member of prev_<b> (ceil(n/64) uint64 words, bit i at word[i/64]>>(i&63));
 * prev_subaccum / prev_accum are the bitcount index members (modules/io/bitcount.cpp:84-123);
 * fixed[5] as in seqset::finalize.  Any output pointer may be NULL. */
int bgx_export_seqset(bgx_ctx* ctx, uint64_t* n_entries, uint32_t* max_entry_len, uint16_t** sizes,
                      uint16_t** shared, uint64_t* prev_bits[4], uint64_t* prev_subaccum[4],
                      uint64_t* prev_accum[4], uint64_t fixed[5]);

/* replaces: mutable_packed_varbit_vector (modules/io/packed_varbit_vector.cpp:174-228), the
 * `elements` member of entry_sizes (which = 0, max_value = max entry length) and shared
 * (which = 1, max_value = max entry length - 1): values packed LSB-first into little-endian uint64
 * words at bits_per_value = bit_length(max_value).  Multi-GPU: each rank packs its own range;
 * ranges are multiples of 512 entries, so the word arrays concatenate. */
int bgx_export_varbit(bgx_ctx* ctx, int32_t which, uint64_t** words, uint64_t* n_words, uint32_t* bits_per_value,
                      uint64_t* max_value);

/* First step beyond the path: the entry lookups of make_readmap -- for every read,
 * in input order, the seqset entry id of its corrected sequence and of its reverse complement
 * (seqset::find_existing_unique as called from parallel_mate_loop_table_builder::operator(),
 * modules/bio_mapred/make_readmap.cpp:137-167; modules/bio_base/seqset.cpp:173-188), UINT64_MAX for
 * a dropped read.  The mate-loop table and the readmap file are not built.  Single GPU only.
 * Arrays are bgx_free()'d by the caller. */
int bgx_lookup_reads(bgx_ctx* ctx, uint64_t* n_reads, uint64_t** fwd_entry, uint64_t** rc_entry);

/* replaces: make_readmap::create_from_reads (modules/bio_mapred/make_readmap.cpp:229-362; rows and their
 * order make_readmap.h:53-76,187-205; sparse_multi_builder modules/io/sparse_multi.cpp:90-113).
 * paired = 0: every read is a record (two rows: the read, its reverse complement).
 * paired != 0: reads 2i and 2i+1 are mates (four rows; a pair with one read dropped is a single
 *   read; the read with the smaller sequence starts the loop, make_readmap.cpp:170-175).
 * Rows sorted by (entry id, type, length, mate length, loop entry):
 *   read_lengths[i], mate_loop_ptr[i] (next row of the read's mate loop: read -> its reverse
 *   complement [-> mate -> mate's reverse complement] -> read, rows of identical reads handed
 *   out in the order of the reference's claim pass, :302-360), is_forward (bit i, uint64 words
 *   LSB-first), and the two bitcount vectors of `read_ids` as {bits, subaccum, accum}:
 *   source_to_mid over the seqset entries, dest_to_mid over the rows.
 * These are the payload members of the readmap spiral file (read_lengths and mate_loop_ptr go
 * through packed_varbit_vector on the host).  Single GPU only.  Parity: the unpaired form is
 * pinned to the reference's golden readmap; the paired form to a transcription of the reference
 * code (no reference-built paired readmap in the current row order exists in its tree). */
int bgx_build_readmap(bgx_ctx* ctx, int32_t paired, uint64_t* n_rows, uint16_t** read_lengths, uint64_t** mate_loop_ptr,
                      uint64_t** is_forward, uint64_t* read_ids_source[3], uint64_t* read_ids_dest[3]);

/* Debug/parity hook: entry i as ASCII (entries are <= BGX_MAX_READ_LEN bases). */
int bgx_export_entries_ascii(bgx_ctx* ctx, uint64_t first, uint64_t count, char** bases,
                             uint64_t** offs);

/* ---- seqset merge (`biograph merge`, modules/biograph/biograph_merge.cpp:199-290) ------
 * One input seqset as its file members hold it: entry sizes and the four prev bit vectors (bitcount
 * `bits` layout, ceil(n_entries/64) words each; `fixed` follows from their popcounts as in
 * seqset::finalize, modules/bio_base/seqset.cpp:113-129, and is re-derived here). */
typedef struct bgx_seqset_part {
  uint64_t n_entries;
  const uint16_t* sizes;        /* n_entries */
  const uint64_t* prev_bits[4]; /* prev_A .. prev_T */
} bgx_seqset_part;

/* replaces, in one call: seqset_flat_builder::build for every input (modules/bio_base/
 * seqset_flat.cpp:232-290), make_mergemap::build + fill_mergemap (make_mergemap.cpp:22-44,188-259) and
 * seqset_merger::build (seqset_merger.cpp:53-78,109-197).  Every input is flattened on the GPU (entry
 * sequences rebuilt from the prev bits by pointer doubling), the union of all entries is sorted and
 * prefix-deduplicated by the kernels bgx_build_seqset uses, and the tables of the merged seqset are
 * computed; the context is then in the "built" state: bgx_export_seqset / bgx_export_varbit /
 * bgx_export_entries_ascii return the merged seqset.
 * parallel_splits: seqset_merger runs merge_range over generate_chunks(0, n, g_parallel_splits)
 *   (modules/io/parallel.cpp:13,60-83) and the chunking decides on which entry of a run of entries with
 *   a common prefix a prev bit lands; 0 = the reference's 100000 (byte-identical prev members to
 *   `biograph merge`), 1 = builder::build_chunks' placement (what `biograph create` would write for the
 *   union of the reads).  Entry set, sizes, shared and fixed do not depend on it.
 * Single GPU; at most 64 inputs; fewer than 2^31 entries in total.  An input that is not a seqset
 * (prev bit totals != entries, a size of 0, not closed under pop_front) is an error. */
int bgx_merge_seqsets(bgx_ctx* ctx, const bgx_seqset_part* parts, uint32_t n_parts, uint64_t parallel_splits);

/* replaces: seqset_mergemap_builder / make_mergemap::fill_mergemap for input `part`: the `merged_entries`
 * bitcount (n_bits = merged entries; bit x set iff merged entry x, or a prefix of it, is an entry of the
 * input; input entry i is merged entry find_count(i)), as {bits, subaccum, accum}.  *n_set = set bits =
 * entries of the input (seqset_merger.cpp:33).  Arrays are bgx_free()'d by the caller. */
int bgx_export_mergemap(bgx_ctx* ctx, uint32_t part, uint64_t* merged_entries[3], uint64_t* n_bits, uint64_t* n_set);

/* replaces: the read_ids part of make_readmap::fast_migrate (modules/bio_mapred/make_readmap.cpp:459-486):
 * a bit vector indexed by the entries of input `part` (a readmap's read_ids/source_to_mid `bits`, n_old
 * = entries of the input) re-targeted to merged entry ids, as {bits, subaccum, accum} of n_bits = merged
 * entries.  Every other readmap member is copied unchanged by fast_migrate (:487-520). */
int bgx_migrate_bits(bgx_ctx* ctx, uint32_t part, const uint64_t* old_bits, uint64_t n_old, uint64_t* migrated[3],
                     uint64_t* n_bits);

/* replaces: seqset_flat::get(i) (modules/bio_base/seqset_flat.h:117-140) for entries [first, first+count)
 * of input `part` after bgx_merge_seqsets (a merge of ONE input is a plain flatten): ASCII, concatenated,
 * entry first+i = bases[offs[i] .. offs[i+1]). */
int bgx_export_flat_ascii(bgx_ctx* ctx, uint32_t part, uint64_t first, uint64_t count, char** bases, uint64_t** offs);

/* ---- multi-GPU (no reference analogue: the reference is one process) -------------------
 * One context per GPU, driven by one process per GPU or by one host thread per GPU of a single
 * process (bgx_bs::multi_session).  Rank 0 obtains an id with bgx_dist_unique_id and hands
 * it to every rank out of band (torch.distributed broadcast, MPI, a file); every rank then calls
 * bgx_dist_init before adding its share of the reads.  From then on bgx_count_kmers, bgx_correct,
 * bgx_build_seqset, bgx_run and bgx_destroy are COLLECTIVE: every rank must call them.  K-mer
 * instances travel to the owner of their hash partition, suffix records to the owner of their prefix
 * range, through peer-mapped device memory (CUDA IPC / peer access over NVLink, an SM copy kernel;
 * NCCL for the all-gathers and small collectives); each rank ends up with a contiguous range of the final seqset whose
 * length is a multiple of 512 entries (except the last), so the per-rank tables returned by
 * bgx_export_seqset concatenate in rank order into exactly the single-GPU tables.
 * bgx_export_kmers returns the k-mers this rank owns; bgx_export_corrected its own reads. */
int bgx_dist_unique_id(uint8_t id[128]);
int bgx_dist_init(bgx_ctx* ctx, int32_t world_size, int32_t rank, const uint8_t id[128]);
/* layout[0] entries held by this rank, [1] entries over all ranks, [2] global index of this rank's
 * first entry, [3..5] number of uint64 words per base in prev_bits / prev_subaccum / prev_accum as
 * returned by bgx_export_seqset on this rank. */
int bgx_seqset_layout(bgx_ctx* ctx, uint64_t layout[6]);

/* Whole path on resident reads: count -> correct -> seqset.  Equivalent to the three calls. */
int bgx_run(bgx_ctx* ctx);

/* Drops everything derived from the reads (tables, corrected reads, seqset) but keeps the
 * uploaded reads resident, so the path can be re-run (benchmark loop). */
int bgx_reset_results(bgx_ctx* ctx);
/* Drops the reads too. */
int bgx_clear_reads(bgx_ctx* ctx);

/* replaces: runtime_stats / SPLOG stage counters (modules/io/runtime_stats.h, qc/create_stats.json).
 * Writes a JSON object (stage times in ms measured with CUDA events, counts, algorithmic bytes
 * per kernel family) into buf; returns non-zero if cap is too small. */
int bgx_stats_json(bgx_ctx* ctx, char* buf, size_t cap);

/* Measurement hooks (no reference analogue).  bgx_timer_start/stop bracket a region with CUDA
 * events on the context's own stream (the stream every kernel and copy of this library is issued
 * on), so host-side gaps between enqueues are inside the measured span.  bgx_launch_count is the
 * process-wide number of kernels this library has launched. */
int bgx_timer_start(bgx_ctx* ctx);
int bgx_timer_stop(bgx_ctx* ctx, double* elapsed_ms);
uint64_t bgx_launch_count(void);

/* Host-only test hooks (no GPU needed): the bijective k-mer hash of the counting passes and its
 * inverse, and the plan of a count (hash-range batches, log2 of the hash partitions per batch) for
 * k_local instances on this rank, k_share owned per rank, n_ranks GPUs of total_mem bytes each. */
uint64_t bgx_debug_khash(uint64_t x, int32_t k, int32_t inverse);
void bgx_debug_count_plan(uint64_t k_local, uint64_t k_share, int32_t n_ranks, uint64_t total_mem,
                          uint64_t batch_reads, uint64_t n_reads, uint64_t* batches, int32_t* part_bits);

/* Test hook for the device-wide primitives (no reference analogue): stable LSD radix sort of
 * n (key, value) pairs held in HOST arrays on key bits [begin_bit, end_bit), in place, run by the
 * same kernels the seqset stage uses. */
int bgx_debug_sort_pairs(bgx_ctx* ctx, uint64_t* keys, uint64_t* vals, uint64_t n, int begin_bit, int end_bit);

#ifdef __cplusplus
}
#endif
#endif /* BGX_H_ */

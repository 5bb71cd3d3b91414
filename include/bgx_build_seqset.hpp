// bgx_build_seqset.hpp -- C++ host façade over the bgx C ABI (include/bgx.h) with the class and
// method names of the reference's build_seqset stages, so that SEQSETMain
// (modules/biograph/biograph_create.cpp:431-950) keeps its call sites.  Header-only, C++17, no
// CUDA types: link with -lbgx.  INTEGRATION.md shows the edits to biograph_create.cpp.
//
//   reference (CPU, threads + temp files)                      this façade (one GPU per context)
//   --------------------------------------------------------   -----------------------------------------
//   build_seqset::kmer_counter + prob_pass_processor::add      kmer_counter::prob_pass_processor::add
//     (bs/kmer_counter.h:123-155, 297-326)                       -> bgx_add_reads_ascii (batched)
//   run_kmerize_subtask (bio_mapred/kmerize_bf.h:81-84)        run_kmerize_subtask -> bgx_count_kmers
//   kmer_set (bio_mapred/kmer_set.h)                           kmer_set (sorted canonical k-mers + flags)
//   build_seqset::correct_reads::correct                       correct_reads::correct_all -> bgx_correct
//     (bs/correct_reads.h:14-22, .cpp:154-231)
//   build_seqset::expander::sort_and_dedup / expand            expander (rounds collapse into one closure
//     (bs/expand.h:9-46)                                         walk on the GPU; counts are reported)
//   build_seqset::builder::build_chunks / make_seqset          builder -> bgx_build_seqset / bgx_export_*
//     (bs/builder.h:9-15)
//   seqset_for_reads (bio_base/seqset_testutil.h:13)           seqset_for_reads
//   spiral_file_create_mmap + seqset ctor/finalize             seqset_file_writer (stored ZIP64, members in the
//     (io/spiral_file_mmap.cpp, bio_base/seqset.cpp:19-44)       reference's order and byte layout)
//
// Errors: the reference throws io_exception; every failing C-ABI call is rethrown here as
// bgx_bs::io_exception carrying bgx_last_error().
#pragma once

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "bgx.h"

namespace bgx_bs {

using kmer_t = uint64_t;                                   // modules/bio_base/dna_sequence.h:13
using progress_handler_t = std::function<void(double)>;    // modules/io/progress.h
inline void null_progress_handler(double) {}

struct io_exception : std::runtime_error {                 // modules/io/io.h
  using std::runtime_error::runtime_error;
};

namespace detail {
inline void ck(int rc) {
  if (rc) throw io_exception(bgx_last_error());
}
template <typename T>
struct host_array {  // library-owned host buffer, released with bgx_free
  T* p = nullptr;
  uint64_t n = 0;
  host_array() = default;
  host_array(const host_array&) = delete;
  host_array& operator=(const host_array&) = delete;
  host_array(host_array&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; }
  host_array& operator=(host_array&& o) noexcept {
    if (this != &o) { bgx_free(p); p = o.p; n = o.n; o.p = nullptr; }
    return *this;
  }
  ~host_array() { bgx_free(p); }
  const T& operator[](uint64_t i) const { return p[i]; }
};
}  // namespace detail

// `biograph create` flags that reach the path (biograph_create.cpp:258-335); defaults as the CLI
struct count_kmer_options {        // bs/kmer_counter.h: count_kmer_options (subset that affects results)
  unsigned kmer_size = 30;
  unsigned min_count = 5;          // kmerize_bf_params::min_count
  int device = 0;
};

struct read_correction_params {    // modules/bio_mapred/read_correction.h:9-60 (fields the fast path reads)
  float trim_after_portion = 0.7f;
  unsigned frc_max_corrections = 8;
  unsigned frc_min_good_run = 2;
};

// One build = one GPU context, shared by the stage objects below (the reference shares a
// part_repo and a kmer_set between its stages the same way).
class session {
 public:
  explicit session(const count_kmer_options& ko = count_kmer_options(),
                   const read_correction_params& rp = read_correction_params()) {
    bgx_options o;
    bgx_default_options(&o);
    o.kmer_size = (int32_t)ko.kmer_size;
    o.min_kmer_count = (int32_t)ko.min_count;
    o.max_corrections = (int32_t)rp.frc_max_corrections;
    o.min_good_run = (int32_t)rp.frc_min_good_run;
    o.trim_after_portion = rp.trim_after_portion;
    o.device = ko.device;
    m_ko = ko;
    detail::ck(bgx_create(&o, &m_ctx));
  }
  const count_kmer_options& kmer_options() const { return m_ko; }
  ~session() { bgx_destroy(m_ctx); }
  session(const session&) = delete;
  session& operator=(const session&) = delete;
  bgx_ctx* ctx() const { return m_ctx; }
  // multi-GPU: one session per rank; see bgx.h
  static std::vector<uint8_t> unique_id() {
    std::vector<uint8_t> id(128);
    detail::ck(bgx_dist_unique_id(id.data()));
    return id;
  }
  void dist_init(int world_size, int rank, const std::vector<uint8_t>& id) {
    detail::ck(bgx_dist_init(m_ctx, world_size, rank, id.data()));
  }
  std::mutex& add_mutex() { return m_add_mu; }
  // the seqset stage runs once per session, whichever stage object asks first
  void ensure_seqset_built() {
    if (!m_built) { detail::ck(bgx_build_seqset(m_ctx)); m_built = true; }
  }

 private:
  bool m_built = false;
  bgx_ctx* m_ctx = nullptr;
  count_kmer_options m_ko;
  std::mutex m_add_mu;
};

// One build over several GPUs of ONE process -- what a single-process caller like SEQSETMain needs
// (modules/main/main.cpp:255-261 forks nothing but a logger): one session per device, each driven by its
// own host thread, joined into one sharded build (NCCL communicator + peer mappings inside the library).
//   multi_session ms(4, ko, rp);
//   ms.parallel([&](int rank, session& s) { /* add this rank's share of the reads, run the stages */ });
// Every stage call is collective: all ranks must make the same calls in the same order.
class multi_session {
 public:
  multi_session(int n_gpus, const count_kmer_options& ko = count_kmer_options(),
                const read_correction_params& rp = read_correction_params(), const std::vector<int>& devices = {}) {
    if (n_gpus < 1 || (n_gpus & (n_gpus - 1))) throw io_exception("multi_session: the GPU count must be a power of two");
    for (int r = 0; r < n_gpus; ++r) {
      count_kmer_options k = ko;
      k.device = devices.empty() ? r : devices[(size_t)r];
      m_s.emplace_back(new session(k, rp));
    }
    if (n_gpus > 1) {
      const std::vector<uint8_t> id = session::unique_id();
      parallel([&](int rank, session& s) { s.dist_init(n_gpus, rank, id); });   // ncclCommInitRank: all ranks at once
    }
  }
  // destroying a sharded context is collective (bgx.h): every rank goes away on its own thread
  ~multi_session() {
    std::vector<std::thread> th;
    for (auto& s : m_s) th.emplace_back([&s] { s.reset(); });
    for (auto& t : th) t.join();
  }
  int size() const { return (int)m_s.size(); }
  session& rank(int r) { return *m_s[(size_t)r]; }
  // f(rank, session) on every rank, each from its own host thread; the first exception is rethrown
  void parallel(const std::function<void(int, session&)>& f) {
    std::vector<std::thread> th;
    std::vector<std::string> err(m_s.size());
    for (int r = 0; r < size(); ++r)
      th.emplace_back([&, r] {
        try { f(r, *m_s[(size_t)r]); } catch (const std::exception& e) { err[(size_t)r] = e.what()[0] ? e.what() : "error"; }
      });
    for (auto& t : th) t.join();
    for (const std::string& e : err)
      if (!e.empty()) throw io_exception(e);
  }

 private:
  std::vector<std::unique_ptr<session>> m_s;
};

// ---- k-mer counting ---------------------------------------------------------------------------------
class kmer_counter {
 public:
  struct element {  // bs/kmer_counter.h:127-137
    kmer_t kmer = ~kmer_t(0);
    uint32_t fwd_count = 0, rev_count = 0;
    bool fwd_starts_read = false, rev_starts_read = false;
  };
  explicit kmer_counter(session& s) : m_s(s) {}
  void start_prob_pass() {}   // no probabilistic pre-pass on the GPU (result-invisible)
  void close_prob_pass() {}
  // Per-thread processor (bs/kmer_counter.h:349-357).  add() buffers reads and hands full batches
  // to the device; many processors may run concurrently on one counter, one add() at a time each.
  class prob_pass_processor {
   public:
    explicit prob_pass_processor(kmer_counter& k, size_t batch_bases = size_t(64) << 20) : m_k(k), m_batch(batch_bases) {
      m_offs.push_back(0);
    }
    // a destructor must not throw (an append error here would be std::terminate): callers that care
    // flush explicitly first, as seqset_for_reads and bgx-create do (bs/kmer_counter.cpp:214-223)
    ~prob_pass_processor() {
      try { flush_all(); } catch (...) {}
    }
    void add(const char* seq, size_t len) {  // the reference takes a string_view
      m_bases.append(seq, len);
      m_offs.push_back(m_bases.size());
      if (m_bases.size() >= m_batch) flush_all();
    }
    void add(const std::string& seq) { add(seq.data(), seq.size()); }
    void flush_all() {  // bs/kmer_counter.cpp:214-223
      if (m_offs.size() > 1) {
        std::lock_guard<std::mutex> l(m_k.m_s.add_mutex());  // the C ABI serialises appends per context
        detail::ck(bgx_add_reads_ascii(m_k.m_s.ctx(), m_bases.data(), m_offs.data(), m_offs.size() - 1));
      }
      m_bases.clear();
      m_offs.assign(1, 0);
    }

   private:
    kmer_counter& m_k;
    size_t m_batch;
    std::string m_bases;
    std::vector<uint64_t> m_offs;
  };
  // kmer_counter::extract_exact_counts (bs/kmer_counter.h:110-120): every k-mer with
  // fwd+rev >= min_count, ascending
  void extract_exact_counts(const std::function<void(const element*, const element*)>& output_f, uint32_t min_count = 1) {
    uint64_t n = 0;
    uint64_t* k = nullptr;
    uint32_t *f = nullptr, *r = nullptr;
    uint8_t* fl = nullptr;
    detail::ck(bgx_export_kmers(m_s.ctx(), min_count, &n, &k, &f, &r, &fl));
    std::vector<element> el(n);
    for (uint64_t i = 0; i < n; ++i) {
      el[i].kmer = k[i];
      el[i].fwd_count = f[i];
      el[i].rev_count = r[i];
      el[i].fwd_starts_read = fl[i] & BGX_FLAG_FWD_STARTS_READ;
      el[i].rev_starts_read = fl[i] & BGX_FLAG_REV_STARTS_READ;
    }
    bgx_free(k); bgx_free(f); bgx_free(r); bgx_free(fl);
    output_f(el.data(), el.data() + n);
  }
  void close() {}
  session& sess() { return m_s; }

 private:
  session& m_s;
};

// kmer_set (modules/bio_mapred/kmer_set.h:14-186): the sorted solid k-mers; index = rank
class kmer_set {
 public:
  static constexpr unsigned k_fwd_starts_read = BGX_FLAG_FWD_STARTS_READ;
  static constexpr unsigned k_rev_starts_read = BGX_FLAG_REV_STARTS_READ;
  size_t size() const { return m_kmers.size(); }
  unsigned kmer_size() const { return m_kmer_size; }
  static constexpr size_t k_not_present = ~size_t(0);
  size_t find_table_index(kmer_t canonical) const {  // kmer_set.cpp:296-360
    auto it = std::lower_bound(m_kmers.begin(), m_kmers.end(), canonical);
    return (it != m_kmers.end() && *it == canonical) ? size_t(it - m_kmers.begin()) : k_not_present;
  }
  unsigned get_flags(size_t index) const { return m_flags[index]; }
  kmer_t operator[](size_t index) const { return m_kmers[index]; }

 private:
  friend std::unique_ptr<kmer_set> run_kmerize_subtask(kmer_counter*, progress_handler_t);
  std::vector<kmer_t> m_kmers;
  std::vector<uint8_t> m_flags;
  unsigned m_kmer_size = 0;
};

// run_kmerize_subtask (modules/bio_mapred/kmerize_bf.h:81-84): exact counts, min-count filter,
// k-mer set.  (The reference also returns count/histogram manifests for the QC report; the
// histogram can be rebuilt from extract_exact_counts.)
inline std::unique_ptr<kmer_set> run_kmerize_subtask(kmer_counter* counter, progress_handler_t progress = null_progress_handler) {
  bgx_ctx* c = counter->sess().ctx();
  detail::ck(bgx_count_kmers(c));
  progress(0.9);
  std::unique_ptr<kmer_set> ks(new kmer_set());
  ks->m_kmer_size = counter->sess().kmer_options().kmer_size;
  uint64_t n = 0;
  uint64_t* k = nullptr;
  uint32_t *f = nullptr, *r = nullptr;
  uint8_t* fl = nullptr;
  // kmer_passes (kmerize_bf.cpp:290-318): fwd + rev >= min_count
  detail::ck(bgx_export_kmers(c, counter->sess().kmer_options().min_count, &n, &k, &f, &r, &fl));
  ks->m_kmers.assign(k, k + n);
  ks->m_flags.assign(fl, fl + n);
  bgx_free(k); bgx_free(f); bgx_free(r); bgx_free(fl);
  progress(1.0);
  return ks;
}

// The reference's own signature (modules/bio_mapred/kmerize_bf.h:81-84):
//   pair<unique_ptr<kmer_set>, vector<manifest>> run_kmerize_subtask(kmerize_bf_params, manifest reads,
//                                                                    kmer_counter*, progress)
// The reads are already resident (they arrived through prob_pass_processor::add), so the reads manifest
// is only carried along; the three result manifests -- k-mer counts, count histogram, overrepresented
// k-mers (kmerize_bf.cpp:444-473) -- come back as in-memory tables.
struct kmerize_bf_params {             // modules/bio_mapred/kmerize_bf.h:8-48 (fields that reach this path)
  size_t kmer_size = 0;
  size_t min_count = 4;
  size_t ref_size = 0, memory_bound = 0, num_threads = 0;
  float skew_cutoff = 0.0f;            // 0.0: the skew filter is off (kmerize_bf.h:38)
  size_t overrep = 0;                  // 0: overrepresentation filtering is off
};
struct manifest {                      // modules/io/manifest.h, reduced to what the stage hands on
  std::string tag;
  size_t num_records = 0;
  std::vector<std::pair<uint64_t, uint64_t>> records;   // histogram: (count, number of k-mers)
};
inline std::pair<std::unique_ptr<kmer_set>, std::vector<manifest>> run_kmerize_subtask(
    const kmerize_bf_params& params, const manifest& /*reads*/, kmer_counter* counter,
    progress_handler_t progress = null_progress_handler) {
  if (params.kmer_size && params.kmer_size != counter->sess().kmer_options().kmer_size)
    throw io_exception("run_kmerize_subtask: kmer_size differs from the counter's");
  if (params.min_count != counter->sess().kmer_options().min_count)
    throw io_exception("run_kmerize_subtask: min_count differs from the counter's");
  std::unique_ptr<kmer_set> ks = run_kmerize_subtask(counter, progress);
  manifest counts, hist, overrep;
  counts.tag = "kmers";
  counts.num_records = ks->size();
  hist.tag = "kmer_histogram";
  overrep.tag = "overrep";
  // the histogram of fwd + rev over the k-mers that passed (kmer_quality_report.html's input)
  std::vector<uint64_t> h;
  counter->extract_exact_counts(
      [&](const kmer_counter::element* b, const kmer_counter::element* e) {
        for (; b != e; ++b) {
          const uint64_t t = (uint64_t)b->fwd_count + b->rev_count;
          if (t >= h.size()) h.resize(std::min<uint64_t>(t, 1u << 20) + 1);
          ++h[std::min<uint64_t>(t, h.size() - 1)];
        }
      },
      (uint32_t)params.min_count);
  for (uint64_t c = 0; c < h.size(); ++c)
    if (h[c]) hist.records.emplace_back(c, h[c]);
  hist.num_records = hist.records.size();
  std::vector<manifest> out;
  out.push_back(std::move(counts));
  out.push_back(std::move(hist));
  out.push_back(std::move(overrep));
  return std::make_pair(std::move(ks), std::move(out));
}

// ---- read correction ------------------------------------------------------------------------------------
struct corrected_read {  // modules/bio_base/corrected_read.h (fields this path fills)
  std::string corrected;   // empty = read dropped
  unsigned corrections = 0;
};
struct unaligned_read {  // modules/bio_base/unaligned_read.h:26-49 (the field correction reads)
  int pair_number = 0;
  std::string sequence;
};

class correct_reads {
 public:
  correct_reads(session& s, kmer_set& /*ks*/, const read_correction_params& /*params: in the session*/) : m_s(s) {}
  void add_initial_repo(progress_handler_t = null_progress_handler) {}  // compression only
  // All resident reads at once; replaces the parallel_for over temp read files
  // (biograph_create.cpp:858-903).  Afterwards correct(i, cr) reads back read i.
  void correct_all() {
    detail::ck(bgx_correct(m_s.ctx()));
    uint64_t n = 0, nb = 0;
    uint16_t* lens = nullptr;
    char* bases = nullptr;
    uint8_t* corr = nullptr;
    detail::ck(bgx_export_corrected(m_s.ctx(), &n, &lens, &bases, &nb, &corr, nullptr, nullptr));
    m_lens.assign(lens, lens + n);
    m_corr.assign(corr, corr + n);
    m_bases.assign(bases, bases + nb);
    m_offs.assign(n + 1, 0);
    for (uint64_t i = 0; i < n; ++i) m_offs[i + 1] = m_offs[i] + lens[i];
    bgx_free(lens); bgx_free(bases); bgx_free(corr);
  }
  size_t size() const { return m_lens.size(); }
  unsigned corrected_length(size_t read_index) const { return m_lens[read_index]; }   // 0 = dropped
  // the result of resident read number read_index (append order): false = dropped
  bool correct(size_t read_index, corrected_read& cr) const {
    cr.corrected.assign(m_bases.data() + m_offs[read_index], m_lens[read_index]);
    cr.corrections = m_corr[read_index];
    return m_lens[read_index] != 0;
  }
  // The reference's own signature (bs/correct_reads.h:14-22): thread-safe, called from a parallel_for
  // over the read files (biograph_create.cpp:858-903).  Correction is a pure function of the read's
  // bases (given the k-mer set and the parameters), so the answer for `r` is the answer of the resident
  // read with the same sequence: the first call runs correct_all() if nobody has, and indexes the
  // resident reads by sequence.  A read that was never added throws.
  bool correct(const unaligned_read& r, corrected_read& cr) {
    {
      std::lock_guard<std::mutex> l(m_mu);
      if (m_lens.empty()) correct_all();
      if (m_by_seq.empty() && !m_lens.empty()) index_inputs();
    }
    auto it = m_by_seq.find(r.sequence);
    if (it == m_by_seq.end()) throw io_exception("correct_reads::correct: read was not added to the k-mer counter");
    return correct(it->second, cr);
  }

 private:
  void index_inputs() {
    uint64_t n = 0, nb = 0;
    uint16_t* lens = nullptr;
    char* bases = nullptr;
    detail::ck(bgx_export_reads(m_s.ctx(), &n, &lens, &bases, &nb));
    uint64_t off = 0;
    for (uint64_t i = 0; i < n; ++i) {
      m_by_seq.emplace(std::string(bases + off, lens[i]), (size_t)i);
      off += lens[i];
    }
    bgx_free(lens);
    bgx_free(bases);
  }
  session& m_s;
  std::mutex m_mu;
  std::vector<uint16_t> m_lens;
  std::vector<uint8_t> m_corr;
  std::vector<uint64_t> m_offs;
  std::string m_bases;
  std::unordered_map<std::string, size_t> m_by_seq;
};

// ---- expand / sort / dedup ---------------------------------------------------------------------------------
// The reference runs three sort_and_dedup rounds with stride-7/255 and stride-1/6 expansions in
// between (biograph_create.cpp:921-931); they are an I/O schedule for temp files, and the final
// set is their fixed point.  The GPU build reaches it with one closure walk, so the first call
// does the whole build and every call reports the matching count.
class expander {
 public:
  expander(session& s, bool /*keep_tmp*/) : m_s(s) {}
  size_t sort_and_dedup(const std::string& /*already_sorted_pass*/, const std::string& /*new_entries_pass*/,
                        const std::string& result_sorted_pass, const std::string& /*result_expanded_pass*/,
                        unsigned /*stride*/, unsigned /*count*/, progress_handler_t progress = null_progress_handler) {
    ensure_built();
    progress(1.0);
    return result_sorted_pass == "init_sorted" ? stat("entries_round1") : stat("entries");
  }
  size_t expand(const std::string&, const std::string&, unsigned, unsigned, progress_handler_t progress = null_progress_handler) {
    ensure_built();
    progress(1.0);
    return stat("walk_new_records");
  }

 private:
  void ensure_built() { m_s.ensure_seqset_built(); }
  size_t stat(const char* key) {
    char buf[1 << 16];
    detail::ck(bgx_stats_json(m_s.ctx(), buf, sizeof(buf)));
    std::string pat = std::string("\"") + key + "\":";
    const char* p = strstr(buf, pat.c_str());
    return p ? (size_t)strtod(p + pat.size(), nullptr) : 0;
  }
  session& m_s;
};

// ---- output: the seqset tables and the spiral file ---------------------------------------------------------
struct seqset_tables {  // what seqset's members hold (modules/bio_base/seqset.cpp:19-44)
  uint64_t num_entries = 0;
  uint32_t max_entry_len = 0;
  uint64_t fixed[5] = {0, 0, 0, 0, 0};
  detail::host_array<uint64_t> sizes_elements, shared_elements;  // packed_varbit_vector `elements`
  uint32_t sizes_bits = 0, shared_bits = 0;
  uint64_t sizes_max = 0, shared_max = 0;
  detail::host_array<uint64_t> prev_bits[4], prev_subaccum[4], prev_accum[4];
};

// "Spiral file" writer: an uncompressed ZIP64 archive whose members are the reference's, in the
// reference's creation order (modules/io/spiral_file.h:9-27, seqset.cpp:19-44), framed byte for byte
// as spiral_file_create_mmap frames them through the vendored minizip
// (modules/io/spiral_file_mmap.cpp:361-450: every member is opened with zip64 = 1, method 0):
//   local header   : version needed 45, flag 0, method 0, dos date 0, crc, sizes, name, and a 20-byte
//                    extra field 0x0001 / 16 / uncompressed / compressed (vendor/minizip/zip.c:1089-1150).
//                    Sizes below 4 GiB are patched into the 32-bit fields and the extra field keeps its
//                    zeros; from 4 GiB on the 32-bit fields stay 0xFFFFFFFF and the extra field holds
//                    the sizes (zip.c:1857-1895).
//   crc            : JSON members (create_path_contents, non-raw) carry their CRC-32; array members
//                    (create_path, raw: the data is written through an mmap afterwards) carry 0
//                    ("don't bother to fill crc", spiral_file_mmap.cpp:421-423) -- which is why generic
//                    unzip tools report CRC errors on a .bg and readers go by offset.
//   central header : version made by 0 / needed 20, raised to 45 / 45 for a member that needs ZIP64; a
//                    0x0001 extra field with only the values that overflow 32 bits (zip.c:1753-1810)
//   end            : ZIP64 end-of-central-directory record + locator only when the central directory
//                    starts at or beyond 4 GiB (zip.c:1980-2030), then the classic end record.
// file_info.json carries a timestamp/uuid/command line in the reference, so whole-file identity is
// impossible even between two reference runs; the parity contract is every other member byte for byte.
class seqset_file_writer {
 public:
  explicit seqset_file_writer(const std::string& path) : m_fd(::open(path.c_str(), O_CREAT | O_RDWR | O_TRUNC, 0666)) {
    if (m_fd < 0) throw io_exception("Could not open zip for writing: " + path + ": " + strerror(errno));
  }
  ~seqset_file_writer() { if (m_fd >= 0) ::close(m_fd); }
  seqset_file_writer(const seqset_file_writer&) = delete;
  seqset_file_writer& operator=(const seqset_file_writer&) = delete;

  // spiral_file_create_state::create_membuf -> create_path: an array member (crc field 0)
  void add(const std::string& name, const void* data, uint64_t size) {
    const uint64_t off = begin_member(name, size, 0);
    write_at(off, data, size);
  }
  // create_json -> create_path_contents: a JSON member (real CRC-32)
  void add(const std::string& name, const std::string& text) {
    const uint64_t off = begin_member(name, text.size(), crc32(text.data(), text.size()));
    write_at(off, text.data(), text.size());
  }
  // create_path without contents: the member's bytes are zeros (a hole) until written with write_at;
  // returns the offset of its data in the file
  uint64_t reserve(const std::string& name, uint64_t size) { return begin_member(name, size, 0); }
  void write_at(uint64_t offset, const void* data, uint64_t size) {
    const char* p = static_cast<const char*>(data);
    while (size) {
      const ssize_t n = ::pwrite(m_fd, p, (size_t)std::min<uint64_t>(size, 1ull << 30), (off_t)offset);
      if (n <= 0) throw io_exception(std::string("write to zip: ") + strerror(errno));
      p += n; offset += (uint64_t)n; size -= (uint64_t)n;
    }
  }
  // zipClose_64; returns the size of the file
  uint64_t finish() {
    const uint64_t cd = m_end;
    std::string dir;
    for (const entry& e : m_entries) {
      const bool big_size = e.size >= 0xffffffffull, big_off = e.offset >= 0xffffffffull;
      const bool z64 = big_size || big_off;
      std::string x;  // 0x0001 extra field: only what overflows
      if (big_size) { put64(x, e.size); put64(x, e.size); }
      if (big_off) put64(x, e.offset);
      put32(dir, 0x02014b50); put16(dir, z64 ? 45 : 0); put16(dir, z64 ? 45 : 20); put16(dir, 0); put16(dir, 0); put32(dir, 0);
      put32(dir, e.crc); put32(dir, big_size ? 0xffffffffu : (uint32_t)e.size); put32(dir, big_size ? 0xffffffffu : (uint32_t)e.size);
      put16(dir, (uint16_t)e.name.size()); put16(dir, x.empty() ? 0 : (uint16_t)(x.size() + 4)); put16(dir, 0);
      put16(dir, 0); put16(dir, 0); put32(dir, 0); put32(dir, big_off ? 0xffffffffu : (uint32_t)e.offset);
      dir += e.name;
      if (!x.empty()) { put16(dir, 1); put16(dir, (uint16_t)x.size()); dir += x; }
    }
    std::string tail;
    const uint64_t n = m_entries.size();
    if (cd >= 0xffffffffull) {
      const uint64_t z64_pos = cd + dir.size();
      put32(tail, 0x06064b50); put64(tail, 44); put16(tail, 0); put16(tail, 45); put32(tail, 0); put32(tail, 0);
      put64(tail, n); put64(tail, n); put64(tail, dir.size()); put64(tail, cd);
      put32(tail, 0x07064b50); put32(tail, 0); put64(tail, z64_pos); put32(tail, 1);
    }
    put32(tail, 0x06054b50); put16(tail, 0); put16(tail, 0);
    put16(tail, n >= 0xffff ? 0xffff : (uint16_t)n); put16(tail, n >= 0xffff ? 0xffff : (uint16_t)n);
    put32(tail, (uint32_t)dir.size()); put32(tail, cd >= 0xffffffffull ? 0xffffffffu : (uint32_t)cd); put16(tail, 0);
    write_at(cd, dir.data(), dir.size());
    write_at(cd + dir.size(), tail.data(), tail.size());
    m_end = cd + dir.size() + tail.size();
    ::close(m_fd);
    m_fd = -1;
    return m_end;
  }

 private:
  struct entry { std::string name; uint64_t offset /* of the local header */, size; uint32_t crc; };
  // local header (final, "patched" form) + room for `size` bytes; returns where the data goes
  uint64_t begin_member(const std::string& name, uint64_t size, uint32_t crc) {
    entry e{name, m_end, size, crc};
    const bool big = size >= 0xffffffffull;
    std::string h;
    put32(h, 0x04034b50); put16(h, 45); put16(h, 0); put16(h, 0); put32(h, 0);
    put32(h, crc); put32(h, big ? 0xffffffffu : (uint32_t)size); put32(h, big ? 0xffffffffu : (uint32_t)size);
    put16(h, (uint16_t)name.size()); put16(h, 20);
    h += name;
    put16(h, 1); put16(h, 16); put64(h, big ? size : 0); put64(h, big ? size : 0);
    write_at(m_end, h.data(), h.size());
    const uint64_t data_off = m_end + h.size();
    m_end = data_off + size;
    if (::ftruncate(m_fd, (off_t)m_end) < 0) throw io_exception(std::string("ftruncate to extend zip: ") + strerror(errno));
    m_entries.push_back(e);
    return data_off;
  }
  static uint32_t crc32(const void* data, uint64_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
      for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : (c >> 1);
        table[i] = c;
      }
      init = true;
    }
    uint32_t c = 0xffffffffu;
    const uint8_t* p = static_cast<const uint8_t*>(data);
    for (uint64_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c ^ 0xffffffffu;
  }
  static void put16(std::string& o, uint16_t v) { o.append(reinterpret_cast<const char*>(&v), 2); }
  static void put32(std::string& o, uint32_t v) { o.append(reinterpret_cast<const char*>(&v), 4); }
  static void put64(std::string& o, uint64_t v) { o.append(reinterpret_cast<const char*>(&v), 8); }
  int m_fd;
  uint64_t m_end = 0;
  std::vector<entry> m_entries;
};

class builder {
 public:
  explicit builder(session& s) : m_s(s) {}
  // builder::build_chunks (bs/builder.cpp:8-164): sizes, shared, prev bits of the "complete" pass
  void build_chunks(const std::string& /*pass_name*/ = "complete", bool /*keep_tmp*/ = true,
                    progress_handler_t progress = null_progress_handler) {
    m_s.ensure_seqset_built();
    progress(1.0);
  }
  // builder::make_seqset (bs/builder.cpp:207-263) + seqset::finalize: the tables in host memory
  seqset_tables tables() {
    seqset_tables t;
    uint64_t *pb[4], *ps[4], *pa[4];
    // sizes / shared travel as packed_varbit_vector elements (bgx_export_varbit below), not as uint16 arrays
    detail::ck(bgx_export_seqset(m_s.ctx(), &t.num_entries, &t.max_entry_len, nullptr, nullptr, pb, ps, pa, t.fixed));
    uint64_t lay[6];
    detail::ck(bgx_seqset_layout(m_s.ctx(), lay));
    for (int b = 0; b < 4; ++b) {
      t.prev_bits[b].p = pb[b]; t.prev_bits[b].n = lay[3];
      t.prev_subaccum[b].p = ps[b]; t.prev_subaccum[b].n = lay[4];
      t.prev_accum[b].p = pa[b]; t.prev_accum[b].n = lay[5];
    }
    detail::ck(bgx_export_varbit(m_s.ctx(), 0, &t.sizes_elements.p, &t.sizes_elements.n, &t.sizes_bits, &t.sizes_max));
    detail::ck(bgx_export_varbit(m_s.ctx(), 1, &t.shared_elements.p, &t.shared_elements.n, &t.shared_bits, &t.shared_max));
    return t;
  }
  // writes the seqset spiral file (uuid: file_info.json's, the BioGraph id); returns the tables it wrote
  seqset_tables make_seqset(const std::string& path, progress_handler_t progress = null_progress_handler,
                            const std::string& uuid = "") {
    seqset_tables t = tables();
    seqset_file_writer w(path);
    char ts[64];
    time_t now = time(nullptr);
    strftime(ts, sizeof(ts), "%a %b %e %H:%M:%S %Y", localtime(&now));
    w.add("file_info.json", std::string("{\"build_host\":\"bgx\",\"build_is_clean\":true,\"build_revision\":\"") + bgx_version() +
                                "\",\"build_timestamp\":0,\"build_timestamp_text\":\"\",\"build_user\":\"\",\"command_line\":[],"
                                "\"create_timestamp\":" + std::to_string((long long)now) + ",\"create_timestamp_text\":\"" + ts +
                                "\",\"uuid\":\"" + uuid + "\"}");
    w.add("part_info.json", part_info("seqset", 1, 1, 0));  // seqset::seqset_version 1.1.0 (seqset.cpp:12)
    w.add("seqset.json", "{\"num_entries\":" + std::to_string(t.num_entries) + "}");
    w.add("fixed", t.fixed, sizeof(t.fixed));
    add_varbit(w, "entry_sizes", t.sizes_elements, t.sizes_bits, t.num_entries, t.sizes_max);
    add_varbit(w, "shared", t.shared_elements, t.shared_bits, t.num_entries, t.shared_max);
    for (int b = 0; b < 4; ++b) {
      std::string dir = std::string("prev_") + "ACGT"[b] + "/";
      w.add(dir + "part_info.json", part_info("bitcount", 1, 0, 0));  // bitcount.cpp:10
      w.add(dir + "bitcount.json", "{\"nbits\":" + std::to_string(t.num_entries) + "}");
      w.add(dir + "bits", t.prev_bits[b].p, t.prev_bits[b].n * 8);
      w.add(dir + "subaccum", t.prev_subaccum[b].p, t.prev_subaccum[b].n * 8);
      w.add(dir + "accum", t.prev_accum[b].p, t.prev_accum[b].n * 8);
    }
    w.finish();
    progress(1.0);
    return t;
  }

 private:
  static std::string part_info(const char* type, int major, int minor, int patch) {
    return std::string("{\"part_type\":\"") + type + "\",\"version\":{\"build\":\"\",\"major\":" + std::to_string(major) +
           ",\"minor\":" + std::to_string(minor) + ",\"patch\":" + std::to_string(patch) + ",\"pre\":\"\"}}";
  }
  static void add_varbit(seqset_file_writer& w, const std::string& name, const detail::host_array<uint64_t>& el, uint32_t bits,
                         uint64_t count, uint64_t max_value) {
    w.add(name + "/part_info.json", part_info("packed_varbit_vector", 1, 0, 0));  // packed_varbit_vector.cpp:7
    w.add(name + "/packed_varbit_vector.json", "{\"bits_per_value\":" + std::to_string(bits) + ",\"element_count\":" +
                                                   std::to_string(count) + ",\"max_value\":" + std::to_string(max_value) + "}");
    w.add(name + "/elements", el.p, el.n * 8);
  }
  session& m_s;
};

// seqset_for_reads (modules/bio_base/seqset_testutil.h:13): reads in, seqset file out, no
// correction -- the simplest whole-stage-3 operator.  Returns the tables; writes `path` if given.
inline seqset_tables seqset_for_reads(const std::vector<std::string>& reads, const std::string& path = "", int device = 0) {
  count_kmer_options ko;
  ko.device = device;
  session s(ko);
  {
    kmer_counter kc(s);
    kmer_counter::prob_pass_processor p(kc);
    for (const std::string& r : reads) p.add(r);
  }
  detail::ck(bgx_seed_uncorrected(s.ctx()));
  builder b(s);
  b.build_chunks();
  return path.empty() ? b.tables() : b.make_seqset(path);
}

// ---- input: reading spiral files back ----------------------------------------------------------------------
// spiral_file_open_mmap (modules/io/spiral_file_mmap.cpp:82-160): members are stored (method 0) and read
// by offset; the CRC fields of array members are not valid (:421-423), so nothing is verified.  Handles
// the classic and the ZIP64 end records and the 0x0001 extra fields the reference's minizip writes.
class spiral_file_reader {
 public:
  struct member { std::string name; uint64_t offset /* of the data */, size; };
  explicit spiral_file_reader(const std::string& path) : m_path(path), m_fd(::open(path.c_str(), O_RDONLY)) {
    if (m_fd < 0) throw io_exception("Could not open spiral file " + path + ": " + strerror(errno));
    const off_t end = ::lseek(m_fd, 0, SEEK_END);
    if (end < 22) throw io_exception(path + " is not a spiral file (too short)");
    m_size = (uint64_t)end;
    const uint64_t fsz = (uint64_t)end, tail_n = std::min<uint64_t>(fsz, 65536 + 22 + 20);
    std::string tail = pread_str(fsz - tail_n, tail_n);
    size_t e = std::string::npos;
    for (size_t i = tail.size() - 22 + 1; i-- > 0;)
      if (get32(tail, i) == 0x06054b50u) { e = i; break; }
    if (e == std::string::npos) throw io_exception(path + " is not a spiral file (no end of central directory)");
    uint64_t n = get16(tail, e + 10), cd_size = get32(tail, e + 12), cd_off = get32(tail, e + 16);
    if (e >= 20 && get32(tail, e - 20) == 0x07064b50u) {  // ZIP64 locator -> ZIP64 end record
      const uint64_t z = get64(tail, e - 20 + 8);
      const std::string r = pread_str(z, 56);
      if (get32(r, 0) != 0x06064b50u) throw io_exception(path + ": bad ZIP64 end of central directory");
      n = get64(r, 32); cd_size = get64(r, 40); cd_off = get64(r, 48);
    }
    const std::string cd = pread_str(cd_off, cd_size);
    size_t p = 0;
    for (uint64_t i = 0; i < n; ++i) {
      if (p + 46 > cd.size() || get32(cd, p) != 0x02014b50u) throw io_exception(path + ": bad central directory");
      if (get16(cd, p + 10) != 0) throw io_exception(path + ": compressed member (a spiral file stores its members)");
      uint64_t usize = get32(cd, p + 24), lho = get32(cd, p + 42);
      const uint16_t nl = get16(cd, p + 28), xl = get16(cd, p + 30), cl = get16(cd, p + 32);
      if (p + 46 + (size_t)nl + xl > cd.size()) throw io_exception(path + ": bad central directory");
      const std::string name = cd.substr(p + 46, nl);
      size_t x = p + 46 + nl;
      const size_t xe = x + xl;
      while (x + 4 <= xe) {  // 0x0001: only the values that overflowed, in this order
        const uint16_t id = get16(cd, x), len = get16(cd, x + 2);
        if (id == 1) {
          size_t q = x + 4;
          if (usize == 0xffffffffull) { usize = get64(cd, q); q += 8; }
          if (get32(cd, p + 20) == 0xffffffffu) q += 8;
          if (lho == 0xffffffffull) lho = get64(cd, q);
        }
        x += 4 + (size_t)len;
      }
      const std::string lh = pread_str(lho, 30);
      if (get32(lh, 0) != 0x04034b50u) throw io_exception(path + ": bad local header of " + name);
      m_members.push_back(member{name, lho + 30 + get16(lh, 26) + get16(lh, 28), usize});
      m_index[name] = m_members.size() - 1;
      p = xe + cl;
    }
  }
  ~spiral_file_reader() { if (m_fd >= 0) ::close(m_fd); }
  spiral_file_reader(const spiral_file_reader&) = delete;
  spiral_file_reader& operator=(const spiral_file_reader&) = delete;
  const std::vector<member>& members() const { return m_members; }   // creation order
  bool has(const std::string& name) const { return m_index.count(name) != 0; }
  const member& find(const std::string& name) const {
    auto it = m_index.find(name);
    if (it == m_index.end()) throw io_exception("Path not found in spiral file " + m_path + ": " + name);
    return m_members[it->second];
  }
  std::string read(const std::string& name) const { const member& m = find(name); return pread_str(m.offset, m.size); }
  template <typename T>
  std::vector<T> read_array(const std::string& name) const {
    const member& m = find(name);
    if (m.offset > m_size || m.size > m_size - m.offset) throw io_exception("corrupt spiral file " + m_path + ": member " + name + " lies beyond the end of the file");
    std::vector<T> v(m.size / sizeof(T));
    pread_into(m.offset, v.data(), v.size() * sizeof(T));
    return v;
  }
  // the value of a top-level string / integer field of a (compact) JSON member
  static std::string json_string_field(const std::string& json, const std::string& key) {
    const size_t k = json.find("\"" + key + "\":\"");
    if (k == std::string::npos) return "";
    const size_t b = k + key.size() + 4, e = json.find('"', b);
    return json.substr(b, e - b);
  }
  static uint64_t json_uint_field(const std::string& json, const std::string& key) {
    const size_t k = json.find("\"" + key + "\":");
    if (k == std::string::npos) throw io_exception("field " + key + " missing in " + json);
    return strtoull(json.c_str() + k + key.size() + 3, nullptr, 10);
  }

 private:
  template <typename T>
  static T get(const std::string& s, size_t i) {
    if (i + sizeof(T) > s.size()) throw io_exception("corrupt spiral file: a header field lies beyond its record");
    T v;
    memcpy(&v, s.data() + i, sizeof(T));
    return v;
  }
  static uint16_t get16(const std::string& s, size_t i) { return get<uint16_t>(s, i); }
  static uint32_t get32(const std::string& s, size_t i) { return get<uint32_t>(s, i); }
  static uint64_t get64(const std::string& s, size_t i) { return get<uint64_t>(s, i); }
  void pread_into(uint64_t off, void* dst, uint64_t n) const {
    char* p = static_cast<char*>(dst);
    while (n) {
      const ssize_t r = ::pread(m_fd, p, (size_t)std::min<uint64_t>(n, 1ull << 30), (off_t)off);
      if (r <= 0) throw io_exception("short read from spiral file " + m_path);
      p += r; off += (uint64_t)r; n -= (uint64_t)r;
    }
  }
  std::string pread_str(uint64_t off, uint64_t n) const {
    if (off > m_size || n > m_size - off) throw io_exception("corrupt spiral file " + m_path + ": a record lies beyond the end of the file");
    std::string s(n, '\0');
    pread_into(off, &s[0], n);
    return s;
  }
  std::string m_path;
  int m_fd;
  uint64_t m_size = 0;
  std::vector<member> m_members;
  std::unordered_map<std::string, size_t> m_index;
};

// seqset_file / seqset (modules/bio_base/seqset.cpp:46-111): the members a merge reads.  entry_sizes is a
// packed_varbit_vector from seqset version 1.1.0 on and a raw uint8 array before (:58-62).
class seqset_file {
 public:
  explicit seqset_file(const std::string& path) : m_path(path) {
    spiral_file_reader r(path);
    m_uuid = spiral_file_reader::json_string_field(r.read("file_info.json"), "uuid");
    m_command_line = r.read("file_info.json");
    m_entries = spiral_file_reader::json_uint_field(r.read("seqset.json"), "num_entries");
    m_sizes.resize(m_entries);
    if (r.has("entry_sizes/elements")) {
      const std::string meta = r.read("entry_sizes/packed_varbit_vector.json");
      const unsigned bits = (unsigned)spiral_file_reader::json_uint_field(meta, "bits_per_value");
      if (spiral_file_reader::json_uint_field(meta, "element_count") != m_entries) throw io_exception(path + ": entry_sizes has the wrong element count");
      const std::vector<uint64_t> el = r.read_array<uint64_t>("entry_sizes/elements");
      const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
      for (uint64_t i = 0; i < m_entries; ++i) {  // packed_varbit_vector::get (packed_varbit_vector.cpp:80-139)
        const uint64_t pos = i * bits, q = pos >> 6, sh = pos & 63;
        uint64_t v = el[q] >> sh;
        if (sh + bits > 64) v |= el[q + 1] << (64 - sh);
        m_sizes[i] = (uint16_t)(v & mask);
      }
    } else {
      const std::vector<uint8_t> raw = r.read_array<uint8_t>("entry_sizes");
      if (raw.size() < m_entries) throw io_exception(path + ": entry_sizes is too short");
      for (uint64_t i = 0; i < m_entries; ++i) m_sizes[i] = raw[i];
    }
    for (int b = 0; b < 4; ++b) {
      m_prev[b] = r.read_array<uint64_t>(std::string("prev_") + "ACGT"[b] + "/bits");
      if (m_prev[b].size() < (m_entries + 63) / 64) throw io_exception(path + ": prev bits are too short");
    }
    for (uint16_t v : m_sizes) m_max_read_len = std::max<unsigned>(m_max_read_len, v);
  }
  const std::string& path() const { return m_path; }
  const std::string& uuid() const { return m_uuid; }
  const std::string& file_info() const { return m_command_line; }
  uint64_t size() const { return m_entries; }
  unsigned max_read_len() const { return m_max_read_len; }
  bgx_seqset_part part() const {
    bgx_seqset_part p;
    p.n_entries = m_entries;
    p.sizes = m_sizes.data();
    for (int b = 0; b < 4; ++b) p.prev_bits[b] = m_prev[b].data();
    return p;
  }

 private:
  std::string m_path, m_uuid, m_command_line;
  uint64_t m_entries = 0;
  unsigned m_max_read_len = 0;
  std::vector<uint16_t> m_sizes;
  std::vector<uint64_t> m_prev[4];
};

// seqset_flat_builder + make_mergemap + seqset_merger (modules/bio_base/seqset_flat.h, make_mergemap.h,
// seqset_merger.h; driven by MergeSEQSETMain::do_merge, modules/biograph/biograph_merge.cpp:199-290) in ONE
// GPU call: the reference writes a .flat and a .mergemap temp file per input and then merges; here the
// flattened entries and the mergemaps stay on the device.
class seqset_merger {
 public:
  struct mergemap_tables {  // the `merged_entries` bitcount of seqset_mergemap (seqset_mergemap.cpp:5-20)
    uint64_t n_bits = 0, n_set = 0;
    detail::host_array<uint64_t> bits, subaccum, accum;
  };
  // parallel_splits: see bgx_merge_seqsets (0 = g_parallel_splits of the reference binary)
  seqset_merger(session& s, const std::vector<const seqset_file*>& inputs, uint64_t parallel_splits = 0)
      : m_s(s), m_inputs(inputs), m_splits(parallel_splits) {
    if (inputs.empty()) throw io_exception("seqset_merger: no inputs");
  }
  // make_mergemap::build + seqset_merger::build
  void build(progress_handler_t progress = null_progress_handler) {
    std::vector<bgx_seqset_part> parts;
    for (const seqset_file* f : m_inputs) parts.push_back(f->part());
    detail::ck(bgx_merge_seqsets(m_s.ctx(), parts.data(), (uint32_t)parts.size(), m_splits));
    m_done = true;
    progress(1.0);
  }
  // make_mergemap::total_merged_entries
  size_t total_merged_entries() {
    need();
    uint64_t lay[6];
    detail::ck(bgx_seqset_layout(m_s.ctx(), lay));
    return (size_t)lay[1];
  }
  // make_mergemap::fill_mergemap(input_id, builder)
  mergemap_tables fill_mergemap(unsigned input_id) {
    need();
    mergemap_tables t;
    uint64_t* out[3];
    detail::ck(bgx_export_mergemap(m_s.ctx(), input_id, out, &t.n_bits, &t.n_set));
    t.bits.p = out[0]; t.bits.n = (t.n_bits + 63) / 64;
    t.subaccum.p = out[1]; t.subaccum.n = (t.n_bits + 511) / 512;
    t.accum.p = out[2]; t.accum.n = (t.n_bits + 1 + 511) / 512;
    return t;
  }
  // ... written as the reference's .mergemap spiral file (seqset_mergemap_builder, seqset_mergemap.cpp:5-20)
  void write_mergemap(unsigned input_id, const std::string& path, const std::string& merged_seqset_uuid) {
    mergemap_tables t = fill_mergemap(input_id);
    seqset_file_writer w(path);
    w.add("file_info.json", file_info_json(""));
    w.add("part_info.json", part_info_json("mergemap", 1, 0, 0));
    w.add("mergemap.json", "{\"merged_seqset_uuid\":\"" + merged_seqset_uuid + "\",\"orig_seqset_uuid\":\"" + m_inputs[input_id]->uuid() + "\"}");
    w.add("merged_entries/part_info.json", part_info_json("bitcount", 1, 0, 0));
    w.add("merged_entries/bitcount.json", "{\"nbits\":" + std::to_string(t.n_bits) + "}");
    w.add("merged_entries/bits", t.bits.p, t.bits.n * 8);
    w.add("merged_entries/subaccum", t.subaccum.p, t.subaccum.n * 8);
    w.add("merged_entries/accum", t.accum.p, t.accum.n * 8);
    w.finish();
  }
  // seqset_flat::get(i) of input `input_id`, entries [first, first + count)
  std::vector<std::string> flat_entries(unsigned input_id, uint64_t first, uint64_t count) {
    need();
    detail::host_array<char> bases;
    detail::host_array<uint64_t> offs;
    detail::ck(bgx_export_flat_ascii(m_s.ctx(), input_id, first, count, &bases.p, &offs.p));
    std::vector<std::string> out;
    for (uint64_t i = 0; i < count; ++i) out.emplace_back(bases.p + offs.p[i], bases.p + offs.p[i + 1]);
    return out;
  }
  // the merged seqset spiral file (seqset_merger::build writes it through its create state)
  seqset_tables write_seqset(const std::string& path, const std::string& uuid, progress_handler_t progress = null_progress_handler) {
    need();
    return builder(m_s).make_seqset(path, progress, uuid);
  }
  // make_readmap::fast_migrate (modules/bio_mapred/make_readmap.cpp:46-52,459-520): every member of the old
  // readmap copied in order, read_ids/source_to_mid re-targeted through the input's mergemap, readmap.json
  // pointed at the merged seqset
  void fast_migrate(unsigned input_id, const std::string& old_readmap_path, const std::string& new_readmap_path,
                    const std::string& merged_seqset_uuid) {
    need();
    spiral_file_reader r(old_readmap_path);
    const uint64_t n_old = spiral_file_reader::json_uint_field(r.read("read_ids/source_to_mid/bitcount.json"), "nbits");
    if (n_old != m_inputs[input_id]->size()) throw io_exception(old_readmap_path + " does not belong to " + m_inputs[input_id]->path());
    std::vector<uint64_t> old_bits = r.read_array<uint64_t>("read_ids/source_to_mid/bits");
    old_bits.resize((n_old + 63) / 64);
    uint64_t* out[3];
    uint64_t n_bits = 0;
    detail::ck(bgx_migrate_bits(m_s.ctx(), input_id, old_bits.data(), n_old, out, &n_bits));
    detail::host_array<uint64_t> bits, sub, acc;
    bits.p = out[0]; bits.n = (n_bits + 63) / 64;
    sub.p = out[1]; sub.n = (n_bits + 511) / 512;
    acc.p = out[2]; acc.n = (n_bits + 1 + 511) / 512;
    seqset_file_writer w(new_readmap_path);
    for (const spiral_file_reader::member& m : r.members()) {
      if (m.name == "file_info.json") w.add(m.name, file_info_json(""));
      else if (m.name == "readmap.json") w.add(m.name, "{\"seqset_uuid\":\"" + merged_seqset_uuid + "\"}");
      else if (m.name == "read_ids/source_to_mid/bitcount.json") w.add(m.name, "{\"nbits\":" + std::to_string(n_bits) + "}");
      else if (m.name == "read_ids/source_to_mid/bits") w.add(m.name, bits.p, bits.n * 8);
      else if (m.name == "read_ids/source_to_mid/subaccum") w.add(m.name, sub.p, sub.n * 8);
      else if (m.name == "read_ids/source_to_mid/accum") w.add(m.name, acc.p, acc.n * 8);
      else if (m.name.size() > 5 && m.name.compare(m.name.size() - 5, 5, ".json") == 0) w.add(m.name, r.read(m.name));
      else { const std::string d = r.read(m.name); w.add(m.name, d.data(), d.size()); }
    }
    w.finish();
  }
  static std::string part_info_json(const char* type, int major, int minor, int patch) {
    return std::string("{\"part_type\":\"") + type + "\",\"version\":{\"build\":\"\",\"major\":" + std::to_string(major) +
           ",\"minor\":" + std::to_string(minor) + ",\"patch\":" + std::to_string(patch) + ",\"pre\":\"\"}}";
  }
  static std::string file_info_json(const std::string& uuid) {
    return std::string("{\"build_host\":\"bgx\",\"build_is_clean\":true,\"build_revision\":\"") + bgx_version() +
           "\",\"build_timestamp\":0,\"build_timestamp_text\":\"\",\"build_user\":\"\",\"command_line\":[],\"create_timestamp\":" +
           std::to_string((long long)time(nullptr)) + ",\"create_timestamp_text\":\"\",\"uuid\":\"" + uuid + "\"}";
  }

 private:
  void need() const { if (!m_done) throw io_exception("seqset_merger: call build() first"); }
  session& m_s;
  std::vector<const seqset_file*> m_inputs;
  uint64_t m_splits;
  bool m_done = false;
};

// make_readmap::do_make (modules/bio_mapred/make_readmap.{h,cpp}; called at biograph_create.cpp:818-831):
// the tables come from bgx_build_readmap, the spiral file is written here in the reference's member
// order (readmap 1.2.0: readmap.cpp:13; sparse_multi 1.0.0; packed_varbit_vector 1.0.0; packed_vector
// 1.0.0).
class make_readmap {
 public:
  struct tables {
    uint64_t n_rows = 0, n_entries = 0;
    detail::host_array<uint16_t> read_lengths;
    detail::host_array<uint64_t> mate_loop_ptr, is_forward;
    detail::host_array<uint64_t> source[3], dest[3];  // bits, subaccum, accum
  };
  static tables build(session& s, bool is_paired = false) {
    tables t;
    uint64_t* src[3];
    uint64_t* dst[3];
    detail::ck(bgx_build_readmap(s.ctx(), is_paired ? 1 : 0, &t.n_rows, &t.read_lengths.p, &t.mate_loop_ptr.p, &t.is_forward.p, src, dst));
    uint64_t lay[6];
    detail::ck(bgx_seqset_layout(s.ctx(), lay));
    t.n_entries = lay[1];
    t.read_lengths.n = t.mate_loop_ptr.n = t.n_rows;
    t.is_forward.n = (t.n_rows + 63) / 64;
    const uint64_t nb[2] = {t.n_entries, t.n_rows};
    for (int i = 0; i < 3; ++i) {
      t.source[i].p = src[i];
      t.dest[i].p = dst[i];
    }
    for (int w = 0; w < 2; ++w) {
      detail::host_array<uint64_t>* a = w ? t.dest : t.source;
      a[0].n = (nb[w] + 63) / 64;
      a[1].n = (nb[w] + 511) / 512;
      a[2].n = (nb[w] + 1 + 511) / 512;
    }
    return t;
  }
  static tables do_make(const std::string& readmap_file_path, session& s, const std::string& seqset_uuid, bool is_paired,
                        unsigned max_read_len, progress_handler_t progress = null_progress_handler) {
    tables t = build(s, is_paired);  // paired: reads 2i and 2i+1 of the session are mates
    seqset_file_writer w(readmap_file_path);
    w.add("file_info.json", std::string("{\"build_host\":\"bgx\",\"build_is_clean\":true,\"build_revision\":\"") + bgx_version() +
                                "\",\"build_timestamp\":0,\"build_timestamp_text\":\"\",\"build_user\":\"\",\"command_line\":[],"
                                "\"create_timestamp\":" + std::to_string((long long)time(nullptr)) +
                                ",\"create_timestamp_text\":\"\",\"uuid\":\"\"}");
    w.add("part_info.json", part_info("readmap", 1, 2, 0));
    w.add("readmap.json", "{\"seqset_uuid\":\"" + seqset_uuid + "\"}");
    w.add("read_ids/part_info.json", part_info("sparse_multi", 1, 0, 0));
    const char* dirs[2] = {"read_ids/source_to_mid/", "read_ids/dest_to_mid/"};
    const uint64_t nb[2] = {t.n_entries, t.n_rows};
    for (int k = 0; k < 2; ++k) {
      const detail::host_array<uint64_t>* a = k ? t.dest : t.source;
      w.add(std::string(dirs[k]) + "part_info.json", part_info("bitcount", 1, 0, 0));
      w.add(std::string(dirs[k]) + "bitcount.json", "{\"nbits\":" + std::to_string(nb[k]) + "}");
      w.add(std::string(dirs[k]) + "bits", a[0].p, a[0].n * 8);
      w.add(std::string(dirs[k]) + "subaccum", a[1].p, a[1].n * 8);
      w.add(std::string(dirs[k]) + "accum", a[2].p, a[2].n * 8);
    }
    // mutable_packed_varbit_vector(state, num_reads, max_read_len) / (state, n, n) (make_readmap.cpp:226-227,254-255)
    std::vector<uint64_t> lens(t.n_rows);
    for (uint64_t i = 0; i < t.n_rows; ++i) lens[i] = t.read_lengths.p[i];
    add_varbit(w, "read_lengths", lens.data(), t.n_rows, max_read_len);
    add_varbit(w, "mate_loop_ptr", t.mate_loop_ptr.p, t.n_rows, t.n_rows);
    w.add("is_forward/part_info.json", part_info("packed_vector", 1, 0, 0));
    w.add("is_forward/packed_data", t.is_forward.p, t.is_forward.n * 8);
    w.add("is_forward/packed_vector.json", "{\"value_count\":" + std::to_string(t.n_rows) + ",\"value_width_bits\":1}");
    w.finish();
    progress(1.0);
    return t;
  }

 private:
  static std::string part_info(const char* type, int major, int minor, int patch) {
    return std::string("{\"part_type\":\"") + type + "\",\"version\":{\"build\":\"\",\"major\":" + std::to_string(major) +
           ",\"minor\":" + std::to_string(minor) + ",\"patch\":" + std::to_string(patch) + ",\"pre\":\"\"}}";
  }
  // packed_varbit_vector (modules/io/packed_varbit_vector.cpp:174-228): bits_per_value = bit_length(max_value),
  // values packed LSB-first into little-endian uint64 words
  static void add_varbit(seqset_file_writer& w, const std::string& name, const uint64_t* vals, uint64_t n, uint64_t max_value) {
    unsigned bits = 0;
    while (bits < 64 && (max_value >> bits)) ++bits;
    std::vector<uint64_t> el((n * bits + 63) / 64, 0);
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t pos = i * bits, q = pos >> 6, sh = pos & 63;
      el[q] |= vals[i] << sh;
      if (sh + bits > 64) el[q + 1] |= vals[i] >> (64 - sh);
    }
    w.add(name + "/part_info.json", part_info("packed_varbit_vector", 1, 0, 0));
    w.add(name + "/packed_varbit_vector.json", "{\"bits_per_value\":" + std::to_string(bits) + ",\"element_count\":" +
                                                   std::to_string(n) + ",\"max_value\":" + std::to_string(max_value) + "}");
    w.add(name + "/elements", el.data(), el.size() * 8);
  }
};

}  // namespace bgx_bs

// ref_shim.cpp -- C entry points over the REFERENCE'S OWN classes for the hot path, compiled from the sources where
// they lie under /root/reference (oracle/Makefile, target `_ref`; outputs only under oracle/_ref/).
//
// TEST INFRASTRUCTURE.  Nothing under biograph_b200/ or include/ may load this.  It exists to check the CPU
// restatement (oracle/oracle.cpp) and the CUDA path against what the reference itself computes, and as the
// `cpu_baseline.kind = "reference"` arm of bench.py.
//
// What is the reference's and what is not: every class used below (kmer_counter, kmer_set, correct_reads,
// fast_read_correct, part_repo, expander, builder, seqset, bitcount, packed_varbit_vector, sparse_multi, make_readmap,
// readmap, seqset_flat, make_mergemap, seqset_mergemap, seqset_merger, biograph_dir, spiral_file_create_mem / _mmap,
// spiral_file_open_mem / _mmap) is compiled unmodified from /root/reference/modules/**.  The reference's Bazel build
// and its third-party libraries (Boost, glog, json_spirit, msgpack, htslib) are absent from this image;
// oracle/ref_stubs/ holds minimal stand-ins for the few headers of those that the leaf sources include.  The code in
// THIS file only restates DRIVER sequences -- which class is called when, with the CLI's defaults -- because the
// functions that hold them in the reference read the map-reduce temp files (manifests, msgpack kv streams) that are
// out of scope:
//   * ref_count_kmers / ref_correct / ref_make_seqset: SEQSETMain::run (modules/biograph/biograph_create.cpp:665-779,
//     835-950) and kmerizer::run (modules/bio_mapred/kmerize_bf.cpp:267-430).
//       - import: every read is handed to prob_pass_processor::add as read_importer_state::process does
//         (biograph_create.cpp:119-151); reads are kept in memory instead of the msgpack temp files.
//       - k-mer filter: tot_count >= min_count and the strand-skew cut of kmer_passes with the defaults of
//         kmerize_bf_params (kmerize_bf.cpp:290-318); the overrepresentation filter is off (threshold 0 disables it
//         there too).
//       - the reference genome as a compression dictionary (add_initial_repo) is not used: result-invisible.
//   * ref_make_readmap: make_readmap::do_make as SEQSETMain::do_readmap calls it (biograph_create.cpp:818-826); the
//     corrected-read records reach it through the in-memory manifest stand-in.
//   * ref_merge / ref_fast_migrate: MergeSEQSETMain (modules/biograph/biograph_merge.cpp:196-312) through in-memory
//     spiral files.
//   * ref_open_seqset_file / ref_read_readmap_file / ref_open_biograph: the reference's readers on files that bgx wrote.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "modules/bio_base/fast_read_correct.h"
#include "modules/bio_base/seqset.h"
#include "modules/bio_base/biograph_dir.h"
#include "modules/bio_base/corrected_read.h"
#include "modules/bio_base/make_mergemap.h"
#include "modules/bio_base/readmap.h"
#include "modules/bio_base/seqset_flat.h"
#include "modules/bio_base/seqset_merger.h"
#include "modules/bio_format/fastq.h"
#include "modules/io/file_io.h"
#include "modules/bio_mapred/kmer_set.h"
#include "modules/bio_mapred/make_readmap.h"
#include "modules/build_seqset/builder.h"
#include "modules/build_seqset/correct_reads.h"
#include "modules/build_seqset/expand.h"
#include "modules/build_seqset/kmer_counter.h"
#include "modules/build_seqset/part_repo.h"
#include "modules/io/config.h"
#include "modules/io/parallel.h"
#include "modules/io/spiral_file_mem.h"
#include "modules/io/spiral_file_mmap.h"
#include "modules/io/track_mem.h"

namespace {

thread_local std::string g_err;

struct reads_view {
  const char* bases;
  const int64_t* offs;
  int64_t n;
  string_view operator[](int64_t i) const { return string_view(bases + offs[i], size_t(offs[i + 1] - offs[i])); }
};

// run f(first, last) over chunks of [0, n) on the reference's own thread pool (modules/io/parallel.h): part_repo's
// writers and the k-mer pass processors rely on its per-thread state.
template <class F>
void blocks(int64_t n, int /*threads*/, const F& f) {
  if (n <= 0) return;
  parallel_for(0, size_t(n), [&](size_t a, size_t b) { f(int64_t(a), int64_t(b)); });
}

struct ref_run {
  std::string tmp;
  int threads = 1;
  unsigned k = 30;
  // count stage
  std::vector<uint64_t> c_kmer;
  std::vector<uint32_t> c_fwd, c_rev;
  std::vector<uint8_t> c_flags;
  std::unique_ptr<kmer_set> ks;
  // correct stage
  std::unique_ptr<build_seqset::part_repo> entries;
  std::unique_ptr<build_seqset::part_counts> part_counts;
  std::string cr_seq;
  std::vector<int64_t> cr_offs;
  std::vector<uint8_t> cr_kept;
  // seqset stage
  std::unique_ptr<spiral_file_create_mem> create;
  std::unique_ptr<spiral_file_open_mmap> open_mmap;  // ref_open_seqset_file (declared before ss: outlives it)
  std::unique_ptr<seqset> ss;
  int64_t stats[6] = {0, 0, 0, 0, 0, 0};
  spiral_file_mem_storage storage;
  std::vector<std::string> member_names;
  // merge: one bit vector over the merged entries per input (bit x: merged entry x comes from this input)
  std::vector<std::vector<uint64_t>> mergemaps;
  std::vector<std::unique_ptr<seqset_mergemap>> mergemap_objs;  // kept for ref_fast_migrate
};

template <class F>
int guarded(const F& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
  } catch (...) {
    g_err = "unknown exception";
  }
  return 1;
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// tmp_dir must exist; the reference keeps its partition / repository files there.
void* ref_open(const char* tmp_dir, int threads, uint64_t max_mem_bytes) {
  auto* r = new ref_run;
  r->tmp = tmp_dir;
  r->threads = threads < 1 ? 1 : threads;
  Config::set("temp_root", r->tmp);
  Config::set("resources_root", r->tmp);
  Config::set("path_bulkdata", r->tmp);
  set_thread_count(std::to_string(r->threads));
  if (max_mem_bytes) set_maximum_mem_bytes(max_mem_bytes);
  return r;
}

void ref_close(void* h) { delete static_cast<ref_run*>(h); }

// fast_read_correct on one read against an explicit list of canonical solid k-mers (ascending).
// out must hold len bytes; returns the corrected length, *corrections = substitutions made.
int ref_fast_read_correct(const char* read, int len, const uint64_t* solid, int64_t n_solid, int k,
                          int max_corrections, int min_good_run, char* out, int* corrections) {
  int n_out = -1;
  guarded([&] {
    frc_params params;
    params.max_corrections = max_corrections;
    params.min_good_run = min_good_run;
    params.kmer_size = k;
    params.kmer_lookup_f = [&](kmer_t kmer, frc_kmer* ki) -> bool {
      kmer_t canon = canonicalize(kmer, k, ki->flipped);
      const uint64_t* e = solid + n_solid;
      const uint64_t* it = std::lower_bound(solid, e, uint64_t(canon));
      if (it == e || *it != canon) return false;
      ki->index = it - solid;
      return true;
    };
    frc_output res = fast_read_correct(string_view(read, size_t(len)), params);
    std::string s = res.corrected.as_string();
    memcpy(out, s.data(), s.size());
    if (corrections) *corrections = int(res.corrections);
    n_out = int(s.size());
  });
  return n_out;
}

// Stage 1: kmer_counter (probabilistic pass, exact passes), then the solid kmer_set.
// counter_max_memory_bytes = 0: the create flow's budget (get_maximum_mem_bytes()).  Returns 0 on success.
int ref_count_kmers(void* h, const char* bases, const int64_t* offs, int64_t n, int k, int min_count,
                    uint64_t counter_max_memory_bytes, int force_exact_passes, uint64_t genome_bases) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    reads_view rv{bases, offs, n};
    r->k = k;
    build_seqset::count_kmer_options opts;
    opts.kmer_size = k;
    opts.min_count = min_count;
    // biograph_create.cpp:546: the counter may use the process's whole memory budget (--max-mem, default 48 GiB or
    // the machine's RAM); count_kmer_options' own default of 20 MB is for unit tests and means dozens of exact passes
    opts.max_memory_bytes = counter_max_memory_bytes ? counter_max_memory_bytes : get_maximum_mem_bytes();
    opts.force_exact_passes = force_exact_passes;
    // biograph_create.cpp:549-551 bounds the probabilistic table by 100 x the reference genome's size (--ref);
    // genome_bases = 0: 100 x the bases read, which only makes the table larger (fewer false positives of a filter
    // whose false positives never reach the result).
    opts.max_prob_table_entries = std::max<size_t>(size_t(genome_bases ? genome_bases : offs[n]) * 100, 1024 * 1024);
    const bool timing = getenv("REF_SHIM_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      auto now = std::chrono::steady_clock::now();
      if (timing) fprintf(stderr, "ref_count_kmers: %-18s %.2f s\n", what, std::chrono::duration<double>(now - t_last).count());
      t_last = now;
    };
    build_seqset::kmer_counter counter(opts);
    counter.start_prob_pass();
    lap("start_prob_pass");
    blocks(n, r->threads, [&](int64_t a, int64_t b) {
      build_seqset::kmer_counter::prob_pass_processor p(counter);
      for (int64_t i = a; i < b; ++i) p.add(rv[i]);
    });
    lap("prob pass");
    counter.close_prob_pass();
    lap("close_prob_pass");
    if (timing) fprintf(stderr, "ref_count_kmers: %u exact passes, counter budget %.1f GB\n", counter.exact_passes(), opts.max_memory_bytes / 1e9);
    for (unsigned pass = 0; pass < counter.exact_passes(); ++pass) {
      counter.start_exact_pass(pass);
      blocks(n, r->threads, [&](int64_t a, int64_t b) {
        build_seqset::kmer_counter::exact_pass_processor p(counter);
        for (int64_t i = a; i < b; ++i) p.add(rv[i]);
      });
    }
    lap("exact passes");
    counter.close_exact_passes();
    lap("close_exact_passes");

    // defaults of kmerize_bf_params (modules/bio_mapred/kmerize_bf.h:37-38); the create flow never changes them
    struct { size_t prior_count = 5; float skew_cutoff = 0.0f; } kbf;
    auto passes = [&](const build_seqset::kmer_counter::element& e) {
      size_t tot = size_t(e.fwd_count) + e.rev_count;
      if (tot < size_t(min_count)) return false;
      int32_t mn = std::min(e.fwd_count, e.rev_count);
      float low = float(mn + kbf.prior_count) / float(tot + 2 * kbf.prior_count);
      return !(low < kbf.skew_cutoff);
    };

    std::mutex mu;
    size_t approx = 0;
    r->c_kmer.clear(); r->c_fwd.clear(); r->c_rev.clear(); r->c_flags.clear();
    counter.extract_exact_counts([&](build_seqset::kmer_counter::extract_iterator s,
                                     build_seqset::kmer_counter::extract_iterator e) {
      std::vector<build_seqset::kmer_counter::element> local;
      for (auto it = s; it != e; ++it) local.push_back(*it);
      std::lock_guard<std::mutex> l(mu);
      approx += local.size();
      for (const auto& el : local) {
        r->c_kmer.push_back(el.kmer);
        r->c_fwd.push_back(el.fwd_count);
        r->c_rev.push_back(el.rev_count);
        r->c_flags.push_back(uint8_t((el.fwd_starts_read ? 1 : 0) | (el.rev_starts_read ? 2 : 0)));
      }
    });
    lap("extract counts");
    size_t mem_gb = std::max<size_t>(1, get_maximum_mem_bytes() / 1024 / 1024 / 1024);
    r->ks = make_unique<kmer_set>(
        approx, k, mem_gb,
        [&](const kmer_set::kmer_output_f& output_f, progress_handler_t) {
          counter.extract_exact_counts([&](build_seqset::kmer_counter::extract_iterator s,
                                           build_seqset::kmer_counter::extract_iterator e) {
            for (auto it = s; it != e; ++it) {
              auto el = *it;
              if (!passes(el)) continue;
              unsigned flags = 0;
              if (el.fwd_starts_read) flags |= kmer_set::k_fwd_starts_read;
              if (el.rev_starts_read) flags |= kmer_set::k_rev_starts_read;
              output_f(el.kmer, flags);
            }
          });
        },
        null_progress_handler);
    lap("kmer_set");
    counter.close();
    lap("close");
  });
}

// every element extract_exact_counts yielded (unordered; k-mers below the probabilistic threshold may be absent)
int64_t ref_counts(void* h, const uint64_t** kmers, const uint32_t** fwd, const uint32_t** rev, const uint8_t** flags) {
  auto* r = static_cast<ref_run*>(h);
  *kmers = r->c_kmer.data(); *fwd = r->c_fwd.data(); *rev = r->c_rev.data(); *flags = r->c_flags.data();
  return int64_t(r->c_kmer.size());
}

// the solid set in kmer_set order (ascending), flags bit0 = fwd_starts_read, bit1 = rev_starts_read
int64_t ref_solid_size(void* h) {
  auto* r = static_cast<ref_run*>(h);
  return r->ks ? int64_t(r->ks->size()) : -1;
}
int ref_solid(void* h, uint64_t* kmers, uint8_t* flags) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    size_t i = 0;
    for (auto it = r->ks->begin(); it != r->ks->end(); ++it, ++i) {
      kmers[i] = *it;
      unsigned f = r->ks->get_flags(i);
      flags[i] = uint8_t(((f & kmer_set::k_fwd_starts_read) ? 1 : 0) | ((f & kmer_set::k_rev_starts_read) ? 2 : 0));
    }
  });
}

// Stage 2: build_seqset::correct_reads over every read; seeds go into the part_repo's "initial" pass.
int ref_correct(void* h, const char* bases, const int64_t* offs, int64_t n, int max_corrections, int min_good_run,
                double trim_after_portion, int partition_depth) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    if (!r->ks) throw io_exception("ref_correct: no kmer_set (run ref_count_kmers first)");
    reads_view rv{bases, offs, n};
    if (partition_depth <= 0) {  // biograph_create.cpp:716-722
      partition_depth = n < 10 * 1000 * 1000 ? 2 : n < 100 * 1000 * 1000 ? 3 : 4;
    }
    r->entries = make_unique<build_seqset::part_repo>(partition_depth, r->tmp + "/seq_ref-", r->tmp + "/seq_repo");
    read_correction_params rcp;  // biograph_create.cpp:727-733
    rcp.min_kmer_score = 0;
    rcp.skip_snps = false;
    rcp.exact = (max_corrections == 0);
    // the CLI value is a float (biograph_create.cpp:489-490) widened into the params' double: 0.7f * 150 = 104.99..,
    // where 0.7 * 150 would be 105
    rcp.trim_after_portion = float(trim_after_portion);
    rcp.frc_max_corrections = max_corrections;
    rcp.frc_min_good_run = min_good_run;
    build_seqset::correct_reads cr(*r->entries, *r->ks, rcp);
    r->entries->open_write_pass("initial");
    std::vector<std::string> out{size_t(n), std::string()};
    r->cr_kept.assign(size_t(n), 0);
    blocks(n, r->threads, [&](int64_t a, int64_t b) {
      for (int64_t i = a; i < b; ++i) {
        unaligned_read ur;
        ur.sequence = std::string(rv[i]);
        corrected_read c;
        if (cr.correct(ur, c)) {
          out[size_t(i)] = c.corrected.as_string();
          r->cr_kept[size_t(i)] = 1;
        }
      }
    });
    r->entries->flush();
    r->part_counts = r->entries->release_part_counts("initial");
    r->cr_offs.assign(size_t(n) + 1, 0);
    r->cr_seq.clear();
    for (int64_t i = 0; i < n; ++i) {
      r->cr_seq += out[size_t(i)];
      r->cr_offs[size_t(i) + 1] = int64_t(r->cr_seq.size());
    }
  });
}

int64_t ref_corrected(void* h, const char** seq, const int64_t** offs, const uint8_t** kept) {
  auto* r = static_cast<ref_run*>(h);
  *seq = r->cr_seq.data(); *offs = r->cr_offs.data(); *kept = r->cr_kept.data();
  return int64_t(r->cr_kept.size());
}

// Instead of stage 2: seed the part_repo from given sequences with explicit seed counts, as the reference's own
// seqset_for_reads test helper does (modules/bio_base/seqset_testutil.cpp:24-32) -- nf / nr null: 1 and 1.
int ref_seed(void* h, const char* bases, const int64_t* offs, int64_t n, const int32_t* nf, const int32_t* nr,
             int partition_depth) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    reads_view rv{bases, offs, n};
    r->entries = make_unique<build_seqset::part_repo>(partition_depth > 0 ? partition_depth : 2, r->tmp + "/seq_ref-",
                                                      r->tmp + "/seq_repo");
    r->entries->open_write_pass("initial");
    blocks(n, r->threads, [&](int64_t a, int64_t b) {
      for (int64_t i = a; i < b; ++i) {
        dna_sequence s{std::string(rv[i])};
        r->entries->write(s, nf ? unsigned(nf[i]) : 1, nr ? unsigned(nr[i]) : 1);
      }
    });
    r->entries->flush();
    r->part_counts = r->entries->release_part_counts("initial");
  });
}

// Stage 3: SEQSETMain::make_seqset (biograph_create.cpp:914-950): the four expander calls with the production
// strides, build_chunks, make_seqset into an in-memory spiral file.
static int make_seqset_impl(void* h, const char* path) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    if (!r->entries) throw io_exception("ref_make_seqset: no entries (run ref_correct or ref_seed first)");
    auto& entries = *r->entries;
    entries.flush();
    if (r->part_counts) entries.reset_part_counts("initial", std::move(r->part_counts));
    {
      build_seqset::expander expand(entries, false);
      r->stats[0] = 0;
      r->stats[1] = int64_t(expand.sort_and_dedup("", "initial", "init_sorted", "", 0, 0));
      r->stats[2] = int64_t(expand.expand("init_sorted", "init_expanded", 7, 255));
      r->stats[3] = int64_t(expand.sort_and_dedup("init_sorted", "init_expanded", "pass2_sorted", "pass2_expanded", 1, 6));
      r->stats[4] = 0;
      r->stats[5] = int64_t(expand.sort_and_dedup("pass2_sorted", "pass2_expanded", "complete", "", 0, 0));
    }
    build_seqset::builder b;
    b.build_chunks(entries, "complete", false);
    entries.partitions("complete", false, true);
    r->entries.reset();
    if (path) {
      // as SEQSETMain::make_seqset does: into a spiral FILE (biograph_create.cpp:943-946), then reopened
      {
        spiral_file_create_mmap c(path);
        b.make_seqset(c.create());
      }
      r->open_mmap = make_unique<spiral_file_open_mmap>(path);
      r->ss = make_unique<seqset>(r->open_mmap->open());
    } else {
      r->create = make_unique<spiral_file_create_mem>();
      r->ss = b.make_seqset(r->create->create());
    }
  });
}

int ref_make_seqset(void* h) { return make_seqset_impl(h, nullptr); }
// the same, written by the reference's own spiral_file_create_mmap to `path` (a seqset file as `biograph create`
// leaves it in <out>.bg/seqset), then reopened from there
int ref_make_seqset_file(void* h, const char* path) { return make_seqset_impl(h, path); }

// Opens a seqset spiral FILE with the reference's own reader (spiral_file_open_mmap + seqset::seqset(open state):
// zip directory through the vendored minizip, part types and versions, bitcount / packed_varbit_vector members) --
// for files the facade's writer or bgx-create wrote.  ref_seqset_tables / ref_flat then work on it.
int ref_open_seqset_file(void* h, const char* path) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    r->ss.reset();
    r->open_mmap = make_unique<spiral_file_open_mmap>(path);
    r->ss = make_unique<seqset>(r->open_mmap->open());
  });
}
const char* ref_seqset_uuid(void* h) {
  auto* r = static_cast<ref_run*>(h);
  static thread_local std::string u;
  u = r->ss ? r->ss->uuid() : "";
  return u.c_str();
}

// Opens a readmap spiral FILE with the reference's own reader (readmap::open_anonymous_readmap: sparse_multi,
// packed_varbit_vector, packed_vector members) and reads every row back through its public accessors:
// entry = index_to_entry, len = get_readlength, fwd = get_is_forward, ptr = the row's mate-loop pointer
// (get_rev_comp for a forward row, get_mate_rc for a reverse one: one step along the loop, readmap.cpp:248-290).
// Arrays must hold ref_readmap_rows(path) elements.
int64_t ref_readmap_rows(const char* path) {
  int64_t n = -1;
  guarded([&] { n = int64_t(readmap::open_anonymous_readmap(path)->size()); });
  return n;
}
int ref_read_readmap_file(const char* path, uint64_t* entry, int32_t* len, uint8_t* fwd, uint64_t* ptr,
                          char* seqset_uuid_out /* 64 bytes */) {
  return guarded([&] {
    std::unique_ptr<readmap> rm = readmap::open_anonymous_readmap(path);
    if (!rm->has_pairing_data() || !rm->has_mate_loop()) throw io_exception("readmap without a mate loop table");
    size_t n = rm->size();
    for (size_t i = 0; i < n; ++i) {
      entry[i] = rm->index_to_entry(i);
      len[i] = rm->get_readlength(uint32_t(i));
      fwd[i] = rm->get_is_forward(uint32_t(i)) ? 1 : 0;
      ptr[i] = fwd[i] ? rm->get_rev_comp(uint32_t(i)) : rm->get_mate_rc(uint32_t(i));
    }
    std::string u = rm->metadata().seqset_uuid;
    strncpy(seqset_uuid_out, u.c_str(), 63);
    seqset_uuid_out[63] = 0;
  });
}

// fastq_reader::read (modules/bio_format/fastq.cpp:40-126) over a FASTQ file through the reference's file_reader, as
// read_importer feeds it: the bases of every record it accepts (concatenated, offs[n + 1]) and, where it throws, the
// exception's text (an empty string = the whole file was accepted).  Buffers come from malloc: ref_free.
int ref_read_fastq(const char* path, char** bases, int64_t** offs, int64_t* n_reads, char* error, size_t error_cap) {
  return guarded([&] {
    std::string all, err;
    std::vector<int64_t> o{0};
    try {
      file_reader fr(path);
      fastq_reader rd(fr, false /* as read_importer does: qualities are not kept, read_importer.cpp:633-635 */);
      read_id id;
      unaligned_read r;
      while (rd.read(id, r)) {
        all += r.sequence;
        o.push_back(int64_t(all.size()));
      }
    } catch (const io_exception& e) {
      err = e.what();
    }
    *bases = static_cast<char*>(malloc(all.size() + 1));
    memcpy(*bases, all.data(), all.size());
    *offs = static_cast<int64_t*>(malloc(o.size() * sizeof(int64_t)));
    memcpy(*offs, o.data(), o.size() * sizeof(int64_t));
    *n_reads = int64_t(o.size()) - 1;
    strncpy(error, err.c_str(), error_cap - 1);
    error[error_cap - 1] = 0;
  });
}

// The directory side of `biograph create` (biograph_create.cpp:518, 785-811) through the reference's own biograph_dir:
// creates <dir>/{metadata,coverage,qc,analysis} (call first), or -- with sample_readmap_id set -- writes
// metadata/bg_info.json for one sample (biograph_id is taken from <dir>/seqset by save_metadata itself).
int ref_biograph_dir(const char* dir, const char* accession_id, const char* sample_readmap_id) {
  return guarded([&] {
    biograph_dir bg(dir, CREATE_BGDIR);
    if (sample_readmap_id && *sample_readmap_id) {
      biograph_metadata m = bg.get_metadata();
      m.accession_id = accession_id;
      m.samples[accession_id] = sample_readmap_id;
      m.command_history.push_back("oracle/_ref");
      bg.set_metadata(m);
      bg.save_metadata();
    }
  });
}

// Opens a whole BioGraph directory the way the reference's consumers do (biograph_dir(path, READ_BGDIR): directory
// layout + metadata/bg_info.json; seqset_file(bgdir.seqset()); readmap(seqset, bgdir.find_readmap(""))) -- the readmap
// constructor CHECKs that readmap.json's seqset_uuid is the seqset's -- and reports what it sees as "key=value" lines.
int ref_open_biograph(const char* path, const char* sample /* accession id or readmap id; "" = the only one */, char* out,
                      size_t cap) {
  return guarded([&] {
    biograph_dir bg(path, READ_BGDIR);
    auto ss = std::make_shared<seqset>(bg.seqset());
    std::string rm_path = bg.find_readmap(sample ? sample : "");
    readmap rm(ss, rm_path);
    readmap::pair_stats ps = rm.get_pair_stats();
    std::ostringstream os;
    os << "biograph_id=" << bg.biograph_id() << "\naccession_id=" << bg.accession_id() << "\nversion=" << bg.get_metadata().version
       << "\nsamples=" << bg.samples().size() << "\nsample_accession=" << bg.find_readmap_accession(sample ? sample : "")
       << "\nreadmap_path=" << rm_path << "\nseqset_uuid=" << ss->uuid() << "\nseqset_entries=" << ss->size()
       << "\nmax_read_len=" << ss->max_read_len() << "\nreadmap_rows=" << rm.size() << "\nreadmap_seqset_uuid="
       << rm.metadata().seqset_uuid << "\nnum_bases=" << rm.get_num_bases() << "\npaired_reads=" << ps.paired_reads
       << "\nunpaired_reads=" << ps.unpaired_reads << "\npaired_bases=" << ps.paired_bases << "\nunpaired_bases="
       << ps.unpaired_bases << "\nmin_read_len=" << rm.min_read_len() << "\nreadmap_max_read_len=" << rm.max_read_len() << "\n";
    std::string s = os.str();
    if (s.size() + 1 > cap) throw io_exception("ref_open_biograph: buffer too small");
    memcpy(out, s.c_str(), s.size() + 1);
  });
}

int64_t ref_seqset_size(void* h) {
  auto* r = static_cast<ref_run*>(h);
  return r->ss ? int64_t(r->ss->size()) : -1;
}

// sizes / shared: uint16[n]; prev: 4 x ceil(n/64) uint64 words (bit i of row b = entry i has base b in front);
// fixed: uint64[5]; stats: int64[6] (what the expander calls returned, see ref_make_seqset)
int ref_seqset_tables(void* h, uint16_t* sizes, uint16_t* shared, uint64_t* prev, uint64_t* fixed, int64_t* stats) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    const seqset& ss = *r->ss;
    size_t n = ss.size(), words = (n + 63) / 64;
    memset(prev, 0, 4 * words * sizeof(uint64_t));
    for (size_t i = 0; i < n; ++i) {
      sizes[i] = uint16_t(ss.entry_size(i));
      shared[i] = uint16_t(ss.entry_shared(i));
      for (dna_base b : dna_bases())
        if (ss.entry_has_front(i, b)) prev[size_t(int(b)) * words + i / 64] |= 1ULL << (i % 64);
    }
    for (int b = 0; b < 4; ++b) fixed[b] = ss.entry_push_front(0, dna_base(b));  // get_fixed(b) + count(0) = fixed[b]
    fixed[4] = n;
    if (stats) memcpy(stats, r->stats, sizeof(r->stats));
  });
}

// make_readmap::do_make (modules/bio_mapred/make_readmap.cpp:15-22, as SEQSETMain::do_readmap calls it,
// biograph_create.cpp:818-826) over the seqset of ref_make_seqset: the readmap spiral file is written to `path`.
// Corrected reads are given record by record as the corrected_reads kv stream holds them: record r holds reads
// [rec_offs[r], rec_offs[r+1]) -- one read, or two mates.  The records reach make_readmap through the in-memory
// manifest stand-in (ref_stubs/modules/mapred/manifest_parallel.h).
int ref_make_readmap(void* h, const char* path, const char* bases, const int64_t* offs, const int64_t* rec_offs,
                     int64_t n_rec, int is_paired) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    if (!r->ss) throw io_exception("ref_make_readmap: no seqset (run ref_make_seqset first, before ref_members)");
    auto records = std::make_shared<std::vector<std::pair<std::string, corrected_reads>>>();
    records->reserve(size_t(n_rec));
    for (int64_t rec = 0; rec < n_rec; ++rec) {
      corrected_reads cr;
      for (int64_t i = rec_offs[rec]; i < rec_offs[rec + 1]; ++i) {
        cr.emplace_back();
        cr.back().corrected = dna_sequence(std::string(bases + offs[i], size_t(offs[i + 1] - offs[i])));
      }
      records->emplace_back("r" + std::to_string(rec), std::move(cr));
    }
    manifest m;
    m.ref_stub_set_records(records);
    make_readmap::do_make(path, *r->ss, m, is_paired != 0, r->ss->max_read_len());
  });
}

// `biograph merge`'s seqset path (modules/biograph/biograph_merge.cpp:196-285) over the seqsets of n_in runs that
// have done ref_make_seqset: seqset_flat_builder per input, make_mergemap over the flats, one seqset_mergemap per
// input, seqset_merger::build.  Everything goes through in-memory spiral files instead of the temp files.  The merged
// seqset lands in `out` (ref_seqset_tables / ref_members work on it), the mergemap bits in ref_mergemap.
int ref_merge(void** ins, int n_in, void* out) {
  auto* o = static_cast<ref_run*>(out);
  return guarded([&] {
    std::vector<const seqset*> sets;
    for (int i = 0; i < n_in; ++i) {
      auto* r = static_cast<ref_run*>(ins[i]);
      if (!r->ss) throw io_exception("ref_merge: an input has no seqset");
      sets.push_back(r->ss.get());
    }
    std::vector<std::unique_ptr<seqset_flat>> flats;
    std::vector<const seqset_flat*> flat_ptrs;
    for (const seqset* s : sets) {
      spiral_file_create_mem c;
      seqset_flat_builder b(s);
      b.build(c.create());
      spiral_file_open_mem op(c.close());
      flats.emplace_back(new seqset_flat(op.open(), s));
      flat_ptrs.push_back(flats.back().get());
    }
    o->create = make_unique<spiral_file_create_mem>();
    make_mergemap mm(flat_ptrs);
    mm.build();
    size_t total = mm.total_merged_entries();
    std::vector<std::unique_ptr<seqset_mergemap>>& maps = o->mergemap_objs;
    maps.clear();
    std::vector<const seqset_mergemap*> map_ptrs;
    o->mergemaps.clear();
    for (int i = 0; i < n_in; ++i) {
      spiral_file_create_mem c;
      seqset_mergemap_builder b(c.create(), sets[size_t(i)]->uuid(), o->create->uuid(), total);
      mm.fill_mergemap(unsigned(i), &b);
      spiral_file_open_mem op(c.close());
      maps.emplace_back(new seqset_mergemap(op.open()));
      map_ptrs.push_back(maps.back().get());
      std::vector<uint64_t> bits((total + 63) / 64, 0);
      const bitcount& bc = maps.back()->get_bitcount();
      for (size_t x = 0; x < total; ++x)
        if (bc.get(x)) bits[x / 64] |= 1ULL << (x % 64);
      o->mergemaps.push_back(std::move(bits));
    }
    seqset_merger merger(flat_ptrs, map_ptrs);
    // seqset_merger keeps the merged seqset to itself; reopen what it wrote
    merger.build(o->create->create());
    o->storage = o->create->close();
    o->create.reset();
    o->member_names.clear();
    for (auto& kv : o->storage.paths) o->member_names.push_back(kv.first);
    spiral_file_open_mem op(o->storage);
    o->ss = make_unique<seqset>(op.open());
  });
}
// make_readmap::fast_migrate (modules/bio_mapred/make_readmap.cpp:459-520, as MergeSEQSETMain runs it per sample,
// biograph_merge.cpp:291-312): the readmap FILE of input `input` moves to the merged seqset of `h` (after ref_merge).
int ref_fast_migrate(void* h, int input, const char* old_readmap_path, const char* new_readmap_path) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    if (input < 0 || size_t(input) >= r->mergemap_objs.size()) throw io_exception("ref_fast_migrate: no such input");
    std::unique_ptr<readmap> old_rm = readmap::open_anonymous_readmap(old_readmap_path);
    spiral_file_create_mmap c(new_readmap_path);
    make_readmap::fast_migrate(*old_rm, *r->mergemap_objs[size_t(input)], c.create());
  });
}
int64_t ref_mergemap(void* h, int input, const uint64_t** bits) {
  auto* r = static_cast<ref_run*>(h);
  if (input < 0 || size_t(input) >= r->mergemaps.size()) return -1;
  *bits = r->mergemaps[size_t(input)].data();
  return int64_t(r->mergemaps[size_t(input)].size());
}
// the flat sequence of every entry of the run's seqset, through the reference's seqset_flat (ASCII, offs[n + 1])
int ref_flat(void* h, char** seq, int64_t** offs, int64_t* n_out) {
  auto* r = static_cast<ref_run*>(h);
  return guarded([&] {
    if (!r->ss) throw io_exception("ref_flat: no seqset");
    spiral_file_create_mem c;
    seqset_flat_builder b(r->ss.get());
    b.build(c.create());
    spiral_file_open_mem op(c.close());
    seqset_flat flat(op.open(), r->ss.get());
    std::string all;
    std::vector<int64_t> o{0};
    for (auto it = flat.begin(); it != flat.end(); ++it) {
      all += (*it).as_string();
      o.push_back(int64_t(all.size()));
    }
    *seq = static_cast<char*>(malloc(all.size() + 1));
    memcpy(*seq, all.data(), all.size());
    *offs = static_cast<int64_t*>(malloc(o.size() * sizeof(int64_t)));
    memcpy(*offs, o.data(), o.size() * sizeof(int64_t));
    *n_out = int64_t(o.size()) - 1;
  });
}
void ref_free(void* p) { free(p); }

// the members of the in-memory spiral file exactly as the reference's encoders wrote them (closes the file)
int64_t ref_members(void* h) {
  auto* r = static_cast<ref_run*>(h);
  int64_t n = -1;
  guarded([&] {
    if (r->create) {
      r->ss.reset();
      r->storage = r->create->close();
      r->create.reset();
      r->member_names.clear();
      for (auto& kv : r->storage.paths) r->member_names.push_back(kv.first);
    }
    n = int64_t(r->member_names.size());
  });
  return n;
}
const char* ref_member_name(void* h, int64_t i) { return static_cast<ref_run*>(h)->member_names[size_t(i)].c_str(); }
int64_t ref_member_data(void* h, int64_t i, const char** data) {
  auto* r = static_cast<ref_run*>(h);
  auto& mb = r->storage.paths[r->member_names[size_t(i)]];
  *data = mb.data();
  return int64_t(mb.size());
}

}  // extern "C"

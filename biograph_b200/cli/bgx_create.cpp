// bgx-create -- `biograph create` for the seqset-construction path on a B200: reads in, .bg out.
//
// Takes the flags of SEQSETMain (modules/biograph/biograph_create.cpp:258-335), validates them with
// the reference's rules and messages (:345-430, :483-519), runs the stages through the facade
// (include/bgx_build_seqset.hpp -> C ABI -> CUDA) in the reference's order
//   import -> kmerization -> read_correction -> make_seqset -> make_readmap -> metadata
// (:540-811) and writes the BioGraph directory the way biograph_dir does
// (modules/bio_base/biograph_dir.cpp:11-80):
//   <out>/seqset                         spiral file (ZIP64, members as seqset.cpp:19-44)
//   <out>/coverage/<sha1>.readmap        spiral file, named by its SHA-1 (biograph_create.cpp:820-828)
//   <out>/metadata/bg_info.json          biograph_metadata: version, biograph_id, accession_id, samples
//   <out>/qc/create_stats.json           the counters of :796-808 + stage timings
//   <out>/qc/create_log.txt, <out>/qc/kmer_quality_report.html, <out>/analysis/
// Inputs: FASTQ, plain or gzip (zlib), single, --pair <second file> or --interleaved; BAM (BGZF through
// zlib, records decoded here the way read_importer does through htslib: secondary / supplementary
// records skipped, reverse-strand records reverse-complemented, mates joined by read name) and CRAM 3.0
// (cram_reader.hpp; bases rebuilt from <ref>/source.fasta).  There is no CPU fallback: the stages need
// a CUDA device.
#include <sys/stat.h>
#include <zlib.h>

#include <chrono>
#include <cstdarg>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <unordered_map>

#include "bgx_build_seqset.hpp"
#include "cli_util.hpp"
#include "cram_reader.hpp"

using namespace bgx_cli;

namespace {

const char* kVersion = "7.1.2-dev";  // versions.bzl:12 BIOGRAPH_VERSION of the reference commit

struct Args {
  std::string out, ref, id, format = "auto", tmp, stats_file;
  std::vector<std::string> reads, pairs;
  bool interleaved = false, force = false, allow_long_reads = false, keep_tmp = false;
  std::string min_kmer_count = "5", kmer_size = "30", trim_after_portion = "0.7", max_corrections = "8", min_good_run = "2",
              min_reads = "0.4", warn_reads = "0.7", tmp_encoding = "gzip1", sample_reads = "0", cut_reads, overrep = "0";
  int device = 0;
  bool dump_reads = false;   // test hook: import only, print the reads (no GPU needed for BAM input)
};

// validate_param / validate_float_param (biograph_create.cpp:337-375): same messages
size_t validate_param(const std::string& param, const std::string& value, size_t lo, size_t hi) {
  size_t v;
  try {
    size_t pos = 0;
    v = std::stoull(value, &pos);
  } catch (const std::exception&) {
    throw std::runtime_error(param + " must specify an integer");
  }
  if (v < lo) throw std::runtime_error(fmt("%s must specify an integer >= %zu", param.c_str(), lo));
  if (v > hi) throw std::runtime_error(fmt("%s must specify an integer <= %zu", param.c_str(), hi));
  return v;
}
float validate_float_param(const std::string& param, const std::string& value, float lo, float hi) {
  float v;
  try {
    v = std::stof(value);
  } catch (const std::exception&) {
    throw std::runtime_error(param + " must specify a floating point number");
  }
  if (v < lo) throw std::runtime_error(fmt("%s must specify a floating point number >= %f", param.c_str(), lo));
  if (v > hi) throw std::runtime_error(fmt("%s must specify a floating point number <= %f", param.c_str(), hi));
  return v;
}

void usage() {
  std::cerr << "bgx-create version " << kVersion << " (" << bgx_version() << ")\n\n"
            << "Usage: bgx-create [OPTIONS] --reads <file> --ref <refdir> --out <biograph> [--pair <fastq pairs>] [...]\n\n"
            << "Convert reads to BioGraph format.\n\n"
               "  --out arg                       Output BioGraph name (.bg)\n"
               "  --ref arg                       Reference directory (or FASTA; only its size is used here)\n"
               "  --reads, --in arg               Input file to process (fastq, fastq.gz, bam, cram; - for STDIN)\n"
               "  --format arg (=auto)            Input file format when using STDIN\n"
               "  --interleaved                   Input reads are interleaved (fastq only)\n"
               "  --pair arg                      Second input file containing read pairs (fastq only)\n"
               "  --id arg                        Optional accession ID for this sample\n"
               "  -f, --force                     Overwrite existing BioGraph\n"
               "  --min-kmer-count arg (=5)       Minimum kmer count (min 1)\n"
               "  --kmer-size arg (=30)           The size of kmers to use for kmer generation\n"
               "  --trim-after-portion arg (=0.7) Trim the end of reads until they pass read correction\n"
               "  --max-corrections arg (=8)      Correct up to the specified number of bases\n"
               "  --min-good-run arg (=2)         Minimum number of good bases between corrections\n"
               "  --min-reads arg (=0.4)          Minimum fraction of reads that must survive read correction\n"
               "  --warn-reads arg (=0.7)         Warn when this fraction of reads does not survive read correction\n"
               "  --sample-reads arg (=0)         If non-zero, sample this portion of the input reads\n"
               "  --cut-reads arg                 e.g. 10-100: only use the 10th through the 100th base of each read\n"
               "  --device arg (=0)               CUDA device ordinal\n"
               "  (accepted for compatibility, no effect here: --tmp, --keep-tmp, --threads, --max-mem, --tmp-encoding,\n"
               "   --cache, --stats)\n";
}

Args parse(int argc, char** argv) {
  Args a;
  std::vector<std::string> positional;
  auto need = [&](int& i) -> std::string {
    if (i + 1 >= argc) die(std::string("the required argument for option '") + argv[i] + "' is missing");
    return argv[++i];
  };
  for (int i = 1; i < argc; ++i) {
    std::string o = argv[i], v;
    const size_t eq = o.find('=');
    bool has_v = false;
    if (o.rfind("--", 0) == 0 && eq != std::string::npos) { v = o.substr(eq + 1); o = o.substr(0, eq); has_v = true; }
    auto val = [&]() { return has_v ? v : need(i); };
    if (o == "--out") a.out = val();
    else if (o == "--ref") a.ref = val();
    else if (o == "--reads" || o == "--in") a.reads.push_back(val());
    else if (o == "--pair") a.pairs.push_back(val());
    else if (o == "--format") a.format = val();
    else if (o == "--interleaved") a.interleaved = true;
    else if (o == "--id") a.id = val();
    else if (o == "--force" || o == "-f") a.force = true;
    else if (o == "--min-kmer-count") a.min_kmer_count = val();
    else if (o == "--kmer-size") a.kmer_size = val();
    else if (o == "--trim-after-portion") a.trim_after_portion = val();
    else if (o == "--max-corrections") a.max_corrections = val();
    else if (o == "--min-good-run") a.min_good_run = val();
    else if (o == "--min-reads") a.min_reads = val();
    else if (o == "--warn-reads") a.warn_reads = val();
    else if (o == "--allow-long-reads") a.allow_long_reads = true;
    else if (o == "--tmp-encoding") a.tmp_encoding = val();
    else if (o == "--overrep-threshold") a.overrep = val();
    else if (o == "--sample-reads") a.sample_reads = val();
    else if (o == "--cut-reads") a.cut_reads = val();
    else if (o == "--stats") a.stats_file = val();
    else if (o == "--tmp") a.tmp = val();
    else if (o == "--keep-tmp") a.keep_tmp = true;
    else if (o == "--threads" || o == "--max-mem" || o == "--cache" || o == "--debug" || o == "--sys-err-thresh" || o == "--rnd-err-thresh" ||
             o == "--dump-kmers") { if (o != "--cache" && o != "--debug") (void)val(); }
    else if (o == "--device") a.device = atoi(val().c_str());
    else if (o == "--dump-reads") a.dump_reads = true;
    else if (o == "--help" || o == "-h") { usage(); exit(0); }
    else if (o.rfind("-", 0) == 0 && o != "-") die("unrecognised option '" + o + "'");
    else positional.push_back(o);
  }
  // positional: in, ref, out (biograph_create.cpp:330-332)
  size_t pi = 0;
  if (a.reads.empty() && pi < positional.size()) a.reads.push_back(positional[pi++]);
  if (a.ref.empty() && pi < positional.size()) a.ref = positional[pi++];
  if (a.out.empty() && pi < positional.size()) a.out = positional[pi++];
  if (a.out.empty()) die("the option '--out' is required but missing");
  if (a.reads.empty()) die("the option '--reads' is required but missing");
  return a;
}

// ---- FASTQ input (modules/bio_format/fastq.cpp:40-126 semantics live in bgx_add_reads_fastq) ---------------
struct LineReader {   // plain or gzip, through zlib
  gzFile f = nullptr;
  std::string name;
  uint64_t linenum = 0;   // fastq_reader::m_linenum of the record parser reading from here
  explicit LineReader(const std::string& path) : name(path) {
    f = path == "/dev/stdin" ? gzdopen(0, "rb") : gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Could not open " + path + " for reading");
    gzbuffer(f, 1 << 20);
  }
  ~LineReader() { if (f) gzclose(f); }
  // appends up to max_bytes of whole lines to out; false at end of file
  bool read_chunk(std::string& out, size_t max_bytes) {
    const size_t start = out.size();
    out.resize(start + max_bytes);
    const int n = gzread(f, &out[start], (unsigned)max_bytes);
    if (n < 0) throw std::runtime_error("read error in " + name);
    out.resize(start + (size_t)n);
    return n > 0;
  }
};

// The GPU parser takes well-formed text only: records of exactly four lines, the whole file ending in a newline.
// Anything else -- a malformed record, blank lines between records (which the reference skips, fastq.cpp:45-58), an
// unterminated last line -- raises this, and the caller goes on with the host record parser from the first byte that
// was not handed over yet, which gives the reference's verdict: the same reads, or the same message with the line
// number counted over the whole file.
struct NeedHostParser {};

// hands whole FASTQ records (4 lines) to the GPU parser; the tail of a chunk that is not a whole
// record stays in `carry`.  `lines_done` counts the lines that went to the device.
uint64_t feed_fastq_text(bgx_bs::session& s, std::string& carry, bool last, uint64_t* lines_done, unsigned unit_lines) {
  // unit_lines: 4 = whole records; 8 = whole pairs of an interleaved file, so that the device never holds half a pair
  size_t lines = 0, cut = 0;
  for (size_t i = 0; i < carry.size(); ++i)
    if (carry[i] == '\n' && (++lines % unit_lines) == 0) cut = i + 1;
  uint64_t n = 0;
  if (cut) {
    std::lock_guard<std::mutex> l(s.add_mutex());
    if (bgx_add_reads_fastq(s.ctx(), carry.data(), cut, &n) != 0) throw NeedHostParser();   // nothing was appended
    carry.erase(0, cut);
    *lines_done += lines - lines % unit_lines;
  }
  (void)last;   // what is left at the end of the file is the caller's to hand to the host parser
  return n;
}

// One FASTQ record on the host (paired files, filters, the read dump), with fastq_reader::read's rules and messages
// (modules/bio_format/fastq.cpp:40-126; the importer reads without keeping qualities, read_importer.cpp:633-635):
// blank lines before a record are skipped, every line is checked, and errors carry the reader's line count.
//   rec[0] = id line, rec[1] = bases, rec[2] = '+' line, rec[3] = qualities.
// Lines end at '\n' and lose one trailing '\r' (io.cpp:177-196).  A last line without a newline is handed out without
// being consumed (io.cpp:215-223), so the reader sees it again as the next record's id line and fails there: the
// reference cannot import a FASTQ whose last line is unterminated, and neither can this.
struct FastqLines {
  LineReader& r;
  std::string& buf;
  size_t& pos;
  uint64_t& linenum;
  // false: end of input and nothing left
  bool readline(std::string& line) {
    for (;;) {
      const size_t nl = buf.find('\n', pos);
      if (nl != std::string::npos) {
        line.assign(buf, pos, nl - pos);
        pos = nl + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return true;
      }
      buf.erase(0, pos);
      pos = 0;
      if (!r.read_chunk(buf, 8 << 20)) {
        if (buf.empty()) return false;
        line = buf;   // unterminated last line: not consumed
        return true;
      }
    }
  }
};

bool next_record(LineReader& r, std::string& buf, size_t& pos, std::string rec[4]) {
  FastqLines in{r, buf, pos, r.linenum};
  auto fail = [&](const char* what) { throw std::runtime_error("line " + std::to_string(r.linenum) + ": " + what); };
  for (;;) {   // until a non-blank line or the end
    ++r.linenum;
    if (!in.readline(rec[0])) return false;
    if (!rec[0].empty()) break;
  }
  if (rec[0].size() < 2) fail("Sequence id too short");
  if (rec[0][0] != '@') fail("Sequence id missing @");
  if (!in.readline(rec[1])) fail("End of file while reading sequence line");
  ++r.linenum;
  if (rec[1].empty()) fail("Expecting sequence, found empty line");
  if (rec[1].find_first_not_of("ACGTN") != std::string::npos) fail("Sequence contains unexpected characters");
  if (!in.readline(rec[2])) fail("End of file while reading + line");
  ++r.linenum;
  if (rec[2].empty()) fail("Expecting +, found empty line");
  if (rec[2][0] != '+') fail("Expecting + as first char of line");
  if (!in.readline(rec[3])) fail("End of file while reading quality line");
  ++r.linenum;
  if (rec[3].size() != rec[1].size()) fail("Quality line not same length as sequence");
  return true;
}

// ---- small utilities ---------------------------------------------------------------------------------------
// ---- BAM input (modules/build_seqset/read_importer.cpp:182-266,483-575; htslib there, zlib here) --------------
// A BAM file is a series of BGZF blocks = gzip members, which zlib's gzread inflates back to back:
//   "BAM\1", l_text, text, n_ref, { l_name, name, l_ref } x n_ref, then records
//   block_size | refID pos l_read_name mapq bin n_cigar_op flag l_seq next_refID next_pos tlen |
//   read_name (NUL terminated) | cigar (4 x n_cigar_op) | seq (4-bit codes, (l_seq + 1) / 2 bytes) | qual | tags
using BamRecord = bgx_cli::CramRecord;   // read name, bases as sequenced, BAM flags

struct BamReader {
  gzFile f = nullptr;
  std::string name;
  std::vector<uint8_t> buf;
  uint64_t n_records = 0;
  uint64_t records() const { return n_records; }
  explicit BamReader(const std::string& path) : name(path) {
    f = path == "/dev/stdin" ? gzdopen(0, "rb") : gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Unable to open file " + path);            // read_importer.cpp:490-492
    gzbuffer(f, 1 << 20);
    char magic[4];
    if (!fill(magic, 4) || memcmp(magic, "BAM\1", 4) != 0) throw std::runtime_error(path + " is not a valid BAM file.");   // :513-515
    const int32_t l_text = i32();
    skip((uint32_t)l_text);
    const int32_t n_ref = i32();
    for (int32_t i = 0; i < n_ref; ++i) {
      const int32_t l_name = i32();
      skip((uint32_t)l_name + 4);
    }
  }
  ~BamReader() { if (f) gzclose(f); }
  bool next(BamRecord& r) {
    int32_t block_size;
    const int got = gzread(f, &block_size, 4);
    if (got == 0) return false;                                                    // sam_read1 == -1: EOF
    if (got != 4 || block_size < 32) throw std::runtime_error("sam_read1 returned -2 when reading " + name);   // :545-548
    buf.resize((size_t)block_size);
    if (!fill(buf.data(), (size_t)block_size)) throw std::runtime_error("sam_read1 returned -2 when reading " + name);
    const uint8_t* b = buf.data();
    const uint32_t l_read_name = b[8], n_cigar = rd16(b + 12), l_seq = rd32(b + 16);
    r.flag = rd16(b + 14);
    const size_t seq_off = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar;
    if (l_read_name == 0 || seq_off + (l_seq + 1) / 2 + l_seq > (size_t)block_size)
      throw std::runtime_error("sam_read1 returned -4 when reading " + name);
    r.qname.assign(reinterpret_cast<const char*>(b + 32), l_read_name - 1);
    r.seq.resize(l_seq);
    static const char nt16[] = "=ACMGRSVTWYHKDBN";                                 // htslib seq_nt16_str
    for (uint32_t i = 0; i < l_seq; ++i) r.seq[i] = nt16[(b[seq_off + (i >> 1)] >> ((~i & 1) << 2)) & 0xf];   // bam_seqi
    if (r.flag & 0x10) {                                                           // BAM_FREVERSE: reverse_complement_iupac_string
      static const char comp[] = "TVGH..CD..M.KN...YSA.BW.R.";                     // modules/bio_base/dna_base_set.cpp:78-79
      std::reverse(r.seq.begin(), r.seq.end());
      for (char& c : r.seq) {
        const int idx = c - 'A';
        if (idx >= 0 && idx < 26) c = comp[idx];
      }
    }
    ++n_records;
    return true;
  }

 private:
  static uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
  static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
  bool fill(void* dst, size_t n) {
    char* p = static_cast<char*>(dst);
    while (n) {
      const int r = gzread(f, p, (unsigned)std::min<size_t>(n, 1u << 30));
      if (r <= 0) return false;
      p += r; n -= (size_t)r;
    }
    return true;
  }
  int32_t i32() {
    int32_t v;
    if (!fill(&v, 4)) throw std::runtime_error(name + " is not a valid BAM file.");
    return v;
  }
  void skip(uint32_t n) {
    buf.resize(n);
    if (n && !fill(buf.data(), n)) throw std::runtime_error(name + " is not a valid BAM file.");
  }
};

// read_importer_base::bam_process_line + bam_output_unpaired (read_importer.cpp:182-203,388-470): secondary and
// supplementary records are skipped; a record flagged as paired waits for its mate by read name, mates leave
// together; whatever never met a mate is a single read at the end.  pair(a, b) / single(a) receive the reads;
// returns the number of reads imported.  The k-mer counter only knows ACGT and N (dna_base(char) throws on
// anything else, modules/bio_base/dna_base.h:38-56), so other IUPAC codes are refused with its message.
template <typename Reader, typename PairFn, typename SingleFn>
uint64_t import_alignments(Reader& in, const std::string& path, bool* got_paired, PairFn pair, SingleFn single) {
  BamRecord rec;
  std::unordered_map<std::string, std::string> pair_cache;   // qname -> the mate seen first
  uint64_t n = 0;
  auto check = [](const std::string& seq) {
    for (char c : seq)
      if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') throw std::runtime_error(fmt("Failed conversion of dna_base, c = '%c'", c));
  };
  while (in.next(rec)) {
    if (rec.flag & (0x100 | 0x800)) continue;                // BAM_FSECONDARY, BAM_FSUPPLEMENTARY
    check(rec.seq);
    ++n;
    if (rec.flag & 0x1) {                                    // BAM_FPAIRED
      *got_paired = true;
      auto it = pair_cache.find(rec.qname);
      if (it != pair_cache.end()) {
        pair(rec.seq, it->second);                           // add_paired_read(qname, this record, the cached mate)
        pair_cache.erase(it);
      } else {
        pair_cache.emplace(rec.qname, rec.seq);
      }
    } else {
      single(rec.seq);
    }
  }
  if (in.records() == 0) std::cerr << "WARNING: " << path << ": no records present\n";
  // deterministic order for the leftovers (the reference walks a hash map)
  std::vector<std::string> names;
  for (const auto& kv : pair_cache) names.push_back(kv.first);
  std::sort(names.begin(), names.end());
  for (const std::string& q : names) single(pair_cache[q]);
  return n;
}

// --sample-reads and --cut-reads, per record (a single read or a pair of mates)
struct RecordFilter {
  double sample = 0, accum = 0.5;        // read_importer_state::process, biograph_create.cpp:125-132 (m_sample_accum = 0.5, :191)
  unsigned cut_start = 0, cut_end = 0;   // read_batch::cut_reads, bs/read_importer.cpp:157-169; 0-0 = off
  bool active() const { return sample != 0 || cut_end != 0; }
  bool keep() {
    if (sample == 0) return true;
    accum += sample;
    if (accum > 1) { accum -= 1; return true; }
    return false;
  }
  std::string cut(const std::string& seq) const {
    if (!cut_end) return seq;
    const size_t this_end = std::min<size_t>(cut_end, seq.size());
    if (this_end <= cut_start)   // CHECK_GT(this_end, start) aborts the reference here
      throw std::runtime_error(fmt("Check failed: this_end > start (%zu vs. %u): a read is too short for --cut-reads", this_end, cut_start));
    return seq.substr(cut_start, this_end - cut_start);
  }
};

// validate_cut_param (biograph_create.cpp:376-412)
std::pair<unsigned, unsigned> validate_cut_param(const std::string& param, const std::string& value) {
  if (value.empty()) return {0, 0};
  const size_t it = value.find('-');
  if (it == std::string::npos) throw std::runtime_error(param + " must specify a range separated by a dash");
  unsigned v[2];
  const std::string part[2] = {value.substr(0, it), value.substr(it + 1)};
  for (int i = 0; i < 2; ++i) {
    try {
      v[i] = (unsigned)std::stoul(part[i]);
    } catch (const std::exception&) {
      throw std::runtime_error(param + " must specify a numerical range; couldn't parse " + part[i] + " as a number");
    }
  }
  if (v[1] <= v[0]) throw std::runtime_error(fmt("%s must specify a nonzero range; %u must be less than %u", param.c_str(), v[0], v[1]));
  return {v[0], v[1]};
}

// Where imported records go (read_importer_base::read_batch::add_paired_read / add_unpaired_read).  `text`, when
// set, takes whole-record FASTQ text instead (the GPU parser): used for plain single-file input without filters.
struct ReadSink {
  std::function<void(const std::string&, const std::string&)> pair;
  std::function<void(const std::string&)> single;
  std::function<uint64_t(std::string& carry, bool last, uint64_t* lines_done)> text;
};

// the import stage over every --reads / --pair argument (SEQSETMain::run, biograph_create.cpp:575-627); returns
// the number of reads read (before sampling, as m_read_count); *exit_code != 0: a refusal was printed
uint64_t import_inputs(const Args& a, RecordFilter& filt, bool* got_paired, ReadSink& sink, int* exit_code) {
  uint64_t read_count = 0;
  auto put_pair = [&](const std::string& x, const std::string& y) { if (filt.keep()) sink.pair(filt.cut(x), filt.cut(y)); };
  auto put_single = [&](const std::string& x) { if (filt.keep()) sink.single(filt.cut(x)); };
  // read_importer.cpp:688-693: the last read of an interleaved file with an odd number of reads is counted and dropped
  auto odd_interleaved = [&] { std::cerr << "Warning: interleaved fastq specified, but read an odd number of reads.\n"; };
  for (size_t i = 0; i < a.reads.size(); ++i) {
    std::string in_reads = a.reads[i] == "-" ? "/dev/stdin" : a.reads[i];
    const std::string in_pairs = a.pairs.empty() ? "" : a.pairs[i];
    std::string in_format = a.format;
    if (in_format == "auto") {  // :583-606
      if (a.interleaved) { std::cerr << "--interleaved specified. Assuming fastq format.\n"; in_format = "fastq"; }
      else if (!in_pairs.empty()) { std::cerr << "--pair specified. Assuming fastq format.\n"; in_format = "fastq"; }
      else if (ends_with(in_reads, ".bam")) in_format = "bam";
      else if (ends_with(in_reads, ".cram")) in_format = "cram";
      else if (ends_with(in_reads, ".fq") || ends_with(in_reads, ".fq.gz") || ends_with(in_reads, ".fastq") || ends_with(in_reads, ".fastq.gz")) in_format = "fastq";
      else if (in_reads == "/dev/stdin") in_format = "bam";
      else {
        std::cerr << "Cannot determine the input file type of " << in_reads << ".\n"
                  << "Input file does not end in .bam .cram .fq .fastq .fq.gz or .fastq.gz.\nPlease specify --format.\n";
        *exit_code = 1;
        return read_count;
      }
    }
    if (in_format == "bam") {
      BamReader in(in_reads);
      read_count += import_alignments(in, in_reads, got_paired, put_pair, put_single);
    } else if (in_format == "cram") {
      // CRAM_OPT_REFERENCE = <ref dir>/source.fasta (read_importer.cpp:498-509)
      CramReader in(in_reads, a.ref);
      read_count += import_alignments(in, in_reads, got_paired, put_pair, put_single);
      if (getenv("BGX_CRAM_STATS"))
        for (const auto& kv : in.stats()) std::cerr << "cram: " << kv.first << ": " << kv.second << "\n";
    } else if (!in_pairs.empty()) {
      // two files in step: mates leave together
      LineReader r1(in_reads), r2(in_pairs);
      std::string b1, b2, rec1[4], rec2[4];
      size_t p1 = 0, p2 = 0;
      bool second_done = false;
      for (;;) {
        // read_importer.cpp:680-725: a record of the first file takes the next record of the second as its mate while
        // there is one; whatever either file holds beyond the other is imported as unpaired reads
        const bool h1 = next_record(r1, b1, p1, rec1);
        if (!h1) break;
        if (!second_done && next_record(r2, b2, p2, rec2)) {
          *got_paired = true;
          put_pair(rec1[1], rec2[1]);
          read_count += 2;
        } else {
          second_done = true;
          put_single(rec1[1]);
          ++read_count;
        }
      }
      while (!second_done && next_record(r2, b2, p2, rec2)) {   // "in_file2 may contain more unpaired reads"
        put_single(rec2[1]);
        ++read_count;
      }
    } else if (sink.text && !filt.active() && (a.interleaved || !*got_paired)) {
      // whole-file chunks go to the GPU parser (split, validate, 2-bit pack on the device)
      LineReader r(in_reads);
      std::string carry;
      uint64_t n_file = 0;
      try {
        while (r.read_chunk(carry, 64 << 20)) n_file += sink.text(carry, false, &r.linenum);
        n_file += sink.text(carry, true, &r.linenum);
        if (!carry.empty()) throw NeedHostParser();   // a cut-off record, blank lines, an unterminated last line
      } catch (const NeedHostParser&) {
        // from here on record by record on the host: `carry` holds everything the device has not taken
        std::string rec[4], rec2[4];
        size_t p = 0;
        while (next_record(r, carry, p, rec)) {
          if (a.interleaved) {
            if (!next_record(r, carry, p, rec2)) { odd_interleaved(); ++n_file; break; }
            sink.pair(rec[1], rec2[1]);
            n_file += 2;
          } else {
            sink.single(rec[1]);
            ++n_file;
          }
        }
      }
      if (a.interleaved) *got_paired = *got_paired || n_file >= 2;
      read_count += n_file;
    } else {
      // one file, record by record on the host (filters, the read dump, or single reads after paired input)
      LineReader r(in_reads);
      std::string b, rec[4], rec2[4];
      size_t p = 0;
      while (next_record(r, b, p, rec)) {
        if (a.interleaved) {
          if (!next_record(r, b, p, rec2)) { odd_interleaved(); ++read_count; break; }
          *got_paired = true;
          put_pair(rec[1], rec2[1]);
          read_count += 2;
        } else {
          put_single(rec[1]);
          ++read_count;
        }
      }
    }
  }
  return read_count;
}

uint64_t reference_bases(const std::string& ref) {
  // the reference directory holds the FASTA as source.fasta (biograph reference); a FASTA path works too
  std::vector<std::string> cand = {ref, ref + "/source.fasta", ref + "/reference.fasta"};
  for (const std::string& p : cand) {
    struct stat st;
    if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
    std::ifstream in(p);
    std::string line;
    uint64_t n = 0;
    while (std::getline(in, line))
      if (!line.empty() && line[0] != '>') n += line.size() - (line.back() == '\r');
    return n;
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    Args a = parse(argc, argv);
    // biograph_create.cpp:483-498
    const size_t kmer_size = validate_param("--kmer-size", a.kmer_size, 16, 32);
    const size_t min_kmer_count = validate_param("--min-kmer-count", a.min_kmer_count, 1, 10000000);
    // the range check lets 32 through, the k-mer counter's constructor does not (bs/kmer_counter.cpp:52-54)
    if (kmer_size > 31) throw std::runtime_error("A maximum kmer size of 31 is supported for read correction");
    const float min_corrected_reads = validate_float_param("min-reads", a.min_reads, 0.0f, 1.0f);
    const float warn_corrected_reads = validate_float_param("warn-reads", a.warn_reads, 0.0f, 1.0f);
    const float trim_after_portion = validate_float_param("trim-after-portion", a.trim_after_portion, 0.0f, 1.0f);
    const unsigned max_corrections = (unsigned)validate_param("max-corrections", a.max_corrections, 0, 32);
    const unsigned min_good_run = (unsigned)validate_param("min-good-run", a.min_good_run, 0, 64);
    if (validate_param("overrep-threshold", a.overrep, 0, 10000000) != 0) die("--overrep-threshold other than 0 is not supported by bgx-create");
    RecordFilter filt;
    filt.sample = validate_float_param("sample-reads", a.sample_reads, 0.0f, 1.0f);
    {
      const std::pair<unsigned, unsigned> cut = validate_cut_param("cut-reads", a.cut_reads);
      filt.cut_start = cut.first;
      filt.cut_end = cut.second;
    }
    if (a.allow_long_reads) die("--allow-long-reads is not supported by bgx-create (reads are at most 255 bases)");
    static const std::set<std::string> formats = {"bam", "cram", "fastq", "auto"};
    if (!formats.count(a.format)) die("Invalid input format '" + a.format + "'");
    if (!a.pairs.empty() && a.pairs.size() != a.reads.size())
      die("If pair files are present, there must be the same number of them as read files.");
    if (a.dump_reads) {
      // test hook: the import stage alone, no GPU -- one line per imported read or pair ("a\tb"), then the count
      bool paired = false;
      int rc = 0;
      ReadSink sink;
      sink.pair = [](const std::string& x, const std::string& y) { std::cout << x << "\t" << y << "\n"; };
      sink.single = [](const std::string& x) { std::cout << x << "\n"; };
      const uint64_t n = import_inputs(a, filt, &paired, sink, &rc);
      if (rc) return rc;
      std::cout << "# reads " << n << " paired " << (paired ? 1 : 0) << "\n";
      return 0;
    }
    struct stat st;
    if (!a.force && stat(a.out.c_str(), &st) == 0) die("Refusing to overwrite '" + a.out + "'. Use --force to override.");

    // biograph_dir(m_out, CREATE_BGDIR) (biograph_dir.cpp:25-37)
    mkdir(a.out.c_str(), 0777);
    for (const char* d : {"metadata", "coverage", "qc", "analysis"}) mkdir((a.out + "/" + d).c_str(), 0777);
    if (stat((a.out + "/qc").c_str(), &st) != 0)
      throw std::runtime_error("Attempted to create " + a.out + " but the resulting biograph was not valid. Cannot continue.");
    if (a.id.empty()) {  // fs::canonical(m_out).stem()
      std::string base = a.out;
      while (base.size() > 1 && base.back() == '/') base.pop_back();
      const size_t sl = base.find_last_of('/');
      if (sl != std::string::npos) base = base.substr(sl + 1);
      const size_t dot = base.find_last_of('.');
      a.id = dot == std::string::npos || dot == 0 ? base : base.substr(0, dot);
    }
    if (a.stats_file.empty()) a.stats_file = a.out + "/qc/create_stats.json";
    std::ofstream log(a.out + "/qc/create_log.txt");
    auto splog = [&](const std::string& m) {
      const time_t now = time(nullptr);
      char ts[32];
      strftime(ts, sizeof(ts), "%Y-%m-%d %H:%M:%S", localtime(&now));
      log << ts << " " << m << "\n";
      log.flush();
    };
    {
      std::string cmd;
      for (int i = 0; i < argc; ++i) cmd += std::string(i ? " " : "") + argv[i];
      splog("bgx-create " + std::string(kVersion) + " (" + bgx_version() + "): " + cmd);
    }

    bgx_bs::count_kmer_options ko;
    ko.kmer_size = (unsigned)kmer_size;
    ko.min_count = (unsigned)min_kmer_count;
    ko.device = a.device;
    bgx_bs::read_correction_params rcp;
    rcp.trim_after_portion = trim_after_portion;
    rcp.frc_max_corrections = max_corrections;
    rcp.frc_min_good_run = min_good_run;
    bgx_bs::session sess(ko, rcp);
    Stages stages;
    const auto t_total = std::chrono::steady_clock::now();

    // ---- import (:540-663) ----------------------------------------------------------------------------------
    stages.start();
    std::cerr << "Importing reads\n";
    bgx_bs::kmer_counter counter(sess);
    counter.start_prob_pass();
    bool got_paired = false;
    uint64_t read_count = 0;
    {
      // Mates go in back to back: reads 2i, 2i + 1 of the session are mates once anything is paired.  A single
      // read among pairs (a BAM orphan, an unpaired file after a paired one) gets a one-base partner, which
      // correction drops (shorter than a k-mer), leaving the read a single one for make_readmap (bgx.h: "a pair
      // with one read dropped is a single read").  Pairs after plain single reads would shift that layout.
      bgx_bs::kmer_counter::prob_pass_processor proc(counter);
      uint64_t plain_singles = 0;
      ReadSink sink;
      // read_importer_state::process (biograph_create.cpp:133-139; --allow-long-reads is refused above)
      auto check_len = [](const std::string& x) {
        if (x.size() > 255)
          throw std::runtime_error(fmt("Encountered read of length %ld, which is larger than the maximum read length %d", (long)x.size(), 255));
      };
      sink.pair = [&](const std::string& x, const std::string& y) {
        check_len(x);
        check_len(y);
        if (plain_singles) throw std::runtime_error("paired reads after unpaired ones are not supported by bgx-create: put the paired input first");
        proc.add(x);
        proc.add(y);
      };
      sink.single = [&](const std::string& x) {
        check_len(x);
        proc.add(x);
        if (got_paired) proc.add(std::string("A")); else ++plain_singles;
      };
      sink.text = [&](std::string& carry, bool last, uint64_t* lines_done) {
        proc.flush_all();   // keep the order of the reads across inputs
        if (a.interleaved && plain_singles) throw std::runtime_error("paired reads after unpaired ones are not supported by bgx-create: put the paired input first");
        const uint64_t n = feed_fastq_text(sess, carry, last, lines_done, a.interleaved ? 8 : 4);
        if (!a.interleaved) plain_singles += n;
        return n;
      };
      int rc = 0;
      read_count = import_inputs(a, filt, &got_paired, sink, &rc);
      if (rc) return rc;
      proc.flush_all();
    }
    if (!a.pairs.empty() && !got_paired) throw std::runtime_error("Pair files specified but no pairs were successfully imported");
    if (read_count == 0) throw std::runtime_error("\nNo reads were imported, exiting.");
    counter.close_prob_pass();
    std::cerr << "\nTotal reads imported: " << read_count << std::endl;
    splog(fmt("%lu reads imported", (unsigned long)read_count));
    stages.end("import");

    // ---- kmerization (:665-700) ---------------------------------------------------------------------------------
    stages.start();
    std::cerr << "\nGenerating kmers\n";
    bgx_bs::kmerize_bf_params kp;
    kp.kmer_size = kmer_size;
    kp.min_count = min_kmer_count;
    auto kres = bgx_bs::run_kmerize_subtask(kp, bgx_bs::manifest(), &counter);
    std::unique_ptr<bgx_bs::kmer_set> ks = std::move(kres.first);
    splog(fmt("%lu kmers passed the minimum count of %lu", (unsigned long)ks->size(), (unsigned long)min_kmer_count));
    {
      std::ofstream html(a.out + "/qc/kmer_quality_report.html");
      html << "<html><head><title>k-mer count histogram</title></head><body><h1>k-mer count histogram</h1>\n"
           << "<p>" << ks->size() << " " << kmer_size << "-mers with count &gt;= " << min_kmer_count << "</p>\n<table><tr><th>count</th><th>k-mers</th></tr>\n";
      for (const auto& rec : kres.second[1].records) html << "<tr><td>" << rec.first << "</td><td>" << rec.second << "</td></tr>\n";
      // the points in the form the reference's report carries them (kmerize_bf.cpp:451-454), for tools that scrape it
      html << "</table>\n<script>var kmer_histogram = [";
      for (const auto& rec : kres.second[1].records) html << "{'x':" << rec.first << ",'y':" << rec.second << "},";
      html << "];</script></body></html>\n";
    }
    stages.end("kmerization");

    // ---- read correction (:727-778) -----------------------------------------------------------------------------
    stages.start();
    std::cerr << "\nCorrecting reads\n";
    bgx_bs::correct_reads cr(sess, *ks, rcp);
    cr.add_initial_repo();
    cr.correct_all();
    uint64_t num_corrected_reads = 0, num_corrected_bases = 0;
    unsigned max_read_len = 0;
    for (size_t i = 0; i < cr.size(); ++i) {
      const unsigned len = cr.corrected_length(i);
      if (len) { ++num_corrected_reads; num_corrected_bases += len; max_read_len = std::max(max_read_len, len); }
    }
    const uint64_t ref_size = reference_bases(a.ref);
    const float cov_estimate = ref_size ? num_corrected_bases * 1. / ref_size : 0.f;
    splog(fmt("%0.2fx estimated corrected coverage", cov_estimate));
    const float corrected_pct = num_corrected_reads * 1. / read_count;
    if (corrected_pct < min_corrected_reads) {
      const std::string msg = fmt("Fewer than %2.0f%% of reads (set by --min-reads) were kept after correction (%lu / %lu remain). Cannot continue.",
                                  min_corrected_reads * 100.0, (unsigned long)num_corrected_reads, (unsigned long)read_count);
      splog(msg);
      throw std::runtime_error(msg);
    }
    if (corrected_pct < warn_corrected_reads) {
      const std::string msg = fmt("Warning: Fewer than %2.0f%% of reads (set by --warn-reads) survived correction (%lu / %lu remain)",
                                  warn_corrected_reads * 100.0, (unsigned long)num_corrected_reads, (unsigned long)read_count);
      splog(msg);
      std::cerr << msg << "\n";
    } else {
      splog(fmt("%lu / %lu reads survived read correction.", (unsigned long)num_corrected_reads, (unsigned long)read_count));
    }
    stages.end("read_correction");

    // ---- make_seqset (:914-950) -------------------------------------------------------------------------------------
    stages.start();
    std::cerr << "\nGenerating BioGraph\n";
    bgx_bs::expander expand(sess, a.keep_tmp);
    const size_t n1 = expand.sort_and_dedup("", "initial", "init_sorted", "init_expanded", 7, 255);
    splog(fmt("Initial sort and dedup: %lu entries", (unsigned long)n1));
    const size_t n_final = expand.sort_and_dedup("pass2_sorted", "pass2_expanded", "complete", "", 0, 0);
    splog(fmt("Final sort and dedup: %lu entries", (unsigned long)n_final));
    bgx_bs::builder b(sess);
    b.build_chunks("complete", a.keep_tmp);
    const std::string uuid = make_uuid();
    bgx_bs::seqset_tables tables = b.make_seqset(a.out + "/seqset", bgx_bs::null_progress_handler, uuid);
    stages.end("make_seqset");

    // ---- make_readmap (:818-831) ----------------------------------------------------------------------------------------
    stages.start();
    std::cerr << "\nCalculating coverage...\n";
    const std::string tmp_readmap = a.out + "/coverage/tmp.readmap";
    bgx_bs::make_readmap::do_make(tmp_readmap, sess, uuid, got_paired, max_read_len);
    const std::string readmap_sha = sha1_file(tmp_readmap);
    if (rename(tmp_readmap.c_str(), (a.out + "/coverage/" + readmap_sha + ".readmap").c_str()) != 0)
      throw std::runtime_error("cannot rename the readmap");
    stages.end("make_readmap");

    // ---- metadata (:785-811) ------------------------------------------------------------------------------------------------
    stages.start();
    {
      std::ofstream os(a.out + "/metadata/bg_info.json");
      os << "{\"accession_id\":" << json_str(a.id) << ",\"biograph_id\":" << json_str(uuid) << ",\"command_history\":[],\"samples\":{"
         << json_str(a.id) << ":" << json_str(readmap_sha) << "},\"version\":" << json_str(kVersion) << "}";
      if (!os.good()) throw std::runtime_error("Could not write to " + a.out + "/metadata/bg_info.json");
    }
    stages.end("metadata");
    {
      std::ofstream os(a.stats_file);
      os.precision(17);
      os << "{\"command\":\"create\",\"version\":" << json_str(kVersion) << ",\"accession_id\":" << json_str(a.id) << ",\"reference\":"
         << json_str(a.ref) << ",\"imported_reads\":" << read_count << ",\"coverage\":" << cov_estimate << ",\"corrected_reads\":"
         << num_corrected_reads << ",\"corrected_bases\":" << num_corrected_bases << ",\"avg_bases_per_read\":"
         << (num_corrected_reads ? num_corrected_bases * 1. / num_corrected_reads : 0.0) << ",\"corrected_pct\":" << corrected_pct
         << ",\"uuid\":" << json_str(uuid) << ",\"entries\":" << tables.num_entries << ",\"timings\":[";
      for (const auto& s : stages.t) os << "{" << json_str(s.first) << ":" << (long)s.second << "},";
      os << "{\"total\":" << (long)std::chrono::duration<double>(std::chrono::steady_clock::now() - t_total).count() << "}]}";
    }
    splog(fmt("%lu entries in the seqset", (unsigned long)tables.num_entries));
    std::cerr << "\n" << a.out << " created.\n";
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
}

// mock_tables_bgx.cpp -- a stand-in for the C ABI that serves FINISHED results from raw files (written by the test
// from the CPU oracle's result), so that the HOST side of the drop-in -- the facade's spiral-file writers, and the whole
// bgx-create executable with its .bg directory, metadata and stats -- runs without a GPU.  What it writes is then
// opened by the reference's OWN readers (oracle/_ref: biograph_dir, spiral_file_open_mmap + seqset, readmap),
// tests/test_ref_reads_facade_files.py.  Test infrastructure; installed as "libbgx.so" in a scratch directory and
// put in front of the real one with LD_LIBRARY_PATH.
//
// The compute entry points (add reads, count, correct, build) succeed without doing anything; the exports serve:
// km_kmers.bin / km_fwd.bin / km_rev.bin / km_flags.bin (every counted k-mer, ascending; filtered by min_count here),
// cr_lens.bin (uint16 per read, 0 = dropped) / cr_bases.bin / cr_corr.bin (uint8 per read), stats.json (optional).
// $BGX_MOCK_TABLES/: meta.txt = "num_entries max_entry_len prev_words sub_words acc_words sizes_bits sizes_max
// shared_bits shared_max"; fixed.bin; prev_bits_<b>.bin, prev_sub_<b>.bin, prev_acc_<b>.bin (b = 0..3);
// sizes_elements.bin, shared_elements.bin (uint64 words).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "bgx.h"

struct bgx_ctx {};
static std::string g_err;

static std::string dir() {
  const char* d = getenv("BGX_MOCK_TABLES");
  return d ? d : ".";
}
static bool slurp(const std::string& name, std::vector<char>* out) {
  FILE* f = fopen((dir() + "/" + name).c_str(), "rb");
  if (!f) { g_err = "mock: cannot open " + name; return false; }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize((size_t)n);
  size_t got = n ? fread(out->data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == (size_t)n;
}
static uint64_t* words_of(const std::string& name, uint64_t* n_words) {
  std::vector<char> b;
  if (!slurp(name, &b)) return nullptr;
  uint64_t* p = (uint64_t*)malloc(b.size() ? b.size() : 8);
  memcpy(p, b.data(), b.size());
  if (n_words) *n_words = b.size() / 8;
  return p;
}
struct meta_t { unsigned long long n, maxlen, pw, sw, aw, sbits, smax, hbits, hmax; };
static bool meta(meta_t* m) {
  std::vector<char> b;
  if (!slurp("meta.txt", &b)) return false;
  b.push_back(0);
  return sscanf(b.data(), "%llu %llu %llu %llu %llu %llu %llu %llu %llu", &m->n, &m->maxlen, &m->pw, &m->sw, &m->aw, &m->sbits,
                &m->smax, &m->hbits, &m->hmax) == 9;
}

extern "C" {
void bgx_default_options(bgx_options* o) { memset(o, 0, sizeof *o); }
const char* bgx_last_error(void) { return g_err.c_str(); }
const char* bgx_version(void) { return "mock-tables"; }
int bgx_create(const bgx_options*, bgx_ctx** out) { *out = new bgx_ctx(); return 0; }
void bgx_destroy(bgx_ctx* c) { delete c; }
void bgx_free(void* p) { free(p); }
int bgx_build_seqset(bgx_ctx*) { return 0; }
int bgx_seqset_layout(bgx_ctx*, uint64_t lay[6]) {
  meta_t m;
  if (!meta(&m)) return 1;
  lay[0] = lay[1] = m.n; lay[2] = 0; lay[3] = m.pw; lay[4] = m.sw; lay[5] = m.aw;
  return 0;
}
int bgx_export_seqset(bgx_ctx*, uint64_t* n_entries, uint32_t* max_entry_len, uint16_t** sizes, uint16_t** shared,
                      uint64_t* prev_bits[4], uint64_t* prev_subaccum[4], uint64_t* prev_accum[4], uint64_t fixed[5]) {
  meta_t m;
  if (!meta(&m)) return 1;
  if (sizes || shared) { g_err = "mock: per-entry arrays are not served"; return 1; }
  if (n_entries) *n_entries = m.n;
  if (max_entry_len) *max_entry_len = (uint32_t)m.maxlen;
  for (int b = 0; b < 4; ++b) {
    const std::string s = std::to_string(b);
    if (prev_bits && !(prev_bits[b] = words_of("prev_bits_" + s + ".bin", nullptr))) return 1;
    if (prev_subaccum && !(prev_subaccum[b] = words_of("prev_sub_" + s + ".bin", nullptr))) return 1;
    if (prev_accum && !(prev_accum[b] = words_of("prev_acc_" + s + ".bin", nullptr))) return 1;
  }
  if (fixed) {
    uint64_t* f = words_of("fixed.bin", nullptr);
    if (!f) return 1;
    memcpy(fixed, f, 40);
    free(f);
  }
  return 0;
}
int bgx_export_varbit(bgx_ctx*, int32_t which, uint64_t** words, uint64_t* n_words, uint32_t* bits_per_value, uint64_t* max_value) {
  meta_t m;
  if (!meta(&m)) return 1;
  *words = words_of(which == 0 ? "sizes_elements.bin" : "shared_elements.bin", n_words);
  if (!*words) return 1;
  *bits_per_value = (uint32_t)(which == 0 ? m.sbits : m.hbits);
  *max_value = which == 0 ? m.smax : m.hmax;
  return 0;
}
// readmap tables: rm_meta.txt = "n_rows"; rm_lens.bin (uint16), rm_ptr.bin, rm_fwd.bin (uint64 words),
// rm_src_<i>.bin / rm_dst_<i>.bin (i = 0 bits, 1 subaccum, 2 accum)
int bgx_build_readmap(bgx_ctx*, int32_t, uint64_t* n_rows, uint16_t** read_lengths, uint64_t** mate_loop_ptr, uint64_t** is_forward,
                      uint64_t* source[3], uint64_t* dest[3]) {
  std::vector<char> b;
  if (!slurp("rm_meta.txt", &b)) return 1;
  b.push_back(0);
  *n_rows = strtoull(b.data(), nullptr, 10);
  *read_lengths = (uint16_t*)words_of("rm_lens.bin", nullptr);
  *mate_loop_ptr = words_of("rm_ptr.bin", nullptr);
  *is_forward = words_of("rm_fwd.bin", nullptr);
  if (!*read_lengths || !*mate_loop_ptr || !*is_forward) return 1;
  for (int i = 0; i < 3; ++i) {
    source[i] = words_of("rm_src_" + std::to_string(i) + ".bin", nullptr);
    dest[i] = words_of("rm_dst_" + std::to_string(i) + ".bin", nullptr);
    if (!source[i] || !dest[i]) return 1;
  }
  return 0;
}
// entry points the facade header references elsewhere; none is reached by the writer test
// ---- bgx-merge: the merged seqset is served like a built one (meta.txt ...); mm_<p>.bin = the mergemap bit words of
// input p over the merged entries.  The inputs bgx-merge hands over are checked against in_<p>_sizes.bin (uint16), so
// that the facade's reading of the input .bg files is part of what is tested.  The bitcount index and the readmap
// migration are computed here from those bits (bitcount.cpp:84-123, make_readmap.cpp:472-486).
static void bitcount3(const std::vector<uint64_t>& bits, uint64_t nbits, uint64_t* out[3]) {
  const uint64_t words = (nbits + 63) / 64, nsub = (nbits + 511) / 512, nacc = (nbits + 1 + 511) / 512;
  uint64_t* b = (uint64_t*)calloc(words + 1, 8);
  uint64_t* sub = (uint64_t*)calloc(nsub + 1, 8);
  uint64_t* acc = (uint64_t*)calloc(nacc + 1, 8);
  memcpy(b, bits.data(), words * 8);
  uint64_t s = 0, total = 0;
  for (uint64_t i = 0; i < words; ++i) {
    if (i % 8 == 0) { acc[i / 8] = total; if (i) sub[i / 8 - 1] = s; s = 0; }
    const uint64_t c = (uint64_t)__builtin_popcountll(bits[i]);
    s = (s << 8) | c;
    total += c;
  }
  for (uint64_t left = words % 8; left; left = (left + 1) % 8) s <<= 8;
  if (nbits) sub[nsub - 1] = s;
  if (nbits % 512 == 0) acc[nbits / 512] = total;
  out[0] = b; out[1] = sub; out[2] = acc;
}
static bool mergemap_bits(uint32_t part, std::vector<uint64_t>* bits, uint64_t* nbits) {
  meta_t m;
  if (!meta(&m)) return false;
  std::vector<char> raw;
  if (!slurp("mm_" + std::to_string(part) + ".bin", &raw)) return false;
  bits->assign((m.n + 63) / 64, 0);
  memcpy(bits->data(), raw.data(), std::min(raw.size(), bits->size() * 8));
  *nbits = m.n;
  return true;
}
int bgx_merge_seqsets(bgx_ctx*, const bgx_seqset_part* parts, uint32_t n, uint64_t) {
  for (uint32_t p = 0; p < n; ++p) {
    std::vector<char> want;
    if (!slurp("in_" + std::to_string(p) + "_sizes.bin", &want)) return 1;
    if (want.size() != parts[p].n_entries * 2 || memcmp(want.data(), parts[p].sizes, want.size()) != 0) {
      g_err = "mock: input " + std::to_string(p) + " was not read from its .bg as written";
      return 1;
    }
  }
  return 0;
}
int bgx_export_mergemap(bgx_ctx*, uint32_t part, uint64_t* out[3], uint64_t* n_bits, uint64_t* n_set) {
  std::vector<uint64_t> bits;
  if (!mergemap_bits(part, &bits, n_bits)) return 1;
  uint64_t set = 0;
  for (uint64_t w : bits) set += (uint64_t)__builtin_popcountll(w);
  *n_set = set;
  bitcount3(bits, *n_bits, out);
  return 0;
}
int bgx_migrate_bits(bgx_ctx*, uint32_t part, const uint64_t* old_bits, uint64_t n_old, uint64_t* out[3], uint64_t* n_bits) {
  std::vector<uint64_t> mm;
  if (!mergemap_bits(part, &mm, n_bits)) return 1;
  std::vector<uint64_t> bits(mm.size(), 0);
  uint64_t e = 0;   // the e-th set bit of the mergemap is where old entry e went
  for (uint64_t x = 0; x < *n_bits; ++x)
    if ((mm[x >> 6] >> (x & 63)) & 1) {
      if (e < n_old && ((old_bits[e >> 6] >> (e & 63)) & 1)) bits[x >> 6] |= 1ull << (x & 63);
      ++e;
    }
  if (e != n_old) { g_err = "mock: the mergemap does not cover the input"; return 1; }
  bitcount3(bits, *n_bits, out);
  return 0;
}
int bgx_export_flat_ascii(bgx_ctx*, uint32_t, uint64_t, uint64_t, char**, uint64_t**) { return 1; }
// Every read handed over is appended to $BGX_MOCK_LOG (if set): "A <bases>" through bgx_add_reads_ascii, "F <bases>"
// through the FASTQ text entry point.
static void log_read(char kind, const char* p, size_t n) {
  const char* path = getenv("BGX_MOCK_LOG");
  if (!path) return;
  FILE* f = fopen(path, "a");
  if (!f) return;
  fprintf(f, "%c %.*s\n", kind, (int)n, p);
  fclose(f);
}
int bgx_add_reads_ascii(bgx_ctx*, const char* bases, const uint64_t* offs, uint64_t n) {
  for (uint64_t r = 0; r < n; ++r) log_read('A', bases + offs[r], (size_t)(offs[r + 1] - offs[r]));
  return 0;
}
// The device parser's contract (biograph_b200/csrc/reads.cu: reads_append_fastq, fq_records_kernel): the text ends in
// a newline; blank lines only after the last record; every record is four lines that pass fastq_reader::read's
// checks; on any error NOTHING is appended.  No '\r' handling.
int bgx_add_reads_fastq(bgx_ctx*, const char* text, uint64_t size, uint64_t* n_reads) {
  if (n_reads) *n_reads = 0;
  if (size == 0) return 0;
  if (text[size - 1] != '\n') { g_err = "Partial line in fastq file (the text must end with a newline)"; return 1; }
  std::vector<std::pair<uint64_t, uint64_t>> lines;   // start, length
  for (uint64_t i = 0, st = 0; i < size; ++i)
    if (text[i] == '\n') { lines.emplace_back(st, i - st); st = i + 1; }
  while (!lines.empty() && lines.back().second == 0) lines.pop_back();
  const uint64_t n = lines.size() / 4;
  for (uint64_t r = 0; r < n; ++r) {
    const auto &id = lines[4 * r], &sq = lines[4 * r + 1], &pl = lines[4 * r + 2], &ql = lines[4 * r + 3];
    bool ok = id.second >= 2 && text[id.first] == '@' && sq.second > 0 && sq.second <= 255 && pl.second > 0 && text[pl.first] == '+' &&
              ql.second == sq.second;
    for (uint64_t i = 0; ok && i < sq.second; ++i) ok = strchr("ACGTN", text[sq.first + i]) != nullptr;
    if (!ok) { g_err = "line " + std::to_string(4 * r + 1) + ": malformed record (mock of the device parser)"; return 1; }
  }
  if (lines.size() % 4) { g_err = "End of file inside a record (mock of the device parser)"; return 1; }
  for (uint64_t r = 0; r < n; ++r) log_read('F', text + lines[4 * r + 1].first, (size_t)lines[4 * r + 1].second);
  if (n_reads) *n_reads = n;
  return 0;
}
int bgx_count_kmers(bgx_ctx*) { return 0; }
int bgx_stats_json(bgx_ctx*, char* buf, size_t cap) {
  std::vector<char> b;
  if (!slurp("stats.json", &b)) b.assign({'{', '}'});
  if (b.size() + 1 > cap) return 1;
  memcpy(buf, b.data(), b.size());
  buf[b.size()] = 0;
  return 0;
}
int bgx_export_kmers(bgx_ctx*, uint32_t min_count, uint64_t* n, uint64_t** kmers, uint32_t** fwd, uint32_t** rev, uint8_t** flags) {
  std::vector<char> k, f, r, fl;
  if (!slurp("km_kmers.bin", &k) || !slurp("km_fwd.bin", &f) || !slurp("km_rev.bin", &r) || !slurp("km_flags.bin", &fl)) return 1;
  const uint64_t total = k.size() / 8;
  const uint64_t* K = (const uint64_t*)k.data();
  const uint32_t *F = (const uint32_t*)f.data(), *R = (const uint32_t*)r.data();
  uint64_t* ok = (uint64_t*)malloc(total * 8 + 8);
  uint32_t *of = (uint32_t*)malloc(total * 4 + 4), *orv = (uint32_t*)malloc(total * 4 + 4);
  uint8_t* ofl = (uint8_t*)malloc(total + 1);
  uint64_t m = 0;
  for (uint64_t i = 0; i < total; ++i)
    if ((uint64_t)F[i] + R[i] >= min_count) { ok[m] = K[i]; of[m] = F[i]; orv[m] = R[i]; ofl[m] = (uint8_t)fl[i]; ++m; }
  *n = m;
  if (kmers) *kmers = ok; else free(ok);
  if (fwd) *fwd = of; else free(of);
  if (rev) *rev = orv; else free(orv);
  if (flags) *flags = ofl; else free(ofl);
  return 0;
}
int bgx_export_reads(bgx_ctx*, uint64_t*, uint16_t**, char**, uint64_t*) { return 1; }
int bgx_correct(bgx_ctx*) { return 0; }
int bgx_seed_uncorrected(bgx_ctx*) { return 1; }
int bgx_export_corrected(bgx_ctx*, uint64_t* n_reads, uint16_t** lens, char** bases, uint64_t* n_bases, uint8_t** corrections,
                         uint16_t** next_fwd, uint16_t** next_rev) {
  std::vector<char> l, b, c;
  if (!slurp("cr_lens.bin", &l) || !slurp("cr_bases.bin", &b) || !slurp("cr_corr.bin", &c)) return 1;
  if (next_fwd || next_rev) { g_err = "mock: seed counts are not served"; return 1; }
  auto dup = [](const std::vector<char>& v) { void* p = malloc(v.size() + 1); memcpy(p, v.data(), v.size()); return p; };
  if (n_reads) *n_reads = l.size() / 2;
  if (n_bases) *n_bases = b.size();
  if (lens) *lens = (uint16_t*)dup(l);
  if (bases) *bases = (char*)dup(b);
  if (corrections) *corrections = (uint8_t*)dup(c);
  return 0;
}
int bgx_dist_unique_id(uint8_t*) { return 1; }
int bgx_dist_init(bgx_ctx*, int32_t, int32_t, const uint8_t*) { return 1; }
}

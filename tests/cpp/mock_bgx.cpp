// mock_bgx.cpp -- a stand-in for the handful of C-ABI entry points bgx_bs::seqset_merger calls, returning canned
// data, so that the facade's HOST logic for a merge (mergemap spiral file, readmap migration file, flat entries)
// can be checked in the GPU-less dev container (tests/test_facade_merge_host.py).  Test infrastructure: never
// linked into the product; the real entry points live in biograph_b200/csrc and are checked on the B200
// (tests/test_zz_merge_gpu.py, tests/test_zz_cli_merge.py).
//
// Canned merge: N = 1000 merged entries; input p has a mergemap bit at x iff x % (p + 2) == 0; migrate_bits
// returns bit x = old bit (x / 2) for even x; flat entries of input p are "ACGT" repeated (i % 5 + 1) times.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "bgx.h"

struct bgx_ctx { int merged = 0; uint32_t n_parts = 0; };
static const uint64_t N = 1000;
static std::string g_err;

static uint64_t* bitcount3(const uint64_t* bits, uint64_t nbits, int which) {
  const uint64_t words = (nbits + 63) / 64, sub = (nbits + 511) / 512, acc = (nbits + 1 + 511) / 512;
  if (which == 0) { uint64_t* o = (uint64_t*)calloc(words ? words : 1, 8); memcpy(o, bits, words * 8); return o; }
  if (which == 1) {
    uint64_t* o = (uint64_t*)calloc(sub ? sub : 1, 8);
    for (uint64_t g = 0; g < sub; ++g) {
      uint64_t s = 0;
      for (int j = 0; j < 8; ++j) { s <<= 8; if (g * 8 + j < words) s |= (uint64_t)__builtin_popcountll(bits[g * 8 + j]); }
      o[g] = s;
    }
    return o;
  }
  uint64_t* o = (uint64_t*)calloc(acc ? acc : 1, 8);
  uint64_t run = 0;
  for (uint64_t g = 0; g < acc; ++g) {
    o[g] = run;
    for (int j = 0; j < 8; ++j) if (g * 8 + j < words) run += (uint64_t)__builtin_popcountll(bits[g * 8 + j]);
  }
  return o;
}

extern "C" {
void bgx_default_options(bgx_options* o) { memset(o, 0, sizeof *o); }
const char* bgx_last_error(void) { return g_err.c_str(); }
const char* bgx_version(void) { return "mock"; }
int bgx_create(const bgx_options*, bgx_ctx** out) { *out = new bgx_ctx(); return 0; }
void bgx_destroy(bgx_ctx* c) { delete c; }
void bgx_free(void* p) { free(p); }
int bgx_merge_seqsets(bgx_ctx* c, const bgx_seqset_part* parts, uint32_t n, uint64_t) {
  if (!parts || n == 0) { g_err = "no inputs"; return 1; }
  c->merged = 1; c->n_parts = n;
  return 0;
}
int bgx_seqset_layout(bgx_ctx*, uint64_t lay[6]) { lay[0] = lay[1] = N; lay[2] = 0; lay[3] = (N + 63) / 64; lay[4] = (N + 511) / 512; lay[5] = (N + 512) / 512; return 0; }
int bgx_export_mergemap(bgx_ctx* c, uint32_t part, uint64_t* out[3], uint64_t* n_bits, uint64_t* n_set) {
  if (!c->merged || part >= c->n_parts) { g_err = "bgx_export_mergemap: no such input"; return 1; }
  uint64_t bits[(N + 63) / 64] = {0};
  uint64_t set = 0;
  for (uint64_t x = 0; x < N; ++x) if (x % (part + 2) == 0) { bits[x >> 6] |= 1ull << (x & 63); ++set; }
  for (int k = 0; k < 3; ++k) out[k] = bitcount3(bits, N, k);
  *n_bits = N; *n_set = set;
  return 0;
}
int bgx_migrate_bits(bgx_ctx* c, uint32_t part, const uint64_t* old_bits, uint64_t n_old, uint64_t* out[3], uint64_t* n_bits) {
  if (!c->merged || part >= c->n_parts) { g_err = "bgx_migrate_bits: no such input"; return 1; }
  uint64_t bits[(N + 63) / 64] = {0};
  for (uint64_t x = 0; x < N && x / 2 < n_old; x += 2) if ((old_bits[(x / 2) >> 6] >> ((x / 2) & 63)) & 1) bits[x >> 6] |= 1ull << (x & 63);
  for (int k = 0; k < 3; ++k) out[k] = bitcount3(bits, N, k);
  *n_bits = N;
  return 0;
}
int bgx_export_flat_ascii(bgx_ctx*, uint32_t, uint64_t first, uint64_t count, char** bases, uint64_t** offs) {
  std::string s;
  uint64_t* o = (uint64_t*)calloc(count + 1, 8);
  for (uint64_t i = 0; i < count; ++i) { o[i] = s.size(); for (uint64_t r = 0; r < (first + i) % 5 + 1; ++r) s += "ACGT"; }
  o[count] = s.size();
  *bases = (char*)malloc(s.size() + 1);
  memcpy(*bases, s.data(), s.size());
  *offs = o;
  return 0;
}
// The import stage of bgx-create: every read handed to the device is appended to the file named by BGX_MOCK_LOG
// ("A <bases>" through bgx_add_reads_ascii, "F <bases>" through the FASTQ text entry point), and the k-mer stage
// stops the run ("mock: the import stage ended").
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
int bgx_add_reads_fastq(bgx_ctx*, const char* text, uint64_t size, uint64_t* n_reads) {
  uint64_t line = 0, start = 0, n = 0;
  for (uint64_t i = 0; i < size; ++i)
    if (text[i] == '\n') {
      if (line % 4 == 1) { log_read('F', text + start, (size_t)(i - start)); ++n; }
      ++line;
      start = i + 1;
    }
  *n_reads = n;
  return 0;
}
int bgx_count_kmers(bgx_ctx*) { g_err = "mock: the import stage ended"; return 1; }
int bgx_stats_json(bgx_ctx*, char* buf, size_t cap) { if (cap) buf[0] = 0; return 1; }
int bgx_export_kmers(bgx_ctx*, uint32_t, uint64_t*, uint64_t**, uint32_t**, uint32_t**, uint8_t**) { return 1; }
int bgx_export_reads(bgx_ctx*, uint64_t*, uint16_t**, char**, uint64_t*) { return 1; }
int bgx_correct(bgx_ctx*) { return 1; }
int bgx_seed_uncorrected(bgx_ctx*) { return 1; }
int bgx_export_corrected(bgx_ctx*, uint64_t*, uint16_t**, char**, uint64_t*, uint8_t**, uint16_t**, uint16_t**) { return 1; }
int bgx_build_seqset(bgx_ctx*) { return 1; }
int bgx_export_seqset(bgx_ctx*, uint64_t*, uint32_t*, uint16_t**, uint16_t**, uint64_t*[4], uint64_t*[4], uint64_t*[4], uint64_t[5]) { return 1; }
int bgx_export_varbit(bgx_ctx*, int32_t, uint64_t**, uint64_t*, uint32_t*, uint64_t*) { return 1; }
int bgx_build_readmap(bgx_ctx*, int32_t, uint64_t*, uint16_t**, uint64_t**, uint64_t**, uint64_t*[3], uint64_t*[3]) { return 1; }
int bgx_dist_unique_id(uint8_t*) { return 1; }
int bgx_dist_init(bgx_ctx*, int32_t, int32_t, const uint8_t*) { return 1; }
}

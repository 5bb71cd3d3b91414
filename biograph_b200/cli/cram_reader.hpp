// cram_reader.hpp -- a CRAM 3.x reader for bgx-create's import stage: records come out as (read name, bases,
// BAM flags), the three things read_importer_base::bam_process_line / bam1_to_unaligned_read use
// (modules/build_seqset/read_importer.cpp:182-266; the reference reads CRAM through htslib's sam_read1 with
// CRAM_OPT_REFERENCE = <ref dir>/source.fasta, :498-509).
//
// Written from the CRAM format specification (version 3.0): file definition, containers, blocks (raw, gzip,
// rANS 4x8 order 0 / 1), compression header (preservation map, data series and tag encodings: EXTERNAL, HUFFMAN,
// BYTE_ARRAY_LEN, BYTE_ARRAY_STOP, BETA, SUBEXP, GAMMA), slices, records and read features; bases of mapped
// records are rebuilt from the reference (FASTA, or the slice's embedded reference) and checked against the
// slice's MD5.  bzip2 / lzma blocks and CRAM 2.x / 3.1 codecs are refused with a message.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace bgx_cli {

namespace cram {

inline std::runtime_error err(const std::string& m) { return std::runtime_error("CRAM: " + m); }

// ---- MD5 (RFC 1321), for the slice reference checksums -------------------------------------------------------
inline void md5(const uint8_t* data, size_t n, uint8_t out[16]) {
  static const uint32_t K[64] = {
      0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
      0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
      0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
      0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
      0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
      0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
  static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                            4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
  uint32_t h[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
  auto block = [&](const uint8_t* p) {
    uint32_t w[16];
    for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[4 * i] | (uint32_t)p[4 * i + 1] << 8 | (uint32_t)p[4 * i + 2] << 16 | (uint32_t)p[4 * i + 3] << 24;
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 64; ++i) {
      uint32_t f;
      int g;
      if (i < 16) { f = (b & c) | (~b & d); g = i; }
      else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
      else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
      else { f = c ^ (b | ~d); g = (7 * i) & 15; }
      const uint32_t t = a + f + K[i] + w[g];
      a = d; d = c; c = b;
      b = b + (t << S[i] | t >> (32 - S[i]));
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
  };
  size_t i = 0;
  for (; i + 64 <= n; i += 64) block(data + i);
  uint8_t tail[128];
  size_t have = n - i;
  memcpy(tail, data + i, have);
  tail[have++] = 0x80;
  while (have % 64 != 56) tail[have++] = 0;
  const uint64_t bits = (uint64_t)n * 8;
  for (int k = 0; k < 8; ++k) tail[have++] = (uint8_t)(bits >> (8 * k));
  for (size_t k = 0; k < have; k += 64) block(tail + k);
  for (int k = 0; k < 4; ++k)
    for (int b = 0; b < 4; ++b) out[4 * k + b] = (uint8_t)(h[k] >> (8 * b));
}

// ---- a byte cursor with the format's variable-length integers ---------------------------------------------------
struct Cursor {
  const uint8_t* p = nullptr;
  const uint8_t* end = nullptr;
  Cursor() = default;
  Cursor(const uint8_t* b, size_t n) : p(b), end(b + n) {}
  void need(size_t n) const { if ((size_t)(end - p) < n) throw err("truncated data"); }
  uint8_t u8() { need(1); return *p++; }
  int32_t i32() { need(4); int32_t v; memcpy(&v, p, 4); p += 4; return v; }
  uint32_t u32() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
  int32_t itf8() {
    need(1);
    const uint8_t v = *p;
    if (v < 0x80) { p += 1; return v; }
    if (v < 0xc0) { need(2); const int32_t r = ((v & 0x3f) << 8) | p[1]; p += 2; return r; }
    if (v < 0xe0) { need(3); const int32_t r = ((v & 0x1f) << 16) | (p[1] << 8) | p[2]; p += 3; return r; }
    if (v < 0xf0) { need(4); const int32_t r = ((v & 0x0f) << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; p += 4; return r; }
    need(5);
    const uint32_t r = ((uint32_t)(v & 0x0f) << 28) | ((uint32_t)p[1] << 20) | ((uint32_t)p[2] << 12) | ((uint32_t)p[3] << 4) | (p[4] & 0x0f);
    p += 5;
    return (int32_t)r;
  }
  int64_t ltf8() {
    need(1);
    const uint8_t v = *p;
    int n = 0;
    while (n < 8 && (v & (0x80 >> n))) ++n;
    need(1 + (size_t)n);
    uint64_t r = n == 8 ? 0 : (uint64_t)(v & (0xff >> (n + 1)));
    for (int i = 0; i < n; ++i) r = (r << 8) | p[1 + i];
    p += 1 + n;
    return (int64_t)r;
  }
  std::vector<int32_t> itf8_array() {
    const int32_t n = itf8();
    if (n < 0) throw err("negative array length");
    std::vector<int32_t> v((size_t)n);
    for (int32_t& x : v) x = itf8();
    return v;
  }
  bool done() const { return p >= end; }
  // the next n bytes as a cursor of their own
  Cursor sub(int32_t n) {
    if (n < 0) throw err("negative length");
    need((size_t)n);
    Cursor s(p, (size_t)n);
    p += n;
    return s;
  }
};

// ---- rANS 4x8 (order 0 and 1), the entropy coder of CRAM 3.0 blocks with method 4 ---------------------------------
struct RansSym { uint16_t F = 0, C = 0; };
struct RansTable {
  RansSym sym[256];
  uint8_t rev[4096];
};
inline void rans_read_table(Cursor& c, RansTable& t) {
  unsigned x = 0;
  int rle = 0;
  int j = c.u8();
  do {
    unsigned F = c.u8();
    if (F >= 128) F = ((F & 127) << 8) | c.u8();
    if (x + F > 4096) throw err("rANS frequency table overflows");
    t.sym[j].F = (uint16_t)F;
    t.sym[j].C = (uint16_t)x;
    memset(t.rev + x, j, F);
    x += F;
    c.need(1);
    if (!rle && j + 1 == *c.p) { j = c.u8(); rle = c.u8(); }
    else if (rle) { --rle; ++j; }
    else j = c.u8();
  } while (j);
}
inline void rans_renorm(uint32_t& r, Cursor& c) {
  while (r < (1u << 23)) {
    if (c.p >= c.end) { r <<= 8; continue; }   // the encoder's flush may leave the last state short of input
    r = (r << 8) | *c.p++;
  }
}
inline std::vector<uint8_t> rans_decode(const uint8_t* in, size_t n) {
  Cursor c(in, n);
  const int order = c.u8();
  const uint32_t csz = c.u32(), usz = c.u32();
  (void)csz;
  std::vector<uint8_t> out(usz);
  if (usz == 0) return out;
  if (order == 0) {
    std::unique_ptr<RansTable> t(new RansTable());
    memset(t->rev, 0, sizeof t->rev);
    rans_read_table(c, *t);
    uint32_t R[4];
    for (uint32_t& r : R) r = c.u32();
    const size_t n4 = usz & ~(size_t)3;
    for (size_t i = 0; i < n4; i += 4) {
      for (int k = 0; k < 4; ++k) {
        const uint32_t m = R[k] & 0xfff;
        const uint8_t s = t->rev[m];
        out[i + k] = s;
        R[k] = t->sym[s].F * (R[k] >> 12) + m - t->sym[s].C;
      }
      for (int k = 0; k < 4; ++k) rans_renorm(R[k], c);
    }
    for (size_t i = n4, k = 0; i < usz; ++i, ++k) out[i] = t->rev[R[k] & 0xfff];
    return out;
  }
  if (order != 1) throw err("unknown rANS order");
  std::vector<std::unique_ptr<RansTable>> t(256);
  {
    int rle = 0;
    int i = c.u8();
    do {
      t[i].reset(new RansTable());
      memset(t[i]->rev, 0, sizeof t[i]->rev);
      rans_read_table(c, *t[i]);
      c.need(1);
      if (!rle && i + 1 == *c.p) { i = c.u8(); rle = c.u8(); }
      else if (rle) { --rle; ++i; }
      else i = c.u8();
    } while (i);
  }
  uint32_t R[4];
  for (uint32_t& r : R) r = c.u32();
  const size_t q = usz >> 2;
  size_t idx[4] = {0, q, 2 * q, 3 * q};
  uint8_t last[4] = {0, 0, 0, 0};
  auto step = [&](int k) {
    const RansTable* tb = t[last[k]].get();
    if (!tb) throw err("rANS order-1 context without a table");
    const uint32_t m = R[k] & 0xfff;
    const uint8_t s = tb->rev[m];
    out[idx[k]++] = s;
    R[k] = tb->sym[s].F * (R[k] >> 12) + m - tb->sym[s].C;
    last[k] = s;
  };
  for (size_t i = 0; i < q; ++i) {
    for (int k = 0; k < 4; ++k) step(k);
    for (int k = 0; k < 4; ++k) rans_renorm(R[k], c);
  }
  while (idx[3] < usz) {   // the remainder belongs to the last quarter
    step(3);
    rans_renorm(R[3], c);
  }
  return out;
}

inline std::vector<uint8_t> gunzip(const uint8_t* in, size_t n, size_t usz) {
  std::vector<uint8_t> out(usz ? usz : 1);
  z_stream z;
  memset(&z, 0, sizeof z);
  if (inflateInit2(&z, 15 + 32) != Z_OK) throw err("zlib init failed");
  z.next_in = const_cast<Bytef*>(in);
  z.avail_in = (uInt)n;
  z.next_out = out.data();
  z.avail_out = (uInt)out.size();
  const int rc = inflate(&z, Z_FINISH);
  const size_t got = z.total_out;
  inflateEnd(&z);
  if (rc != Z_STREAM_END || got != usz) throw err("bad gzip block");
  out.resize(usz);
  return out;
}

struct Block {
  int method = 0, type = 0, id = 0;
  std::vector<uint8_t> data;
};
inline Block read_block(Cursor& c) {
  Block b;
  b.method = c.u8();
  b.type = c.u8();
  b.id = c.itf8();
  const int32_t csz = c.itf8(), usz = c.itf8();
  if (csz < 0 || usz < 0) throw err("negative block size");
  c.need((size_t)csz + 4);
  switch (b.method) {
    case 0: b.data.assign(c.p, c.p + csz); break;
    case 1: b.data = gunzip(c.p, (size_t)csz, (size_t)usz); break;
    case 4: b.data = rans_decode(c.p, (size_t)csz); break;
    case 2: throw err("bzip2-compressed block (not supported by bgx-create)");
    case 3: throw err("lzma-compressed block (not supported by bgx-create)");
    default: throw err("unknown block compression method " + std::to_string(b.method));
  }
  if (b.data.size() != (size_t)usz) throw err("block does not decompress to its stated size");
  c.p += csz + 4;   // + CRC32
  return b;
}

// ---- encodings ----------------------------------------------------------------------------------------------------
struct BitReader {   // the core block, most significant bit first
  const uint8_t* p = nullptr;
  size_t n = 0, bit = 0;
  int get(int nbits) {
    if (nbits < 0 || nbits > 31) throw err("bit field wider than 31 bits");
    uint32_t v = 0;
    for (int i = 0; i < nbits; ++i) {
      if ((bit >> 3) >= n) throw err("core block exhausted");
      v = (v << 1) | ((p[bit >> 3] >> (7 - (bit & 7))) & 1u);
      ++bit;
    }
    return (int)v;
  }
};

struct SliceData {
  BitReader core;
  std::unordered_map<int, Cursor> ext;   // external blocks by content id
  Cursor& external(int id) {
    auto it = ext.find(id);
    if (it == ext.end()) throw err("external block " + std::to_string(id) + " missing");
    return it->second;
  }
};

struct Encoding {
  enum Kind { NONE = 0, EXTERNAL = 1, GOLOMB = 2, HUFFMAN = 3, BYTE_ARRAY_LEN = 4, BYTE_ARRAY_STOP = 5, BETA = 6, SUBEXP = 7, GOLOMB_RICE = 8, GAMMA = 9 };
  int kind = NONE;
  int ext_id = 0;                          // EXTERNAL, BYTE_ARRAY_STOP
  int offset = 0, nbits = 0, k = 0;        // BETA / SUBEXP / GAMMA
  uint8_t stop = 0;                        // BYTE_ARRAY_STOP
  std::vector<int32_t> alphabet, lens;     // HUFFMAN
  std::vector<std::pair<int, int32_t>> codes;   // canonical: (code length, value) in code order; codes_first[len] etc. derived
  std::vector<uint32_t> code_vals;
  std::shared_ptr<Encoding> len_enc, val_enc;   // BYTE_ARRAY_LEN

  static Encoding parse(Cursor& c) {
    Encoding e;
    e.kind = c.itf8();
    const int32_t n = c.itf8();
    if (n < 0) throw err("bad encoding parameter length");
    c.need((size_t)n);
    Cursor p(c.p, (size_t)n);
    c.p += n;
    switch (e.kind) {
      case NONE: break;
      case EXTERNAL: e.ext_id = p.itf8(); break;
      case HUFFMAN: {
        e.alphabet = p.itf8_array();
        e.lens = p.itf8_array();
        if (e.alphabet.size() != e.lens.size() || e.alphabet.empty()) throw err("bad HUFFMAN encoding");
        for (int32_t l : e.lens)
          if (l < 0 || l > 31) throw err("bad HUFFMAN code length");
        // canonical codes: sort by (length, value), consecutive codes, shifted when the length grows
        for (size_t i = 0; i < e.alphabet.size(); ++i) e.codes.emplace_back(e.lens[i], e.alphabet[i]);
        std::sort(e.codes.begin(), e.codes.end());
        uint32_t code = 0;
        int prev_len = e.codes[0].first;
        for (size_t i = 0; i < e.codes.size(); ++i) {
          if (i) { ++code; code = (uint32_t)((uint64_t)code << (e.codes[i].first - prev_len)); prev_len = e.codes[i].first; }
          e.code_vals.push_back(code);
        }
        break;
      }
      case BYTE_ARRAY_LEN:
        e.len_enc = std::make_shared<Encoding>(parse(p));
        e.val_enc = std::make_shared<Encoding>(parse(p));
        break;
      case BYTE_ARRAY_STOP: e.stop = p.u8(); e.ext_id = p.itf8(); break;
      case BETA:
        e.offset = p.itf8(); e.nbits = p.itf8();
        if (e.nbits < 0 || e.nbits > 31) throw err("bad BETA encoding");
        break;
      case SUBEXP:
        e.offset = p.itf8(); e.k = p.itf8();
        if (e.k < 0 || e.k > 30) throw err("bad SUBEXP encoding");
        break;
      case GAMMA: e.offset = p.itf8(); break;
      default: throw err("encoding " + std::to_string(e.kind) + " is not supported by bgx-create");
    }
    return e;
  }

  int32_t get_int(SliceData& s) const {
    switch (kind) {
      case EXTERNAL: return s.external(ext_id).itf8();
      case HUFFMAN: {
        if (codes.size() == 1 && codes[0].first == 0) return codes[0].second;   // one symbol, zero bits
        uint32_t code = 0;
        int len = 0;
        size_t i = 0;
        while (i < codes.size()) {
          const int want = codes[i].first;
          code = (uint32_t)((uint64_t)code << (want - len)) | (uint32_t)s.core.get(want - len);
          len = want;
          for (; i < codes.size() && codes[i].first == len; ++i)
            if (code_vals[i] == code) return codes[i].second;
        }
        throw err("bad HUFFMAN code in the core block");
      }
      case BETA: return (int32_t)((uint32_t)s.core.get(nbits) - (uint32_t)offset);
      case GAMMA: {
        int z = 0;
        while (s.core.get(1) == 0)
          if (++z > 30) throw err("bad GAMMA code in the core block");
        return (int32_t)(((1u << z) | (uint32_t)s.core.get(z)) - (uint32_t)offset);
      }
      case SUBEXP: {
        int i = 0;
        while (s.core.get(1) == 1)
          if (++i + k > 31) throw err("bad SUBEXP code in the core block");
        uint32_t v;
        if (i == 0) v = (uint32_t)s.core.get(k);
        else { const int b = i + k - 1; v = (1u << b) | (uint32_t)s.core.get(b); }
        return (int32_t)(v - (uint32_t)offset);
      }
      default: throw err("data series without a usable integer encoding");
    }
  }
  uint8_t get_byte(SliceData& s) const {
    if (kind == EXTERNAL) return s.external(ext_id).u8();
    return (uint8_t)get_int(s);
  }
  void get_bytes(SliceData& s, std::string& out) const {
    out.clear();
    if (kind == BYTE_ARRAY_STOP) {
      Cursor& c = s.external(ext_id);
      for (;;) {
        const uint8_t b = c.u8();
        if (b == stop) break;
        out.push_back((char)b);
      }
    } else if (kind == BYTE_ARRAY_LEN) {
      const int32_t n = len_enc->get_int(s);
      if (n < 0) throw err("negative byte array length");
      if (val_enc->kind == EXTERNAL) {
        Cursor& c = s.external(val_enc->ext_id);
        c.need((size_t)n);
        out.assign(reinterpret_cast<const char*>(c.p), (size_t)n);
        c.p += n;
      } else {
        for (int32_t i = 0; i < n; ++i) out.push_back((char)val_enc->get_byte(s));
      }
    } else {
      throw err("data series without a usable byte array encoding");
    }
  }
};

inline int series_key(const char* k) { return (k[0] << 8) | k[1]; }

}  // namespace cram

// One record as the importer needs it
struct CramRecord {
  std::string qname, seq;
  uint16_t flag = 0;
};

class CramReader {
 public:
  // ref: a reference directory holding source.fasta (what `biograph create --ref` takes), or a FASTA file; may be
  // empty when every slice embeds its reference or stores its bases verbatim
  CramReader(const std::string& path, const std::string& ref) : m_name(path), m_ref_path(ref) {
    m_f = path == "/dev/stdin" ? stdin : fopen(path.c_str(), "rb");
    if (!m_f) throw std::runtime_error("Unable to open file " + path);
    uint8_t def[26];
    if (fread(def, 1, 26, m_f) != 26 || memcmp(def, "CRAM", 4) != 0) throw std::runtime_error(path + " is not a valid CRAM file.");
    if (def[4] != 3) throw cram::err("version " + std::to_string(def[4]) + "." + std::to_string(def[5]) + " is not supported by bgx-create (3.0 is)");
    if (def[5] != 0) throw cram::err("version 3." + std::to_string(def[5]) + " codecs are not supported by bgx-create (3.0 is)");
    // the first container holds the SAM header: reference names in @SQ order
    std::vector<uint8_t> body;
    Container c;
    if (!read_container(c, body)) throw cram::err("no header container");
    cram::Cursor cur(body.data(), body.size());
    const cram::Block hb = cram::read_block(cur);
    if (hb.type != 0 || hb.data.size() < 4) throw cram::err("bad file header block");
    int32_t tl;
    memcpy(&tl, hb.data.data(), 4);
    const std::string text(reinterpret_cast<const char*>(hb.data.data()) + 4, std::min<size_t>((size_t)std::max(tl, 0), hb.data.size() - 4));
    size_t pos = 0;
    while (pos < text.size()) {
      size_t nl = text.find('\n', pos);
      if (nl == std::string::npos) nl = text.size();
      const std::string line = text.substr(pos, nl - pos);
      pos = nl + 1;
      if (line.compare(0, 3, "@SQ") != 0) continue;
      const size_t sn = line.find("\tSN:");
      if (sn == std::string::npos) continue;
      const size_t e = line.find('\t', sn + 4);
      m_ref_names.push_back(line.substr(sn + 4, e == std::string::npos ? std::string::npos : e - sn - 4));
    }
  }
  ~CramReader() { if (m_f && m_f != stdin) fclose(m_f); }
  CramReader(const CramReader&) = delete;
  CramReader& operator=(const CramReader&) = delete;

  uint64_t records() const { return m_records; }
  // how often each read feature code / block method / encoding kind was met (diagnostics: BGX_CRAM_STATS)
  const std::map<std::string, uint64_t>& stats() const { return m_stats; }

  bool next(CramRecord& r) {
    while (m_out_pos >= m_out.size()) {
      if (!load_container()) return false;
    }
    r = std::move(m_out[m_out_pos++]);
    ++m_records;
    return true;
  }

 private:
  struct Container {
    int32_t length = 0, ref_id = 0, start = 0, span = 0, n_records = 0, n_blocks = 0;
    int64_t record_counter = 0, bases = 0;
    std::vector<int32_t> landmarks;
  };
  struct CompressionHeader {
    bool read_names = true, ap_delta = true, ref_required = true;
    uint8_t sub[5][4];                                    // substitution matrix: [ref base ACGTN][code] -> read base
    std::vector<std::vector<std::string>> tag_dict;       // TD: per line, the 3-byte tag ids
    std::map<int, cram::Encoding> series;                 // data series by two-letter key
    std::map<int, cram::Encoding> tags;                   // tag encodings by (tag << 8 | type)
    const cram::Encoding& get(const char* k) const {
      auto it = series.find(cram::series_key(k));
      if (it == series.end() || it->second.kind == cram::Encoding::NONE) throw cram::err(std::string("data series ") + k + " has no encoding");
      return it->second;
    }
    bool has(const char* k) const {
      auto it = series.find(cram::series_key(k));
      return it != series.end() && it->second.kind != cram::Encoding::NONE;
    }
  };

  bool read_container(Container& c, std::vector<uint8_t>& body) {
    uint8_t hdr[4];
    const size_t got = fread(hdr, 1, 4, m_f);
    if (got == 0) return false;
    if (got != 4) throw cram::err("truncated container header in " + m_name);
    memcpy(&c.length, hdr, 4);
    // the rest of the header is variable-length integers: taken byte by byte, so a pipe works as well as a file
    auto byte = [&]() -> uint8_t {
      const int ch = fgetc(m_f);
      if (ch == EOF) throw cram::err("truncated container header in " + m_name);
      return (uint8_t)ch;
    };
    auto itf8 = [&]() -> int32_t {
      uint8_t b[5];
      b[0] = byte();
      const int extra = b[0] < 0x80 ? 0 : b[0] < 0xc0 ? 1 : b[0] < 0xe0 ? 2 : b[0] < 0xf0 ? 3 : 4;
      for (int i = 1; i <= extra; ++i) b[i] = byte();
      return cram::Cursor(b, (size_t)extra + 1).itf8();
    };
    auto ltf8 = [&]() -> int64_t {
      uint8_t b[9];
      b[0] = byte();
      int extra = 0;
      while (extra < 8 && (b[0] & (0x80 >> extra))) ++extra;
      for (int i = 1; i <= extra; ++i) b[i] = byte();
      return cram::Cursor(b, (size_t)extra + 1).ltf8();
    };
    c.ref_id = itf8(); c.start = itf8(); c.span = itf8(); c.n_records = itf8();
    c.record_counter = ltf8(); c.bases = ltf8(); c.n_blocks = itf8();
    const int32_t n_landmarks = itf8();
    if (n_landmarks < 0) throw cram::err("negative landmark count");
    c.landmarks.resize((size_t)n_landmarks);
    for (int32_t& l : c.landmarks) l = itf8();
    for (int i = 0; i < 4; ++i) byte();   // CRC32 of the header
    if (c.length < 0) throw cram::err("negative container length");
    body.resize((size_t)c.length);
    if (c.length && fread(body.data(), 1, body.size(), m_f) != body.size()) throw cram::err("truncated container in " + m_name);
    return true;
  }

  static CompressionHeader parse_compression_header(const cram::Block& b) {
    CompressionHeader h;
    static const char kDefaultSub[5][4] = {{'C', 'G', 'T', 'N'}, {'A', 'G', 'T', 'N'}, {'A', 'C', 'T', 'N'}, {'A', 'C', 'G', 'N'}, {'A', 'C', 'G', 'T'}};
    memcpy(h.sub, kDefaultSub, sizeof h.sub);
    cram::Cursor c(b.data.data(), b.data.size());
    {  // preservation map
      cram::Cursor p = c.sub(c.itf8());
      const int32_t n = p.itf8();
      for (int32_t i = 0; i < n; ++i) {
        const char k0 = (char)p.u8(), k1 = (char)p.u8();
        if (k0 == 'R' && k1 == 'N') h.read_names = p.u8() != 0;
        else if (k0 == 'A' && k1 == 'P') h.ap_delta = p.u8() != 0;
        else if (k0 == 'R' && k1 == 'R') h.ref_required = p.u8() != 0;
        else if (k0 == 'S' && k1 == 'M') {
          // one byte per reference base (A, C, G, T, N): four 2-bit codes, most significant first, giving for
          // each of the other four bases (in ACGTN order without the reference base) its substitution code
          static const char kOthers[5][4] = {{'C', 'G', 'T', 'N'}, {'A', 'G', 'T', 'N'}, {'A', 'C', 'T', 'N'}, {'A', 'C', 'G', 'N'}, {'A', 'C', 'G', 'T'}};
          for (int r = 0; r < 5; ++r) {
            const uint8_t v = p.u8();
            for (int j = 0; j < 4; ++j) h.sub[r][(v >> (6 - 2 * j)) & 3] = (uint8_t)kOthers[r][j];
          }
        } else if (k0 == 'T' && k1 == 'D') {
          const int32_t len = p.itf8();
          p.need((size_t)len);
          std::vector<std::string> line;
          std::string cur;
          for (int32_t j = 0; j < len; ++j) {
            const uint8_t ch = p.p[j];
            if (ch == 0) {
              line.clear();
              for (size_t q = 0; q + 3 <= cur.size(); q += 3) line.push_back(cur.substr(q, 3));
              h.tag_dict.push_back(line);
              cur.clear();
            } else {
              cur.push_back((char)ch);
            }
          }
          p.p += len;
        } else {
          throw cram::err(std::string("unknown preservation map key ") + k0 + k1);
        }
      }
    }
    {  // data series encodings
      cram::Cursor p = c.sub(c.itf8());
      const int32_t n = p.itf8();
      for (int32_t i = 0; i < n; ++i) {
        const int key = (p.u8() << 8);
        const int key2 = key | p.u8();
        h.series[key2] = cram::Encoding::parse(p);
      }
    }
    {  // tag encodings
      cram::Cursor p = c.sub(c.itf8());
      const int32_t n = p.itf8();
      for (int32_t i = 0; i < n; ++i) {
        const int32_t key = p.itf8();
        h.tags[key] = cram::Encoding::parse(p);
      }
    }
    return h;
  }

  // ---- reference sequences ------------------------------------------------------------------------------------------
  const std::string& reference(int32_t ref_id) {
    auto it = m_refs.find(ref_id);
    if (it != m_refs.end()) return it->second;
    if (ref_id < 0 || (size_t)ref_id >= m_ref_names.size()) throw cram::err("reference id " + std::to_string(ref_id) + " is not in the header");
    if (!m_fasta_loaded) load_fasta();
    auto f = m_fasta.find(m_ref_names[(size_t)ref_id]);
    if (f == m_fasta.end()) throw cram::err("reference sequence " + m_ref_names[(size_t)ref_id] + " is not in " + m_fasta_path);
    return m_refs.emplace(ref_id, std::move(f->second)).first->second;
  }
  void load_fasta() {
    m_fasta_loaded = true;
    std::vector<std::string> cand = {m_ref_path + "/source.fasta", m_ref_path + "/reference.fasta", m_ref_path};
    for (const std::string& p : cand) {
      std::ifstream in(p);
      if (!in || m_ref_path.empty()) continue;
      std::string line, name;
      std::string* seq = nullptr;
      bool any = false;
      while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
          const size_t e = line.find_first_of(" \t");
          name = line.substr(1, e == std::string::npos ? std::string::npos : e - 1);
          seq = &m_fasta[name];
          any = true;
        } else if (seq) {
          for (char ch : line) seq->push_back((char)toupper((unsigned char)ch));
        }
      }
      if (any) { m_fasta_path = p; return; }
    }
    throw cram::err("a reference FASTA is needed to decode " + m_name + " (looked for " + m_ref_path + "/source.fasta; pass --ref)");
  }

  // ---- one container: compression header + slices -------------------------------------------------------------------------
  bool load_container() {
    m_out.clear();
    m_out_pos = 0;
    Container c;
    std::vector<uint8_t> body;
    if (!read_container(c, body)) return false;
    if (c.n_records == 0 && c.ref_id == -1 && c.start == 4542278) return true;   // the EOF container
    if (c.n_blocks == 0 || body.empty()) return true;
    cram::Cursor cur(body.data(), body.size());
    const cram::Block chb = cram::read_block(cur);
    if (chb.type != 1) throw cram::err("container does not start with a compression header");
    const CompressionHeader h = parse_compression_header(chb);
    for (const auto& kv : h.series) ++m_stats["series encoding " + std::to_string(kv.second.kind)];
    for (const auto& kv : h.tags) ++m_stats["tag encoding " + std::to_string(kv.second.kind)];
    while (!cur.done()) decode_slice(h, cur);
    return true;
  }

  void decode_slice(const CompressionHeader& h, cram::Cursor& cur) {
    const cram::Block shb = cram::read_block(cur);
    if (shb.type != 2) throw cram::err("expected a slice header block");
    cram::Cursor s(shb.data.data(), shb.data.size());
    const int32_t ref_id = s.itf8(), start = s.itf8(), span = s.itf8(), n_records = s.itf8();
    const int64_t record_counter = s.ltf8();
    const int32_t n_blocks = s.itf8();
    const std::vector<int32_t> content_ids = s.itf8_array();
    const int32_t embedded_ref = s.itf8();
    uint8_t md5sum[16];
    s.need(16);
    memcpy(md5sum, s.p, 16);
    std::vector<cram::Block> blocks;
    for (int32_t i = 0; i < n_blocks; ++i) {
      blocks.push_back(cram::read_block(cur));
      ++m_stats["block method " + std::to_string(blocks.back().method)];
    }
    cram::SliceData sd;
    const std::string* embedded = nullptr;
    std::string embedded_store;
    for (const cram::Block& b : blocks) {
      if (b.type == 5) { sd.core.p = b.data.data(); sd.core.n = b.data.size(); }
      else if (b.type == 4) {
        sd.ext[b.id] = cram::Cursor(b.data.data(), b.data.size());
        if (embedded_ref >= 0 && b.id == embedded_ref) { embedded_store.assign(b.data.begin(), b.data.end()); embedded = &embedded_store; }
      }
    }
    // reference for a single-reference slice: checked against the slice's MD5 once
    const std::string* ref = nullptr;
    int64_t ref_base = 1;   // 1-based position of ref[0]
    if (ref_id >= 0) {
      if (embedded) { ref = embedded; ref_base = start; }
      else if (h.ref_required) {
        ref = &reference(ref_id);
        bool zero = true;
        for (uint8_t b : md5sum) zero = zero && b == 0;
        if (!zero && start >= 1 && (size_t)(start - 1 + span) <= ref->size()) {
          uint8_t got[16];
          cram::md5(reinterpret_cast<const uint8_t*>(ref->data()) + (start - 1), (size_t)span, got);
          if (memcmp(got, md5sum, 16) != 0)
            throw cram::err("the reference does not match " + m_name + " (MD5 of " + m_ref_names[(size_t)ref_id] + ":" + std::to_string(start) + "+" +
                            std::to_string(span) + " differs): wrong --ref?");
        }
      }
    }

    const size_t first = m_out.size();
    m_out.resize(first + (size_t)n_records);
    std::vector<int32_t> next_frag((size_t)n_records, -1);
    std::vector<uint8_t> detached((size_t)n_records, 0);
    int64_t last_ap = start;   // 64-bit: corrupt deltas must not overflow
    std::string scratch;
    for (int32_t i = 0; i < n_records; ++i) {
      CramRecord& r = m_out[first + (size_t)i];
      const int32_t bf = h.get("BF").get_int(sd);
      const int32_t cf = h.get("CF").get_int(sd);
      int32_t ri = ref_id;
      if (ref_id == -2) ri = h.get("RI").get_int(sd);
      const int32_t rl = h.get("RL").get_int(sd);
      int64_t ap = h.get("AP").get_int(sd);
      if (h.ap_delta) { ap += last_ap; last_ap = ap; }
      h.get("RG").get_int(sd);
      if (h.read_names) h.get("RN").get_bytes(sd, r.qname);
      if (cf & 0x2) {                              // detached: mate information stored explicitly
        detached[(size_t)i] = 1;
        h.get("MF").get_int(sd);
        if (!h.read_names) h.get("RN").get_bytes(sd, r.qname);
        h.get("NS").get_int(sd);
        h.get("NP").get_int(sd);
        h.get("TS").get_int(sd);
      } else if (cf & 0x4) {                       // the mate is a later record of this slice
        next_frag[(size_t)i] = h.get("NF").get_int(sd);
      }
      {                                            // tags: decoded to keep the streams in step, then dropped
        const int32_t tl = h.get("TL").get_int(sd);
        if (tl < 0 || (size_t)tl >= h.tag_dict.size()) throw cram::err("tag line out of range");
        for (const std::string& t : h.tag_dict[(size_t)tl]) {
          const int32_t key = ((uint8_t)t[0] << 16) | ((uint8_t)t[1] << 8) | (uint8_t)t[2];
          auto it = h.tags.find(key);
          if (it == h.tags.end()) throw cram::err("tag without an encoding");
          it->second.get_bytes(sd, scratch);
        }
      }
      r.flag = (uint16_t)bf;
      r.seq.assign((size_t)std::max(rl, 0), 'N');
      if (!(bf & 0x4)) {                           // mapped: reference bases + read features
        const std::string* rr = ref;
        int64_t rbase = ref_base;
        if (ref_id == -2 && ri >= 0 && h.ref_required) { rr = &reference(ri); rbase = 1; }
        const int32_t fn = h.get("FN").get_int(sd);
        int32_t rpos = 0;                          // bases of the read done so far
        int64_t gpos = ap;                         // 1-based reference position of the next read base
        auto ref_base_at = [&](int64_t g) -> char {
          if (!rr) return 'N';
          const int64_t o = g - rbase;
          return o >= 0 && (size_t)o < rr->size() ? (*rr)[(size_t)o] : 'N';
        };
        auto copy_ref_until = [&](int32_t upto) {  // read positions [rpos, upto) match the reference
          for (; rpos < upto && rpos < rl; ++rpos, ++gpos) r.seq[(size_t)rpos] = ref_base_at(gpos);
        };
        int64_t fpos = 0;
        for (int32_t f = 0; f < fn; ++f) {
          const uint8_t code = h.get("FC").get_byte(sd);
          fpos += h.get("FP").get_int(sd);         // 1-based position in the read, delta coded
          copy_ref_until((int32_t)std::min<int64_t>(std::max<int64_t>(fpos - 1, 0), rl));
          ++m_stats[std::string("feature ") + (char)code];
          switch (code) {
            case 'B': {                            // a base and its quality
              const uint8_t b = h.get("BA").get_byte(sd);
              h.get("QS").get_byte(sd);
              if (rpos < rl) r.seq[(size_t)rpos] = (char)b;
              ++rpos; ++gpos;
              break;
            }
            case 'X': {                            // substitution, through the matrix
              const int32_t bs = h.get("BS").get_int(sd);
              const char rb = ref_base_at(gpos);
              const int ridx = rb == 'A' ? 0 : rb == 'C' ? 1 : rb == 'G' ? 2 : rb == 'T' ? 3 : 4;
              if (rpos < rl) r.seq[(size_t)rpos] = (char)h.sub[ridx][bs & 3];
              ++rpos; ++gpos;
              break;
            }
            case 'I': {                            // insertion
              h.get("IN").get_bytes(sd, scratch);
              for (char ch : scratch) { if (rpos < rl) r.seq[(size_t)rpos] = ch; ++rpos; }
              break;
            }
            case 'i': {                            // single-base insertion
              const uint8_t b = h.get("BA").get_byte(sd);
              if (rpos < rl) r.seq[(size_t)rpos] = (char)b;
              ++rpos;
              break;
            }
            case 'S': {                            // soft clip
              h.get("SC").get_bytes(sd, scratch);
              for (char ch : scratch) { if (rpos < rl) r.seq[(size_t)rpos] = ch; ++rpos; }
              break;
            }
            case 'b': {                            // a stretch of bases
              h.get("BB").get_bytes(sd, scratch);
              for (char ch : scratch) { if (rpos < rl) r.seq[(size_t)rpos] = ch; ++rpos; ++gpos; }
              break;
            }
            case 'D': gpos += h.get("DL").get_int(sd); break;     // deletion
            case 'N': gpos += h.get("RS").get_int(sd); break;     // reference skip
            case 'P': h.get("PD").get_int(sd); break;             // padding
            case 'H': h.get("HC").get_int(sd); break;             // hard clip
            case 'Q': h.get("QS").get_byte(sd); break;            // a quality score alone
            case 'q': h.get("QQ").get_bytes(sd, scratch); break;  // a stretch of quality scores
            default: throw cram::err(std::string("unknown read feature '") + (char)code + "'");
          }
        }
        copy_ref_until(rl);
        h.get("MQ").get_int(sd);
        if (cf & 0x1) for (int32_t q = 0; q < rl; ++q) h.get("QS").get_byte(sd);
      } else {                                     // unmapped: the bases are stored verbatim
        for (int32_t q = 0; q < rl; ++q) r.seq[(size_t)q] = (char)h.get("BA").get_byte(sd);
        if (cf & 0x1) for (int32_t q = 0; q < rl; ++q) h.get("QS").get_byte(sd);
      }
      if (cf & 0x8) r.seq.clear();                 // "no sequence" (SEQ is '*')
    }
    // names of records stored without one: mates inside the slice share a generated name
    for (int32_t i = 0; i < n_records; ++i) {
      CramRecord& r = m_out[first + (size_t)i];
      if (r.qname.empty()) r.qname = "cram:" + std::to_string(record_counter + i);
      const int32_t nf = next_frag[(size_t)i];
      if (nf >= 0) {
        const int64_t j = (int64_t)i + nf + 1;
        if (j >= n_records) throw cram::err("mate link leaves the slice");
        CramRecord& m = m_out[first + (size_t)j];
        if (m.qname.empty()) m.qname = r.qname;
        // the mate fields of BAM flags are implied for attached mates: both ends are paired reads
        r.flag |= 0x1;
        m.flag |= 0x1;
      }
    }
    // the sequence as the read was sequenced: reverse-strand records are stored reverse-complemented
    for (int32_t i = 0; i < n_records; ++i) {
      CramRecord& r = m_out[first + (size_t)i];
      if (r.flag & 0x10) {
        static const char comp[] = "TVGH..CD..M.KN...YSA.BW.R.";           // modules/bio_base/dna_base_set.cpp:78-79
        std::reverse(r.seq.begin(), r.seq.end());
        for (char& ch : r.seq) {
          const int idx = ch - 'A';
          if (idx >= 0 && idx < 26) ch = comp[idx];
        }
      }
    }
  }

  std::string m_name, m_ref_path, m_fasta_path;
  FILE* m_f = nullptr;
  std::vector<std::string> m_ref_names;
  std::unordered_map<std::string, std::string> m_fasta;
  bool m_fasta_loaded = false;
  std::unordered_map<int32_t, std::string> m_refs;
  std::vector<CramRecord> m_out;
  size_t m_out_pos = 0;
  uint64_t m_records = 0;
  std::map<std::string, uint64_t> m_stats;
};

}  // namespace bgx_cli

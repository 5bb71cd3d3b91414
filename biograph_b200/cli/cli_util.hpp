// cli_util.hpp -- helpers shared by bgx-create and bgx-merge: messages, SHA-1 (readmaps are named by
// their SHA-1, modules/biograph/biograph_create.cpp:820-828, biograph_merge.cpp:306-308), uuids, JSON
// strings, stage timings (runtime_stats start_stage / end_stage).
#pragma once

#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace bgx_cli {

[[noreturn]] inline void die(const std::string& msg) {
  std::cerr << msg << "\n";
  exit(1);
}

inline std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return buf;
}

inline bool ends_with(const std::string& s, const std::string& e) { return s.size() >= e.size() && s.compare(s.size() - e.size(), e.size(), e) == 0; }


inline std::string sha1_file(const std::string& path) {
  uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  auto block = [&](const uint8_t* p) {
    uint32_t w[80];
    for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 80; ++i) { uint32_t t = w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16]; w[i] = t << 1 | t >> 31; }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
    for (int i = 0; i < 80; ++i) {
      uint32_t f, k;
      if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
      else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
      else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
      else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
      uint32_t t = (a << 5 | a >> 27) + f + e + k + w[i];
      e = d; d = c; c = b << 30 | b >> 2; b = a; a = t;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
  };
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot read " + path);
  std::vector<uint8_t> buf(1 << 20);
  uint8_t tail[128];
  uint64_t total = 0;
  size_t have = 0;
  for (;;) {
    in.read(reinterpret_cast<char*>(buf.data()), buf.size());
    const size_t n = (size_t)in.gcount();
    if (!n) break;
    total += n;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) block(buf.data() + i);
    have = n - i;
    memcpy(tail, buf.data() + i, have);
    if (n < buf.size()) break;   // the buffer is a multiple of 64, so only the last chunk leaves a partial block
  }
  tail[have++] = 0x80;
  while (have % 64 != 56) tail[have++] = 0;
  const uint64_t bits = total * 8;
  for (int i = 7; i >= 0; --i) tail[have++] = (uint8_t)(bits >> (8 * i));
  for (size_t i = 0; i < have; i += 64) block(tail + i);
  char out[41];
  snprintf(out, sizeof(out), "%08x%08x%08x%08x%08x", h[0], h[1], h[2], h[3], h[4]);
  return out;
}

inline std::string make_uuid() {
  std::random_device rd;
  std::mt19937_64 g(((uint64_t)rd() << 32) ^ rd() ^ (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count());
  uint64_t a = g(), b = g();
  a = (a & ~0xF000ull) | 0x4000ull;                 // version 4
  b = (b & ~(3ull << 62)) | (2ull << 62);           // variant 1
  return fmt("%08x-%04x-%04x-%04x-%012llx", (unsigned)(a >> 32), (unsigned)((a >> 16) & 0xffff), (unsigned)(a & 0xffff),
             (unsigned)(b >> 48), (unsigned long long)(b & 0xffffffffffffull));
}

inline std::string json_str(const std::string& s) {
  std::string o = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\') { o += '\\'; o += c; }
    else if (c == '\n') o += "\\n";
    else o += c;
  }
  return o + "\"";
}

struct Stages {  // m_stats.start_stage / end_stage: seconds per stage, in order
  std::vector<std::pair<std::string, double>> t;
  std::chrono::steady_clock::time_point t0;
  void start() { t0 = std::chrono::steady_clock::now(); }
  void end(const std::string& name) { t.emplace_back(name, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); }
};

}  // namespace bgx_cli

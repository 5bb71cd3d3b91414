// correct.cu -- k-mer based read correction and suffix-seed counts on the GPU.
//
// Replaces (reference, CPU): fast_read_correct / correct_internal
// (modules/bio_base/fast_read_correct.cpp:94-182, :16-90) and correct_reads::correct
// (bs/correct_reads.cpp:154-231, kmer_starts_read :308-311).
//
// One thread per read.  The reference's recursive DFS is run iteratively with an explicit
// frame stack; a partial result is (length, substitutions[<=16]) instead of a copied sequence,
// because a corrected read is always the input with a few substituted bases, truncated.
// k-mer membership = one probe sequence in the 8-byte-slot solid hash set (L2 resident for
// bacterial genomes, one DRAM sector per probe otherwise).
#include <algorithm>

#include "ctx.h"

namespace bgx {
namespace {

constexpr int kMaxCorr = 16;
constexpr int kMaxWords = 9;  // 255 bases -> 8 words + 1 pad

struct Res {
  int len;
  int ncorr;
  uint8_t pos[kMaxCorr];  // logical positions of substitutions
  uint8_t base[kMaxCorr];
};

struct Frame {
  uint64_t kmer;   // k-mer before `start` on ENTER; k-mer at the error point afterwards
  int start;       // logical start of this frame's input
  int run;         // bases extended without correction
  int e;           // logical index of the skipped (bad) base
  int b;           // substitution being tried
  int budget;
  int best_size;
  int best_b;
  Res best;
};

struct Params {
  int k;
  int max_corr;
  int min_run;
  double trim;
  const unsigned long long* set;
  uint64_t set_mask;
};

// logical view of a stretch of the read: forward from `origin`, or reverse-complemented
// walking down from `origin`
struct Input {
  const uint64_t* w;
  const uint32_t* m;  // nullptr when the read set has no N
  int origin;
  int n;
  bool rev;
  __device__ __forceinline__ int get(int i) const {  // 0..3, 4 = 'N'
    int pos = rev ? origin - i : origin + i;
    if (m != nullptr && ((m[pos >> 5] >> (31 - (pos & 31))) & 1u)) return 4;
    int c = (int)((w[pos >> 5] >> (62 - 2 * (pos & 31))) & 3u);
    return rev ? 3 - c : c;
  }
};

// kmer_lookup_f (bs/correct_reads.cpp:163-171): canonicalise, probe the solid set.
// Returns the stored key|flags word, or kEmptyKey when absent; *flipped as canonicalize.
__device__ __forceinline__ unsigned long long solid_find(const Params& P, uint64_t kmer, bool* flipped) {
  bool fl;
  uint64_t canon = canonicalize(kmer, P.k, fl);
  *flipped = fl;
  uint64_t slot = mix64(canon) & P.set_mask;
  for (;;) {
    unsigned long long cur = __ldg(&P.set[slot]);
    if (cur == kEmptyKey) return kEmptyKey;
    if ((cur & kKmerMask) == canon) return cur;
    slot = (slot + 1) & P.set_mask;
  }
}
__device__ __forceinline__ bool solid_has(const Params& P, uint64_t kmer) {
  bool fl;
  return solid_find(P, kmer, &fl) != kEmptyKey;
}

__device__ __forceinline__ uint64_t shift_in(uint64_t kmer, int b, uint64_t mask) {
  return ((kmer << 2) | (uint64_t)b) & mask;
}

// correct_internal (fast_read_correct.cpp:16-90), iteratively.  Returns the result for the
// whole input `in` starting from k-mer `kmer0`.
__device__ void correct_internal(const Params& P, const Input& in, uint64_t kmer0, int min_run0, int budget0,
                                 bool require_run_at_end, Frame* st, Res* out) {
  const uint64_t kmask = kmer_low_mask(P.k);
  int depth = 0;
  st[0].start = 0;
  st[0].kmer = kmer0;
  st[0].budget = budget0;
  Res ret;
  ret.len = 0;
  ret.ncorr = 0;
  enum { ENTER, TRY, RETURN } mode = ENTER;
  for (;;) {
    if (mode == ENTER) {
      Frame& f = st[depth];
      int it = f.start, run = 0;
      uint64_t kmer = f.kmer;
      bool finished = false;
      int c = in.get(it);
      if (c != 4) {
        uint64_t nk = shift_in(kmer, c, kmask);
        while (solid_has(P, nk)) {
          ++run;
          ++it;
          if (it == in.n) { finished = true; break; }
          kmer = nk;
          c = in.get(it);
          if (c == 4) break;
          nk = shift_in(kmer, c, kmask);
        }
      }
      int min_run_here = depth == 0 ? min_run0 : P.min_run;
      if (finished || run < min_run_here || f.budget == 0) {
        ret.len = run;
        ret.ncorr = 0;
        if (depth == 0) break;
        --depth;
        mode = RETURN;
        continue;
      }
      f.run = run;
      f.kmer = kmer;
      f.e = it;
      f.b = 0;
      f.best_size = 0;
      f.best_b = 0;
      mode = TRY;
    }
    if (mode == RETURN) {
      Frame& f = st[depth];
      if (!(require_run_at_end && ret.len < P.min_run) && ret.len >= f.best_size) {
        f.best_size = ret.len + 1;  // a later base wins only if strictly longer
        f.best_b = f.b;
        f.best = ret;
      }
      ++f.b;
      mode = TRY;
    }
    // mode == TRY
    {
      Frame& f = st[depth];
      bool pushed = false;
      while (f.b < 4) {
        uint64_t tk = shift_in(f.kmer, f.b, kmask);
        if (!solid_has(P, tk)) { ++f.b; continue; }
        if (f.e + 1 != in.n) {
          Frame& g = st[depth + 1];
          g.start = f.e + 1;
          g.kmer = tk;
          g.budget = f.budget - 1;
          ++depth;
          mode = ENTER;
          pushed = true;
          break;
        }
        // the substituted base is the last one: empty continuation
        if (!require_run_at_end || 0 >= P.min_run) {
          if (0 >= f.best_size) {
            f.best_size = 1;
            f.best_b = f.b;
            f.best.len = 0;
            f.best.ncorr = 0;
          }
        }
        ++f.b;
      }
      if (pushed) continue;
      if (f.best_size) {
        ret = f.best;
        ret.pos[ret.ncorr] = (uint8_t)f.e;
        ret.base[ret.ncorr] = (uint8_t)f.best_b;
        ret.ncorr += 1;
        ret.len = f.run + 1 + f.best.len;
      } else {
        ret.len = f.run;
        ret.ncorr = 0;
      }
      if (depth == 0) break;
      --depth;
      mode = RETURN;
    }
  }
  *out = ret;
}

__device__ __forceinline__ void set_base(uint64_t* w, int pos, int b) {
  int sh = 62 - 2 * (pos & 31);
  w[pos >> 5] = (w[pos >> 5] & ~(3ULL << sh)) | ((uint64_t)b << sh);
}

__device__ __forceinline__ uint64_t local_window(const uint64_t* w, int a) {
  int q = a >> 5;
  unsigned s = (unsigned)(a & 31) * 2;
  return s ? ((w[q] << s) | (w[q + 1] >> (64 - s))) : w[q];
}

__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

__global__ void __launch_bounds__(128) correct_kernel(const uint64_t* __restrict__ words,
                                                      const uint32_t* __restrict__ nmask,
                                                      const uint32_t* __restrict__ word_off,
                                                      const uint16_t* __restrict__ lens, uint32_t n_reads, Params P,
                                                      uint64_t* __restrict__ store, uint64_t rc_word_base,
                                                      uint16_t* __restrict__ clen, uint8_t* __restrict__ ncorr,
                                                      uint16_t* __restrict__ next_fwd, uint16_t* __restrict__ next_rev,
                                                      uint32_t* __restrict__ seed_cnt,
                                                      unsigned long long* __restrict__ totals) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t w[kMaxWords];
  uint32_t m[kMaxWords];
  Frame st[kMaxCorr + 1];
  int out_len = 0, corrections = 0, nf = 0, nr = 0;
  const int k = P.k;
  const uint64_t kmask = kmer_low_mask(k);
  int L = 0, nw = 0;
  uint32_t base = 0;
  if (r < n_reads) {
    L = lens[r];
    nw = (L + 31) >> 5;
    base = word_off[r];
#pragma unroll
    for (int i = 0; i < kMaxWords; ++i) {
      w[i] = i < nw ? words[base + i] : 0;
      m[i] = (nmask != nullptr && i < nw) ? nmask[base + i] : 0;
    }
  }
  bool ok = r < n_reads && L >= k;
  if (ok) {
    Input whole{w, nmask ? m : nullptr, 0, L, false};
    // scan right to the first solid k-mer (fast_read_correct.cpp:108-123)
    int it = 0, left = k;
    uint64_t kmer = 0;
    for (;;) {
      if (!left && solid_has(P, kmer)) break;
      if (it == L) { ok = false; break; }
      int c = whole.get(it);
      ++it;
      if (c == 4) { left = k; continue; }
      kmer = shift_in(kmer, c, kmask);
      if (left) --left;
    }
    int budget = P.max_corr;
    if (ok && it != k) {
      // left side: correct the reverse complement of read[0, kmer_start) (:135-167)
      int kmer_start = it - k;
      Input lin{w, nmask ? m : nullptr, kmer_start - 1, kmer_start, true};
      Res lres;
      correct_internal(P, lin, revcomp_kmer(kmer, k), 0, budget, false, st, &lres);
      if (lres.len != kmer_start) {
        ok = false;  // left correction failed (:150-153)
      } else {
        for (int i = 0; i < lres.ncorr; ++i) {
          int pos = kmer_start - 1 - lres.pos[i];
          set_base(w, pos, 3 - lres.base[i]);
          if (nmask) m[pos >> 5] &= ~(1u << (31 - (pos & 31)));
        }
        budget -= lres.ncorr;
        corrections += lres.ncorr;
      }
    }
    if (ok) {
      out_len = it;
      if (it != L) {
        Input rin{w, nmask ? m : nullptr, it, L - it, false};
        Res rres;
        correct_internal(P, rin, kmer, 0, budget, true, st, &rres);
        for (int i = 0; i < rres.ncorr; ++i) set_base(w, it + rres.pos[i], rres.base[i]);
        corrections += rres.ncorr;
        out_len = it + rres.len;
      }
      // drop rule: corrected.size() < unsigned(trim_after_portion * len) (bs/correct_reads.cpp:174-178)
      unsigned needed = (unsigned)(P.trim * (double)L);
      if ((unsigned)out_len < needed) ok = false;
    }
  }
  if (r < n_reads) {
    if (!ok) { out_len = 0; corrections = 0; }
    // truncate and store forward + reverse-complement copies
    int nwc = (out_len + 31) >> 5;
    for (int i = 0; i < nw; ++i) {
      uint64_t v = 0;
      if (i < nwc) {
        v = w[i];
        int rem = out_len - 32 * i;
        if (rem < 32) v &= top_bases_mask(rem);
      }
      w[i] = v;
      store[base + i] = v;
    }
    for (int i = nw; i < kMaxWords; ++i) w[i] = 0;
    for (int q = 0; q < nw; ++q) {
      uint64_t v = 0;
      if (q < nwc) {
        int mcount = min(32, out_len - 32 * q);
        int lo_pos = out_len - 32 * q - mcount;
        uint64_t win = local_window(w, lo_pos);
        uint64_t x = win >> (64 - 2 * mcount);
        v = revcomp_kmer(x, mcount) << (64 - 2 * mcount);
      }
      store[rc_word_base + base + q] = v;
    }
    if (ok) {
      // next_fwd_read / next_rev_read (bs/correct_reads.cpp:195-210): walk k-mers until one
      // (other than the first) whose flag says a read starts there.
      int nkc = out_len - k + 1;
      for (int p = 0; p < nkc; ++p) {
        if (nf > 0) {
          uint64_t km = local_window(w, p) >> (64 - 2 * k);
          bool fl;
          unsigned long long e = solid_find(P, km, &fl);
          if (e & (fl ? kRevFlag : kFwdFlag)) break;
        }
        ++nf;
      }
      for (int p = nkc - 1; p >= 0; --p) {
        if (nr > 0) {
          uint64_t km = local_window(w, p) >> (64 - 2 * k);
          bool fl;
          unsigned long long e = solid_find(P, km, &fl);
          if (e & (fl ? kFwdFlag : kRevFlag)) break;  // as_flipped()
        }
        ++nr;
      }
    }
    clen[r] = (uint16_t)out_len;
    ncorr[r] = (uint8_t)corrections;
    next_fwd[r] = (uint16_t)nf;
    next_rev[r] = (uint16_t)nr;
    seed_cnt[r] = (uint32_t)(nf + nr);
  }
  unsigned kept = warp_sum(ok ? 1u : 0u);
  unsigned kb = warp_sum((unsigned)out_len);
  unsigned sd = warp_sum((unsigned)(nf + nr));
  unsigned cr = warp_sum((unsigned)corrections);
  unsigned tr = warp_sum((ok && out_len < L) ? 1u : 0u);
  if (lane_id() == 0) {
    if (kept) atomicAdd(&totals[0], (unsigned long long)kept);
    if (kb) atomicAdd(&totals[1], (unsigned long long)kb);
    if (sd) atomicAdd(&totals[2], (unsigned long long)sd);
    if (cr) atomicAdd(&totals[3], (unsigned long long)cr);
    if (tr) atomicAdd(&totals[4], (unsigned long long)tr);
  }
}

// seqset_for_reads seeding: the read as it is + its reverse complement, one seed each
__global__ void __launch_bounds__(128) passthrough_kernel(const uint64_t* __restrict__ words,
                                                          const uint32_t* __restrict__ word_off,
                                                          const uint16_t* __restrict__ lens, uint32_t n_reads,
                                                          uint64_t* __restrict__ store, uint64_t rc_word_base,
                                                          uint16_t* __restrict__ clen, uint8_t* __restrict__ ncorr,
                                                          uint16_t* __restrict__ next_fwd, uint16_t* __restrict__ next_rev,
                                                          unsigned long long* __restrict__ totals) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  int L = 0;
  if (r < n_reads) {
    L = lens[r];
    const int nw = (L + 31) >> 5;
    const uint32_t base = word_off[r];
    uint64_t w[kMaxWords];
#pragma unroll
    for (int i = 0; i < kMaxWords; ++i) w[i] = i < nw ? words[base + i] : 0;
    for (int i = 0; i < nw; ++i) store[base + i] = w[i];
    for (int q = 0; q < nw; ++q) {
      int mcount = min(32, L - 32 * q);
      int lo_pos = L - 32 * q - mcount;
      uint64_t x = local_window(w, lo_pos) >> (64 - 2 * mcount);
      store[rc_word_base + base + q] = revcomp_kmer(x, mcount) << (64 - 2 * mcount);
    }
    clen[r] = (uint16_t)L;
    ncorr[r] = 0;
    next_fwd[r] = L ? 1 : 0;
    next_rev[r] = L ? 1 : 0;
  }
  unsigned kept = warp_sum(L ? 1u : 0u), kb = warp_sum((unsigned)L);
  if (lane_id() == 0 && kept) {
    atomicAdd(&totals[0], (unsigned long long)kept);
    atomicAdd(&totals[1], (unsigned long long)kb);
    atomicAdd(&totals[2], 2ULL * kept);
  }
}

}  // namespace

void stage_seed_uncorrected(Context* c) {
  BGX_CHECK(!c->has_n, "bgx_seed_uncorrected: reads must not contain N");
  cudaStream_t s = c->stream;
  ScopedStage st_all(c, "correct_total");
  const uint64_t n = c->n_reads;
  c->store.alloc(2 * c->n_words + 1, s);
  c->clen.alloc(std::max<uint64_t>(n, 1), s);
  c->ncorr.alloc(std::max<uint64_t>(n, 1), s);
  c->next_fwd.alloc(std::max<uint64_t>(n, 1), s);
  c->next_rev.alloc(std::max<uint64_t>(n, 1), s);
  DevBuf<unsigned long long> totals(3, s);
  BGX_CUDA(cudaMemsetAsync(totals.p, 0, 3 * sizeof(unsigned long long), s));
  BGX_CUDA(cudaMemsetAsync(c->store.p + 2 * c->n_words, 0, sizeof(uint64_t), s));
  if (n)
    KLAUNCH(passthrough_kernel)<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(c->words.p, c->word_off.p, c->lens.p, (uint32_t)n,
                                                                  c->store.p, c->n_words, c->clen.p, c->ncorr.p,
                                                                  c->next_fwd.p, c->next_rev.p, totals.p);
  BGX_CUDA(cudaGetLastError());
  unsigned long long h[3];
  BGX_CUDA(cudaMemcpyAsync(h, totals.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  c->n_kept = h[0];
  c->kept_bases = h[1];
  c->n_seeds = h[2];
  c->corrected = true;
  c->built = false;
  st_all.stop();
  c->set_stat("reads_kept", (double)h[0]);
  c->set_stat("corrected_bases", (double)h[1]);
  c->set_stat("seeds", (double)h[2]);
}

void stage_correct(Context* c) {
  BGX_CHECK(c->counted, "bgx_correct: call bgx_count_kmers first");
  BGX_CHECK(c->opt.max_corrections <= kMaxCorr, "max_corrections > 16 is not supported");
  cudaStream_t s = c->stream;
  ScopedStage st_all(c, "correct_total");
  const uint64_t n = c->n_reads;
  c->store.alloc(2 * c->n_words + 1, s);
  c->clen.alloc(n, s);
  c->ncorr.alloc(n, s);
  c->next_fwd.alloc(n, s);
  c->next_rev.alloc(n, s);
  DevBuf<uint32_t> seed_cnt(n, s);
  DevBuf<unsigned long long> totals(5, s);
  BGX_CUDA(cudaMemsetAsync(totals.p, 0, 5 * sizeof(unsigned long long), s));
  BGX_CUDA(cudaMemsetAsync(c->store.p + 2 * c->n_words, 0, sizeof(uint64_t), s));
  Params P;
  P.k = c->opt.kmer_size;
  P.max_corr = c->opt.max_corrections;
  P.min_run = c->opt.min_good_run;
  P.trim = (double)c->opt.trim_after_portion;  // float widened to double (biograph_create.cpp:489-490,731)
  P.set = c->solid.p;
  P.set_mask = c->solid_slots - 1;
  {
    ScopedStage st(c, "correct_kernel");
    KLAUNCH(correct_kernel)<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(c->words.p, c->has_n ? c->nmask.p : nullptr,
                                                              c->word_off.p, c->lens.p, (uint32_t)n, P, c->store.p,
                                                              c->n_words, c->clen.p, c->ncorr.p, c->next_fwd.p,
                                                              c->next_rev.p, seed_cnt.p, totals.p);
    BGX_CUDA(cudaGetLastError());
    st.stop();
  }
  unsigned long long h[5];
  BGX_CUDA(cudaMemcpyAsync(h, totals.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
  c->n_kept = h[0];
  c->kept_bases = h[1];
  c->n_seeds = h[2];
  c->corrected = true;
  c->built = false;
  st_all.stop();
  c->set_stat("reads_kept", (double)h[0]);
  c->set_stat("corrected_bases", (double)h[1]);
  c->set_stat("seeds", (double)h[2]);
  c->set_stat("substitutions", (double)h[3]);
  c->set_stat("reads_truncated", (double)h[4]);
  // B/4 + K*32 + B_out/4
  c->set_stat("alg_bytes_correct", (double)c->n_bases / 4 + 32.0 * (double)c->n_kmer_instances + (double)h[1] / 4);
}

}  // namespace bgx

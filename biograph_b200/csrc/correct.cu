// correct.cu -- k-mer based read correction and suffix-seed counts on the GPU.
//
// Replaces (reference, CPU): fast_read_correct / correct_internal
// (modules/bio_base/fast_read_correct.cpp:94-182, :16-90) and correct_reads::correct
// (bs/correct_reads.cpp:154-231, kmer_starts_read :308-311).
//
// Two kernels.
//   probe_kernel  : a warp per read (4 reads in turn).  The lanes look up ALL k-mers of the read at once (32
//                   independent probe sequences per step instead of one dependent chain), giving
//                   the read's "solid mask" plus the two starts-read flag masks.  A read whose
//                   k-mers are all solid is its own correction (fast_read_correct extends to the
//                   end without a substitution): its store words, rc copy and seed counts are
//                   written right here.  A read with no solid k-mer is dropped here.  The rest
//                   go on the slow list with their solid mask.
//   correct_kernel: one thread per slow read.  The reference's recursive DFS is run iteratively
//                   with an explicit frame stack; a partial result is (length, substitutions[<=16])
//                   instead of a copied sequence, because a corrected read is always the input
//                   with a few substituted bases, truncated.  A k-mer window that no substitution
//                   of the current path touches is answered from the solid mask, so only the
//                   <= k windows after each tried substitution probe the set.
// k-mer membership = one 256-bit load of the k-mer's home bucket in the solid hash set (32-byte
// buckets of four key|flags slots; a full bucket without the key sends the probe on to the next
// one): one DRAM sector per probe once the set outgrows L2.
#include <algorithm>

#include "ctx.h"

namespace bgx {
namespace {

// --max-corrections goes up to 32 on the reference's command line (biograph_create.cpp:486); the frame
// stack of the DFS is sized by it, so the kernel exists for 16 (the default of 8 fits) and for 32
constexpr int kMaxCorrLimit = 32;
constexpr int kMaxWords = 9;  // 255 bases -> 8 words + 1 pad

template <int MAXC>
struct ResT {
  int len;
  int ncorr;
  uint8_t pos[MAXC];  // logical positions of substitutions
  uint8_t base[MAXC];
};

template <int MAXC>
struct FrameT {
  uint64_t kmer;   // k-mer before `start` on ENTER; k-mer at the error point afterwards
  int start;       // logical start of this frame's input
  int run;         // bases extended without correction
  int e;           // logical index of the skipped (bad) base
  int b;           // substitution being tried
  int budget;
  int best_size;
  int best_b;
  ResT<MAXC> best;
};

struct Params {
  int k;
  int max_corr;
  int min_run;
  double trim;
  const unsigned long long* set;
  uint64_t set_mask;   // buckets - 1 (a bucket = 4 slots = one 32-byte sector)
  int set_shift;       // 2k - log2(buckets): the home bucket is the TOP bits of the (2k-bit) hash
  int pf;              // L2 prefetch-size hint of the bucket loads (ld_bucket4)
};

// solid mask of one read: bit (p & 31) of word (p >> 5) = the k-mer starting at read position p is
// in the solid set (windows holding an 'N' are 0)
struct SolidMask {
  const uint32_t* bits;
  __device__ __forceinline__ bool at(int p) const { return (bits[p >> 5] >> (p & 31)) & 1u; }
  // number of consecutive set bits at p, p+1, ... (at most lim)
  __device__ __forceinline__ int run_up(int p, int lim) const {
    int cnt = 0;
    while (cnt < lim) {
      const uint32_t inv = ~(bits[p >> 5] >> (p & 31));
      const int avail = min(32 - (p & 31), lim - cnt);
      const int ones = inv ? __ffs(inv) - 1 : 32;
      if (ones < avail) return cnt + ones;
      cnt += avail;
      p += avail;
    }
    return cnt;
  }
  // ... at p, p-1, ... (at most lim; p - lim + 1 >= 0)
  __device__ __forceinline__ int run_down(int p, int lim) const {
    int cnt = 0;
    while (cnt < lim) {
      const uint32_t inv = ~(bits[p >> 5] << (31 - (p & 31)));
      const int avail = min((p & 31) + 1, lim - cnt);
      const int ones = inv ? __clz(inv) : 32;
      if (ones < avail) return cnt + ones;
      cnt += avail;
      p -= avail;
    }
    return cnt;
  }
};

// logical view of a stretch of the read: forward from `origin`, or reverse-complemented
// walking down from `origin`
struct Input {
  const uint64_t* w;
  const uint32_t* m;  // nullptr when the read set has no N
  int origin;
  int n;
  bool rev;
  // read position of the first base of the k-mer formed by shifting in logical base i
  __device__ __forceinline__ int kmer_pos(int i, int k) const { return rev ? origin - i : origin + i - k + 1; }
  // the k-mer of logical bases i-k .. i-1, straight from the read words (valid when none of them
  // is a substitution of the current path)
  __device__ __forceinline__ uint64_t kmer_before(int i, int k) const {
    const int a = rev ? origin - i + 1 : origin + i - k;
    const int q = a >> 5;
    const unsigned s = (unsigned)(a & 31) * 2;
    const uint64_t win = s ? ((w[q] << s) | (w[q + 1] >> (64 - s))) : w[q];
    const uint64_t x = win >> (64 - 2 * k);
    return rev ? revcomp_kmer(x, k) : x;
  }
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
  uint64_t b = khash(canon, P.k) >> P.set_shift;
  for (;;) {
    unsigned long long kk[4];
    ld_bucket4(P.set, b, kk, P.pf);
    bool more;
    const unsigned long long e = bucket4_match(kk, canon, &more);
    if (!more) return e;
    b = (b + 1) & P.set_mask;
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
// A window is "fresh" when no substitution of the current DFS path lies inside it: its membership
// is the probe kernel's mask bit.  The substitutions of the path are the frames' e (ascending),
// so only the innermost one can be within k of the running position.
template <int MAXC>
__device__ void correct_internal(const Params& P, const Input& in, const SolidMask& sm, uint64_t kmer0, int min_run0,
                                 int budget0, bool require_run_at_end, FrameT<MAXC>* st, ResT<MAXC>* out) {
  using Frame = FrameT<MAXC>;
  using Res = ResT<MAXC>;
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
      const int last_sub = depth ? st[depth - 1].e : -P.k;  // innermost substitution of the path
      int c = in.get(it);
      if (c != 4) {
        uint64_t nk = shift_in(kmer, c, kmask);
        for (;;) {
          if (it - last_sub >= P.k) {
            // every window from here on is free of substitutions: the probe pass already answered
            // them all (a window holding an 'N' has a 0 bit, and ends the run at the same base as
            // the base-by-base loop would), so the run is a bit scan of the solid mask
            const int lim = in.n - it;
            const int p0 = in.kmer_pos(it, P.k);
            const int cnt = in.rev ? sm.run_down(p0, lim) : sm.run_up(p0, lim);
            if (cnt > 0) {
              run += cnt;
              it += cnt;
              if (it == in.n) finished = true;
              else kmer = in.kmer_before(it, P.k);
            }
            break;
          }
          if (!(it - last_sub >= P.k ? sm.at(in.kmer_pos(it, P.k)) : solid_has(P, nk))) break;
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
      const int orig = in.get(f.e);  // the base the extension stopped on: already known not to extend
      while (f.b < 4) {
        uint64_t tk = shift_in(f.kmer, f.b, kmask);
        if (f.b == orig || !solid_has(P, tk)) { ++f.b; continue; }
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

// totals: [0] reads kept [1] bases kept [2] seeds [3] substitutions [4] reads truncated
// Pass 1: a warp takes kProbeRPW reads, one at a time; see the file comment.
// MAXIT = ceil(max k-mers per read / 32).  Counters are summed per block (the totals and the
// slow-list cursor would otherwise be one same-address atomic per read).
constexpr int kProbeThreads = 256;
constexpr int kProbeWarps = kProbeThreads / 32;
constexpr int kProbeRPW = 4;
template <int MAXIT, bool HAS_N>
__global__ void __launch_bounds__(kProbeThreads, (MAXIT <= 4 ? 3 : 1)) probe_kernel(const uint64_t* __restrict__ words,
                                                              const uint32_t* __restrict__ nmask,
                                                              const uint32_t* __restrict__ word_off,
                                                              const uint16_t* __restrict__ lens, uint32_t n_reads, Params P,
                                                              uint64_t* __restrict__ store, uint64_t rc_word_base,
                                                              uint16_t* __restrict__ clen, uint8_t* __restrict__ ncorr,
                                                              uint16_t* __restrict__ next_fwd, uint16_t* __restrict__ next_rev,
                                                              unsigned long long* __restrict__ totals,
                                                              uint32_t* __restrict__ slow_list,
                                                              uint32_t* __restrict__ slow_mask /*[slot][MAXIT]*/,
                                                              unsigned int* __restrict__ n_slow) {
  __shared__ unsigned int blk_kept, blk_bases, blk_seeds, blk_slow, blk_slow_base;
  if (threadIdx.x == 0) blk_kept = blk_bases = blk_seeds = blk_slow = 0;
  __syncthreads();
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const int k = P.k;
  uint32_t slow_r[kProbeRPW], slow_m[kProbeRPW], slow_i[kProbeRPW];  // slow reads of this warp: read, my mask word, block-local slot
  int n_my_slow = 0;
  unsigned w_kept = 0, w_bases = 0, w_seeds = 0;  // lane 0 only

#pragma unroll 1
  for (int q = 0; q < kProbeRPW; ++q) {
    const uint32_t r = (blockIdx.x * kProbeWarps + warp) * kProbeRPW + q;
    if (r >= n_reads) break;  // warp-uniform
    const int L = lens[r];
    const int nw = (L + 31) >> 5;
    const uint32_t base = word_off[r];
    const int nk = L >= k ? L - k + 1 : 0;
    // lane i holds word i of the read (lane nw: the word after it; the store ends with a pad word)
    const uint64_t wv = (int)lane <= nw ? words[base + lane] : 0;
    uint32_t mv = 0;
    if (HAS_N) mv = (int)lane <= nw ? nmask[base + lane] : 0;

    // ---- all k-mers at once: first probes of every step issued back to back ----------------------
    uint64_t canon[MAXIT], slot[MAXIT];
    unsigned long long cur[MAXIT][4];   // the home bucket of every k-mer: one 256-bit load each
    bool live[MAXIT], flip[MAXIT];
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int p = it * 32 + (int)lane;
      const uint64_t hi = __shfl_sync(0xffffffffu, wv, it), lo = __shfl_sync(0xffffffffu, wv, it + 1);
      const unsigned s = lane * 2;
      const uint64_t win = s ? ((hi << s) | (lo >> (64 - s))) : hi;
      bool has_n = false;
      if (HAS_N) {
        const uint32_t mh = __shfl_sync(0xffffffffu, mv, it), ml = __shfl_sync(0xffffffffu, mv, it + 1);
        const uint32_t mwin = lane ? ((mh << lane) | (ml >> (32 - lane))) : mh;
        has_n = (mwin >> (32 - k)) != 0;
      }
      live[it] = p < nk && !has_n;
      bool fl;
      canon[it] = canonicalize(win >> (64 - 2 * k), k, fl);
      flip[it] = fl;
      slot[it] = khash(canon[it], k) >> P.set_shift;
      if (live[it]) ld_bucket4(P.set, slot[it], cur[it], P.pf);
      else cur[it][0] = cur[it][1] = cur[it][2] = cur[it][3] = kEmptyKey;
    }
    uint32_t my_mask = 0;       // lane it keeps the solid mask word of step it
    bool all_solid = nk > 0, any_solid = false;
    int first_f = -1, last_g = -1;  // first p >= 1 whose k-mer starts a read (as seen), last p <= nk-2 (rc view)
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int p = it * 32 + (int)lane;
      bool more;
      unsigned long long e = bucket4_match(cur[it], canon[it], &more);
      // home bucket full without the key: rare at load factor <= 1/2.  (Running these follow-up probes
      // in rounds over all steps at once was measured slower: 5.5 against 4.4 ms on E. coli 100x.)
      while (live[it] && more) {
        slot[it] = (slot[it] + 1) & P.set_mask;
        ld_bucket4(P.set, slot[it], cur[it], P.pf);
        e = bucket4_match(cur[it], canon[it], &more);
      }
      const bool found = live[it] && e != kEmptyKey;
      const unsigned bm = __ballot_sync(0xffffffffu, found);
      if ((int)lane == it) my_mask = bm;
      const int in_range = min(32, max(0, nk - it * 32));
      const unsigned want = in_range >= 32 ? 0xffffffffu : ((1u << in_range) - 1);
      all_solid = all_solid && bm == want;
      any_solid = any_solid || bm != 0;
      // kmer_starts_read as the read sees it (bs/correct_reads.cpp:195-210, :308-311)
      unsigned fm = __ballot_sync(0xffffffffu, found && (e & (flip[it] ? kRevFlag : kFwdFlag)) != 0 && p >= 1);
      unsigned gm = __ballot_sync(0xffffffffu, found && (e & (flip[it] ? kFwdFlag : kRevFlag)) != 0 && p <= nk - 2);
      if (first_f < 0 && fm) first_f = it * 32 + __ffs(fm) - 1;
      if (gm) last_g = it * 32 + 31 - __clz(gm);
    }

    if (all_solid) {
      // the read is its own correction: store it forward and reverse-complemented
      if ((int)lane < nw) {
        uint64_t v = wv;
        const int rem = L - 32 * (int)lane;
        if (rem < 32) v &= top_bases_mask(rem);
        store[base + lane] = v;
      }
      {
        // rc word j = revcomp of read[L - 32j - mcount, L - 32j)
        const int j = (int)lane;
        const int mcount = min(32, max(0, L - 32 * j));
        const int lo_pos = max(0, L - 32 * j - mcount);
        const uint64_t hi = __shfl_sync(0xffffffffu, wv, lo_pos >> 5), lo = __shfl_sync(0xffffffffu, wv, (lo_pos >> 5) + 1);
        const unsigned s = (unsigned)(lo_pos & 31) * 2;
        const uint64_t win = s ? ((hi << s) | (lo >> (64 - s))) : hi;
        const int mc = max(mcount, 1);  // lanes past the read compute a dummy
        if (j < nw) store[rc_word_base + base + j] = revcomp_kmer(win >> (64 - 2 * mc), mc) << (64 - 2 * mc);
      }
      if (lane == 0) {
        const int nf = first_f >= 0 ? first_f : nk;
        const int nr = last_g >= 0 ? nk - 1 - last_g : nk;
        clen[r] = (uint16_t)L;
        ncorr[r] = 0;
        next_fwd[r] = (uint16_t)nf;
        next_rev[r] = (uint16_t)nr;
        w_kept += 1;
        w_bases += (unsigned)L;
        w_seeds += (unsigned)(nf + nr);
      }
    } else if (!any_solid) {
      // shorter than k, or no solid k-mer to anchor on (fast_read_correct.cpp:108-123): dropped
      if ((int)lane < nw) {
        store[base + lane] = 0;
        store[rc_word_base + base + lane] = 0;
      }
      if (lane == 0) {
        clen[r] = 0;
        ncorr[r] = 0;
        next_fwd[r] = 0;
        next_rev[r] = 0;
      }
    } else {
      unsigned idx = 0;
      if (lane == 0) idx = atomicAdd(&blk_slow, 1u);
      slow_r[n_my_slow] = r;
      slow_m[n_my_slow] = my_mask;
      slow_i[n_my_slow] = __shfl_sync(0xffffffffu, idx, 0);
      ++n_my_slow;
    }
  }
  if (lane == 0 && w_kept) {
    atomicAdd(&blk_kept, w_kept);
    atomicAdd(&blk_bases, w_bases);
    atomicAdd(&blk_seeds, w_seeds);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (blk_kept) {
      atomicAdd(&totals[0], (unsigned long long)blk_kept);
      atomicAdd(&totals[1], (unsigned long long)blk_bases);
      atomicAdd(&totals[2], (unsigned long long)blk_seeds);
    }
    blk_slow_base = blk_slow ? atomicAdd(n_slow, blk_slow) : 0u;
  }
  __syncthreads();
  const unsigned sbase = blk_slow_base;
#pragma unroll
  for (int q = 0; q < kProbeRPW; ++q) {
    if (q < n_my_slow) {
      const unsigned idx = sbase + slow_i[q];
      if (lane == 0) slow_list[idx] = slow_r[q];
      if ((int)lane < MAXIT) slow_mask[(size_t)idx * MAXIT + lane] = slow_m[q];
    }
  }
}

// Pass 2: one thread per read on the slow list.
template <int MAXC>
__global__ void __launch_bounds__(128) correct_kernel(const uint64_t* __restrict__ words,
                                                      const uint32_t* __restrict__ nmask,
                                                      const uint32_t* __restrict__ word_off,
                                                      const uint16_t* __restrict__ lens, Params P,
                                                      const uint32_t* __restrict__ slow_list,
                                                      const uint32_t* __restrict__ slow_mask, int mask_words,
                                                      const unsigned int* __restrict__ n_slow_p,
                                                      uint64_t* __restrict__ store, uint64_t rc_word_base,
                                                      uint16_t* __restrict__ clen, uint8_t* __restrict__ ncorr,
                                                      uint16_t* __restrict__ next_fwd, uint16_t* __restrict__ next_rev,
                                                      unsigned long long* __restrict__ totals) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_slow = *n_slow_p;
  const bool active = j < n_slow;
  const uint32_t r = active ? slow_list[j] : 0;
  uint64_t w[kMaxWords];
  uint32_t m[kMaxWords];
  uint32_t smw[kMaxWords - 1];
  using Frame = FrameT<MAXC>;
  using Res = ResT<MAXC>;
  Frame st[MAXC + 1];
  int out_len = 0, corrections = 0, nf = 0, nr = 0;
  const int k = P.k;
  int L = 0, nw = 0;
  uint32_t base = 0;
  if (active) {
    L = lens[r];
    nw = (L + 31) >> 5;
    base = word_off[r];
#pragma unroll
    for (int i = 0; i < kMaxWords; ++i) {
      w[i] = i < nw ? words[base + i] : 0;
      m[i] = (nmask != nullptr && i < nw) ? nmask[base + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < kMaxWords - 1; ++i) smw[i] = i < mask_words ? slow_mask[(size_t)j * mask_words + i] : 0;
  }
  const SolidMask sm{smw};
  bool ok = active;  // on the list: L >= k and at least one solid k-mer
  if (ok) {
    // the first solid k-mer (fast_read_correct.cpp:108-123): lowest set mask bit
    int kmer_start = 0;
    while (!sm.at(kmer_start)) ++kmer_start;
    const int it = kmer_start + k;
    const uint64_t kmer = local_window(w, kmer_start) >> (64 - 2 * k);
    int budget = P.max_corr;
    if (kmer_start != 0) {
      // left side: correct the reverse complement of read[0, kmer_start) (:135-167)
      Input lin{w, nmask ? m : nullptr, kmer_start - 1, kmer_start, true};
      Res lres;
      correct_internal(P, lin, sm, revcomp_kmer(kmer, k), 0, budget, false, st, &lres);
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
        correct_internal(P, rin, sm, kmer, 0, budget, true, st, &rres);
        for (int i = 0; i < rres.ncorr; ++i) set_base(w, it + rres.pos[i], rres.base[i]);
        corrections += rres.ncorr;
        out_len = it + rres.len;
      }
      // drop rule: corrected.size() < unsigned(trim_after_portion * len) (bs/correct_reads.cpp:174-178)
      unsigned needed = (unsigned)(P.trim * (double)L);
      if ((unsigned)out_len < needed) ok = false;
    }
  }
  if (active) {
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
  reads_ready(c);
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
  reads_ready(c);
  BGX_CHECK(c->counted, "bgx_correct: call bgx_count_kmers first");
  BGX_CHECK(c->opt.max_corrections <= kMaxCorrLimit, "max_corrections > 32 is not supported");
  cudaStream_t s = c->stream;
  ScopedStage st_all(c, "correct_total");
  const uint64_t n = c->n_reads;
  c->store.alloc(2 * c->n_words + 1, s);
  c->clen.alloc(std::max<uint64_t>(n, 1), s);
  c->ncorr.alloc(std::max<uint64_t>(n, 1), s);
  c->next_fwd.alloc(std::max<uint64_t>(n, 1), s);
  c->next_rev.alloc(std::max<uint64_t>(n, 1), s);
  DevBuf<unsigned long long> totals(5, s);
  BGX_CUDA(cudaMemsetAsync(totals.p, 0, 5 * sizeof(unsigned long long), s));
  BGX_CUDA(cudaMemsetAsync(c->store.p + 2 * c->n_words, 0, sizeof(uint64_t), s));
  Params P;
  P.k = c->opt.kmer_size;
  P.max_corr = c->opt.max_corrections;
  P.min_run = c->opt.min_good_run;
  P.trim = (double)c->opt.trim_after_portion;  // float widened to double (biograph_create.cpp:489-490,731)
  P.set = c->solid.p;
  P.set_mask = c->solid_slots / 4 - 1;
  P.set_shift = 2 * P.k;
  P.pf = 0;
  if (const char* e = getenv("BGX_PROBE_PF")) P.pf = atoi(e);  // experiment hook
  for (uint64_t b = c->solid_slots / 4; b > 1; b >>= 1) --P.set_shift;
  const int max_kmers = std::max<int>((int)c->max_len - P.k + 1, 1);
  const int mask_words = max_kmers <= 128 ? 4 : 8;
  BGX_CHECK(max_kmers <= 256, "read longer than 255 bases");
  DevBuf<uint32_t> slow_list(std::max<uint64_t>(n, 1), s), slow_mask(std::max<uint64_t>(n, 1) * mask_words, s);
  DevBuf<unsigned int> n_slow(1, s);
  BGX_CUDA(cudaMemsetAsync(n_slow.p, 0, sizeof(unsigned int), s));
  unsigned int h_slow = 0;
  {
    ScopedStage st(c, "correct_probe");
    const unsigned grid = (unsigned)((n + kProbeWarps * kProbeRPW - 1) / (kProbeWarps * kProbeRPW));
    const uint32_t* nm = c->has_n ? c->nmask.p : nullptr;
    if (n == 0) {
      // a rank without reads (sharded build): nothing to launch, but it still takes part in the exchanges
    } else
#define BGX_PROBE(MAXIT, HASN)                                                                                          \
  note_launch();                                                                                                        \
  probe_kernel<MAXIT, HASN><<<grid, kProbeThreads, 0, s>>>(c->words.p, nm, c->word_off.p, c->lens.p, (uint32_t)n, P, \
                                                                     c->store.p, c->n_words, c->clen.p, c->ncorr.p,   \
                                                                     c->next_fwd.p, c->next_rev.p, totals.p,          \
                                                                     slow_list.p, slow_mask.p, n_slow.p)
    if (mask_words == 4 && c->has_n) { BGX_PROBE(4, true); }
    else if (mask_words == 4) { BGX_PROBE(4, false); }
    else if (c->has_n) { BGX_PROBE(8, true); }
    else { BGX_PROBE(8, false); }
#undef BGX_PROBE
    BGX_CUDA(cudaGetLastError());
    st.stop();
  }
  {
    ScopedStage st(c, "correct_kernel");
    // sized for the worst case (every read slow); threads past the device-side count leave at once
    BGX_CUDA(cudaMemcpyAsync(&h_slow, n_slow.p, sizeof(h_slow), cudaMemcpyDeviceToHost, s));
    BGX_CUDA(cudaStreamSynchronize(s));
#define BGX_CORRECT(MAXC)                                                                                                   \
  note_launch();                                                                                                            \
  correct_kernel<MAXC><<<(h_slow + 127) / 128, 128, 0, s>>>(c->words.p, c->has_n ? c->nmask.p : nullptr, c->word_off.p, \
                                                                      c->lens.p, P, slow_list.p, slow_mask.p, mask_words, n_slow.p, \
                                                                      c->store.p, c->n_words, c->clen.p, c->ncorr.p,               \
                                                                      c->next_fwd.p, c->next_rev.p, totals.p)
    if (h_slow && P.max_corr <= 16) { BGX_CORRECT(16); }
    else if (h_slow) { BGX_CORRECT(32); }
#undef BGX_CORRECT
    BGX_CUDA(cudaGetLastError());
    st.stop();
  }
  c->set_stat("reads_slow_path", (double)h_slow);
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

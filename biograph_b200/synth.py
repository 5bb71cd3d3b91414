"""Deterministic synthetic read sets for the parity tests and bench.py.

Not part of the hot path: plain numpy on the host."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def random_genome(length, seed, repeat_frac=0.0, repeat_seed=None):
    """i.i.d. uniform bases; optionally overwrite repeat_frac of the length with copies of earlier
    segments (1-10 kb) to create repeats."""
    rng = np.random.Generator(np.random.MT19937(seed))
    g = _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
    if repeat_frac > 0:
        r2 = np.random.Generator(np.random.MT19937(seed + 1 if repeat_seed is None else repeat_seed))
        done = 0
        while done < repeat_frac * length:
            seg = int(r2.integers(1000, 10001))
            if seg * 2 >= length:
                break
            src = int(r2.integers(0, length - seg))
            dst = int(r2.integers(0, length - seg))
            g[dst:dst + seg] = g[src:src + seg].copy()
            done += seg
    return g


def unpack_genome(packed, length):
    """2-bit packed (4 bases/byte, first base in the high bits) -> ASCII uint8 array."""
    p = np.asarray(packed, dtype=np.uint8)
    codes = np.stack([(p >> 6) & 3, (p >> 4) & 3, (p >> 2) & 3, p & 3], axis=1).reshape(-1)[:length]
    return _ACGT[codes]


def simulate_reads(genome, n_reads, read_len=150, error_rate=0.005, seed=20260101, paired=True, frag_mean=400,
                   frag_sd=40, n_rate=0.0):
    """Returns (uint8 array [n_reads, read_len] of ASCII bases).  Uniform start/strand, i.i.d.
    substitution errors to a different base, optional 'N' calls."""
    rng = np.random.Generator(np.random.MT19937(seed))
    G = len(genome)
    idx = np.arange(read_len, dtype=np.int64)
    if paired:
        nfrag = (n_reads + 1) // 2
        flen = np.clip(np.rint(rng.normal(frag_mean, frag_sd, nfrag)).astype(np.int64), read_len, G)
        start = (rng.random(nfrag) * (G - flen + 1)).astype(np.int64)
        r1 = genome[start[:, None] + idx]
        r2 = _COMP[genome[(start + flen - 1)[:, None] - idx]]  # rc of the fragment's far end
        flip = rng.random(nfrag) < 0.5
        a = np.where(flip[:, None], r2, r1)
        b = np.where(flip[:, None], r1, r2)
        reads = np.empty((2 * nfrag, read_len), dtype=np.uint8)
        reads[0::2] = a
        reads[1::2] = b
        reads = reads[:n_reads]
    else:
        start = (rng.random(n_reads) * (G - read_len + 1)).astype(np.int64)
        reads = genome[start[:, None] + idx]
        flip = rng.random(n_reads) < 0.5
        rc = _COMP[reads[:, ::-1]]
        reads = np.where(flip[:, None], rc, reads)
    reads = np.ascontiguousarray(reads)
    total = reads.size
    if error_rate > 0:
        nerr = rng.binomial(total, error_rate)
        pos = np.unique(rng.integers(0, total, size=nerr))
        flat = reads.reshape(-1)
        code = np.zeros(256, dtype=np.uint8)
        code[_ACGT] = np.arange(4, dtype=np.uint8)
        newc = (code[flat[pos]] + rng.integers(1, 4, size=len(pos), dtype=np.uint8)) & 3
        flat[pos] = _ACGT[newc]
    if n_rate > 0:
        nn = rng.binomial(total, n_rate)
        pos = rng.integers(0, total, size=nn)
        reads.reshape(-1)[pos] = ord("N")
    return reads


def as_buffer(reads2d):
    """[n, L] uint8 -> (bytes-like uint8 array, int64 offsets[n+1])"""
    n, L = reads2d.shape
    return np.ascontiguousarray(reads2d).reshape(-1), np.arange(n + 1, dtype=np.int64) * L

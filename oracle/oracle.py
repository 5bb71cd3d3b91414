"""ctypes binding for the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
The product package (biograph_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        p = C.POINTER
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_max_threads.restype = C.c_int
        L.orc_rev_comp.restype = C.c_uint64
        L.orc_rev_comp.argtypes = [C.c_uint64, C.c_int]
        L.orc_count_kmers.restype = C.c_int64
        L.orc_count_kmers.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                      p(C.c_void_p), p(C.c_void_p), p(C.c_void_p), p(C.c_void_p)]
        L.orc_count_kmers2.restype = C.c_int64
        L.orc_count_kmers2.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                       p(C.c_void_p), p(C.c_void_p), p(C.c_void_p), p(C.c_void_p)]
        L.orc_fast_read_correct.restype = C.c_int
        L.orc_fast_read_correct.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                            C.c_int, C.c_char_p, p(C.c_int)]
        L.orc_correct_reads.restype = C.c_int64
        L.orc_correct_reads.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_seqset_closed_form.restype = C.c_int64
        L.orc_seqset_closed_form.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_int, p(C.c_void_p),
                                             p(C.c_void_p), p(C.c_void_p), C.c_void_p]
        L.orc_seqset_staged.restype = C.c_int64
        L.orc_seqset_staged.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                        p(C.c_void_p), p(C.c_void_p), p(C.c_void_p), C.c_void_p, C.c_void_p]
        L.orc_bitcount_finalize.restype = C.c_uint64
        L.orc_bitcount_finalize.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_varbit_words.restype = C.c_uint64
        L.orc_varbit_words.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_varbit_pack.restype = C.c_int
        L.orc_varbit_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        _LIB = L
    return _LIB


def max_threads():
    return lib().orc_max_threads()


def pack_reads(reads):
    """list of str/bytes -> (bytes buffer, int64 offsets[n+1])"""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    offs = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=offs[1:])
    return b"".join(bs), offs


def _take(ptr, n, dtype):
    n = int(n)
    if n == 0:
        arr = np.zeros(0, dtype=dtype)
    else:
        arr = np.frombuffer((C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr.value), dtype=dtype).copy()
    lib().orc_free(ptr)
    return arr


def count_kmers(reads, k, threads=0, prefilter_min=0):
    """The distinct canonical k-mers, ascending, with exact fwd/rev counts and flag bits
    (bit0 fwd_starts_read, bit1 rev_starts_read).  prefilter_min > 0: the reference's two-stage form
    (probabilistic 2-bit pass first; only k-mers that pass it are counted exactly) -- same solid set."""
    buf, offs = reads if isinstance(reads, tuple) else pack_reads(reads)
    pk, pf, pr, pfl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = lib().orc_count_kmers2(buf, offs.ctypes.data, len(offs) - 1, k, threads, prefilter_min, C.byref(pk),
                               C.byref(pf), C.byref(pr), C.byref(pfl))
    if n < 0:
        raise ValueError("orc_count_kmers failed")
    return {"kmers": _take(pk, n, np.uint64), "fwd": _take(pf, n, np.uint32), "rev": _take(pr, n, np.uint32),
            "flags": _take(pfl, n, np.uint8)}


def solid_set(counts, min_count):
    """kmerizer kmer_passes: keep iff fwd+rev >= min_count (modules/bio_mapred/kmerize_bf.cpp:290-318)."""
    tot = counts["fwd"].astype(np.uint64) + counts["rev"].astype(np.uint64)
    m = tot >= min_count
    return {k_: v[m] for k_, v in counts.items()}


def fast_read_correct(read, solid_kmers, k, max_corrections=2, min_good_run=2):
    solid_kmers = np.ascontiguousarray(solid_kmers, dtype=np.uint64)
    rb = read.encode() if isinstance(read, str) else read
    out = C.create_string_buffer(max(1, len(rb)))
    corr = C.c_int(0)
    n = lib().orc_fast_read_correct(rb, len(rb), solid_kmers.ctypes.data, len(solid_kmers), k, max_corrections,
                                    min_good_run, out, C.byref(corr))
    return out.raw[:n].decode(), corr.value


def correct_reads(reads, solid, k, max_corrections=8, min_good_run=2, trim_after_portion=0.7, threads=0):
    """Returns dict(seq=bytes, offs=int64[n+1], kept=u8[n], corrections=i32[n], next_fwd, next_rev)."""
    buf, offs = reads if isinstance(reads, tuple) else pack_reads(reads)
    n = len(offs) - 1
    kmers = np.ascontiguousarray(solid["kmers"], dtype=np.uint64)
    flags = np.ascontiguousarray(solid["flags"], dtype=np.uint8)
    out = C.create_string_buffer(max(1, len(buf)))
    out_offs = np.zeros(n + 1, dtype=np.int64)
    kept = np.zeros(n, dtype=np.uint8)
    corr = np.zeros(n, dtype=np.int32)
    nf = np.zeros(n, dtype=np.int32)
    nr = np.zeros(n, dtype=np.int32)
    nk = lib().orc_correct_reads(buf, offs.ctypes.data, n, kmers.ctypes.data, flags.ctypes.data, len(kmers), k,
                                 max_corrections, min_good_run, trim_after_portion, threads, out,
                                 out_offs.ctypes.data, kept.ctypes.data, corr.ctypes.data, nf.ctypes.data,
                                 nr.ctypes.data)
    return {"seq": out.raw[:out_offs[n]], "offs": out_offs, "kept": kept, "corrections": corr, "next_fwd": nf,
            "next_rev": nr, "n_kept": nk}


def corrected_list(cr):
    s, o = cr["seq"], cr["offs"]
    return [s[o[i]:o[i + 1]].decode() for i in range(len(o) - 1) if cr["kept"][i]]


def _seqset_result(n, ps, psh, ppv, fixed):
    if n < 0:
        raise RuntimeError("oracle: Missing expansion?")
    words = (n + 63) // 64
    sizes = _take(ps, n, np.uint16)
    shared = _take(psh, n, np.uint16)
    prev = _take(ppv, 4 * words, np.uint64).reshape(4, words) if n else np.zeros((4, 0), np.uint64)
    return {"n": int(n), "sizes": sizes, "shared": shared, "prev": prev, "fixed": fixed}


def seqset_closed_form(reads, threads=0):
    buf, offs = reads if isinstance(reads, tuple) else pack_reads(reads)
    ps, psh, ppv = C.c_void_p(), C.c_void_p(), C.c_void_p()
    fixed = np.zeros(5, dtype=np.uint64)
    n = lib().orc_seqset_closed_form(buf, offs.ctypes.data, len(offs) - 1, threads, C.byref(ps), C.byref(psh),
                                     C.byref(ppv), fixed.ctypes.data)
    return _seqset_result(n, ps, psh, ppv, fixed)


def seqset_staged(reads, next_fwd=None, next_rev=None, threads=0):
    buf, offs = reads if isinstance(reads, tuple) else pack_reads(reads)
    ps, psh, ppv = C.c_void_p(), C.c_void_p(), C.c_void_p()
    fixed = np.zeros(5, dtype=np.uint64)
    stats = np.zeros(6, dtype=np.int64)
    nf = None if next_fwd is None else np.ascontiguousarray(next_fwd, dtype=np.int32)
    nr = None if next_rev is None else np.ascontiguousarray(next_rev, dtype=np.int32)
    n = lib().orc_seqset_staged(buf, offs.ctypes.data, len(offs) - 1, None if nf is None else nf.ctypes.data,
                                None if nr is None else nr.ctypes.data, threads, C.byref(ps), C.byref(psh),
                                C.byref(ppv), fixed.ctypes.data, stats.ctypes.data)
    r = _seqset_result(n, ps, psh, ppv, fixed)
    r["stats"] = stats
    return r


def bitcount_finalize(bits, nbits):
    bits = np.ascontiguousarray(bits, dtype=np.uint64)
    sub = np.zeros(max(1, (nbits + 511) // 512), dtype=np.uint64)
    acc = np.zeros((nbits + 1 + 511) // 512, dtype=np.uint64)
    tot = lib().orc_bitcount_finalize(bits.ctypes.data, nbits, sub.ctypes.data, acc.ctypes.data)
    return sub[:(nbits + 511) // 512], acc, int(tot)


def varbit_pack(vals, max_value):
    vals = np.ascontiguousarray(vals, dtype=np.uint16)
    words = lib().orc_varbit_words(len(vals), max_value)
    out = np.zeros(words, dtype=np.uint64)
    bits = lib().orc_varbit_pack(vals.ctypes.data, len(vals), max_value, out.ctypes.data)
    return out, bits


# ---- pure-python helpers shared by tests (small cases only) -------------------------------------
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def revcomp(s):
    return "".join(_COMP[c] for c in reversed(s))


def tseq(s):
    """modules/bio_base/dna_testutil.cpp:14-33: each char -> 'C' + 8 bits LSB-first (T=1,A=0) + 'C'."""
    out = []
    for ch in s.encode():
        out.append("C")
        for i in range(8):
            out.append("T" if ch & (1 << i) else "A")
        out.append("C")
    return "".join(out)


def tseq_rc(s):
    return revcomp(tseq(s))


def entries_closed_form_py(reads):
    """Pure-python closed form returning the entry strings; tiny inputs only."""
    S = set()
    for r in reads:
        for x in (r, revcomp(r)):
            for i in range(len(x)):
                S.add(x[i:])
    S = sorted(S)
    return [s for j, s in enumerate(S) if not (j + 1 < len(S) and S[j + 1].startswith(s))]

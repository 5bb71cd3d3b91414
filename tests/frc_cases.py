"""Analytic expectations of modules/bio_base/fast_read_correct_test.cpp:108-258, restated so the
same cases can be run against the oracle (CPU) and the CUDA corrector (GPU).

Parameters as in the reference test: kmer_size 30, max_corrections 2, min_good_run 2; the k-mer
set is every k-mer of the error-free sequence."""
from oracle import oracle as O

K, MAXC, RUN = 30, 2, 2
LONG = O.tseq("aBcDeFgHiJkLmNoPqRsTuVwXyZ")
SIZES = [30, 31, 32, 33, 58, 59, 60, 61, 62, 88, 89, 90, 91, 92, 118, 119, 120, 121, 122, 260]


def enc(km):
    v = 0
    for c in km:
        v = v * 4 + "ACGT".index(c)
    return v


def kmer_set_of(seq):
    import numpy as np
    s = set()
    for i in range(len(seq) - K + 1):
        km = seq[i:i + K]
        s.add(min(enc(km), enc(O.revcomp(km))))
    return np.array(sorted(s), dtype=np.uint64)


def add_error(seq, orig_seq, pos, mode):
    c = "N" if mode == "N" else "ACGT"["ACGT".index(orig_seq[pos]) ^ 1]
    return seq[:pos] + c + seq[pos + 1:]


def cases(size, mode, with_three=False):
    """yields (name, read, expected_corrected, expected_corrections)"""
    seq = LONG[:size]
    n = len(seq)
    yield ("no_errors", seq, seq, 0)
    for i in range(n - RUN):  # single
        exp, ec = seq, 1
        if i < K and i >= n - K:
            exp, ec = "", 0
        yield (f"single[{i}]", add_error(seq, seq, i, mode), exp, ec)
    for i in range(n - RUN, n):  # single_trunc
        exp = seq[:i]
        if i < K and i >= n - K:
            exp = ""
        yield (f"single_trunc[{i}]", add_error(seq, seq, i, mode), exp, 0)
    for i in range(n):  # two_errors
        for j in range(i + 1, n):
            err = add_error(add_error(seq, seq, i, mode), seq, j, mode)
            exp, ec = seq, 2
            if i < K and (j - i - 1) < K and j >= n - K:
                exp, ec = "", 0
            elif (j - i - 1) < RUN:
                exp = "" if i < K else seq[:i]
                ec = 0
            elif j >= n - RUN:
                exp, ec = seq[:j], 1
            yield (f"two[{i},{j}]", err, exp, ec)
    if with_three:
        for i in range(n):
            for j in range(i + 1, n):
                for k in range(j + 1, n):
                    err = add_error(add_error(add_error(seq, seq, i, mode), seq, j, mode), seq, k, mode)
                    if i < K and (j - i - 1) < K and (k - j - 1) < K:
                        exp, ec = "", 0
                    elif (j - i - 1) < RUN:
                        exp, ec = ("" if i < K else seq[:i]), 0
                    elif (k - j - 1) < RUN:
                        exp, ec = ("", 0) if j < K else (seq[:j], 1)
                    else:
                        exp, ec = seq[:k], 2
                    yield (f"three[{i},{j},{k}]", err, exp, ec)

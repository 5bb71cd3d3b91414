"""Known answers of the reference's own modules/build_seqset/correct_reads_test.cpp:121-223, restated:
which suffixes correct_reads::correct seeds for a read (the read's first next_fwd suffixes and
the first next_rev suffixes of its reverse complement), given an explicit k-mer set whose
starts-read flags mark the first k-mer of every added sequence and of its reverse complement
(add_kmers, :38-46; start_correction :48-66).  k = 2 * k_dna_test_sequence_length = 20.

Every case: (name, kmer_seqs, reads, expected seed set, exact).  exact=False is the reference's
ContainsAllOf (the *_in_repo cases; the initial repo itself is compression only)."""
from oracle import oracle as O

K = 20
t, trc, rc = O.tseq, O.tseq_rc, O.revcomp

CASES = [
    ("simple_fwd", [t("bcde"), "GA" + t("bcde") + "A"], ["GA" + t("bcde") + "A"],
     {"GA" + t("bcde") + "A", "A" + t("bcde") + "A", rc("GA" + t("bcde") + "A")}, True),
    ("simple_rc", [t("bcde"), "A" + t("bcde") + "GA"], ["A" + t("bcde") + "GA"],
     {rc("A" + t("bcde") + "GA"), rc("A" + t("bcde") + "G"), "A" + t("bcde") + "GA"}, True),
    ("fwd_in_repo", [t("abcd"), t("ghij")], [t("abcd"), t("ghij")],
     {t("abcd"), trc("abcd"), t("ghij"), trc("ghij")}, False),
    ("almost_in_repo",
     ["G" + t("abcd"), ("G" + t("abcd"))[1:1 + len(t("abcd")) - 1], t("ghij") + "G",
      (t("ghij") + "G")[1:1 + len(t("ghij")) - 1]],
     ["G" + t("abcd"), t("ghij") + "G"],
     {"G" + t("abcd"), trc("abcd") + "C", t("ghij") + "G", "C" + trc("ghij")}, True),
]


def kmer_set_for(seqs, k=K):
    """(sorted canonical k-mers, flags) as correct_reads_test::start_correction builds them"""
    import numpy as np

    def enc(s):
        v = 0
        for ch in s:
            v = (v << 2) | "ACGT".index(ch)
        return v

    def canon(s):
        a, b = enc(s), enc(rc(s))
        return min(a, b)

    kmers, first = set(), set()
    for s in seqs:
        assert len(s) >= k
        for i in range(len(s) - k + 1):
            kmers.add(canon(s[i:i + k]))
        first.add(enc(s[:k]))
        first.add(enc(rc(s)[:k]))
    ks = sorted(kmers)

    def dec(v):
        return "".join("ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))

    flags = []
    for v in ks:
        f = 0
        if v in first:
            f |= 1  # k_fwd_starts_read
        if enc(rc(dec(v))) in first:
            f |= 2  # k_rev_starts_read
        flags.append(f)
    return {"kmers": np.array(ks, dtype=np.uint64), "flags": np.array(flags, dtype=np.uint8)}


def seeds_of(read, nf, nr):
    r = rc(read)
    return {read[i:] for i in range(nf)} | {r[i:] for i in range(nr)}

"""Pins the CPU oracle against the reference's own golden vectors and known answers
(SURVEY.md 8c).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O


@pytest.fixture(scope="module")
def golden_pipeline(golden_reads):
    counts = O.count_kmers(golden_reads, 30)
    solid = O.solid_set(counts, 5)
    cr = O.correct_reads(golden_reads, solid, 30, max_corrections=8, min_good_run=2, trim_after_portion=0.7)
    return counts, solid, cr


def test_golden_stage_counts(golden_reads, golden_pipeline):
    # golden/e_coli_10000snp.bg/qc/create_log.txt, create_stats.json (version-stable counts)
    counts, solid, cr = golden_pipeline
    assert int(counts["fwd"].sum() + counts["rev"].sum()) == 60000
    assert len(solid["kmers"]) == 7108
    assert cr["n_kept"] == 8444
    lens = np.diff(cr["offs"])
    assert int(lens.sum()) == 288464
    kept = cr["kept"].astype(bool)
    in_len = np.array([len(r) for r in golden_reads])
    trunc = kept & (lens < in_len)
    assert int(trunc.sum()) == 2219 and int((in_len - lens)[trunc].sum()) == 7076
    assert int(cr["corrections"].sum()) == 0


def _check_members(ss, golden):
    assert ss["n"] == 19935
    assert np.array_equal(ss["fixed"], golden["fixed"])
    assert np.array_equal(ss["sizes"].astype(np.uint8), golden["entry_sizes"])
    assert np.array_equal(ss["shared"].astype(np.uint8), golden["shared"])
    for b, ch in enumerate("ACGT"):
        assert np.array_equal(ss["prev"][b], golden[f"prev_{ch}_bits"])
        sub, acc, tot = O.bitcount_finalize(ss["prev"][b], ss["n"])
        assert np.array_equal(sub, golden[f"prev_{ch}_subaccum"])
        assert np.array_equal(acc, golden[f"prev_{ch}_accum"])
        assert tot == int(golden["fixed"][b + 1] - golden["fixed"][b])


def test_golden_seqset_closed_form(golden, golden_pipeline):
    _, _, cr = golden_pipeline
    _check_members(O.seqset_closed_form((cr["seq"], cr["offs"])), golden)


def test_golden_seqset_staged(golden, golden_pipeline):
    _, _, cr = golden_pipeline
    st = O.seqset_staged((cr["seq"], cr["offs"]), cr["next_fwd"], cr["next_rev"])
    _check_members(st, golden)
    # this commit's seeding/expansion round sizes (the 2018 log's differ and are non-normative)
    assert list(st["stats"]) == [23076, 12099, 9435, 15614, 21090, 19935]


def test_golden_corrections_invariant(golden_reads, golden_pipeline):
    # the golden input only triggers truncations, so max-corrections 0 or 2 give the same reads
    _, solid, cr = golden_pipeline
    for mc in (0, 2):
        cr2 = O.correct_reads(golden_reads, solid, 30, max_corrections=mc)
        assert cr2["seq"] == cr["seq"] and np.array_equal(cr2["offs"], cr["offs"])


# ---- bs/builder_test.cpp:52-122 ------------------------------------------------------------
SEQ1 = ["AAAATTAC", "AAATTAC", "AATTAC", "AATTTTAG", "AC", "AG", "ATTAC", "ATTTTAG", "CTAAAATTAC", "GTAATTTTAG",
        "TAAAATTAC", "TAATTTTAG", "TAC", "TAG", "TTAC", "TTAG", "TTTAG", "TTTTAG"]


def _entries(reads, ss):
    """Reconstruct entry strings from the python closed form and check the tables against them."""
    E = O.entries_closed_form_py(reads)
    assert ss["n"] == len(E)
    assert list(ss["sizes"]) == [len(e) for e in E]
    lcp = lambda a, b: next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
    assert list(ss["shared"]) == [0] + [lcp(E[i - 1], E[i]) for i in range(1, len(E))]
    prev = np.zeros((4, (len(E) + 63) // 64), dtype=np.uint64)
    for e in E:
        p = e[1:]
        i = next(i for i, f in enumerate(E) if f.startswith(p))
        prev["ACGT".index(e[0])][i // 64] |= np.uint64(1) << np.uint64(i % 64)
    assert np.array_equal(ss["prev"], prev)
    return E


def test_builder_seq1():
    reads = [O.tseq("a")]
    for ss in (O.seqset_closed_form(reads), O.seqset_staged(reads)):
        assert _entries(reads, ss) == SEQ1


@pytest.mark.parametrize("reads,size", [
    ([O.tseq("abcdefg")], 129),
    ([O.tseq("abcd"), O.tseq("cdef"), O.tseq_rc("efgh")], 152),
    ([O.tseq("ab"), O.tseq("bc"), O.tseq("cd"), O.tseq("be")], 91),
    ([O.tseq("AB"), O.tseq("BC"), O.tseq("CD"), O.tseq("BE")], 91),
    ([O.tseq("abc"), O.tseq("cde")], 89),
    ([O.tseq("abc"), O.tseq("efg")], 99),
])
def test_builder_sizes(reads, size):
    for ss in (O.seqset_closed_form(reads), O.seqset_staged(reads)):
        assert ss["n"] == size
        _entries(reads, ss)
        # verify_seqset (bs/builder_test.cpp:18-40): fixed totals the entry count
        assert int(ss["fixed"][4]) == size


def test_staged_with_seed_of_one_closes():
    # one seed per strand forces the expansion rounds to do all the work (bs/expand_test.cpp:87-91)
    rng = np.random.default_rng(5)
    reads = ["".join("ACGT"[i] for i in rng.integers(0, 4, 120)) for _ in range(40)]
    reads += [r[10:90] for r in reads[:10]]
    ones = np.ones(len(reads), dtype=np.int32)
    a = O.seqset_closed_form(reads)
    b = O.seqset_staged(reads, ones, ones)
    assert a["n"] == b["n"] and np.array_equal(a["sizes"], b["sizes"]) and np.array_equal(a["prev"], b["prev"])
    assert np.array_equal(a["shared"], b["shared"])


# ---- bs/kmer_counter_test.cpp:88-121 brute force ---------------------------------------------
def _brute_counts(reads, k):
    d = {}
    for r in reads:
        for i in range(len(r) - k + 1):
            km = r[i:i + k]
            if "N" in km:
                continue
            rc = O.revcomp(km)
            flipped = rc < km
            canon = rc if flipped else km
            e = d.setdefault(canon, [0, 0, 0])
            e[1 if flipped else 0] += 1
            first, last = i == 0, i == len(r) - k
            if flipped:
                first, last = last, first
            e[2] |= (1 if first else 0) | (2 if last else 0)
    return d


def _enc(km):
    v = 0
    for c in km:
        v = v * 4 + "ACGT".index(c)
    return v


@pytest.mark.parametrize("k", [16, 21, 30, 31])
def test_kmer_counts_bruteforce(k):
    rng = np.random.default_rng(k)
    base = "".join("ACGT"[i] for i in rng.integers(0, 4, 400))
    reads = []
    for _ in range(300):
        s = int(rng.integers(0, 250))
        r = base[s:s + int(rng.integers(k - 2, 150))]
        if rng.random() < 0.5:
            r = O.revcomp(r)
        r = list(r)
        for j in range(len(r)):
            if rng.random() < 0.01:
                r[j] = "N"
        reads.append("".join(r))
    reads += [base[:60]] * 300  # overflow past 255 (bs/kmer_counter_test.cpp "overflow")
    c = O.count_kmers(reads, k)
    d = _brute_counts(reads, k)
    assert len(d) == len(c["kmers"])
    exp = sorted((_enc(km), v[0], v[1], v[2]) for km, v in d.items())
    got = list(zip(c["kmers"].tolist(), c["fwd"].tolist(), c["rev"].tolist(), c["flags"].tolist()))
    assert got == exp
    assert max(c["fwd"].max(), c["rev"].max()) > 255


# ---- encoders -------------------------------------------------------------------------------
def test_varbit_pack_roundtrip():
    rng = np.random.default_rng(1)
    for maxv in (1, 34, 35, 150, 255, 256, 1000):
        vals = rng.integers(0, maxv + 1, 1001).astype(np.uint16)
        out, bits = O.varbit_pack(vals, maxv)
        assert bits == int(maxv).bit_length()
        raw = int.from_bytes(out.tobytes(), "little")
        got = [(raw >> (i * bits)) & ((1 << bits) - 1) for i in range(len(vals))]
        assert got == vals.tolist()
    out, bits = O.varbit_pack(np.arange(10, dtype=np.uint16), 255)
    assert out.tobytes()[:10] == bytes(range(10))  # byte-aligned fast path layout


@pytest.mark.parametrize("nbits", [1, 63, 64, 511, 512, 513, 1024, 5000])
def test_bitcount_layout(nbits):
    rng = np.random.default_rng(nbits)
    bits = np.zeros((nbits + 63) // 64, dtype=np.uint64)
    idx = np.flatnonzero(rng.random(nbits) < 0.4)
    for i in idx:
        bits[i // 64] |= np.uint64(1) << np.uint64(i % 64)
    sub, acc, tot = O.bitcount_finalize(bits, nbits)
    assert tot == len(idx)
    assert len(acc) == (nbits + 1 + 511) // 512 and len(sub) == (nbits + 511) // 512
    pc = [bin(int(w)).count("1") for w in bits]
    for g in range(len(sub)):
        assert int(acc[g]) == sum(pc[:8 * g])
        grp = pc[8 * g:8 * g + 8] + [0] * 8
        assert int(sub[g]) == int.from_bytes(bytes(grp[:8]), "big")
    if nbits % 512 == 0:
        assert int(acc[-1]) == tot

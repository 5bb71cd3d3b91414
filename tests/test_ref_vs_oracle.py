"""The CPU oracle against the REFERENCE'S OWN code (oracle/_ref/libref.so: kmer_counter, kmer_set,
fast_read_correct, build_seqset::correct_reads, part_repo, expander, builder, seqset, bitcount and
packed_varbit_vector compiled from the sources under /root/reference -- see oracle/ref_shim.cpp), on the same inputs.
This is what pins the restatement to the reference itself rather than to its fixtures only.  CPU only; skipped where
the library was not built (it needs the reference checkout at build time, not at run time)."""
import numpy as np
import pytest

from biograph_b200 import synth
from oracle import oracle as O
from oracle import ref as R
from tests import frc_cases as F

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref.so not built (no reference checkout)")


def reads_of(genome_len, n_reads, read_len, err, seed, n_rate=0.0, ragged=False, repeat_frac=0.0):
    genome = synth.random_genome(genome_len, seed=seed, repeat_frac=repeat_frac)
    r2 = synth.simulate_reads(genome, n_reads, read_len=read_len, error_rate=err, seed=seed + 1, paired=True,
                              frag_mean=max(read_len + 20, 2 * read_len), frag_sd=10)
    rng = np.random.default_rng(seed + 2)
    out = []
    for row in r2:
        s = bytes(row).decode()
        if n_rate:
            s = "".join("N" if rng.random() < n_rate else c for c in s)
        if ragged:
            s = s[:int(rng.integers(1, read_len + 1))]
        out.append(s)
    return out


def compare_pipeline(reads, k=30, min_count=5, max_corrections=8, min_good_run=2, trim=0.7, threads=4):
    ref = R.create(reads, k, min_count, max_corrections, min_good_run, trim, threads=threads, with_members=True)
    oc = O.count_kmers(reads, k)
    osol = O.solid_set(oc, min_count)
    # the solid set: same k-mers in the same (ascending) order
    assert np.array_equal(osol["kmers"], ref["solid"]["kmers"])
    # exact counts and flags of every k-mer the reference's exact tables hold at or above min_count (k-mers the
    # probabilistic pass filtered are not guaranteed to be there below it, build_seqset/kmer_counter.h:121-123)
    c = ref["counts"]
    m = (c["fwd"].astype(np.int64) + c["rev"]) >= min_count
    assert np.array_equal(c["kmers"][m], osol["kmers"])
    for f in ("fwd", "rev", "flags"):
        assert np.array_equal(c[f][m], osol[f]), f
    # every k-mer the reference did count exactly has the oracle's numbers, below min_count too
    idx = np.searchsorted(oc["kmers"], c["kmers"])
    assert np.array_equal(oc["kmers"][idx], c["kmers"])
    for f in ("fwd", "rev", "flags"):
        assert np.array_equal(oc[f][idx], c[f]), f
    # corrected reads: which survive, and their bases
    ocr = O.correct_reads(reads, osol, k, max_corrections, min_good_run, trim)
    rcr = ref["corrected"]
    assert np.array_equal(ocr["kept"], rcr["kept"])
    assert np.array_equal(ocr["offs"], rcr["offs"])
    assert ocr["seq"] == rcr["seq"]
    # the seqset: every table
    rss = ref["seqset"]
    for form in (O.seqset_staged((ocr["seq"], ocr["offs"]), ocr["next_fwd"], ocr["next_rev"]),
                 O.seqset_closed_form((ocr["seq"], ocr["offs"]))):
        assert form["n"] == rss["n"]
        for t in ("sizes", "shared", "prev", "fixed"):
            assert np.array_equal(form[t], rss[t]), t
    check_members(ref["members"], rss)
    return ref


def check_members(members, ss):
    """the payload members as the reference's encoders wrote them against the oracle's encoders"""
    n = ss["n"]
    assert np.array_equal(np.frombuffer(members["fixed"], dtype=np.uint64), ss["fixed"])
    if n == 0:
        return
    for name, vals, mx in (("entry_sizes", ss["sizes"], int(ss["sizes"].max())),
                           ("shared", ss["shared"], int(ss["sizes"].max()) - 1)):
        got = np.frombuffer(members[name + "/elements"], dtype=np.uint64)
        assert np.array_equal(got, O.varbit_pack(vals, mx)[0]), name
    for b, ch in enumerate("ACGT"):
        sub, acc, tot = O.bitcount_finalize(ss["prev"][b], n)
        bits = np.frombuffer(members[f"prev_{ch}/bits"], dtype=np.uint64)
        assert np.array_equal(bits[:len(ss["prev"][b])], ss["prev"][b]) and not bits[len(ss["prev"][b]):].any()
        assert np.array_equal(np.frombuffer(members[f"prev_{ch}/subaccum"], dtype=sub.dtype), sub)
        assert np.array_equal(np.frombuffer(members[f"prev_{ch}/accum"], dtype=acc.dtype), acc)
        assert tot == int(ss["fixed"][b + 1] - ss["fixed"][b])


def test_builder_test_known_answers_on_the_reference():
    # modules/build_seqset/builder_test.cpp:52-122, run on the reference's own builder through the stand-in build:
    # shows the stand-in headers did not change what the reference computes
    cases = [([O.tseq("a")], 18), ([O.tseq("abcdefg")], 129), ([O.tseq("abcd"), O.tseq("cdef"), O.tseq_rc("efgh")], 152),
             ([O.tseq("ab"), O.tseq("bc"), O.tseq("cd"), O.tseq("be")], 91),
             ([O.tseq("AB"), O.tseq("BC"), O.tseq("CD"), O.tseq("BE")], 91), ([O.tseq("abc"), O.tseq("cde")], 89),
             ([O.tseq("abc"), O.tseq("efg")], 99)]
    for depth in (1, 2, 3):
        for reads, n in cases:
            ss = R.seqset_for_reads(reads, partition_depth=depth, threads=2)
            assert ss["n"] == n
            o = O.seqset_closed_form(reads)
            for t in ("sizes", "shared", "prev", "fixed"):
                assert np.array_equal(o[t], ss[t]), t


def test_golden_through_the_reference(golden, golden_reads):
    # the reference at this commit rebuilds its own 2018 golden BioGraph from the golden reads, and the oracle
    # agrees with it at every stage
    ref = compare_pipeline(golden_reads, threads=4)
    ss = ref["seqset"]
    assert ss["n"] == 19935 and len(ref["solid"]["kmers"]) == 7108 and int(ref["corrected"]["kept"].sum()) == 8444
    assert np.array_equal(ss["fixed"], golden["fixed"])
    assert np.array_equal(ss["sizes"].astype(np.uint8), golden["entry_sizes"])
    assert np.array_equal(ss["shared"].astype(np.uint8), golden["shared"])
    for b, ch in enumerate("ACGT"):
        assert np.array_equal(ss["prev"][b], golden[f"prev_{ch}_bits"])
        assert ref["members"][f"prev_{ch}/subaccum"] == golden[f"prev_{ch}_subaccum"].tobytes()
        assert ref["members"][f"prev_{ch}/accum"] == golden[f"prev_{ch}_accum"].tobytes()


@pytest.mark.parametrize("cfg", [
    dict(genome_len=6000, n_reads=3000, read_len=100, err=0.01, seed=1),
    dict(genome_len=4000, n_reads=2500, read_len=150, err=0.02, seed=2, n_rate=0.002),
    dict(genome_len=3000, n_reads=2500, read_len=80, err=0.01, seed=3, ragged=True),
    dict(genome_len=5000, n_reads=3000, read_len=120, err=0.005, seed=4, repeat_frac=0.3),
    dict(genome_len=2500, n_reads=2000, read_len=250, err=0.01, seed=5),
])
def test_random_pipelines(cfg):
    compare_pipeline(reads_of(**cfg))


@pytest.mark.parametrize("k,min_count,maxc,run,trim", [(16, 3, 2, 2, 0.7), (24, 5, 0, 2, 0.7), (31, 2, 8, 3, 0.5),
                                                      (29, 4, 4, 1, 0.9), (30, 1, 32, 0, 0.1), (30, 5, 1, 4, 1.0)])
def test_parameter_space(k, min_count, maxc, run, trim):
    # (--trim-after-portion 0.0 is not covered: with it the reference CHECK-fails at correct_reads.cpp:212 on the
    # first read that has no solid k-mer at all)
    reads = reads_of(3000, 2000, 100, 0.015, seed=100 + k + maxc, n_rate=0.001)
    compare_pipeline(reads, k, min_count, maxc, run, trim)


def test_trim_portion_zero_check_fails_in_the_reference():
    # --trim-after-portion 0 passes the CLI's range check, but a read without any solid k-mer then reaches
    # CHECK_GT(next_fwd_read, 0) (correct_reads.cpp:212): the reference dies there (here: the stand-in CHECK throws).
    # The CUDA path drops such a read, like every read whose corrected length is below the needed one.
    reads = reads_of(3000, 2000, 100, 0.015, seed=162, n_rate=0.001)
    with pytest.raises(RuntimeError, match="correct_reads.cpp:212"):
        R.create(reads, 30, 1, 32, 0, 0.0)


def test_reference_refuses_k32_in_correction():
    # the CLI accepts --kmer-size 16..32 (biograph_create.cpp:483) but kmer_counter's constructor refuses 32
    # (bs/kmer_counter.cpp:52-54): the product's bgx_create limit of 31 is the reference's effective limit
    reads = reads_of(2000, 500, 100, 0.0, seed=3)
    with R.Run(2) as r:
        with pytest.raises(RuntimeError, match="maximum kmer size of 31"):
            r.count_kmers(reads, 32, 2)


def test_heavy_hitter_past_255():
    # one k-mer seen far more than 255 times: the uint8 table overflows into the overflow table
    # (build_seqset/kmer_counter.cpp:629-697)
    rng = np.random.default_rng(9)
    core = "".join("ACGT"[i] for i in rng.integers(0, 4, 40))
    reads = [core] * 700 + [O.revcomp(core)] * 300 + reads_of(2000, 800, 60, 0.01, seed=10)
    ref = compare_pipeline(reads, 30, 5, 2, 2, 0.7)
    c = ref["counts"]
    assert int((c["fwd"].astype(np.int64) + c["rev"]).max()) >= 1000


def test_degenerate_inputs():
    # reads shorter than k, all-N reads, nothing solid: an empty seqset is an error in finalize for neither side
    reads = ["ACGT" * 5, "N" * 50, "".join("ACGT"[i] for i in np.random.default_rng(1).integers(0, 4, 80))]
    oc = O.count_kmers(reads, 30)
    with R.Run(2) as r:
        c, s = r.count_kmers(reads, 30, 5)
        assert len(s["kmers"]) == 0 == len(O.solid_set(oc, 5)["kmers"])
        cr = r.correct(reads, 8, 2, 0.7)
        assert int(cr["kept"].sum()) == 0


def test_fast_read_correct_analytic_cases_on_the_reference():
    # the analytic cases of modules/bio_base/fast_read_correct_test.cpp:108-258 through the reference's own
    # fast_read_correct, next to the oracle's answer
    n = 0
    for size in (30, 33, 61, 92):
        ks = F.kmer_set_of(F.LONG[:size])
        for mode in ("N", "subst"):
            for name, read, exp, ec in F.cases(size, mode):
                got = R.fast_read_correct(read, ks, F.K, F.MAXC, F.RUN)
                assert got == (exp, ec), (size, mode, name)
                assert got == O.fast_read_correct(read, ks, F.K, F.MAXC, F.RUN)
                n += 1
    assert n > 500


def test_fast_read_correct_random_reads():
    rng = np.random.default_rng(77)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, 1500))
    ks = set()
    for i in range(len(genome) - 24 + 1):
        km = genome[i:i + 24]
        ks.add(min(F.enc(km), F.enc(O.revcomp(km))))
    ks = np.array(sorted(ks), dtype=np.uint64)
    for t in range(1500):
        a = int(rng.integers(0, len(genome) - 130))
        read = list(genome[a:a + int(rng.integers(24, 130))])
        if t & 1:
            read = list(O.revcomp("".join(read)))
        for _ in range(int(rng.integers(0, 5))):
            read[int(rng.integers(0, len(read)))] = "ACGTN"[int(rng.integers(0, 5))]
        read = "".join(read)
        maxc, run = int(rng.integers(0, 6)), int(rng.integers(0, 4))
        assert R.fast_read_correct(read, ks, 24, maxc, run) == O.fast_read_correct(read, ks, 24, maxc, run), read


def test_seed_counts_do_not_change_the_seqset():
    # the expander closes the set whatever the seeding (correct_reads.cpp:195-226 is a heuristic): seeds 1/1, the
    # oracle's seed counts and full seeding give the same reference-built seqset
    reads = reads_of(3000, 1200, 90, 0.0, seed=21)
    rng = np.random.default_rng(5)
    base = R.seqset_for_reads(reads, threads=2)
    lens = np.array([len(r) for r in reads])
    for nf, nr in ((lens, lens), (rng.integers(1, 20, len(reads)), rng.integers(1, 20, len(reads)))):
        ss = R.seqset_for_reads(reads, np.minimum(nf, lens), np.minimum(nr, lens), threads=2, partition_depth=3)
        assert ss["n"] == base["n"]
        for t in ("sizes", "shared", "prev", "fixed"):
            assert np.array_equal(ss[t], base[t]), t
    o = O.seqset_closed_form(reads)
    for t in ("sizes", "shared", "prev", "fixed"):
        assert np.array_equal(o[t], base[t]), t


def adversarial_reads():
    """k-mers that are their own reverse complement, homopolymer and short-period reads, plus ordinary reads"""
    rng = np.random.default_rng(77)
    half = "".join("ACGT"[i] for i in rng.integers(0, 4, 15))
    pal30 = half + O.revcomp(half)                        # a 30-base reverse-complement palindrome
    assert O.revcomp(pal30) == pal30
    flank = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, n))  # noqa: E731
    l, r_ = flank(20), flank(20)
    reads = []
    for i in range(12):
        reads += [l + pal30 + r_, O.revcomp(l + pal30 + r_), pal30, l[5:] + pal30 + r_[:7]]
    reads += ["A" * 60] * 7 + ["T" * 45] * 6 + ["AC" * 40] * 8 + ["GT" * 33] * 5 + ["ACG" * 25] * 9 + ["AAAAAAAAAC" * 6] * 6
    return reads + reads_of(2500, 1500, 90, 0.01, seed=78)


def test_palindromes_homopolymers_and_tandem_repeats():
    """k-mers that are their own reverse complement (even k: neither strand is 'the' canonical one -- fwd / rev counts
    and the flag swap of kmer_count_table.h:54-103 show it), homopolymer and short-period reads (one k-mer counted many
    times per read, suffixes that are prefixes of each other), all of them through the whole flow"""
    reads = adversarial_reads()
    for k, mc in ((30, 5), (16, 4), (24, 3)):
        compare_pipeline(reads, k, mc, 4, 2, 0.6)

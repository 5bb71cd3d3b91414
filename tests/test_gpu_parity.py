"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs,
against the reference's golden fixture, and against the reference tests' known answers.
Bit-exact everywhere (integer / byte / index work)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from tests import frc_cases as F  # noqa: E402


@pytest.fixture(scope="module")
def B():
    import biograph_b200 as B
    if B.load_library().bgx_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu tests need a B200 (there is no CPU fallback)")
    return B


def run_gpu(B, reads, **opts):
    g = B.Bgx(**opts)
    g.add_reads(reads)
    g.count_kmers()
    km = g.export_kmers(1)
    g.correct()
    cr = g.export_corrected()
    g.build_seqset()
    ss = g.export_seqset()
    st = g.stats()
    return g, km, cr, ss, st


def check_seqset_equal(a, b):
    assert a["n"] == b["n"]
    for k in ("sizes", "shared", "prev", "fixed"):
        assert np.array_equal(a[k], b[k]), k


def check_bitcount(ss):
    for b in range(4):
        sub, acc, tot = O.bitcount_finalize(ss["prev"][b], ss["n"])
        assert np.array_equal(sub, ss["subaccum"][b])
        assert np.array_equal(acc, ss["accum"][b])
        assert tot == int(ss["fixed"][b + 1] - ss["fixed"][b])


def full_compare(B, reads, k=30, min_count=5, max_corrections=8, min_good_run=2, trim=0.7, closed_form=False, **kw):
    g, km, cr, ss, st = run_gpu(B, reads, kmer_size=k, min_kmer_count=min_count, max_corrections=max_corrections,
                                min_good_run=min_good_run, trim_after_portion=trim, **kw)
    oc = O.count_kmers(reads, k)
    for f in ("kmers", "fwd", "rev", "flags"):
        assert np.array_equal(oc[f], km[f]), f
    solid = O.solid_set(oc, min_count)
    gs = g.export_kmers(min_count)
    for f in ("kmers", "fwd", "rev", "flags"):
        assert np.array_equal(solid[f], gs[f]), f
    ocr = O.correct_reads(reads, solid, k, max_corrections, min_good_run, trim)
    assert np.array_equal(ocr["kept"], cr["kept"])
    assert np.array_equal(ocr["offs"], cr["offs"])
    assert ocr["seq"] == cr["seq"]
    assert np.array_equal(ocr["corrections"], cr["corrections"])
    assert np.array_equal(ocr["next_fwd"], cr["next_fwd"])
    assert np.array_equal(ocr["next_rev"], cr["next_rev"])
    oss = O.seqset_staged((ocr["seq"], ocr["offs"]), ocr["next_fwd"], ocr["next_rev"])
    check_seqset_equal(oss, ss)
    if closed_form:
        check_seqset_equal(O.seqset_closed_form((ocr["seq"], ocr["offs"])), ss)
    check_bitcount(ss)
    g.close()
    return ss, st


def test_golden_e_coli_10000snp(B, golden, golden_reads):
    """BASELINE config 1: golden/e_coli_10000snp.fq -> every payload member of the reference's own
    golden/e_coli_10000snp.bg/seqset, byte for byte."""
    g, km, cr, ss, st = run_gpu(B, golden_reads)
    assert len(g.export_kmers(5)["kmers"]) == 7108
    assert cr["n_kept"] == 8444 and int(cr["lens"].sum()) == 288464
    assert ss["n"] == 19935
    assert np.array_equal(ss["fixed"], golden["fixed"])
    assert np.array_equal(ss["sizes"].astype(np.uint8), golden["entry_sizes"])
    assert np.array_equal(ss["shared"].astype(np.uint8), golden["shared"])
    for b, ch in enumerate("ACGT"):
        assert np.array_equal(ss["prev"][b], golden[f"prev_{ch}_bits"])
        assert np.array_equal(ss["subaccum"][b], golden[f"prev_{ch}_subaccum"])
        assert np.array_equal(ss["accum"][b], golden[f"prev_{ch}_accum"])
    # entries are strictly sorted, prefix-free, and are exactly dedup(all suffixes) of the corrected reads
    ents = g.export_entries()
    assert ents == sorted(ents) and len(set(ents)) == len(ents)
    g.close()


def test_golden_against_oracle_stages(B, golden_reads):
    full_compare(B, golden_reads, closed_form=True)


# ---- bs/builder_test.cpp known answers (seqset_for_reads: no correction) -----------------------------
def seqset_for_reads(B, reads, **kw):
    """modules/bio_base/seqset_testutil.cpp:19-59 analogue: reads in, seqset out, correction made a
    no-op by using each read min_count times... instead we bypass correction: every k-mer solid."""
    k = 16
    g = B.Bgx(kmer_size=k, min_kmer_count=1, max_corrections=0, trim_after_portion=1.0, **kw)
    g.add_reads(reads)
    g.run()
    ss = g.export_seqset()
    ents = g.export_entries()
    g.close()
    return ss, ents


SEQ1 = ["AAAATTAC", "AAATTAC", "AATTAC", "AATTTTAG", "AC", "AG", "ATTAC", "ATTTTAG", "CTAAAATTAC", "GTAATTTTAG",
        "TAAAATTAC", "TAATTTTAG", "TAC", "TAG", "TTAC", "TTAG", "TTTAG", "TTTTAG"]


@pytest.mark.parametrize("reads,size", [
    ([O.tseq("ab")], None),
    ([O.tseq("abcdefg")], 129),
    ([O.tseq("abcd"), O.tseq("cdef"), O.tseq_rc("efgh")], 152),
    ([O.tseq("ab"), O.tseq("bc"), O.tseq("cd"), O.tseq("be")], 91),
    ([O.tseq("AB"), O.tseq("BC"), O.tseq("CD"), O.tseq("BE")], 91),
    ([O.tseq("abc"), O.tseq("cde")], 89),
    ([O.tseq("abc"), O.tseq("efg")], 99),
])
def test_builder_known_answers(B, reads, size):
    ss, ents = seqset_for_reads(B, reads)
    exp = O.entries_closed_form_py(reads)
    assert ents == exp
    if size is not None:
        assert ss["n"] == size
    check_seqset_equal(O.seqset_closed_form(reads), ss)
    check_bitcount(ss)


# ---- synthetic genomes ---------------------------------------------------------------------------------
def _sim(glen, n, L, err, seed, n_rate=0.0, paired=True):
    from biograph_b200 import synth
    genome = synth.random_genome(glen, seed=seed, repeat_frac=0.05)
    reads = synth.simulate_reads(genome, n, read_len=L, error_rate=err, seed=seed + 1, paired=paired,
                                 frag_mean=max(L + 50, 2 * L), frag_sd=20, n_rate=n_rate)
    buf, offs = synth.as_buffer(reads)
    return (buf.tobytes(), offs)


@pytest.mark.parametrize("glen,n,L,err,nrate,seed", [
    (5000, 3000, 100, 0.005, 0.0, 1),
    (20000, 8000, 150, 0.005, 0.0, 2),
    (20000, 8000, 150, 0.02, 0.002, 3),   # heavy errors + N calls: exercises the DFS and N handling
    (3000, 4000, 60, 0.01, 0.001, 4),
    (50000, 10000, 250, 0.005, 0.0, 5),   # long reads (8 words)
])
def test_synthetic_full_path(B, glen, n, L, err, nrate, seed):
    full_compare(B, _sim(glen, n, L, err, seed, n_rate=nrate), closed_form=(n * L <= 1_500_000))


@pytest.mark.parametrize("k,min_count,maxc,run,trim", [(16, 2, 2, 2, 0.5), (21, 3, 4, 1, 0.7), (31, 5, 8, 3, 0.9),
                                                        (30, 5, 0, 2, 1.0), (25, 4, 16, 2, 0.0), (30, 3, 32, 1, 0.0)])
def test_parameter_sweep(B, k, min_count, maxc, run, trim):
    full_compare(B, _sim(8000, 4000, 120, 0.01, 40 + k, n_rate=0.001), k=k, min_count=min_count, max_corrections=maxc,
                 min_good_run=run, trim=trim)


def test_ragged_and_short_reads(B):
    rng = np.random.default_rng(9)
    from biograph_b200 import synth
    genome = synth.random_genome(6000, seed=99).tobytes().decode()
    reads = []
    for _ in range(5000):
        L = int(rng.integers(1, 256))
        s = int(rng.integers(0, len(genome) - L))
        r = genome[s:s + L]
        if rng.random() < 0.5:
            r = O.revcomp(r)
        reads.append(r)
    reads += ["A", "ACGT", "N" * 40, "ACGTN" * 30, genome[:29], genome[:30], genome[:31]]
    full_compare(B, reads)


def test_repeats_force_big_tie_groups(B):
    # low-complexity sequence: thousands of suffixes share their first 24+ bases -> refinement path
    rng = np.random.default_rng(3)
    unit = "ACGTTGCA"
    base = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
    reads = []
    for i in range(600):
        reads.append(base[i % 100:i % 100 + 60] + "A" * int(rng.integers(40, 120)) + base[200:230 + i % 40])
        reads.append(base[i % 50:i % 50 + 50] + unit * int(rng.integers(5, 20)) + base[100:150])
    ss, st = full_compare(B, reads, min_count=2, closed_form=True)
    assert st.get("tie_big_records_r1", 0) + st.get("tie_big_records_r2", 0) > 0


def test_refinement_path_equals_small_group_path(B):
    reads = _sim(10000, 5000, 150, 0.005, 77)
    ss_a, _ = full_compare(B, reads)
    os.environ["BGX_SMALL_GROUP"] = "1"
    try:
        ss_b, st = full_compare(B, reads)
        assert st.get("tie_big_records_r1", 0) > 0
    finally:
        del os.environ["BGX_SMALL_GROUP"]
    check_seqset_equal(ss_a, ss_b)


@pytest.mark.parametrize("bits", [0, 16, 32, 48, 64])
def test_sort_key_bits_do_not_change_result(B, bits):
    reads = _sim(10000, 5000, 150, 0.005, 78)
    full_compare(B, reads, sort_key_bits=bits)


@pytest.mark.parametrize("batch", [40, 700, 2500, 100000])
def test_batched_counting_does_not_change_result(B, batch):
    """count_batch_reads bounds the k-mer instance buffers (inputs whose instance words do not fit
    HBM): the counting runs in a power-of-two number of HASH-RANGE batches, at least reads / that;
    every batch is partitioned, split and counted in turn.  Counts, flags and everything downstream
    must be unchanged."""
    reads = _sim(10000, 5000, 150, 0.01, 91, n_rate=0.001)
    ss, st = full_compare(B, reads, count_batch_reads=batch)
    want = max(1, -(-5000 // batch))
    assert st["count_batches"] == 1 << (want - 1).bit_length()


def test_batched_counting_with_heavy_hitter(B):
    """a homopolymer run (363 000 instances of one k-mer) overfills its hash partition in the batch that
    owns its hash: the exact-offset re-run of pass 1, and one sub-bin far longer than the others"""
    buf, offs = _sim(4000, 2000, 150, 0.005, 92)
    sim = [buf[offs[i]:offs[i + 1]].decode() for i in range(len(offs) - 1)]
    reads = []
    for i in range(5):
        reads += ["A" * 150] * 600 + sim[400 * i:400 * (i + 1)]
    ss, st = full_compare(B, reads, count_batch_reads=1000)
    assert st["count_batches"] == 8
    assert st.get("count_partition_reruns", 0) >= 1


def test_overfull_count_bins_are_split_finer(B):
    """a sub-bin with more distinct k-mers than its shared-memory table has slots: the split and
    count passes are re-run with finer sub-bins (then larger tables) until every bin fits"""
    reads = _sim(20000, 6000, 150, 0.01, 93)
    os.environ["BGX_BIN_SLOTS_LOG2"] = "9"
    os.environ["BGX_SUB_BITS"] = "0"
    try:
        ss, st = full_compare(B, reads)
        assert st.get("count_bin_reruns", 0) >= 1
    finally:
        del os.environ["BGX_BIN_SLOTS_LOG2"], os.environ["BGX_SUB_BITS"]


def test_k31_read_of_all_t(B):
    """k = 31: a 31-base read of T is one k-mer that is both first and last of its read and has every
    k-mer bit set (no value of the instance word may serve as a 'no item' sentinel)"""
    reads = ["T" * 31] * 7 + ["A" * 31] * 3 + ["ACGT" * 10][:1] + ["T" * 40] * 2
    g, km, cr, ss, st = run_gpu(B, reads, kmer_size=31, min_kmer_count=5)
    oc = O.count_kmers(reads, 31)
    for f in ("kmers", "fwd", "rev", "flags"):
        assert np.array_equal(oc[f], km[f]), f
    assert int(km["fwd"].sum() + km["rev"].sum()) == 7 + 3 + 10 + 2 * 10
    g.close()


def test_packed_input_equals_ascii_input(B):
    from biograph_b200 import bgx
    reads = _sim(8000, 4000, 150, 0.01, 31, n_rate=0.002)
    g1, km1, cr1, ss1, _ = run_gpu(B, reads)
    packed, nmask, woffs, lens = bgx.pack_reads_2bit(reads)
    assert nmask is not None
    g2 = B.Bgx()
    g2.add_reads_packed(packed, nmask, woffs, lens)
    g2.run()
    ss2 = g2.export_seqset()
    check_seqset_equal(ss1, ss2)
    assert g2.export_corrected()["seq"] == cr1["seq"]
    g1.close()
    g2.close()


def test_correct_reads_test_known_answers_gpu(B):
    """The reference's own correct_reads_test.cpp:121-223 (tests/crt_cases.py): which suffixes get
    seeded.  The k-mer set of each case is produced the way the counter would: every add_kmers
    sequence goes in as a read (min_count 1), which marks its first k-mer fwd_starts_read and its
    last one rev_starts_read -- exactly the flags start_correction builds."""
    from tests import crt_cases as T
    for name, kmer_seqs, reads, expected, exact in T.CASES:
        all_reads = list(reads) + [s_ for s_ in kmer_seqs if s_ not in reads]
        g = B.Bgx(kmer_size=T.K, min_kmer_count=1)
        g.add_reads(all_reads)
        g.count_kmers()
        g.correct()
        cr = g.export_corrected()
        seqs = O.corrected_list(cr)
        got = set()
        for i, r in enumerate(reads):
            assert cr["kept"][i] and seqs[i] == r, name
            got |= T.seeds_of(r, int(cr["next_fwd"][i]), int(cr["next_rev"][i]))
        assert (got == expected) if exact else (expected <= got), name
        g.close()


@pytest.mark.parametrize("which", ["golden", "synthetic_with_n"])
def test_lookup_reads_matches_bisect(B, golden_reads, which):
    """bgx_lookup_reads (the entry lookups of make_readmap, make_readmap.cpp:137-167): the id of the first
    entry having the corrected read / its reverse complement as a prefix, checked against a
    bisect over the (parity-checked) sorted entry list."""
    import bisect
    reads = golden_reads if which == "golden" else _sim(6000, 3000, 120, 0.01, 55, n_rate=0.002)
    g, km, cr, ss, st = run_gpu(B, reads)
    ents = g.export_entries()
    fwd, rc = g.lookup_reads()
    kept = cr["kept"]
    assert len(fwd) == len(kept) == len(rc)
    seqs = O.corrected_list(cr)
    j = 0
    n_checked = 0
    for r in range(len(kept)):
        if not kept[r]:
            assert fwd[r] == 2**64 - 1 and rc[r] == 2**64 - 1
            continue
        s_ = seqs[j]
        j += 1
        for seq, got in ((s_, int(fwd[r])), (O.revcomp(s_), int(rc[r]))):
            want = bisect.bisect_left(ents, seq)
            assert want < len(ents) and ents[want].startswith(seq)
            assert got == want
            n_checked += 1
    assert n_checked == 2 * int(kept.sum()) > 0
    g.close()


@pytest.mark.parametrize("which", ["golden", "synthetic_dups"])
def test_readmap_unpaired(B, golden_reads, which):
    """bgx_build_readmap (unpaired) against the CPU restatement (oracle/readmap.py, pinned to the
    reference's golden readmap) and, for the golden reads, against the golden members themselves."""
    from oracle import readmap as RM
    if which == "golden":
        reads = golden_reads
    else:  # many identical reads: long runs of identical rows exercise the claim order
        buf, offs = _sim(3000, 6000, 100, 0.003, 56)
        sim = [buf[offs[i]:offs[i + 1]].decode() for i in range(len(offs) - 1)]
        reads = sim + sim[:500] * 3 + [O.revcomp(r) for r in sim[:300]]
    g, km, cr, ss, st = run_gpu(B, reads)
    fwd, rc = g.lookup_reads()
    got = g.build_readmap()
    kept = cr["kept"].astype(bool)
    want = RM.readmap_tables(fwd[kept], rc[kept], cr["lens"][kept], ss["n"])
    assert got["n_rows"] == want["n_rows"] == 2 * int(kept.sum())
    assert np.array_equal(got["read_lengths"], want["read_lengths"])
    assert np.array_equal(got["mate_loop_ptr"], want["mate_loop_ptr"])
    assert np.array_equal(got["is_forward"], RM.pack_bits(want["is_forward"]))
    for name, nbits in (("source_to_mid", ss["n"]), ("dest_to_mid", want["n_rows"])):
        words = RM.pack_bits(want[name])
        assert np.array_equal(got[name]["bits"], words), name
        sub, acc, _ = O.bitcount_finalize(words, nbits)
        assert np.array_equal(got[name]["subaccum"], sub), name
        assert np.array_equal(got[name]["accum"], acc), name
    if which == "golden":
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "e_coli_10000snp_readmap.npz"))
        assert np.array_equal(got["read_lengths"].astype(np.uint8), z["read_lengths"])
        assert np.array_equal(got["mate_loop_ptr"].astype("<u4"), z["mate_loop_ptr|packed_data"].view("<u4"))
        assert np.array_equal(got["is_forward"], z["is_forward|packed_data"].view("<u8"))
        for name in ("source_to_mid", "dest_to_mid"):
            for part in ("bits", "subaccum", "accum"):
                assert np.array_equal(got[name][part], z[f"read_ids|{name}|{part}"].view("<u8")), (name, part)
    g.close()


def test_readmap_paired(B):
    """bgx_build_readmap(paired): reads 2i, 2i+1 are mates.  Against oracle/readmap.py's paired form (which
    tests/test_oracle_readmap.py checks against a literal transcription of the reference's claim pass):
    identical reads with different mates, duplicate pairs, pairs with a dropped read."""
    from oracle import readmap as RM
    buf, offs = _sim(4000, 3000, 100, 0.004, 57)   # paired simulation: reads 2i, 2i+1 are the two ends of a fragment
    sim = [buf[offs[i]:offs[i + 1]].decode() for i in range(len(offs) - 1)]
    reads = list(sim)
    rng = np.random.default_rng(58)
    for _ in range(400):   # same first read, another pair's mate
        i, j = int(rng.integers(0, len(sim) // 2)), int(rng.integers(0, len(sim) // 2))
        reads += [sim[2 * i], sim[2 * j + 1]]
    for _ in range(400):   # exact duplicate pairs, and pairs given the other way round
        i = int(rng.integers(0, len(sim) // 2))
        reads += [sim[2 * i], sim[2 * i + 1], sim[2 * i + 1], sim[2 * i]]
    reads += ["ACGT" * 5, sim[0], sim[1], "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"]  # pairs with a read the corrector drops
    assert len(reads) % 2 == 0
    g, km, cr, ss, st = run_gpu(B, reads)
    kept = cr["kept"].astype(bool)
    assert (~kept).any() and kept.any()
    fwd, rc = g.lookup_reads()
    cols = RM.pair_records(fwd, rc, cr["lens"], kept)
    want = RM.readmap_tables_paired(*cols, ss["n"])
    got = g.build_readmap(paired=True)
    assert st is not None and got["n_rows"] == want["n_rows"] > 0
    assert int((cols[5] > 0).sum()) > 1000   # real pairs present
    assert np.array_equal(got["read_lengths"], want["read_lengths"])
    assert np.array_equal(got["is_forward"], RM.pack_bits(want["is_forward"]))
    assert np.array_equal(got["mate_loop_ptr"], want["mate_loop_ptr"])
    for name, nbits in (("source_to_mid", ss["n"]), ("dest_to_mid", want["n_rows"])):
        words = RM.pack_bits(want[name])
        assert np.array_equal(got[name]["bits"], words), name
        sub, acc, _ = O.bitcount_finalize(words, nbits)
        assert np.array_equal(got[name]["subaccum"], sub) and np.array_equal(got[name]["accum"], acc), name
    # unpaired build of the same reads still matches the unpaired restatement
    got_u = g.build_readmap()
    want_u = RM.readmap_tables(fwd[kept], rc[kept], cr["lens"][kept], ss["n"])
    assert np.array_equal(got_u["mate_loop_ptr"], want_u["mate_loop_ptr"])
    g.close()


def _fastq(reads, blank_tail=0):
    out = []
    for i, r in enumerate(reads):
        out.append(f"@read{i}/1\n{r}\n+\n{'I' * len(r)}\n")
    return ("".join(out) + "\n" * blank_tail).encode()


def test_readmap_read_props_of_the_reference_test(B):
    """modules/bio_base/readmap_test.cpp:53-168 (TEST(readmap, read_props_tests)) against the GPU readmap:
    the reference's own pairs, and for every read and its reverse complement the properties the test
    asserts -- unique entry, entry <-> read index, length, is_forward, get_rev_comp, and mates that point
    at each other (get_mate / has_mate as readmap.cpp:207-268 derive them from the mate loop)."""
    import bisect
    T = O.tseq
    pairs = [(T("READ1"), T("ANOTHER1")), (T("NEWB"), T("BROTHER")), (T("SOLO"), ""), (T("PREFIXread"), T("PREFIXmate")),
             (T("readSUFFIX"), T("mateSUFFIX")), (T("PREFIXreadSUFFIX"), T("PREFIXmateSUFFIX")), (T("read"), T("mate")),
             (T("XreadS"), T("XmateS"))]
    reads = [r for p in pairs for r in p]
    g = B.Bgx()
    g.add_reads(reads)
    g.seed_uncorrected()
    g.build_seqset()
    ents = g.export_entries()
    rm = g.build_readmap(paired=True)
    n_rows = rm["n_rows"]
    bit = lambda words, i: (int(words[i >> 6]) >> (i & 63)) & 1
    src = [i for i in range(len(ents)) if bit(rm["source_to_mid"]["bits"], i)]          # entries that hold reads
    opens = [i for i in range(n_rows) if bit(rm["dest_to_mid"]["bits"], i)] + [n_rows]  # first row of each of them
    assert len(src) == len(opens) - 1
    loop = rm["mate_loop_ptr"]
    fwd_of = lambda i: bool(bit(rm["is_forward"], i))

    def entry_to_index(e):
        k = bisect.bisect_left(src, e)
        assert k < len(src) and src[k] == e          # get_bit(entry)
        return opens[k], opens[k + 1]

    def index_to_entry(i):
        return src[bisect.bisect_right(opens, i) - 1]

    def find_unique(seq):
        lo = bisect.bisect_left(ents, seq)
        hits = [e for e in range(lo, min(lo + 3, len(ents))) if ents[e].startswith(seq)]
        assert len(hits) == 1                        # entry_read.end() - entry_read.begin() == 1
        return hits[0]

    def row_of(seq):
        e = find_unique(seq)
        a, b = entry_to_index(e)
        rows = [i for i in range(a, b) if rm["read_lengths"][i] == len(seq)]
        assert rows
        return e, rows[-1] if b - a != 1 else a       # the test keeps the last row of that length

    def rev_comp(i):                                  # readmap::get_rev_comp
        for _ in range(1 if fwd_of(i) else 3):
            i = int(loop[i])
        return i

    def props(read, mate, fwd):
        e, i = row_of(read)
        assert fwd_of(i) == fwd and rm["read_lengths"][i] == len(read)
        assert index_to_entry(i) == e and ents[e][:len(read)] == read
        rc = rev_comp(i)
        assert ents[index_to_entry(rc)][:rm["read_lengths"][rc]] == O.revcomp(read)
        if mate:
            me, mi = row_of(mate)
            assert fwd_of(mi) == fwd and rm["read_lengths"][mi] == len(mate)
            has_mate = lambda x: int(loop[int(loop[x])]) != x
            assert has_mate(i) and has_mate(mi)
            assert int(loop[int(loop[i])]) == mi and int(loop[int(loop[mi])]) == i   # get_mate both ways
            assert index_to_entry(mi) == me and ents[me][:len(mate)] == mate
        else:
            assert int(loop[int(loop[i])]) == i       # no mate: the loop is read <-> reverse complement

    for read, mate in pairs:
        props(read, mate, True)
        props(O.revcomp(read), O.revcomp(mate) if mate else "", False)
    g.close()



def test_fastq_import_equals_ascii_import(B, golden_reads):
    """bgx_add_reads_fastq (fastq_reader::read semantics, modules/bio_format/fastq.cpp:40-126): same reads,
    same result as bgx_add_reads_ascii; ragged lengths, N calls, several appends, trailing blank lines."""
    buf, offs = _sim(6000, 3000, 120, 0.01, 61, n_rate=0.003)
    sim = [buf[offs[i]:offs[i + 1]].decode() for i in range(len(offs) - 1)]
    reads = sim + [r[:40 + i % 60] for i, r in enumerate(sim[:500])] + ["N" * 50, "ACGTN" * 7]
    g1, km1, cr1, ss1, _ = run_gpu(B, reads)
    g2 = B.Bgx()
    assert g2.add_reads_fastq(_fastq(reads[:1000])) == 1000
    assert g2.add_reads_fastq(_fastq(reads[1000:], blank_tail=3)) == len(reads) - 1000
    assert g2.add_reads_fastq(b"\n\n") == 0
    g2.run()
    check_seqset_equal(ss1, g2.export_seqset())
    cr2 = g2.export_corrected()
    assert cr2["seq"] == cr1["seq"] and np.array_equal(cr2["offs"], cr1["offs"])
    g3 = B.Bgx()
    assert g3.add_reads_fastq(_fastq(golden_reads)) == 10000
    g3.run()
    assert g3.export_seqset()["n"] == 19935
    for g in (g1, g2, g3):
        g.close()


@pytest.mark.parametrize("text,msg", [
    (b"@r\nACGT\n+\nIIII", "Partial line in fastq file"),
    (b"@r\nACGT\n+\n", "End of file while reading quality line"),
    (b"@r\nACGT\n", "End of file while reading + line"),
    (b"r1\nACGT\n+\nIIII\n", "line 1: Sequence id missing @"),
    (b"@\nACGT\n+\nIIII\n", "line 1: Sequence id too short"),
    (b"@r\nACGT\n+\nIIII\n@s\nACXT\n+\nIIII\n", "line 6: Sequence contains unexpected characters"),
    (b"@r\nacgt\n+\nIIII\n", "line 2: Sequence contains unexpected characters"),
    (b"@r\nACGT\n-\nIIII\n", "line 3: Expecting + as first char of line"),
    (b"@r\nACGT\n+\nIII\n", "line 4: Quality line not same length as sequence"),
    (b"@r\n\n+\nII\n", "line 2: Expecting sequence, found empty line"),
    (b"@r\nACGT\n+\nIIII\n\n@s\nACGT\n+\nIIII\n", "blank line between records"),
])
def test_fastq_errors_follow_the_reference(B, text, msg):
    g = B.Bgx()
    with pytest.raises(B.BgxError) as e:
        g.add_reads_fastq(text)
    assert msg in str(e.value)
    g.close()


def test_async_upload_equals_sync_upload(B):
    """bgx_add_reads_packed_async: chunked copy on a second stream, pass 1 of counting launched per
    chunk; also appended after a synchronous batch, and followed by stages other than count."""
    from biograph_b200 import bgx
    reads = _sim(60000, 300000, 150, 0.005, 33)   # >= 2^18 reads: the chunked path is taken
    packed, nmask, woffs, lens = bgx.pack_reads_2bit(reads)
    assert nmask is None
    g1 = B.Bgx()
    g1.add_reads_packed(packed, None, woffs, lens)
    g1.run()
    ss1 = g1.export_seqset()
    km1 = g1.export_kmers(1)
    g2 = B.Bgx()
    g2.add_reads_packed_ptr(packed.ctypes.data, None, woffs.ctypes.data, lens.ctypes.data, len(lens), overlap=True)
    g2.run()
    check_seqset_equal(ss1, g2.export_seqset())
    km2 = g2.export_kmers(1)
    for f in ("kmers", "fwd", "rev", "flags"):
        assert np.array_equal(km1[f], km2[f]), f
    # clearing right after an async append (copies still in flight) must be safe
    g4 = B.Bgx()
    g4.add_reads_packed_ptr(packed.ctypes.data, None, woffs.ctypes.data, lens.ctypes.data, len(lens), overlap=True)
    g4.clear_reads()
    g4.add_reads_packed_ptr(packed.ctypes.data, None, woffs.ctypes.data, lens.ctypes.data, len(lens), overlap=True)
    g4.run()
    check_seqset_equal(ss1, g4.export_seqset())
    g4.close()
    # batched counting and a second (async) append on top of resident reads
    g3 = B.Bgx(count_batch_reads=100000)
    g3.add_reads_packed_ptr(packed.ctypes.data, None, woffs.ctypes.data, lens.ctypes.data, len(lens), overlap=True)
    g3.run()
    check_seqset_equal(ss1, g3.export_seqset())
    for g in (g1, g2, g3):
        g.close()


def test_incremental_add_reads(B):
    reads = _sim(8000, 4000, 150, 0.01, 32)
    buf, offs = reads
    g1, _, _, ss1, _ = run_gpu(B, reads)
    g2 = B.Bgx()
    cut = 1500
    g2.add_reads((buf[:offs[cut]], offs[:cut + 1]))
    g2.add_reads((buf[offs[cut]:], offs[cut:] - offs[cut]))
    g2.run()
    check_seqset_equal(ss1, g2.export_seqset())
    g1.close()
    g2.close()


def test_nothing_survives(B):
    # all k-mers below min count: every read is dropped, the seqset is empty
    reads = _sim(200000, 300, 100, 0.0, 33)
    g = B.Bgx()
    g.add_reads(reads)
    g.run()
    ss = g.export_seqset()
    assert ss["n"] == 0 and int(ss["fixed"][4]) == 0
    g.close()


def test_errors_are_reported(B):
    g = B.Bgx()
    with pytest.raises(B.BgxError):
        g.count_kmers()  # no reads
    with pytest.raises(B.BgxError):
        g.add_reads(["A" * 300])  # longer than 255
    g.add_reads(["ACGT" * 20])
    with pytest.raises(B.BgxError):
        g.correct()  # before count
    with pytest.raises(B.BgxError):
        g.build_seqset()
    g.close()
    with pytest.raises(B.BgxError):
        B.Bgx(kmer_size=32)


# ---- fast_read_correct analytic cases through the CUDA corrector ----------------------------------------------
@pytest.mark.parametrize("mode", ["N", "X"])
@pytest.mark.parametrize("size", [30, 33, 61, 92, 122, 255])
def test_frc_analytic_gpu(B, size, mode):
    """modules/bio_base/fast_read_correct_test.cpp:108-258.  The k-mer set is injected by adding the
    error-free sequence min_count times; every erroneous k-mer occurs far fewer times."""
    seq = F.LONG[:size]
    cs = list(F.cases(size, mode, with_three=(size <= 33)))
    copies = 2000
    reads = [seq] * copies + [c[1] for c in cs]
    g = B.Bgx(kmer_size=F.K, min_kmer_count=copies, max_corrections=F.MAXC, min_good_run=F.RUN, trim_after_portion=0.0)
    g.add_reads(reads)
    g.count_kmers()
    solid = g.export_kmers(copies)
    assert np.array_equal(solid["kmers"], F.kmer_set_of(seq))
    g.correct()
    cr = g.export_corrected()
    s, o = cr["seq"].decode(), cr["offs"]
    for i, (name, read, exp, ec) in enumerate(cs):
        r = copies + i
        got = s[o[r]:o[r + 1]]
        assert got == exp, (size, mode, name)
        if exp:
            assert cr["corrections"][r] == ec, (size, mode, name)
    g.close()


# ---- size-independent properties at a larger size ---------------------------------------------------------------
def test_properties_large(B):
    reads = _sim(400000, 300000, 150, 0.005, 55)
    g = B.Bgx()
    g.add_reads(reads)
    g.run()
    ss = g.export_seqset()
    n = ss["n"]
    assert n > 0 and int(ss["fixed"][4]) == n and int(ss["fixed"][0]) == 0
    assert np.all(ss["shared"][1:] < ss["sizes"][1:]) and ss["shared"][0] == 0   # prefix-free
    assert np.all(ss["shared"][1:] <= ss["sizes"][:-1])
    pc = [int(np.unpackbits(ss["prev"][b].view(np.uint8)).sum()) for b in range(4)]
    assert sum(pc) == n and [int(ss["fixed"][b + 1] - ss["fixed"][b]) for b in range(4)] == pc
    check_bitcount(ss)
    ents = g.export_entries(0, 20000) + g.export_entries(n - 20000, 20000)
    assert all(a < b and not b.startswith(a) for a, b in zip(ents[:19999], ents[1:20000]))
    # idempotence: re-running on the resident reads gives the same tables
    g.reset_results()
    g.run()
    check_seqset_equal(ss, g.export_seqset())
    # and the staged oracle agrees at this size
    cr = g.export_corrected()
    oss = O.seqset_staged((cr["seq"], cr["offs"]), cr["next_fwd"], cr["next_rev"])
    check_seqset_equal(oss, ss)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,begin,end,kind", [
    (1, 0, 64, "rand"), (2, 0, 64, "rand"), (4095, 0, 64, "rand"), (4096, 0, 64, "rand"), (4097, 16, 64, "rand"),
    (100001, 0, 64, "rand"), (1 << 20, 16, 64, "few"), (3000003, 0, 32, "sorted"), (777777, 8, 40, "const"),
])
def test_radix_sort_pairs_is_a_stable_sort(B, n, begin, end, kind):
    """the onesweep LSD radix sort (prims.cu) against numpy's stable argsort, incl. ragged last tiles,
    skewed digits (look-back with empty digit bins) and sub-ranges of key bits"""
    rng = np.random.default_rng(n + begin)
    if kind == "rand":
        keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    elif kind == "few":
        keys = rng.choice(rng.integers(0, 1 << 62, size=37, dtype=np.uint64), size=n)
    elif kind == "sorted":
        keys = np.sort(rng.integers(0, 1 << 40, size=n, dtype=np.uint64))
    else:
        keys = np.full(n, 0x0123456789ABCDEF, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint64)
    mask = np.uint64(((1 << (end - begin)) - 1) << begin) if end - begin < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    k2, v2 = keys.copy(), vals.copy()
    with B.Bgx() as g:
        g.debug_sort_pairs(k2, v2, begin, end)
    assert np.array_equal(v2, vals[order])
    assert np.array_equal(k2, keys[order])

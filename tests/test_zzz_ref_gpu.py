"""The CUDA path (through the C ABI) against the REFERENCE'S OWN code on the same inputs: oracle/_ref/libref.so is
the reference's kmer_counter / kmer_set / correct_reads / expander / builder / seqset compiled from its sources
(oracle/ref_shim.cpp).  Bit-exact: solid k-mers with their counts and flags, the surviving corrected reads, every
seqset table and the encoded payload members.  Sorts last (the library is a built artefact that travels to the GPU
box; where it is absent these tests skip and the oracle comparisons of test_gpu_parity.py stand alone)."""
import numpy as np
import pytest

from oracle import ref as R
from tests.test_ref_vs_oracle import reads_of

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref.so not built")]


@pytest.fixture(scope="module")
def B():
    import biograph_b200 as B
    if B.load_library().bgx_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu tests need a B200 (there is no CPU fallback)")
    return B


def compare(B, reads, k=30, min_count=5, max_corrections=8, min_good_run=2, trim=0.7):
    ref = R.create(reads, k, min_count, max_corrections, min_good_run, trim, threads=8, with_members=True)
    g = B.Bgx(kmer_size=k, min_kmer_count=min_count, max_corrections=max_corrections, min_good_run=min_good_run,
              trim_after_portion=trim)
    try:
        g.add_reads(reads)
        g.run()
        gs = g.export_kmers(min_count)
        cr = g.export_corrected()
        ss = g.export_seqset()
        vb = [g.export_varbit(0), g.export_varbit(1)]
    finally:
        g.close()
    c = ref["counts"]
    m = (c["fwd"].astype(np.int64) + c["rev"]) >= min_count
    assert np.array_equal(ref["solid"]["kmers"], gs["kmers"])
    for f in ("kmers", "fwd", "rev", "flags"):
        assert np.array_equal(c[f][m], gs[f]), f
    rcr = ref["corrected"]
    assert np.array_equal(rcr["kept"], cr["kept"]) and np.array_equal(rcr["offs"], cr["offs"]) and rcr["seq"] == cr["seq"]
    rss = ref["seqset"]
    assert rss["n"] == ss["n"]
    for t in ("sizes", "shared", "prev", "fixed"):
        assert np.array_equal(rss[t], ss[t]), t
    # the payload members as the reference's encoders wrote them
    mem = ref["members"]
    if ss["n"]:
        assert np.array_equal(np.frombuffer(mem["entry_sizes/elements"], dtype=np.uint64), vb[0]["elements"])
        assert np.array_equal(np.frombuffer(mem["shared/elements"], dtype=np.uint64), vb[1]["elements"])
        for b, ch in enumerate("ACGT"):
            assert mem[f"prev_{ch}/subaccum"] == np.ascontiguousarray(ss["subaccum"][b]).tobytes()
            assert mem[f"prev_{ch}/accum"] == np.ascontiguousarray(ss["accum"][b]).tobytes()
    return ss


def test_golden_reads_gpu_vs_reference(B, golden_reads):
    assert compare(B, golden_reads)["n"] == 19935


@pytest.mark.parametrize("cfg", [
    dict(genome_len=20000, n_reads=8000, read_len=100, err=0.01, seed=31),
    dict(genome_len=8000, n_reads=5000, read_len=150, err=0.02, seed=32, n_rate=0.002),
    dict(genome_len=6000, n_reads=5000, read_len=80, err=0.01, seed=33, ragged=True),
    dict(genome_len=10000, n_reads=6000, read_len=120, err=0.005, seed=34, repeat_frac=0.3),
])
def test_random_reads_gpu_vs_reference(B, cfg):
    compare(B, reads_of(**cfg))


@pytest.mark.parametrize("k,min_count,maxc,run,trim", [(16, 3, 2, 2, 0.7), (24, 5, 0, 2, 0.7), (31, 2, 8, 3, 0.5)])
def test_parameters_gpu_vs_reference(B, k, min_count, maxc, run, trim):
    compare(B, reads_of(5000, 4000, 100, 0.015, seed=200 + k, n_rate=0.001), k, min_count, maxc, run, trim)


@pytest.mark.parametrize("k,min_count", [(30, 5), (16, 4)])
def test_adversarial_reads_gpu_vs_reference(B, k, min_count):
    # reverse-complement palindromes as k-mers, homopolymers, short-period repeats
    from tests.test_ref_vs_oracle import adversarial_reads
    compare(B, adversarial_reads(), k, min_count, 4, 2, 0.6)


def test_larger_sample_gpu_vs_reference(B):
    # 120 k reads of a 600 kb genome at 30x: 1.2 M entries
    compare(B, reads_of(600000, 120000, 150, 0.005, seed=41))


def readmap_vs_reference(B, reads, paired):
    """bgx_build_readmap against the reference's own make_readmap (modules/bio_mapred/make_readmap.cpp compiled into
    oracle/_ref) over the reference's own seqset of the same reads: every payload member of the readmap file"""
    from tests.test_ref_readmap import check, varbit_pack64
    from oracle import oracle as O
    g = B.Bgx()
    try:
        g.add_reads(reads)
        g.run()
        cr = g.export_corrected()
        ss = g.export_seqset()
        got = g.build_readmap(paired=paired)
    finally:
        g.close()
    kept = cr["kept"].astype(bool)
    seqs = [cr["seq"][cr["offs"][i]:cr["offs"][i + 1]].decode() for i in range(len(kept))]
    rec, ro = [], [0]
    step = 2 if paired else 1
    for i in range(0, len(seqs), step):   # one record per pair (or read); a dropped mate leaves a one-read record
        c = [seqs[j] for j in range(i, min(i + step, len(seqs))) if kept[j]]
        if c:
            rec += c
            ro.append(len(rec))
    with R.Run(8) as r:
        r.count_kmers(reads, 30, 5)
        rcr = r.correct(reads)
        assert rcr["seq"] == cr["seq"] and np.array_equal(rcr["kept"], cr["kept"])
        rss = r.make_seqset()
        assert rss["n"] == ss["n"]
        mem = r.make_readmap(rec, ro, paired)
    exp = {"read_lengths/elements": O.varbit_pack(got["read_lengths"], int(ss["sizes"].max()))[0],
           "mate_loop_ptr/elements": varbit_pack64(got["mate_loop_ptr"], got["n_rows"]),
           "is_forward/packed_data": got["is_forward"]}
    for name in ("source_to_mid", "dest_to_mid"):
        for part in ("bits", "subaccum", "accum"):
            exp[f"read_ids/{name}/{part}"] = got[name][part]
    check(mem, exp)
    return got["n_rows"]


def _paired_reads(seed):
    from biograph_b200 import synth
    genome = synth.random_genome(5000, seed=seed)
    r2 = synth.simulate_reads(genome, 4000, read_len=100, error_rate=0.004, seed=seed + 1, paired=True, frag_mean=250, frag_sd=20)
    sim = [bytes(row).decode() for row in r2]
    rng = np.random.default_rng(seed + 2)
    reads = list(sim)
    for _ in range(300):   # same first read with another pair's mate; exact duplicate pairs
        i, j = int(rng.integers(0, len(sim) // 2)), int(rng.integers(0, len(sim) // 2))
        reads += [sim[2 * i], sim[2 * j + 1], sim[2 * i], sim[2 * i + 1]]
    reads += ["ACGT" * 5, sim[0], sim[1], "T" * 35]   # pairs with a read the corrector drops
    return reads


@pytest.mark.parametrize("paired", [False, True])
def test_readmap_gpu_vs_reference(B, paired):
    assert readmap_vs_reference(B, _paired_reads(91), paired) > 5000


def test_merge_gpu_vs_reference(B):
    """bgx_merge_seqsets against the reference's own merge classes (seqset_flat, make_mergemap, seqset_mergemap,
    seqset_merger of oracle/_ref): three seqsets the reference's builder made, merged by both -- every table of the
    merged seqset (prev bits under the reference's chunk rule, > 100 000 merged entries) and the mergemap bits"""
    from tests.test_ref_merge import read_sets
    sets = read_sets(41, 3, 12000, 90000, 60, err=0.05)
    runs = [R.Run(8) for _ in sets]
    out = R.Run(8)
    try:
        parts = []
        for r, rd in zip(runs, sets):
            r.seed(rd)
            t = r.make_seqset()
            parts.append({"sizes": t["sizes"], "prev": t["prev"]})
        want, maps = out.merge_from(runs)
        assert want["n"] > 100000
        with B.Bgx() as g:
            g.merge_seqsets(parts)       # parallel_splits = 0: the reference's g_parallel_splits
            ss = g.export_seqset()
            assert ss["n"] == want["n"]
            for t in ("sizes", "shared", "prev", "fixed"):
                assert np.array_equal(ss[t], want[t]), t
            for p in range(len(sets)):
                mm = g.export_mergemap(p)
                assert np.array_equal(mm["bits"], maps[p]), f"mergemap {p}"
                assert mm["n_set"] == len(parts[p]["sizes"])
    finally:
        for r in runs + [out]:
            r.close()

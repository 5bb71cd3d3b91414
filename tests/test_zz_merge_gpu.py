"""bgx_merge_seqsets through the C ABI on the B200: flatten + merge + tables + mergemaps +
readmap migration against the reference's own merge output and the merge oracle.  (Named to sort after
the path's own parity tests: `pytest -x` reaches the build path first.)

  * golden: family_lambda.bg = `biograph merge` of proband + father + mother (tests/golden/
    ref_seqsets.npz, ref_merge_readmaps.npz): EVERY payload member of the merged seqset byte for byte
    (fixed, varbit elements of entry_sizes / shared, bits / subaccum / accum of the four prev bitcounts)
    and the source_to_mid members of the three migrated readmaps.
  * the reference's merge tests (seqset_merger_test.cpp, make_mergemap_test.cpp) and random inputs
    against oracle/merge.py; parallel_splits = 1 against a GPU build over the union of the reads."""
import json
import os
import random

import numpy as np
import pytest

from oracle import merge as M
from oracle import oracle as O
from oracle.readmap import pack_bits
from tests import refseqset as RS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def B():
    import biograph_b200 as B
    return B


def part_tables(entries, nsplits=1):
    tb = M.merge_tables(entries, nsplits)
    return {"n": tb["n"], "sizes": tb["sizes"], "prev": np.stack([pack_bits(tb["prev"][b]) for b in range(4)])}


def seqset_entries(reads):
    return [e.encode() for e in O.entries_closed_form_py(reads)]


def check_against_oracle(B, part_entries, nsplits, rng):
    parts = [part_tables(e, rng.choice([1, 3, 100000])) for e in part_entries]
    old01 = [np.array([rng.random() < 0.5 for _ in e], dtype=np.uint8) for e in part_entries]
    merged, bits = M.make_mergemap(part_entries)
    tb = M.merge_tables(merged, nsplits)
    with B.Bgx() as g:
        g.merge_seqsets(parts, parallel_splits=nsplits)
        ss = g.export_seqset()
        assert ss["n"] == tb["n"]
        assert np.array_equal(ss["sizes"], tb["sizes"]) and np.array_equal(ss["shared"], tb["shared"])
        assert np.array_equal(ss["fixed"], tb["fixed"])
        for b in range(4):
            want = pack_bits(tb["prev"][b])
            sub, acc, _ = O.bitcount_finalize(want, tb["n"])
            assert np.array_equal(ss["prev"][b], want), b
            assert np.array_equal(ss["subaccum"][b], sub) and np.array_equal(ss["accum"][b], acc)
        assert [e.encode() for e in g.export_entries(0, ss["n"])] == merged
        for p, ent in enumerate(part_entries):
            assert [e.encode() for e in g.export_flat(p, 0, len(ent))] == ent          # seqset_flat::get
            mm = g.export_mergemap(p)
            assert mm["n_set"] == len(ent) and mm["nbits"] == tb["n"]
            assert np.array_equal(mm["bits"], pack_bits(bits[p]))
            sub, acc, _ = O.bitcount_finalize(pack_bits(bits[p]), tb["n"])
            assert np.array_equal(mm["subaccum"], sub) and np.array_equal(mm["accum"], acc)
            mg = g.migrate_bits(p, pack_bits(old01[p]), len(ent))
            assert np.array_equal(mg["bits"], pack_bits(M.migrate_source_bits(old01[p], bits[p])))


def test_golden_family_lambda_every_member(B):
    names = ["proband_lambda", "father_lambda", "mother_lambda"]   # --in order of the golden merge (qc/merge_log.txt)
    parts = [RS.tables(nm) for nm in names]
    name = "family_lambda"
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_merge_readmaps.npz"))
    with B.Bgx() as g:
        g.merge_seqsets(parts)   # parallel_splits = 0: the reference's g_parallel_splits
        ss = g.export_seqset()
        vs, vh = g.export_varbit(0), g.export_varbit(1)
        n = ss["n"]
        assert n == json.loads(RS.member(name, "seqset.json"))["num_entries"] == 103996
        assert ss["fixed"].astype("<u8").tobytes() == RS.member(name, "fixed")
        assert {"bits_per_value": vs["bits_per_value"], "element_count": n, "max_value": vs["max_value"]} == json.loads(
            RS.member(name, "entry_sizes/packed_varbit_vector.json"))
        assert {"bits_per_value": vh["bits_per_value"], "element_count": n, "max_value": vh["max_value"]} == json.loads(
            RS.member(name, "shared/packed_varbit_vector.json"))
        assert vs["elements"].astype("<u8").tobytes() == RS.member(name, "entry_sizes/elements")
        assert vh["elements"].astype("<u8").tobytes() == RS.member(name, "shared/elements")
        for b, ch in enumerate("ACGT"):
            assert ss["prev"][b].astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/bits"), ch
            assert ss["subaccum"][b].astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/subaccum"), ch
            assert ss["accum"][b].astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/accum"), ch
        # the three readmaps the reference migrated (make_readmap::fast_migrate)
        for p, sample in enumerate(("proband", "father", "mother")):
            mm = g.export_mergemap(p)
            assert mm["n_set"] == parts[p]["n"]
            mg = g.migrate_bits(p, z[f"{sample}|old|bits"].view("<u8"), parts[p]["n"])
            assert json.loads(z[f"{sample}|new|bitcount.json"].tobytes()) == {"nbits": mg["nbits"]}
            assert mg["bits"].astype("<u8").tobytes() == z[f"{sample}|new|bits"].tobytes()
            assert mg["subaccum"].astype("<u8").tobytes() == z[f"{sample}|new|subaccum"].tobytes()
            assert mg["accum"].astype("<u8").tobytes() == z[f"{sample}|new|accum"].tobytes()
        # flat entries of an input: its own sorted entry sequences
        mat, sizes = RS.entries_ascii(parts[1])
        got = g.export_flat(1, 1000, 500)
        assert got == [bytes(mat[i, :sizes[i]]).decode() for i in range(1000, 1500)]
        st = g.stats()
        assert st["merge_entries"] == n and st["merge_inputs"] == 3


@pytest.mark.parametrize("nsplits", [1, 7, 100000])
def test_reference_merge_cases(B, nsplits):
    rng = random.Random(nsplits)
    for case in ([[O.tseq(x) for x in ("abc", "bcd", "cde", "cdf", "dfg")]],                       # seqset_flat_test.cpp:14-45
                 [[O.tseq("abc"), O.tseq("de")]],                                       # seqset_merger_test.cpp:124-127
                 [[O.tseq("abc"), O.tseq("cde")], [O.tseq("abc"), O.tseq("efg")]],      # :129-133
                 [[O.tseq("ab"), O.tseq("bc"), O.tseq("cd"), O.tseq("be")],             # make_mergemap_test.cpp:130-137
                  [O.tseq("AB"), O.tseq("BC"), O.tseq("CD"), O.tseq("BE")]]):
        check_against_oracle(B, [seqset_entries(r) for r in case], nsplits, rng)


@pytest.mark.parametrize("seed", range(6))
def test_random_parts(B, seed):
    rng = random.Random(100 + seed)
    lo, hi = ((5, 20) if seed < 3 else (20, 140))
    reads = [["".join(rng.choice("ACGT") for _ in range(rng.randint(lo, hi))) for _ in range(rng.randint(10, 20))]
             for _ in range(rng.randint(1, 6))]
    if seed % 2:
        reads[-1] += reads[0][:5] + [r[: max(3, len(r) // 2)] for r in reads[0][5:8]]
    check_against_oracle(B, [seqset_entries(r) for r in reads], rng.choice([1, 2, 5, 50, 100000]), rng)


def test_single_base_entries(B):
    rng = random.Random(5)
    check_against_oracle(B, [[b"A", b"C", b"GT", b"T"], seqset_entries(["ACC", "G"])], 3, rng)
    check_against_oracle(B, [seqset_entries(["A", "C"])], 100000, rng)


def test_merge_with_one_chunk_equals_a_build_over_all_reads(B):
    """make_mergemap_test::merge_and_verify: merge(seqset(A), seqset(B)) is seqset(A + B); with
    parallel_splits = 1 the prev bits sit where builder::build_chunks puts them, so EVERY table equals
    the GPU build over the union of the reads"""
    from biograph_b200 import synth
    genome = synth.random_genome(30000, seed=3)
    ra = [bytes(r).decode() for r in synth.simulate_reads(genome[:20000], 1500, read_len=100, error_rate=0.0, seed=4, paired=False)]
    rb = [bytes(r).decode() for r in synth.simulate_reads(genome[10000:], 1500, read_len=100, error_rate=0.0, seed=5, paired=False)]

    def build(reads):
        with B.Bgx() as g:
            g.add_reads(reads)
            g.seed_uncorrected()
            g.build_seqset()
            return g.export_seqset()
    sa, sb, sab = build(ra), build(rb), build(ra + rb)
    with B.Bgx() as g:
        g.merge_seqsets([sa, sb], parallel_splits=1)
        m = g.export_seqset()
        na, nb = g.export_mergemap(0)["n_set"], g.export_mergemap(1)["n_set"]
    assert (na, nb) == (sa["n"], sb["n"])
    assert m["n"] == sab["n"]
    for k in ("sizes", "shared", "prev", "fixed"):
        assert np.array_equal(m[k], sab[k]), k
    for b in range(4):
        assert np.array_equal(m["subaccum"][b], sab["subaccum"][b]) and np.array_equal(m["accum"][b], sab["accum"][b])


def test_invalid_inputs_are_errors(B):
    good = part_tables(seqset_entries([O.tseq("ab")]))
    with B.Bgx() as g:
        bad = {"n": good["n"], "sizes": good["sizes"], "prev": good["prev"].copy()}
        bad["prev"][0][0] ^= np.uint64(1)   # prev bit totals != entries (seqset.cpp:123-126)
        with pytest.raises(B.BgxError, match="Invalid seqset"):
            g.merge_seqsets([bad])
        bad = {"n": good["n"], "sizes": good["sizes"].copy(), "prev": good["prev"]}
        bad["sizes"][3] = 0
        with pytest.raises(B.BgxError, match="entry size"):
            g.merge_seqsets([bad])
        with pytest.raises(B.BgxError, match="bgx_merge_seqsets first"):
            g.export_mergemap(0)
        g.merge_seqsets([good])   # the context is still usable
        assert g.export_seqset()["n"] == good["n"]
        with pytest.raises(B.BgxError, match="no such input"):
            g.export_mergemap(1)

"""The merge restatement (oracle/merge.py) against the reference's own merge output and against the
properties the reference's merge tests check.

  * golden: datasets/lambdaToyData/benchmark/family_lambda.bg is `biograph merge --in proband --in
    father --in mother` (its qc/merge_log.txt) of the three *_lambda.bg next to it.  Flattening the
    three inputs, make_mergemap and seqset_merger::merge_range over generate_chunks(n, 100000) must
    reproduce EVERY payload member of the merged seqset, prev bits included, and fast_migrate every
    member of the three migrated readmaps.
  * modules/bio_base/seqset_merger_test.cpp:93-148 (single_simple, merge2, random parts): merged
    entries == union of the parts' entries minus prefixes.
  * modules/bio_base/make_mergemap_test.cpp:19-99 (merge_single, merge_two, random parts): the merged
    entry count equals seqset_for_reads(all reads), total_bits == part size, and bit x of part p is set
    exactly where a part entry is found in the merged seqset."""
import json
import random

import numpy as np
import pytest

from oracle import merge as M
from oracle import oracle as O
from oracle.readmap import pack_bits
from tests import refseqset as RS

PARTS = ["proband_lambda", "father_lambda", "mother_lambda"]  # --in order of the golden merge (merge_log.txt)


def unpack(words, n):
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n]


def flat_of(name):
    t = RS.tables(name)
    prev01 = [unpack(t["prev"][b], t["n"]) for b in range(4)]
    return M.flat_sequences(t["fixed"], prev01, t["sizes"])


@pytest.fixture(scope="module")
def lambda_merge():
    parts = [flat_of(nm) for nm in PARTS]
    merged, bits = M.make_mergemap(parts)
    return parts, merged, bits


def test_flat_sequences_are_the_sorted_entries():
    ent = flat_of("father_lambda")
    assert all(a < b and not b.startswith(a) for a, b in zip(ent, ent[1:]))


def test_golden_family_merge_every_member(lambda_merge):
    parts, merged, bits = lambda_merge
    name = "family_lambda"
    tb = M.merge_tables(merged)
    n = tb["n"]
    assert n == json.loads(RS.member(name, "seqset.json"))["num_entries"] == 103996
    assert tb["fixed"].astype("<u8").tobytes() == RS.member(name, "fixed")
    mx = int(tb["sizes"].max())
    s_el, s_bits = O.varbit_pack(tb["sizes"], mx)
    h_el, h_bits = O.varbit_pack(tb["shared"], mx - 1)
    assert {"bits_per_value": s_bits, "element_count": n, "max_value": mx} == json.loads(
        RS.member(name, "entry_sizes/packed_varbit_vector.json"))
    assert {"bits_per_value": h_bits, "element_count": n, "max_value": mx - 1} == json.loads(
        RS.member(name, "shared/packed_varbit_vector.json"))
    assert s_el.astype("<u8").tobytes() == RS.member(name, "entry_sizes/elements")
    assert h_el.astype("<u8").tobytes() == RS.member(name, "shared/elements")
    for b, ch in enumerate("ACGT"):
        words = pack_bits(tb["prev"][b])
        sub, acc, _ = O.bitcount_finalize(words, n)
        assert words.astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/bits"), ch
        assert sub.astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/subaccum"), ch
        assert acc.astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/accum"), ch


def test_golden_family_merge_prev_bits_closed_form(lambda_merge):
    """the chunk rule in closed form (what the GPU computes) == the literal merge_range"""
    _, merged, _ = lambda_merge
    t = RS.tables("family_lambda")
    prev = M.merge_prev_closed_form(merged)
    for b in range(4):
        assert np.array_equal(prev[b], unpack(t["prev"][b], t["n"]))
    # and it is NOT the create rule (bit on the first entry): the chunking is result-visible
    first = M.merge_prev_closed_form(merged, nsplits=1)
    assert any(not np.array_equal(first[b], prev[b]) for b in range(4))


def test_mergemap_by_sort_equals_the_queue_merge(lambda_merge):
    parts, merged, bits = lambda_merge
    m2, b2 = M.make_mergemap_sorted(parts)
    assert m2 == merged
    for a, b in zip(bits, b2):
        assert np.array_equal(a, b)
    for p, b in zip(parts, bits):
        assert int(b.sum()) == len(p)  # seqset_merger.cpp:33


@pytest.mark.parametrize("sample", ["proband", "father", "mother"])
def test_golden_migrated_readmaps(lambda_merge, sample):
    parts, merged, bits = lambda_merge
    z = np.load(RS.ROOT + "/tests/golden/ref_merge_readmaps.npz")
    p = ["proband", "father", "mother"].index(sample)
    n_old = json.loads(z[f"{sample}|old|bitcount.json"].tobytes())["nbits"]
    assert n_old == len(parts[p])
    old = unpack(z[f"{sample}|old|bits"].view("<u8"), n_old)
    new = M.migrate_source_bits(old, bits[p])
    assert json.loads(z[f"{sample}|new|bitcount.json"].tobytes()) == {"nbits": len(merged)}
    words = pack_bits(new)
    sub, acc, _ = O.bitcount_finalize(words, len(merged))
    assert words.astype("<u8").tobytes() == z[f"{sample}|new|bits"].tobytes()
    assert sub.astype("<u8").tobytes() == z[f"{sample}|new|subaccum"].tobytes()
    assert acc.astype("<u8").tobytes() == z[f"{sample}|new|accum"].tobytes()


# ---- the reference's merge tests -------------------------------------------------------------------------
def seqset_entries(reads):
    return [e.encode() for e in O.entries_closed_form_py(reads)]


def check_merge(part_reads, nsplits):
    parts = [seqset_entries(r) for r in part_reads]
    merged, bits = M.make_mergemap(parts)
    # seqset_merger_test::verify: union of the parts' entries minus prefixes
    expect = sorted(set(e for p in parts for e in p))
    expect = [e for j, e in enumerate(expect) if not (j + 1 < len(expect) and expect[j + 1].startswith(e))]
    assert merged == expect
    # make_mergemap_test::merge_and_verify: == the seqset of all reads; bits where the part's entries are found
    whole = seqset_entries([r for p in part_reads for r in p])
    assert merged == whole
    for p, b in zip(parts, bits):
        assert int(b.sum()) == len(p)
        want = np.zeros(len(merged), dtype=np.uint8)
        for e in p:
            hits = [i for i, m in enumerate(merged) if m.startswith(e)]
            if hits:
                want[hits[0]] = 1   # seqset::find(slice).begin()
        # a part entry that is a prefix of several merged entries is found at the first of them; the
        # mergemap marks the run it was folded into, which starts there
        assert np.array_equal(b, want)
    # tables: any chunking gives a valid seqset; one chunk gives the create rule
    tb = M.merge_tables(merged, nsplits)
    ss = O.seqset_closed_form([r for p in part_reads for r in p])
    assert np.array_equal(tb["sizes"], ss["sizes"]) and np.array_equal(tb["shared"], ss["shared"])
    assert np.array_equal(tb["fixed"], ss["fixed"])
    closed = M.merge_prev_closed_form(merged, nsplits)
    for b in range(4):
        assert np.array_equal(tb["prev"][b], closed[b])
    one = M.merge_tables(merged, 1)
    for b in range(4):
        assert np.array_equal(pack_bits(one["prev"][b]), ss["prev"][b])
    # every chunking decodes to the same sequences (the bits stay inside the range pop_front widens over)
    assert M.flat_sequences(tb["fixed"], tb["prev"], tb["sizes"]) == merged


def test_seqset_flat_test():  # seqset_flat_test.cpp:14-45: the flat sequences are the seqset's entry sequences
    reads = [O.tseq(x) for x in ("abc", "bcd", "cde", "cdf", "dfg")]
    ss = O.seqset_closed_form(reads)
    prev01 = [unpack(ss["prev"][b], ss["n"]) for b in range(4)]
    assert M.flat_sequences(ss["fixed"], prev01, ss["sizes"]) == seqset_entries(reads)
    check_merge([reads], 100000)


def test_single_simple():  # seqset_merger_test.cpp:124-127
    check_merge([[O.tseq("abc"), O.tseq("de")]], 100000)


def test_merge2():  # seqset_merger_test.cpp:129-133
    check_merge([[O.tseq("abc"), O.tseq("cde")], [O.tseq("abc"), O.tseq("efg")]], 100000)


def test_merge_single():  # make_mergemap_test.cpp:123-128
    check_merge([[O.tseq("ab"), O.tseq("bc"), O.tseq("cd"), O.tseq("be")]], 7)


def test_merge_two():  # make_mergemap_test.cpp:130-137
    check_merge([[O.tseq("ab"), O.tseq("bc"), O.tseq("cd"), O.tseq("be")],
                 [O.tseq("AB"), O.tseq("BC"), O.tseq("CD"), O.tseq("BE")]], 13)


@pytest.mark.parametrize("seed", range(6))
def test_random_parts(seed):  # the coverage passes: 1-5 parts of 10-20 random sequences of 5-20 bases
    rng = random.Random(seed)
    parts = [["".join(rng.choice("ACGT") for _ in range(rng.randint(5, 20))) for _ in range(rng.randint(10, 20))]
             for _ in range(rng.randint(1, 5))]
    check_merge(parts, rng.randint(1, 100))


def test_prefix_across_parts():
    """make_mergemap.cpp:108-124: AB in part 1, ABC in part 2, ABB / ABD in part 3 (as whole entries)"""
    parts = [[b"ACG"], [b"ACGG"], [b"ACGC", b"ACGT"]]
    merged, bits = M.make_mergemap(parts)
    assert merged == [b"ACGC", b"ACGG", b"ACGT"]
    assert [list(b) for b in bits] == [[1, 0, 0], [0, 1, 0], [1, 0, 1]]
    m2, b2 = M.make_mergemap_sorted(parts)
    assert m2 == merged and all(np.array_equal(a, b) for a, b in zip(bits, b2))

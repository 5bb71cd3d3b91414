"""The readmap restatement (oracle/readmap.py) against the REFERENCE'S OWN make_readmap (oracle/_ref: compiled from
modules/bio_mapred/make_readmap.cpp, sparse_multi, bitcount, packed_varbit_vector, packed_vector) -- paired and
unpaired, every payload member of the readmap spiral file byte for byte.  This is the byte-level pin of the PAIRED
row order (make_readmap.h:187-205 and the sequential claim pass, make_readmap.cpp:302-360) that the fixtures in the
reference tree cannot give: its paired readmaps were written by an older build.  CPU only."""
import bisect

import numpy as np
import pytest

from oracle import oracle as O
from oracle import readmap as RM
from oracle import ref as R
from tests.test_oracle_readmap import _random_paired_case

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref.so not built (no reference checkout)")


def expected_members(t, n_entries, max_read_len):
    """the oracle's tables in the file's own encodings (this commit's layout: readmap version 1.2.0)"""
    out = {}
    for name, bits in (("source_to_mid", t["source_to_mid"]), ("dest_to_mid", t["dest_to_mid"])):
        words = RM.pack_bits(bits)
        sub, acc, _ = O.bitcount_finalize(words, len(bits))
        out[f"read_ids/{name}/bits"], out[f"read_ids/{name}/subaccum"], out[f"read_ids/{name}/accum"] = words, sub, acc
    out["read_lengths/elements"] = O.varbit_pack(t["read_lengths"], max_read_len)[0]
    out["mate_loop_ptr/elements"] = varbit_pack64(t["mate_loop_ptr"], t["n_rows"])
    out["is_forward/packed_data"] = RM.pack_bits(t["is_forward"])
    return out


def varbit_pack64(vals, max_value):
    """packed_varbit_vector for values beyond 16 bits (modules/io/packed_varbit_vector.cpp:174-228): bits_per_value =
    bits of max_value, value i at bit i * bits_per_value of a little-endian uint64 array, padded by one word"""
    bits = max(1, int(max_value).bit_length())
    n = len(vals)
    words = (n * bits + 63) // 64 + 1
    out = np.zeros(words, dtype=np.uint64)
    for i, v in enumerate(np.asarray(vals, dtype=np.uint64).tolist()):
        pos = i * bits
        w, o = pos // 64, pos % 64
        out[w] |= np.uint64((v << o) & 0xFFFFFFFFFFFFFFFF)
        if o + bits > 64:
            out[w + 1] |= np.uint64(v >> (64 - o))
    return out


def check(mem, exp):
    for name, arr in exp.items():
        got = np.frombuffer(mem[name], dtype=np.uint64)
        arr = np.asarray(arr).view(np.uint64) if np.asarray(arr).dtype != np.uint64 else np.asarray(arr)
        n = min(len(got), len(arr))
        assert np.array_equal(got[:n], arr[:n]), name
        assert not got[n:].any() and not arr[n:].any(), name + " (tail)"


def run_case(reads, kept, ents, is_paired):
    kept_reads = [r for r, k in zip(reads, kept) if k]

    def lookup(s):
        i = bisect.bisect_left(ents, s)
        assert ents[i].startswith(s)
        return i

    fwd = [lookup(r) if k else 0 for r, k in zip(reads, kept)]
    rc = [lookup(O.revcomp(r)) if k else 0 for r, k in zip(reads, kept)]
    lens = [len(r) for r in reads]
    rec, ro = [], [0]
    if is_paired:
        cols = RM.pair_records(fwd, rc, lens, kept)
        t = RM.readmap_tables_paired(*cols, len(ents))
        for i in range(len(reads) // 2):  # one record per pair; a dropped mate leaves a one-read record
            c = [reads[2 * i + j] for j in range(2) if kept[2 * i + j]]
            if c:
                rec += c
                ro.append(len(rec))
    else:
        k = np.asarray(kept, dtype=bool)
        t = RM.readmap_tables(np.asarray(fwd)[k], np.asarray(rc)[k], np.asarray(lens)[k], len(ents))
        for r in kept_reads:
            rec.append(r)
            ro.append(len(rec))
    with R.Run(2) as run:
        run.seed(kept_reads)
        ss = run.make_seqset()
        assert ss["n"] == len(ents)
        mem = run.make_readmap(rec, ro, is_paired)
    check(mem, expected_members(t, len(ents), int(ss["sizes"].max())))
    return t


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_paired_readmap_byte_for_byte(seed):
    reads, kept, ents = _random_paired_case(seed, 300 + 50 * seed, 400, 40, 0.3, 0.1)
    t = run_case(reads, kept, ents, True)
    assert t["n_rows"] > 1000


def test_paired_readmap_many_identical_pairs():
    # long runs of identical rows: the claim order inside a run is what the sequential pass decides
    reads, kept, ents = _random_paired_case(11, 60, 150, 30, 3.0, 0.05)
    run_case(reads, kept, ents, True)


@pytest.mark.parametrize("seed", [7, 8])
def test_unpaired_readmap_byte_for_byte(seed):
    reads, kept, ents = _random_paired_case(seed, 400, 500, 50, 0.3, 0.1)
    run_case(reads, kept, ents, False)


def test_pairs_built_unpaired_mode_is_refused():
    # is_paired = false with two-read records: the reference throws (make_readmap.cpp:131-136)
    reads, kept, ents = _random_paired_case(3, 50, 200, 30, 0.0, 0.0)
    with R.Run(2) as run:
        run.seed(reads)
        run.make_seqset()
        with pytest.raises(RuntimeError, match="Unexpected read pairing"):
            run.make_readmap(reads, list(range(0, len(reads) + 1, 2)), False)


def test_paired_readmap_degenerate_pairs():
    """pairs whose mates are identical, reverse complements of each other, prefixes of each other, given both ways
    round, and repeated: the canonical orientation (make_readmap.cpp:170-175 compares the SEQUENCES) and the claim order
    inside runs of identical rows"""
    rng = np.random.default_rng(123)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
    a, b, c = genome[10:50], genome[60:95], genome[100:140]
    pal = "ACGTACGTTTAAACGTACGT"                      # its own reverse complement
    assert O.revcomp(pal) == pal
    pairs = [(a, a), (a, O.revcomp(a)), (O.revcomp(a), a), (a, a[:25]), (a[:25], a), (b, c), (c, b), (b, c), (pal, pal),
             (pal, a), (a, pal), (O.revcomp(b), O.revcomp(c)), (c, c), (c, O.revcomp(c))] * 3
    reads = [r for p in pairs for r in p]
    kept = np.ones(len(reads), dtype=bool)
    kept[[5, 12, 13, 40]] = False                     # a few dropped mates, one pair dropped entirely
    ents = sorted(O.entries_closed_form_py([r for r, k in zip(reads, kept) if k]))
    run_case(reads, kept, ents, True)
    run_case(reads, kept, ents, False)

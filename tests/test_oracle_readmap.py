"""The readmap restatement (oracle/readmap.py, unpaired reads) against every payload member of the
reference's own golden readmap, golden/e_coli_10000snp.bg/coverage/<sha1>.readmap
(tests/golden/e_coli_10000snp_readmap.npz).  CPU only."""
import bisect
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import readmap as RM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_golden_readmap_members(golden_reads):
    z = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_10000snp_readmap.npz"))
    solid = O.solid_set(O.count_kmers(golden_reads, 30), 5)
    cr = O.correct_reads(golden_reads, solid, 30)
    seqs = O.corrected_list(cr)
    assert len(seqs) == 8444
    ents = sorted(O.entries_closed_form_py(seqs))
    assert len(ents) == 19935

    def lookup(s):
        i = bisect.bisect_left(ents, s)
        assert ents[i].startswith(s)
        return i

    fwd = [lookup(s) for s in seqs]
    rc = [lookup(O.revcomp(s)) for s in seqs]
    t = RM.readmap_tables(fwd, rc, [len(s) for s in seqs], len(ents))
    assert t["n_rows"] == 16888
    # read_ids: the two bitcount bit vectors, with their rank indexes
    for name, bits in (("source_to_mid", t["source_to_mid"]), ("dest_to_mid", t["dest_to_mid"])):
        words = RM.pack_bits(bits)
        assert np.array_equal(words, z[f"read_ids|{name}|bits"].view("<u8")), name
        sub, acc, _ = O.bitcount_finalize(words, len(bits))
        assert np.array_equal(sub, z[f"read_ids|{name}|subaccum"].view("<u8")), name
        assert np.array_equal(acc, z[f"read_ids|{name}|accum"].view("<u8")), name
    # v3.1.1 layouts: read_lengths raw uint8, mate_loop_ptr 32-bit values, is_forward 1-bit values
    assert np.array_equal(t["read_lengths"].astype(np.uint8), z["read_lengths"])
    assert np.array_equal(t["mate_loop_ptr"].astype("<u4"), z["mate_loop_ptr|packed_data"].view("<u4"))
    assert np.array_equal(RM.pack_bits(t["is_forward"]), z["is_forward|packed_data"].view("<u8"))


def _random_paired_case(seed, n_pairs, genome_len, read_len, dup_frac, drop_frac):
    """corrected pairs from a small genome with many duplicates (identical reads, identical mates of
    different reads, reads that are prefixes of others), entries by the closed form"""
    rng = np.random.default_rng(seed)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, genome_len))
    reads = []
    for _ in range(n_pairs):
        a = int(rng.integers(0, genome_len - 2 * read_len))
        la, lb = int(rng.integers(read_len // 2, read_len + 1)), int(rng.integers(read_len // 2, read_len + 1))
        r1 = genome[a:a + la]
        r2 = O.revcomp(genome[a + read_len // 2: a + read_len // 2 + lb])
        if rng.random() < 0.5:
            r1, r2 = r2, r1
        reads += [r1, r2]
    n_dup = int(dup_frac * n_pairs)
    for _ in range(n_dup):  # same first read, another pair's mate
        i, j = int(rng.integers(0, n_pairs)), int(rng.integers(0, n_pairs))
        reads += [reads[2 * i], reads[2 * j + 1]]
    for _ in range(n_dup):  # exact duplicate pairs
        i = int(rng.integers(0, n_pairs))
        reads += [reads[2 * i], reads[2 * i + 1]]
    kept = rng.random(len(reads)) >= drop_frac
    ents = sorted(O.entries_closed_form_py([r for r, k in zip(reads, kept) if k]))
    return reads, kept, ents


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_paired_parallel_form_equals_literal_transcription(seed):
    reads, kept, ents = _random_paired_case(seed, 300, 400, 40, 0.3, 0.1)

    def lookup(s):
        i = bisect.bisect_left(ents, s)
        assert ents[i].startswith(s)
        return i

    fwd = [lookup(r) if k else 0 for r, k in zip(reads, kept)]
    rc = [lookup(O.revcomp(r)) if k else 0 for r, k in zip(reads, kept)]
    cols = RM.pair_records(fwd, rc, [len(r) for r in reads], kept)
    # canonical orientation == the reference's sequence comparison (make_readmap.cpp:170-175)
    kept_pairs = [(reads[2 * i], reads[2 * i + 1]) for i in range(len(reads) // 2) if kept[2 * i] and kept[2 * i + 1]]
    both = cols[5] > 0
    assert int(both.sum()) == len(kept_pairs)
    for (r1, r2), e_, l_ in zip(kept_pairs, cols[0][both], cols[2][both]):
        small = min(r1, r2)
        assert ents[int(e_)].startswith(small) and int(l_) == len(small)
    a = RM.readmap_tables_paired(*cols, len(ents))
    b = RM.readmap_tables_paired_literal(*cols, len(ents))
    for k_ in ("read_lengths", "source_to_mid", "dest_to_mid", "mate_loop_ptr", "is_forward"):
        assert np.array_equal(a[k_], b[k_]), k_
    # the properties readmap_test.cpp:53-168 checks: loops fwd -> rc [-> mate fwd -> mate rc] -> back
    ptr, fw, ln = a["mate_loop_ptr"].astype(np.int64), a["is_forward"], a["read_lengths"]
    for i in np.flatnonzero(a["type"] == RM.LOOP_START):
        r = ptr[i]
        assert fw[i] and not fw[r] and ln[i] == ln[r]
        assert ents[int(a["entry_id"][r])][:ln[r]] == O.revcomp(ents[int(a["entry_id"][i])][:ln[i]])
        if ptr[r] == i:
            continue
        mt = ptr[r]
        mr = ptr[mt]
        assert fw[mt] and not fw[mr] and ptr[mr] == i and ln[mt] == ln[mr]
        assert ents[int(a["entry_id"][mr])][:ln[mr]] == O.revcomp(ents[int(a["entry_id"][mt])][:ln[mt]])


def test_unpaired_is_the_paired_form_without_mates(golden_reads):
    rng = np.random.default_rng(5)
    n_ent = 5000
    e = np.sort(rng.integers(0, n_ent, 3000)).astype(np.uint64)
    rc = rng.permutation(e)  # any consistent map would do: reuse values so RC runs exist
    # make (rc entry, length) a function of (entry, length) as it is for real reads
    ln = (e % 7 + 30).astype(np.uint64)
    rc = (n_ent - 1 - e).astype(np.uint64)
    z = np.zeros(len(e), np.uint64)
    a = RM.readmap_tables(e, rc, ln, n_ent)
    b = RM.readmap_tables_paired(e, rc, ln, z, z, z, n_ent)
    for k_ in ("read_lengths", "source_to_mid", "dest_to_mid", "mate_loop_ptr", "is_forward"):
        assert np.array_equal(a[k_], b[k_]), k_

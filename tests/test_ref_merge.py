"""The merge restatement (oracle/merge.py) against the REFERENCE'S OWN merge classes (oracle/_ref: seqset_flat,
make_mergemap, seqset_mergemap, seqset_merger compiled from modules/bio_base/*.cpp) on seqsets the reference's own
builder made from random read sets: flat sequences, mergemap bits, and every table of the merged seqset -- the prev
bits with seqset_merger's chunk rule included (more than g_parallel_splits = 100 000 merged entries in the large
case, so that chunks hold more than one entry).  CPU only."""
import numpy as np
import pytest

from oracle import merge as M
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref.so not built (no reference checkout)")


def words(bits01):
    b = np.asarray(bits01, dtype=np.uint8)
    return np.frombuffer(np.packbits(np.pad(b, (0, (-len(b)) % 64)), bitorder="little").tobytes(), dtype=np.uint64)


def read_sets(seed, n_sets, n_reads, genome_len, max_len, err=0.1):
    rng = np.random.default_rng(seed)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, genome_len))
    out = []
    for _ in range(n_sets):
        rs = []
        for _ in range(n_reads):
            a = int(rng.integers(0, genome_len - max_len))
            s = genome[a:a + int(rng.integers(8, max_len + 1))]
            if rng.random() < err:
                p = int(rng.integers(0, len(s)))
                s = s[:p] + "ACGT"[int(rng.integers(0, 4))] + s[p + 1:]
            if rng.random() < 0.5:
                s = O.revcomp(s)
            rs.append(s)
        out.append(rs)
    return out


def merge_case(sets, literal=True):
    runs = [R.Run(4) for _ in sets]
    out = R.Run(4)
    try:
        ents = []
        for r, rd in zip(runs, sets):
            r.seed(rd)
            t = r.make_seqset()
            flat = r.flat()
            # seqset_flat::get(i) is the entry's sequence: sorted, prefix-free, and what the oracle derives from the tables
            assert flat == sorted(flat) and len(flat) == t["n"]
            prev01 = [np.unpackbits(t["prev"][b].view(np.uint8), bitorder="little")[:t["n"]] for b in range(4)]
            assert M.flat_sequences(t["fixed"], prev01, t["sizes"]) == flat
            ents.append(flat)
        tab, maps = out.merge_from(runs)
        merged, bits = M.make_mergemap(ents)
        assert tab["n"] == len(merged)
        assert out.flat() == merged
        for p in range(len(sets)):
            assert np.array_equal(maps[p], words(bits[p])), f"mergemap {p}"
        assert M.make_mergemap_sorted(ents)[0] == merged
        if literal:
            t = M.merge_tables(merged)
            prev = t["prev"]
        else:
            prev = M.merge_prev_closed_form(merged)
        assert np.array_equal(tab["sizes"], np.array([len(e) for e in merged], dtype=np.uint16))
        sh = [0] + [len(os_path_commonprefix(a, b)) for a, b in zip(merged, merged[1:])]
        assert np.array_equal(tab["shared"], np.array(sh, dtype=np.uint16))
        for b in range(4):
            assert np.array_equal(tab["prev"][b], words(prev[b])), f"prev_{'ACGT'[b]}"
        fixed = np.concatenate([[0], np.cumsum([int(np.sum(p)) for p in prev])]).astype(np.uint64)
        assert np.array_equal(tab["fixed"], fixed)
        return tab
    finally:
        for r in runs + [out]:
            r.close()


def os_path_commonprefix(a, b):
    n = min(len(a), len(b))
    i = 0
    while i < n and a[i] == b[i]:
        i += 1
    return a[:i]


@pytest.mark.parametrize("seed,n_sets", [(1, 1), (2, 2), (3, 3), (4, 5)])
def test_merge_random_seqsets(seed, n_sets):
    merge_case(read_sets(seed, n_sets, 300, 2500, 50))


def test_merge_identical_and_disjoint_inputs():
    a, b = read_sets(9, 2, 200, 2000, 40)
    merge_case([a, a])           # identical inputs: every mergemap bit set in both
    other = read_sets(10, 1, 200, 2000, 40)[0]
    merge_case([a, other, b])


def test_merge_beyond_parallel_splits():
    # > 100 000 merged entries: seqset_merger's chunks (generate_chunks over g_parallel_splits) hold several entries,
    # and the prev bit of a candidate lands on the first entry OF THE CHUNK that holds the last entry it prefixes
    tab = merge_case(read_sets(21, 2, 16000, 130000, 60, err=0.05), literal=False)
    assert tab["n"] > 2 * M.K_PARALLEL_SPLITS


def test_fast_migrate_on_the_reference(tmp_path):
    """make_readmap::fast_migrate through the reference's own code: each input's readmap (written by the reference's
    make_readmap) moved onto the merged seqset; the oracle's migrate_source_bits gives the new source bits, every
    other payload member is carried over unchanged, and the migrated file opens against the merged seqset"""
    sets = read_sets(31, 2, 250, 2000, 45)
    runs = [R.Run(2) for _ in sets]
    out = R.Run(2)
    try:
        old_members, ents = [], []
        for p, (r, rd) in enumerate(zip(runs, sets)):
            r.seed(rd)
            r.make_seqset()
            ents.append(r.flat())
            old_members.append(r.make_readmap(rd, list(range(len(rd) + 1)), False, keep_path=str(tmp_path / f"in{p}.readmap")))
        tab, maps = out.merge_from(runs)
        merged, bits = M.make_mergemap(ents)
        for p in range(len(sets)):
            new = out.fast_migrate(p, str(tmp_path / f"in{p}.readmap"), str(tmp_path / f"out{p}.readmap"))
            old = old_members[p]
            old_src = np.unpackbits(np.frombuffer(old["read_ids/source_to_mid/bits"], dtype=np.uint8), bitorder="little")[:len(ents[p])]
            want = words(M.migrate_source_bits(old_src, bits[p]))
            got = np.frombuffer(new["read_ids/source_to_mid/bits"], dtype=np.uint64)
            n = min(len(got), len(want))
            assert np.array_equal(got[:n], want[:n]) and not got[n:].any() and not want[n:].any()
            sub, acc, _ = O.bitcount_finalize(want, len(merged))
            assert np.array_equal(np.frombuffer(new["read_ids/source_to_mid/subaccum"], dtype=np.uint64), sub)
            assert np.array_equal(np.frombuffer(new["read_ids/source_to_mid/accum"], dtype=np.uint64), acc)
            for name in ("read_ids/dest_to_mid/bits", "read_ids/dest_to_mid/subaccum", "read_ids/dest_to_mid/accum",
                         "read_lengths/elements", "mate_loop_ptr/elements", "is_forward/packed_data"):
                assert new[name] == old[name], name
            # every read still points at its own sequence: row i's entry in the merged seqset starts with the read
            rows = R.read_readmap_file(str(tmp_path / f"out{p}.readmap"))
            old_rows = R.read_readmap_file(str(tmp_path / f"in{p}.readmap"))
            for i in range(0, len(rows["entry_id"]), 7):
                a = merged[int(rows["entry_id"][i])][:int(rows["read_lengths"][i])]
                b = ents[p][int(old_rows["entry_id"][i])][:int(old_rows["read_lengths"][i])]
                assert a == b
    finally:
        for r in runs + [out]:
            r.close()

"""CPU restatement of make_readmap::create_from_reads for UNPAIRED reads -- test
infrastructure: imported only by tests/ (same rule as oracle.py).

Follows modules/bio_mapred/make_readmap.cpp:
  :134-188  one LOOP_START row (entry id of the corrected read, length, loop entry = the entry id of
            its reverse complement) and one RC row (entry id of the reverse complement, length,
            loop entry = none) per corrected read
  make_readmap.h:187-205  rows sorted by (entry_id, type, read_length, mate_read_length, loop_entry_id)
  :236-245  sparse_multi over the sorted rows' entry ids (modules/io/sparse_multi.cpp:90-113):
            source_to_mid bit e = entry e has a read, dest_to_mid bit i = row i opens its entry's run
  :249-251  read_lengths[i]
  :262-300  first pass: LOOP_START rows get is_forward = 1 and point at the first matching RC row
  :302-360  second pass, in row order: every LOOP_START claims the next unclaimed matching RC row and
            the RC row points back -- for unpaired reads the j-th of a run of identical
            LOOP_START rows gets the j-th row of the matching run of RC rows.
Pinned against the reference's golden readmap (tests/golden/e_coli_10000snp_readmap.npz, made by
tests/golden/make_golden_readmap.py): every payload member reproduced.

Paired reads: readmap_tables_paired is the parallel form (what the GPU computes), and
readmap_tables_paired_literal transcribes the reference's two passes statement by statement
(:168-188 rows, :262-300 first pass, :302-360 the sequential claim pass with its `claimed` bit
vector); tests check the two against each other.  The paired form is pinned against the reference's
OWN make_readmap, compiled from its sources into oracle/_ref and run on the same records
(tests/test_ref_readmap.py: every payload member byte for byte); the only readmap in the reference
tree that follows the current row order is the unpaired golden (the paired readmaps under datasets/
were written by an older build), so no fixture could."""
import numpy as np

K_NO_LOOP_ENTRY = (1 << 37) - 1
LOOP_START, RC, MATE, MATE_RC = 0, 1, 2, 3


def readmap_tables(fwd_entry, rc_entry, lens, n_entries):
    """fwd_entry, rc_entry, lens: one element per KEPT read, in record order."""
    fwd_entry = np.asarray(fwd_entry, dtype=np.uint64)
    rc_entry = np.asarray(rc_entry, dtype=np.uint64)
    lens = np.asarray(lens, dtype=np.uint64)
    n = len(lens)
    entry = np.concatenate([fwd_entry, rc_entry])
    typ = np.concatenate([np.full(n, LOOP_START, np.uint64), np.full(n, RC, np.uint64)])
    rlen = np.concatenate([lens, lens])
    loop = np.concatenate([rc_entry, np.full(n, K_NO_LOOP_ENTRY, np.uint64)])
    order = np.lexsort((loop, rlen, typ, entry))  # last key is the primary one; mate_read_length is 0 throughout
    entry, typ, rlen, loop = entry[order], typ[order], rlen[order], loop[order]
    m = 2 * n
    # sparse_multi
    src = np.zeros(n_entries, dtype=np.uint8)
    src[entry.astype(np.int64)] = 1
    dst = np.ones(m, dtype=np.uint8)
    if m:
        dst[1:] = entry[1:] != entry[:-1]
    # mate loop: run index of every row among identical rows
    same = np.zeros(m, dtype=bool)
    if m:
        same[1:] = (entry[1:] == entry[:-1]) & (typ[1:] == typ[:-1]) & (rlen[1:] == rlen[:-1]) & (loop[1:] == loop[:-1])
    idx = np.arange(m, dtype=np.int64)
    run_start = np.maximum.accumulate(np.where(same, 0, idx))
    rank = idx - run_start
    # first RC row of (entry, length): rows are sorted, so a search over a combined key does it
    key = (entry << np.uint64(12)) | (typ << np.uint64(10)) | rlen  # lengths < 1024 (k_read_length_bits = 10)
    ptr = np.zeros(m, dtype=np.uint64)
    is_fwd = (typ == LOOP_START).astype(np.uint8)
    ls = np.flatnonzero(typ == LOOP_START)
    want = (loop[ls] << np.uint64(12)) | (np.uint64(RC) << np.uint64(10)) | rlen[ls]
    first_rc = np.searchsorted(key, want, side="left")
    tgt = first_rc + rank[ls]
    assert np.all(key[tgt] == want), "a LOOP_START row has no RC row to claim"
    ptr[ls] = tgt.astype(np.uint64)
    ptr[tgt] = ls.astype(np.uint64)
    return {"n_rows": m, "entry_id": entry, "type": typ, "read_lengths": rlen.astype(np.uint16),
            "source_to_mid": src, "dest_to_mid": dst, "mate_loop_ptr": ptr, "is_forward": is_fwd}


def pack_bits(bits01):
    """uint8 0/1 array -> little-endian uint64 words, bit i at word[i / 64] >> (i & 63)"""
    b = np.asarray(bits01, dtype=np.uint8)
    pad = (-len(b)) % 64
    return np.packbits(np.concatenate([b, np.zeros(pad, np.uint8)]), bitorder="little").view("<u8")


def pair_records(fwd_entry, rc_entry, lens, kept):
    """Reads 2i and 2i+1 are mates (one record per pair, as corrected_reads holds them).  Returns the
    per-record columns (e, rc_e, ln, me, rc_me, ml) in the reference's canonical orientation:
    a pair with one read dropped is a single read (read_pair.size() == 1, :157-158), a pair with
    both dropped is no record, and the read with the smaller sequence is the LOOP_START (:170-175;
    sequence order == (entry id, length) order, because the entry of a read is the first entry
    it is a prefix of)."""
    fwd_entry, rc_entry = np.asarray(fwd_entry, dtype=np.uint64), np.asarray(rc_entry, dtype=np.uint64)
    lens, kept = np.asarray(lens, dtype=np.uint64), np.asarray(kept, dtype=bool)
    assert len(lens) % 2 == 0
    a, b = np.arange(0, len(lens), 2), np.arange(1, len(lens), 2)
    both = kept[a] & kept[b]
    swap = both & ((fwd_entry[a] > fwd_entry[b]) | ((fwd_entry[a] == fwd_entry[b]) & (lens[a] > lens[b])))
    first = np.where(kept[a] & ~swap, a, b)   # the LOOP_START read of the record
    second = np.where(first == a, b, a)
    rec = kept[a] | kept[b]
    first, second, both = first[rec], second[rec], both[rec]
    z = np.zeros(len(first), np.uint64)
    return (fwd_entry[first], rc_entry[first], lens[first], np.where(both, fwd_entry[second], z),
            np.where(both, rc_entry[second], z), np.where(both, lens[second], z))


def _paired_rows(e, rc_e, ln, me, rc_me, ml):
    u = np.uint64
    e, rc_e, ln = np.asarray(e, dtype=u), np.asarray(rc_e, dtype=u), np.asarray(ln, dtype=u)
    me, rc_me, ml = np.asarray(me, dtype=u), np.asarray(rc_me, dtype=u), np.asarray(ml, dtype=u)
    has = ml > 0
    none = np.full(len(e), K_NO_LOOP_ENTRY, u)
    zero = np.zeros(len(e), u)
    k = int(has.sum())
    entry = np.concatenate([e, rc_e, me[has], rc_me[has]])
    typ = np.concatenate([np.full(len(e), LOOP_START, u), np.full(len(e), RC, u), np.full(k, MATE, u), np.full(k, MATE_RC, u)])
    rlen = np.concatenate([ln, ln, ml[has], ml[has]])
    mlen = np.concatenate([zero, np.where(has, ml, zero), zero[has], zero[has]])
    loop = np.concatenate([rc_e, np.where(has, me, none), rc_me[has], none[has]])
    order = np.lexsort((loop, mlen, rlen, typ, entry))  # make_readmap.h:187-205
    return entry[order], typ[order], rlen[order], mlen[order], loop[order]


def _common_tables(entry, rlen, n_entries):
    m = len(entry)
    src = np.zeros(n_entries, dtype=np.uint8)
    src[entry.astype(np.int64)] = 1
    dst = np.ones(m, dtype=np.uint8)
    if m:
        dst[1:] = entry[1:] != entry[:-1]
    return src, dst


def readmap_tables_paired(e, rc_e, ln, me, rc_me, ml, n_entries):
    """Parallel form.  One element per record; ml == 0: the record is a single read."""
    u = np.uint64
    entry, typ, rlen, mlen, loop = _paired_rows(e, rc_e, ln, me, rc_me, ml)
    m = len(entry)
    src, dst = _common_tables(entry, rlen, n_entries)
    idx = np.arange(m, dtype=np.int64)
    key = (entry << u(12)) | (typ << u(10)) | rlen  # (entry, type, length): what find_first_of searches for
    ptr = np.zeros(m, dtype=np.int64)
    # LOOP_START -> RC: the j-th row of a run of identical LOOP_START rows takes the j-th RC row of (loop, length)
    ls = np.flatnonzero(typ == LOOP_START)
    rc_idx = np.searchsorted(key, (loop[ls] << u(12)) | (u(RC) << u(10)) | rlen[ls]) + (ls - np.searchsorted(key, key[ls]))
    assert np.all((typ[rc_idx] == RC) & (entry[rc_idx] == loop[ls]) & (rlen[rc_idx] == rlen[ls]))
    ptr[ls] = rc_idx
    solo = loop[rc_idx] == u(K_NO_LOOP_ENTRY)
    ptr[rc_idx[solo]] = ls[solo]
    # RC -> MATE: the MATE rows of (loop, mate length) go to their claimers in LOOP_START row order
    ls_p, rc_p = ls[~solo], rc_idx[~solo]
    first_mate = np.searchsorted(key, (loop[rc_p] << u(12)) | (u(MATE) << u(10)) | mlen[rc_p])
    o = np.lexsort((ls_p, first_mate))
    ls_p, rc_p, first_mate = ls_p[o], rc_p[o], first_mate[o]
    j = np.arange(len(o), dtype=np.int64)
    rank = j - np.searchsorted(first_mate, first_mate)
    mate_idx = first_mate + rank
    assert np.all((typ[mate_idx] == MATE) & (entry[mate_idx] == loop[rc_p]) & (rlen[mate_idx] == mlen[rc_p]))
    ptr[rc_p] = mate_idx
    # MATE -> MATE_RC: same claimers in the same order; MATE_RC -> LOOP_START
    mrc_idx = np.searchsorted(key, (loop[mate_idx] << u(12)) | (u(MATE_RC) << u(10)) | mlen[rc_p]) + rank
    assert np.all((typ[mrc_idx] == MATE_RC) & (entry[mrc_idx] == loop[mate_idx]) & (rlen[mrc_idx] == mlen[rc_p]))
    ptr[mate_idx] = mrc_idx
    ptr[mrc_idx] = ls_p
    is_fwd = ((typ == LOOP_START) | (typ == MATE)).astype(np.uint8)
    return {"n_rows": m, "entry_id": entry, "type": typ, "read_lengths": rlen.astype(np.uint16), "source_to_mid": src,
            "dest_to_mid": dst, "mate_loop_ptr": ptr.astype(np.uint64), "is_forward": is_fwd, "idx": idx}


def readmap_tables_paired_literal(e, rc_e, ln, me, rc_me, ml, n_entries):
    """make_readmap.cpp:262-360 statement by statement (python loops: small inputs only)."""
    import bisect
    entry, typ, rlen, mlen, loop = _paired_rows(e, rc_e, ln, me, rc_me, ml)
    m = len(entry)
    src, dst = _common_tables(entry, rlen, n_entries)
    rows = [(int(entry[i]), int(typ[i]), int(rlen[i]), int(mlen[i]), int(loop[i])) for i in range(m)]
    ptr = [0] * m
    is_fwd = [0] * m

    def find_first_of(t, ent, length):
        return bisect.bisect_left(rows, (ent, t, length, 0, 0))

    for i, (ent, t, length, mlength, lp) in enumerate(rows):  # first pass
        if t == LOOP_START:
            is_fwd[i] = 1
            ptr[i] = find_first_of(RC, lp, length)
        elif t == RC:
            if lp != K_NO_LOOP_ENTRY:
                ptr[i] = find_first_of(MATE, lp, mlength)
        elif t == MATE:
            is_fwd[i] = 1
            ptr[i] = find_first_of(MATE_RC, lp, length)
    claimed = [False] * m

    def claim_next(try_idx, t, ent, length):
        assert try_idx < m and rows[try_idx][:3] == (ent, t, length)
        while claimed[try_idx]:  # claim_next_available
            try_idx += 1
        claimed[try_idx] = True
        assert try_idx < m and rows[try_idx][:3] == (ent, t, length)
        return try_idx

    for i, (ent, t, length, mlength, lp) in enumerate(rows):  # second pass, in row order
        if t != LOOP_START:
            continue
        rc_i = claim_next(ptr[i], RC, lp, length)
        ptr[i] = rc_i
        r = rows[rc_i]
        if r[4] == K_NO_LOOP_ENTRY:
            ptr[rc_i] = i
            continue
        mate_i = claim_next(ptr[rc_i], MATE, r[4], r[3])
        ptr[rc_i] = mate_i
        mr = rows[mate_i]
        mrc_i = claim_next(ptr[mate_i], MATE_RC, mr[4], r[3])
        ptr[mate_i] = mrc_i
        ptr[mrc_i] = i
    return {"n_rows": m, "entry_id": entry, "type": typ, "read_lengths": rlen.astype(np.uint16), "source_to_mid": src,
            "dest_to_mid": dst, "mate_loop_ptr": np.array(ptr, dtype=np.uint64), "is_forward": np.array(is_fwd, dtype=np.uint8)}

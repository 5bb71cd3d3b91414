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
tests/golden/make_golden_readmap.py): every payload member reproduced.  Paired reads (the MATE /
MATE_RC rows and the four-row loops) are not restated."""
import numpy as np

K_NO_LOOP_ENTRY = (1 << 37) - 1
LOOP_START, RC = 0, 1


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

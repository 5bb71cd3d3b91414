"""The readmap restatement (oracle/readmap.py, unpaired reads) against every payload member of the
reference's own golden readmap, golden/e_coli_10000snp.bg/coverage/<sha1>.readmap
(tests/golden/e_coli_10000snp_readmap.npz).  CPU only."""
import bisect
import os

import numpy as np

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

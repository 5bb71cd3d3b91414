"""Parity against seqsets the REFERENCE itself built (tests/golden/ref_seqsets.npz: lambda phage
150 bp x4, HIV 250 bp x5; v1.1.0 layout).  A seqset is a fixed point of its own construction:
its entries, fed back as uncorrected reads (seqset_for_reads seeding, one record per read and per
reverse complement), must reproduce EVERY payload member byte for byte -- fixed, the varbit
`elements` of entry_sizes and shared, and bits / subaccum / accum of the four prev bitcounts."""
import json

import numpy as np
import pytest

from oracle import oracle as O
from tests import refseqset as RS


# family_lambda is the output of `biograph merge` (seqset_merger, out of scope), whose prev bits
# sit on the LAST entry a popped sequence prefixes; `biograph create` (bs/builder.cpp:85-107) puts
# them on the FIRST.  Its entry set, sizes, shared and fixed are still the closure's.
MERGE_BUILT = {"family_lambda"}


def _check_members(name, n, fixed, sizes_el, shared_el, meta_s, meta_h, prev, sub, acc):
    assert n == json.loads(RS.member(name, "seqset.json"))["num_entries"]
    assert fixed.astype("<u8").tobytes() == RS.member(name, "fixed")
    assert meta_s == json.loads(RS.member(name, "entry_sizes/packed_varbit_vector.json"))
    assert meta_h == json.loads(RS.member(name, "shared/packed_varbit_vector.json"))
    assert sizes_el.astype("<u8").tobytes() == RS.member(name, "entry_sizes/elements")
    assert shared_el.astype("<u8").tobytes() == RS.member(name, "shared/elements")
    for b, ch in enumerate("ACGT"):
        assert json.loads(RS.member(name, f"prev_{ch}/bitcount.json")) == {"nbits": n}
        if name in MERGE_BUILT:
            continue
        assert prev[b].astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/bits"), f"prev_{ch}/bits"
        assert sub[b].astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/subaccum"), f"prev_{ch}/subaccum"
        assert acc[b].astype("<u8").tobytes() == RS.member(name, f"prev_{ch}/accum"), f"prev_{ch}/accum"


def test_fixture_entries_are_sorted_and_prefix_free():
    t = RS.tables("family_lambda")
    mat, sizes = RS.entries_ascii(t)
    seqs = [bytes(mat[i, :sizes[i]]) for i in range(0, t["n"], 97)]
    assert seqs == sorted(seqs)
    a = [bytes(mat[i, :sizes[i]]) for i in range(2000)]
    for x, y, sh in zip(a, a[1:], t["shared"][1:2000]):
        assert x < y and not y.startswith(x)
        lcp = next((i for i, (p, q) in enumerate(zip(x, y)) if p != q), min(len(x), len(y)))
        assert lcp == sh


@pytest.mark.parametrize("name", ["family_lambda", "father_lambda", "ERR732130"])
def test_oracle_reproduces_reference_built_seqset(name):
    t = RS.tables(name)
    buf, offs = RS.as_reads(t)
    ss = O.seqset_staged((buf.tobytes(), offs), np.ones(t["n"], np.int32), np.ones(t["n"], np.int32))
    mx = int(ss["sizes"].max())
    s_el, s_bits = O.varbit_pack(ss["sizes"], mx)
    h_el, h_bits = O.varbit_pack(ss["shared"], mx - 1)
    fin = [O.bitcount_finalize(ss["prev"][b], ss["n"]) for b in range(4)]
    _check_members(name, ss["n"], ss["fixed"], s_el, h_el,
                   {"bits_per_value": s_bits, "element_count": ss["n"], "max_value": mx},
                   {"bits_per_value": h_bits, "element_count": ss["n"], "max_value": mx - 1},
                   ss["prev"], [f[0] for f in fin], [f[1] for f in fin])


@pytest.mark.gpu
@pytest.mark.parametrize("name", RS.names())
def test_gpu_reproduces_reference_built_seqset(name):
    import biograph_b200 as B
    t = RS.tables(name)
    buf, offs = RS.as_reads(t)
    with B.Bgx() as g:
        g.add_reads((buf, offs))
        g.seed_uncorrected()
        g.build_seqset()
        ss = g.export_seqset()
        vs, vh = g.export_varbit(0), g.export_varbit(1)
    _check_members(name, ss["n"], ss["fixed"], vs["elements"], vh["elements"],
                   {"bits_per_value": vs["bits_per_value"], "element_count": ss["n"], "max_value": vs["max_value"]},
                   {"bits_per_value": vh["bits_per_value"], "element_count": ss["n"], "max_value": vh["max_value"]},
                   ss["prev"], ss["subaccum"], ss["accum"])

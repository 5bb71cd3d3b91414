"""The oracle's seeding rule (next_fwd / next_rev, bs/correct_reads.cpp:195-226) against the known
answers of the reference's own correct_reads_test.cpp (tests/crt_cases.py).  CPU only."""
import pytest

from oracle import oracle as O
from tests import crt_cases as T


@pytest.mark.parametrize("case", T.CASES, ids=[c[0] for c in T.CASES])
def test_correct_reads_test_known_answers(case):
    name, kmer_seqs, reads, expected, exact = case
    solid = T.kmer_set_for(kmer_seqs)
    cr = O.correct_reads(reads, solid, T.K)
    assert list(cr["kept"]) == [1] * len(reads)
    got = set()
    for i, r in enumerate(O.corrected_list(cr)):
        assert r == reads[i]  # every k-mer is in the set: nothing to correct
        got |= T.seeds_of(r, int(cr["next_fwd"][i]), int(cr["next_rev"][i]))
    if exact:
        assert got == expected
    else:
        assert expected <= got

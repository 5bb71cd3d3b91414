"""fast_read_correct analytic cases against the oracle (pins substitution correction)."""
import pytest

from oracle import oracle as O
from tests import frc_cases as F


@pytest.mark.parametrize("mode", ["N", "X"])
@pytest.mark.parametrize("size", F.SIZES)
def test_frc_analytic(size, mode):
    ks = F.kmer_set_of(F.LONG[:size])
    n = 0
    for name, read, exp, ec in F.cases(size, mode):
        got, gc = O.fast_read_correct(read, ks, F.K, F.MAXC, F.RUN)
        assert (got, gc) == (exp, ec), (size, mode, name)
        n += 1
    assert n > 0


@pytest.mark.parametrize("mode", ["N", "X"])
def test_frc_three_errors_small(mode):
    for size in (33, 62):
        ks = F.kmer_set_of(F.LONG[:size])
        for name, read, exp, ec in F.cases(size, mode, with_three=True):
            if not name.startswith("three"):
                continue
            got, gc = O.fast_read_correct(read, ks, F.K, F.MAXC, F.RUN)
            assert (got, gc) == (exp, ec), (size, mode, name)

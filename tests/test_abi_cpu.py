"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/bgx.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def B():
    import __graft_entry__ as ge
    import biograph_b200 as B
    if not os.path.exists(B.lib_path()):
        ge.build()
    return B


def test_header_symbols_are_exported(B):
    hdr = open(os.path.join(ROOT, "include", "bgx.h")).read()
    declared = set(re.findall(r"\b(bgx_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"bgx_ctx", "bgx_options"}
    assert len(declared) >= 18
    L = B.load_library()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/bgx.h but not exported by libbgx.so"
    assert set(B.bgx.EXPORTS) == declared


def test_options_struct_layout(B):
    import ctypes as C
    o = B.Options()
    B.load_library().bgx_default_options(C.byref(o))
    assert (o.kmer_size, o.min_kmer_count, o.max_corrections, o.min_good_run) == (30, 5, 8, 2)
    assert abs(o.trim_after_portion - 0.7) < 1e-6 and o.sort_key_bits == 0 and o.count_batch_reads == 0
    assert C.sizeof(B.Options) == 32


def test_no_cpu_fallback(B):
    L = B.load_library()
    if L.bgx_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(B.BgxError, match="no CUDA device"):
        B.Bgx()


def test_product_does_not_touch_oracle():
    """The product package must never import, link or call oracle/ (parity would be void)."""
    pkg = os.path.join(ROOT, "biograph_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dp, f)


def test_pack_reads_2bit_layout(B):
    from biograph_b200 import bgx
    reads = ["ACGT" * 8 + "TTGCA", "N", "ACGTNACGT", "G" * 64]
    packed, nmask, woffs, lens = bgx.pack_reads_2bit(reads)
    assert list(lens) == [37, 1, 9, 64] and list(woffs) == [0, 2, 3, 4, 6]
    assert packed[0] == 0b00011011 and packed[8] == 0b11111001 and packed[9] == 0b00000000
    assert nmask[2] == 1 << 31 and nmask[3] == 1 << (31 - 4) and nmask[0] == 0
    assert packed[8 * 4] == 0b10101010


def test_synth_reads_are_deterministic(B):
    from biograph_b200 import synth
    g = synth.random_genome(5000, seed=1, repeat_frac=0.05)
    a = synth.simulate_reads(g, 1001, read_len=100, seed=3, frag_mean=250, frag_sd=20)
    b = synth.simulate_reads(g, 1001, read_len=100, seed=3, frag_mean=250, frag_sd=20)
    assert a.shape == (1001, 100) and np.array_equal(a, b)
    assert set(np.unique(a)) <= set(b"ACGT")
    err = synth.simulate_reads(g, 1000, read_len=100, seed=3, error_rate=0.0, paired=False)
    gs = g.tobytes().decode()
    from oracle import oracle as O
    for r in err[:50]:
        s = r.tobytes().decode()
        assert s in gs or O.revcomp(s) in gs


def test_khash_is_a_bijection_with_its_inverse(B):
    """the counting passes partition by the top bits of khash and drop them from the instance words; the
    k-mer comes back through khash_inv (host copies of the device functions, common.cuh)"""
    import random
    L = B.load_library()
    rng = random.Random(5)
    for k in (16, 21, 30, 31):
        seen = set()
        for _ in range(20000):
            x = rng.getrandbits(2 * k)
            h = L.bgx_debug_khash(x, k, 0)
            assert h < (1 << (2 * k)) and L.bgx_debug_khash(h, k, 1) == x
            seen.add(h >> (2 * k - 10))
        assert len(seen) == 1024          # every top-10-bit partition is hit
    # consecutive k-mers (a shift by one base) land in unrelated partitions
    parts = [L.bgx_debug_khash(i << 2, 30, 0) >> 52 for i in range(1, 4097)]
    assert len(set(parts)) > 200


@pytest.mark.parametrize("name,k_local,n,mem_gb,want_batches,want_part_bits", [
    ("E. coli 100x, 1 GPU", 398_406_294, 1, 183, 1, 7),
    ("chr20 30x, 1 GPU", 1_559_548_914, 1, 183, 1, 8),
    ("chr20 30x per rank, 2 GPUs", 1_559_548_914, 2, 183, 1, 8),
    ("chr20 30x per rank, 4 GPUs", 1_559_548_914, 4, 183, 1, 8),
    ("chr20 30x per rank, 8 GPUs", 1_559_548_914, 8, 183, 1, 9),
    ("GRCh38 30x over 8 GPUs", 9_377_500_000, 8, 183, 8, 9),
    ("8 x chr20 on one GPU (the N=8 parity build)", 12_476_391_312, 1, 183, 4, 8),
])
def test_count_plan(B, name, k_local, n, mem_gb, want_batches, want_part_bits):
    """hash-range batches and partitions per batch for the inputs of BASELINE.json's configs"""
    import ctypes as C
    L = B.load_library()
    batches, pb = C.c_uint64(), C.c_int32()
    L.bgx_debug_count_plan(k_local, k_local, n, mem_gb << 30, 0, 1, C.byref(batches), C.byref(pb))
    assert (batches.value, pb.value) == (want_batches, want_part_bits), name
    # count_batch_reads asks for at least reads / that many batches, rounded up to a power of two
    L.bgx_debug_count_plan(k_local, k_local, n, mem_gb << 30, 1000, 5000, C.byref(batches), C.byref(pb))
    assert batches.value == 8

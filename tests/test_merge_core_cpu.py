"""The bodies of the seqset-merge kernels (biograph_b200/csrc/merge_core.cuh), run serially on the CPU
by tests/cpp/merge_core_test.cpp in the order merge.cu launches them, against the merge oracle and the
reference's own merge output (family_lambda.bg).  The dev container has no GPU: this is how the device
logic is checked here; tests/test_merge_gpu.py runs the real kernels through the C ABI on the B200."""
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import merge as M
from oracle import oracle as O
from oracle.readmap import pack_bits
from tests import refseqset as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "merge_core_test")


def build_harness():
    src = os.path.join(ROOT, "tests", "cpp", "merge_core_test.cpp")
    hdr = os.path.join(ROOT, "biograph_b200", "csrc", "merge_core.cuh")
    cmn = os.path.join(ROOT, "biograph_b200", "csrc", "common.cuh")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(p) for p in (src, hdr, cmn)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I/usr/local/cuda/include", src, "-o", BIN])
    return BIN


def part_tables(entries, nsplits=1):
    """seqset tables of a sorted prefix-free entry list (oracle: any chunking decodes to the same entries)"""
    tb = M.merge_tables(entries, nsplits)
    return {"n": tb["n"], "sizes": tb["sizes"], "prev": np.stack([pack_bits(tb["prev"][b]) for b in range(4)])}


def run_harness(tmp_path, parts, old_bits, nsplits):
    exe = build_harness()
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(np.array([len(parts), nsplits], dtype="<u8").tobytes())
        for p, ob in zip(parts, old_bits):
            n = p["n"]
            f.write(np.array([n], dtype="<u8").tobytes())
            sz = np.ascontiguousarray(p["sizes"], dtype="<u2").tobytes()
            f.write(sz + b"\0" * ((-len(sz)) % 8))
            for b in range(4):
                f.write(np.ascontiguousarray(p["prev"][b], dtype="<u8").tobytes())
            f.write(np.ascontiguousarray(ob, dtype="<u8").tobytes())
    r = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    raw = open(fout, "rb").read()
    off = 0

    def take(dtype, count, pad=False):
        nonlocal off
        nb = count * np.dtype(dtype).itemsize
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += (nb + 7) // 8 * 8 if pad else nb
        return a
    n = int(take("<u8", 1)[0])
    words = (n + 63) // 64
    out = {"n": n, "sizes": take("<u2", n, True), "shared": take("<u2", n, True),
           "prev": take("<u8", 4 * words).reshape(4, words), "missing": int(take("<u8", 1)[0])}
    out["mergemap"] = [take("<u8", words) for _ in parts]
    out["migrated"] = [take("<u8", words) for _ in parts]
    flat = raw[off:].split(b"\n")[:-1]
    out["flat"], o = [], 0
    for p in parts:
        out["flat"].append(flat[o:o + p["n"]])
        o += p["n"]
    return out


def check_against_oracle(tmp_path, part_entries, nsplits, rng):
    parts = [part_tables(e, rng.choice([1, 3, 100000])) for e in part_entries]
    old01 = [np.array([rng.random() < 0.5 for _ in e], dtype=np.uint8) for e in part_entries]
    out = run_harness(tmp_path, parts, [pack_bits(o) for o in old01], nsplits)
    assert out["missing"] == 0
    for got, want in zip(out["flat"], part_entries):
        assert got == want                                        # seqset_flat::get
    merged, bits = M.make_mergemap(part_entries)
    tb = M.merge_tables(merged, nsplits)
    assert out["n"] == tb["n"]
    assert np.array_equal(out["sizes"], tb["sizes"]) and np.array_equal(out["shared"], tb["shared"])
    for b in range(4):
        assert np.array_equal(out["prev"][b], pack_bits(tb["prev"][b])), b
    for p in range(len(parts)):
        assert np.array_equal(out["mergemap"][p], pack_bits(bits[p]))
        assert np.array_equal(out["migrated"][p], pack_bits(M.migrate_source_bits(old01[p], bits[p])))
    return out


def seqset_entries(reads):
    return [e.encode() for e in O.entries_closed_form_py(reads)]


def test_harness_builds():
    assert os.path.exists(build_harness())


@pytest.mark.parametrize("nsplits", [1, 7, 100000])
def test_reference_merge_cases(tmp_path, nsplits):
    rng = random.Random(nsplits)
    # seqset_merger_test.cpp:124-133, make_mergemap_test.cpp:123-137
    for case in ([[O.tseq(x) for x in ("abc", "bcd", "cde", "cdf", "dfg")]],                       # seqset_flat_test.cpp:14-45
                 [[O.tseq("abc"), O.tseq("de")]],
                 [[O.tseq("abc"), O.tseq("cde")], [O.tseq("abc"), O.tseq("efg")]],
                 [[O.tseq("ab"), O.tseq("bc"), O.tseq("cd"), O.tseq("be")], [O.tseq("AB"), O.tseq("BC"), O.tseq("CD"), O.tseq("BE")]]):
        check_against_oracle(tmp_path, [seqset_entries(r) for r in case], nsplits, rng)


@pytest.mark.parametrize("seed", range(8))
def test_random_parts(tmp_path, seed):
    rng = random.Random(100 + seed)
    lo, hi = ((5, 20) if seed < 4 else (20, 140))   # short: many prefix runs; long: entries over several words
    reads = [["".join(rng.choice("ACGT") for _ in range(rng.randint(lo, hi))) for _ in range(rng.randint(10, 20))]
             for _ in range(rng.randint(1, 6))]
    if seed % 2:   # overlapping parts: equal entries and prefixes across inputs
        reads[-1] += reads[0][:5] + [r[: max(3, len(r) // 2)] for r in reads[0][5:8]]
    check_against_oracle(tmp_path, [seqset_entries(r) for r in reads], rng.choice([1, 2, 5, 50, 100000]), rng)


def test_single_base_entries(tmp_path):
    """entries of one base pop to the empty sequence, which prefixes everything"""
    rng = random.Random(5)
    check_against_oracle(tmp_path, [[b"A", b"C", b"GT", b"T"], seqset_entries(["ACC", "G"])], 3, rng)
    check_against_oracle(tmp_path, [seqset_entries(["A", "C"])], 100000, rng)
    check_against_oracle(tmp_path, [seqset_entries(["A"]), seqset_entries(["TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", "C"])], 2, rng)


def test_golden_family_lambda(tmp_path):
    """the reference's own merge: family_lambda.bg from proband + father + mother, every table, and the
    three migrated readmaps' source_to_mid bits"""
    names = ["proband_lambda", "father_lambda", "mother_lambda"]
    parts = [RS.tables(nm) for nm in names]
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_merge_readmaps.npz"))
    old = [z[f"{s}|old|bits"].view("<u8") for s in ("proband", "father", "mother")]
    out = run_harness(tmp_path, parts, old, 100000)
    fam = RS.tables("family_lambda")
    assert out["missing"] == 0 and out["n"] == fam["n"] == 103996
    assert np.array_equal(out["sizes"], fam["sizes"]) and np.array_equal(out["shared"], fam["shared"])
    for b in range(4):
        assert out["prev"][b].astype("<u8").tobytes() == RS.member("family_lambda", f"prev_{'ACGT'[b]}/bits")
    for p, s in enumerate(("proband", "father", "mother")):
        assert out["migrated"][p].astype("<u8").tobytes() == z[f"{s}|new|bits"].tobytes()
        assert int(np.unpackbits(out["mergemap"][p].view(np.uint8)).sum()) == parts[p]["n"]
    # flat entries of an input are its sorted entry sequences
    f = out["flat"][1]
    assert len(f) == parts[1]["n"] and all(a < b for a, b in zip(f, f[1:]))


@pytest.mark.parametrize("mode", ["short", "word", "long", "max", "dup", "many"])
def test_entry_shapes(tmp_path, mode):
    """entries of 1-12 bases, at the 32-base word boundaries, of the maximum length, identical inputs, and up
    to 64 inputs (the limit of one merge)"""
    for seed in range(6):
        rng = random.Random(sum(map(ord, mode)) * 100 + seed)
        nparts = rng.randint(20, 64) if mode == "many" else rng.randint(1, 6)
        base = "".join(rng.choice("ACGT") for _ in range(400))
        reads = []
        for _ in range(nparts):
            rs = []
            for _ in range(rng.randint(1, 12) if mode == "many" else rng.randint(3, 25)):
                L = {"short": rng.randint(1, 12), "word": rng.choice([31, 32, 33, 63, 64, 65, 95, 96, 97]),
                     "long": rng.randint(100, 200), "max": rng.choice([254, 255, 128, 129])}.get(mode, rng.randint(5, 80))
                if rng.random() < 0.6:
                    s = rng.randint(0, len(base) - L)
                    rs.append(base[s:s + L])
                else:
                    rs.append("".join(rng.choice("ACGT" if rng.random() < 0.8 else "AC") for _ in range(L)))
            reads.append(rs)
        if mode == "dup" and nparts > 1:
            reads[1] = list(reads[0])
        check_against_oracle(tmp_path, [seqset_entries(r) for r in reads], rng.choice([1, 2, 3, 7, 64, 100000]), rng)


def test_five_hiv_seqsets(tmp_path):
    """the five reference-built HIV seqsets (datasets/hiv/biograph/*.bg: 78 k - 559 k entries of up to 250 bases, the
    inputs of the reference's legacy merge test, modules/biograph/biograph_merge_test.cpp:170-190) merged in one go:
    1.02 M input entries, heavy duplication between the samples"""
    import bisect
    names = ["ERR381524", "ERR732129", "ERR732130", "ERR732131", "ERR732132"]
    parts = [RS.tables(nm) for nm in names]
    old = [pack_bits(np.arange(p["n"]) % 3 == 0) for p in parts]
    out = run_harness(tmp_path, parts, old, 100000)
    assert out["missing"] == 0
    flats = []
    for p in parts:
        prev01 = [np.unpackbits(p["prev"][b].view(np.uint8), bitorder="little")[:p["n"]] for b in range(4)]
        flats.append(M.flat_sequences(p["fixed"], prev01, p["sizes"]))
    for got, want in zip(out["flat"], flats):
        assert got == want
    merged, bits = M.make_mergemap_sorted(flats)
    n = len(merged)
    assert out["n"] == n and 558849 < n < sum(p["n"] for p in parts)
    assert np.array_equal(out["sizes"], np.array([len(e) for e in merged], dtype=np.uint16))
    lcp = lambda a, b: next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
    step = 97
    assert [int(out["shared"][i]) for i in range(1, n, step)] == [lcp(merged[i - 1], merged[i]) for i in range(1, n, step)]
    prev = M.merge_prev_closed_form(merged)
    for b in range(4):
        assert np.array_equal(out["prev"][b], pack_bits(prev[b])), b
    for p in range(len(parts)):
        assert np.array_equal(out["mergemap"][p], pack_bits(bits[p]))
        old01 = (np.arange(parts[p]["n"]) % 3 == 0).astype(np.uint8)
        assert np.array_equal(out["migrated"][p], pack_bits(M.migrate_source_bits(old01, bits[p])))

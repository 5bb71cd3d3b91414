"""Corrupted inputs to the host-side readers (CRAM, BAM, spiral files) must end in an error message, never in a
crash: bit flips, overwritten size fields and truncations of good files, a seeded sample of what
tools/fuzz_readers.py runs at length (and under ASan / UBSan builds).  No GPU."""
import json
import os
import random
import struct
import subprocess
import zipfile

import numpy as np
import pytest

from tests import refseqset as RS
from tests.test_bam_import import bam_record, bgzf_block

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "biograph_b200", "bgx-create")
MERGE = os.path.join(ROOT, "biograph_b200", "bgx-merge")


def corrupt(raw, rng, lo=0):
    b = bytearray(raw)
    mode = rng.randrange(4)
    if mode == 0:
        for _ in range(rng.randint(1, 4)):
            b[rng.randrange(lo, len(b))] ^= 1 << rng.randrange(8)
    elif mode == 1:
        b[rng.randrange(lo, min(len(b), lo + 3000))] = rng.randrange(256)
    elif mode == 2:
        b = b[:rng.randrange(lo + 4, len(b))]
    else:
        i = rng.randrange(lo, len(b) - 8)
        b[i:i + 4] = struct.pack("<I", rng.choice([0, 0xFFFFFFFF, 0x7FFFFFFF, 1 << 30]))
    return bytes(b)


def ends_cleanly(args):
    r = subprocess.run(args, capture_output=True, timeout=120)
    assert r.returncode in (0, 1), (r.returncode, r.stderr[-300:])
    return r.returncode


def test_cram(tmp_path):
    z = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_test_cram.npz"))
    raw = z["cram"].tobytes()
    os.makedirs(tmp_path / "ref")
    (tmp_path / "ref" / "source.fasta").write_text(">F\n" + str(z["ref"]) + "\n")
    rng = random.Random(41)
    rcs = [0, 0]
    for _ in range(60):
        (tmp_path / "f.cram").write_bytes(corrupt(raw, rng, lo=26))
        rcs[ends_cleanly([EXE, "--dump-reads", "--reads", str(tmp_path / "f.cram"), "--ref", str(tmp_path / "ref"), "--out", "/x"])] += 1
    assert rcs[1] > 10   # most corruptions are noticed


def test_bam(tmp_path):
    rng = random.Random(42)
    recs = [bam_record(f"q{i // 2}", "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 120))),
                       (0x1 | 0x40) if i % 2 == 0 else (0x1 | 0x80 | 0x10)) for i in range(200)]
    text = b"@HD\tVN:1.6\n"
    raw = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", 5) + b"chr1\0" + struct.pack("<i", 1000) + b"".join(recs)
    for _ in range(60):
        b = corrupt(raw, rng, lo=4)   # the record layer: the gzip layer has a CRC of its own
        with open(tmp_path / "f.bam", "wb") as f:
            for i in range(0, len(b), 5000):
                f.write(bgzf_block(b[i:i + 5000]))
            f.write(bgzf_block(b""))
        ends_cleanly([EXE, "--dump-reads", "--reads", str(tmp_path / "f.bam"), "--out", "/x"])


def test_spiral_file(tmp_path):
    members = ["seqset.json", "part_info.json", "fixed", "entry_sizes/packed_varbit_vector.json", "entry_sizes/elements",
               "shared/packed_varbit_vector.json", "shared/elements"] + [f"prev_{b}/{m}" for b in "ACGT" for m in ("bitcount.json", "bits", "subaccum", "accum")]
    good = tmp_path / "good"
    with zipfile.ZipFile(good, "w", zipfile.ZIP_STORED) as z:
        z.writestr("file_info.json", '{"uuid":"u"}')
        for fn in members:
            z.writestr(fn, RS.member("ERR732130", fn))
    raw = good.read_bytes()
    dirs = []
    for k in "ab":
        d = tmp_path / f"{k}.bg"
        for sub in ("metadata", "coverage", "qc"):
            os.makedirs(d / sub)
        (d / "metadata" / "bg_info.json").write_text(json.dumps({"accession_id": k, "biograph_id": "id-" + k, "command_history": [],
                                                                 "samples": {k: "00"}, "version": "x"}))
        dirs.append(str(d))
    (tmp_path / "b.bg" / "seqset").write_bytes(raw)
    rng = random.Random(43)
    for _ in range(60):
        (tmp_path / "a.bg" / "seqset").write_bytes(corrupt(raw, rng, lo=len(raw) - 3000))   # central directory / end records
        ends_cleanly([MERGE, "--list-inputs", "--out", str(tmp_path / "m.bg"), "--in"] + dirs)
    (tmp_path / "a.bg" / "metadata" / "bg_info.json").write_text('{"accession_id": "a", "samples": {')
    r = subprocess.run([MERGE, "--list-inputs", "--out", str(tmp_path / "m.bg"), "--in"] + dirs, capture_output=True, text=True)
    assert r.returncode == 1 and "Could not parse biograph metadata" in r.stderr

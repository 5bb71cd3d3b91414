"""bgx-create with BAM input on the GPU: the reads of a BAM file must give the BioGraph the
same reads give as FASTQ -- unpaired against the reference's golden .bg, paired (mates joined by read name,
half of the records stored reverse-complemented on the reverse strand, secondary records in between)
against the --pair run.  The importer itself is checked without a GPU in tests/test_bam_import.py."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests import refseqset as RS
from tests.test_bam_import import bam_record, rc, write_bam

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "biograph_b200", "bgx-create")


def run(args):
    return subprocess.run([EXE] + args, capture_output=True, text=True, timeout=600)


def fastq_text(reads, tag="r"):
    return "".join(f"@{tag}{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads))


def spiral_members(bg):
    info = json.loads((bg / "metadata" / "bg_info.json").read_text())
    sha = list(info["samples"].values())[0]
    return RS.SpiralZip(bg / "seqset"), RS.SpiralZip(bg / "coverage" / f"{sha}.readmap")


def test_unpaired_bam_gives_the_golden_biograph(tmp_path, golden, golden_reads):
    recs = []
    for i, r in enumerate(golden_reads):
        if i % 3 == 0:
            recs.append(bam_record(f"r{i}", rc(r), 0x10))      # reverse strand: stored reverse-complemented
        else:
            recs.append(bam_record(f"r{i}", r, 0))
        if i % 50 == 0:
            recs.append(bam_record(f"r{i}", r[:20], 0x100))      # a secondary alignment of it: skipped
    write_bam(tmp_path / "g.bam", recs, block=60000)
    out = tmp_path / "g.bg"
    r = run(["--reads", str(tmp_path / "g.bam"), "--out", str(out)])
    assert r.returncode == 0, r.stdout + r.stderr
    st = json.loads((out / "qc" / "create_stats.json").read_text())
    assert (st["imported_reads"], st["corrected_reads"], st["corrected_bases"]) == (10000, 8444, 288464)
    z, zr = spiral_members(out)
    assert json.loads(z.read("seqset.json")) == {"num_entries": 19935}
    assert z.read("fixed") == np.asarray(golden["fixed"], dtype="<u8").tobytes()
    for b in "ACGT":
        for part in ("bits", "subaccum", "accum"):
            assert z.read(f"prev_{b}/{part}") == np.asarray(golden[f"prev_{b}_{part}"]).tobytes()
    gz = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_10000snp_readmap.npz"))
    for d in ("source_to_mid", "dest_to_mid"):
        for part in ("bits", "subaccum", "accum"):
            assert zr.read(f"read_ids/{d}/{part}") == gz[f"read_ids|{d}|{part}"].tobytes(), (d, part)
    r = run(["--reads", str(tmp_path / "missing.cram"), "--out", str(tmp_path / "o.bg")])
    assert r.returncode == 1 and "Unable to open file" in r.stderr


def test_paired_bam_equals_pair_files(tmp_path, golden_reads):
    a, b = golden_reads[0::2], golden_reads[1::2]
    (tmp_path / "a.fq").write_text(fastq_text(a, "p"))
    (tmp_path / "b.fq").write_text(fastq_text(b, "p"))
    r1 = run(["--reads", str(tmp_path / "a.fq"), "--pair", str(tmp_path / "b.fq"), "--out", str(tmp_path / "p.bg")])
    assert r1.returncode == 0, r1.stderr
    # the same pairs as BAM records: the second mate of every pair first (on the reverse strand, so stored
    # reverse-complemented), then the first mate -- the importer hands a pair on when its later record arrives,
    # later record first, so the session sees (a_i, b_i) exactly as with the two FASTQ files
    recs = []
    for i, (x, y) in enumerate(zip(a, b)):
        recs += [bam_record(f"q{i}", rc(y), 0x1 | 0x10 | 0x80), bam_record(f"q{i}", x, 0x1 | 0x40)]
        if i % 97 == 0:
            recs.append(bam_record(f"q{i}", x[:25], 0x1 | 0x800))   # a supplementary piece: skipped
    write_bam(tmp_path / "p.bam", recs, block=60000)
    r2 = run(["--reads", str(tmp_path / "p.bam"), "--out", str(tmp_path / "b.bg")])
    assert r2.returncode == 0, r2.stderr
    zs = [spiral_members(tmp_path / d) for d in ("p.bg", "b.bg")]
    for k in (0, 1):
        assert zs[0][k].namelist() == zs[1][k].namelist()
        for n in zs[0][k].namelist():
            if n not in ("file_info.json", "readmap.json"):
                assert zs[0][k].read(n) == zs[1][k].read(n), n
    rows = json.loads(zs[1][1].read("mate_loop_ptr/packed_varbit_vector.json"))["element_count"]
    assert rows == 16888

"""bgx-create (biograph_b200/cli/bgx_create.cpp): the `biograph create` flags, the reference's flag
validation, and the BioGraph directory it writes (biograph_dir layout, modules/bio_base/biograph_dir.cpp),
on the reference's golden input: plain FASTQ, gzip FASTQ, --pair and --interleaved."""
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests import refseqset as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "biograph_b200", "bgx-create")


def run(args, **kw):
    return subprocess.run([EXE] + args, capture_output=True, text=True, timeout=600, **kw)


def fastq_text(reads, tag="r"):
    return "".join(f"@{tag}{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads))


def test_flag_validation_messages(tmp_path):
    """messages of validate_param / validate_float_param and the refusals of SEQSETMain::run
    (biograph_create.cpp:337-375, 483-519); none of these needs a GPU"""
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    fq = tmp_path / "a.fq"
    fq.write_text(fastq_text(["ACGT" * 10]))
    out = str(tmp_path / "o.bg")
    for extra, msg in [(["--kmer-size", "99"], "--kmer-size must specify an integer <= 32"),
                       (["--kmer-size", "abc"], "--kmer-size must specify an integer"),
                       (["--min-kmer-count", "0"], "--min-kmer-count must specify an integer >= 1"),
                       (["--max-corrections", "33"], "max-corrections must specify an integer <= 32"),
                       (["--min-reads", "1.5"], "min-reads must specify a floating point number <= 1.000000"),
                       (["--trim-after-portion", "x"], "trim-after-portion must specify a floating point number"),
                       (["--format", "sam"], "Invalid input format 'sam'")]:
        r = run(["--reads", str(fq), "--out", out] + extra)
        assert r.returncode == 1 and msg in r.stderr, (extra, r.stderr)
    r = run(["--reads", str(fq), "--pair", str(fq), "--reads", str(fq), "--out", out])
    assert r.returncode == 1 and "there must be the same number of them as read files" in r.stderr
    os.mkdir(out)
    r = run(["--reads", str(fq), "--out", out])
    assert r.returncode == 1 and f"Refusing to overwrite '{out}'. Use --force to override." in r.stderr
    r = run(["--reads", str(tmp_path / "x.bam"), "--out", str(tmp_path / "p.bg")])
    assert r.returncode == 1


@pytest.mark.gpu
def test_create_golden_bg_directory(tmp_path, golden, golden_reads):
    fq = tmp_path / "e_coli_10000snp.fq"
    fq.write_text(fastq_text(golden_reads))
    with gzip.open(tmp_path / "e_coli_10000snp.fq.gz", "wt") as f:
        f.write(fastq_text(golden_reads))
    out = tmp_path / "golden.bg"
    r = run(["--reads", str(fq), "--ref", "/nonexistent/ref", "--out", str(out), "--id", "test_accession_id"])
    assert r.returncode == 0, r.stdout + r.stderr
    # biograph_dir layout
    for d in ("metadata", "coverage", "qc", "analysis"):
        assert (out / d).is_dir()
    info = json.loads((out / "metadata" / "bg_info.json").read_text())
    assert sorted(info) == ["accession_id", "biograph_id", "command_history", "samples", "version"]
    assert info["accession_id"] == "test_accession_id" and list(info["samples"]) == ["test_accession_id"]
    sha = info["samples"]["test_accession_id"]
    rm_path = out / "coverage" / f"{sha}.readmap"
    assert rm_path.exists() and hashlib.sha1(rm_path.read_bytes()).hexdigest() == sha   # biograph_create.cpp:826-827
    # golden/e_coli_10000snp.bg/qc/create_stats.json (normative counters)
    st = json.loads((out / "qc" / "create_stats.json").read_text())
    assert (st["command"], st["imported_reads"], st["corrected_reads"], st["corrected_bases"]) == ("create", 10000, 8444, 288464)
    assert abs(st["avg_bases_per_read"] - 34.16200852676457) < 1e-9 and abs(st["corrected_pct"] - 0.8444) < 1e-6
    assert st["uuid"] == info["biograph_id"] and [list(t)[0] for t in st["timings"]] == [
        "import", "kmerization", "read_correction", "make_seqset", "make_readmap", "metadata", "total"]
    assert (out / "qc" / "create_log.txt").stat().st_size > 0 and (out / "qc" / "kmer_quality_report.html").exists()
    # the seqset: golden members byte for byte, uuid in file_info.json
    z = RS.SpiralZip(out / "seqset")
    assert json.loads(z.read("file_info.json"))["uuid"] == info["biograph_id"]
    assert json.loads(z.read("seqset.json")) == {"num_entries": 19935}
    assert z.read("fixed") == np.asarray(golden["fixed"], dtype="<u8").tobytes()
    for b in "ACGT":
        for part in ("bits", "subaccum", "accum"):
            assert z.read(f"prev_{b}/{part}") == np.asarray(golden[f"prev_{b}_{part}"]).tobytes()
    # the readmap: golden members (unpaired)
    gz = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_10000snp_readmap.npz"))
    zr = RS.SpiralZip(rm_path)
    assert json.loads(zr.read("readmap.json")) == {"seqset_uuid": info["biograph_id"]}
    for d in ("source_to_mid", "dest_to_mid"):
        for part in ("bits", "subaccum", "accum"):
            assert zr.read(f"read_ids/{d}/{part}") == gz[f"read_ids|{d}|{part}"].tobytes(), (d, part)
    assert zr.read("is_forward/packed_data") == gz["is_forward|packed_data"].tobytes()

    # gzip input: same seqset payload
    out2 = tmp_path / "golden_gz.bg"
    r = run(["--in", str(tmp_path / "e_coli_10000snp.fq.gz"), "--out", str(out2)])
    assert r.returncode == 0, r.stdout + r.stderr
    z2 = RS.SpiralZip(out2 / "seqset")
    for n in z.namelist():
        if n != "file_info.json":
            assert z.read(n) == z2.read(n), n
    assert json.loads((out2 / "metadata" / "bg_info.json").read_text())["accession_id"] == "golden_gz"   # stem of --out

    # refusing to overwrite, then --force
    assert run(["--reads", str(fq), "--out", str(out)]).returncode == 1
    assert run(["--reads", str(fq), "--out", str(out), "--force"]).returncode == 0


@pytest.mark.gpu
def test_create_paired_inputs(tmp_path, golden_reads):
    """--pair a b and --interleaved give the same BioGraph; mates are reads 2i, 2i + 1"""
    a, b = golden_reads[0::2], golden_reads[1::2]
    (tmp_path / "a.fq").write_text(fastq_text(a, "p"))
    (tmp_path / "b.fq").write_text(fastq_text(b, "p"))
    inter = [x for pair in zip(a, b) for x in pair]
    (tmp_path / "i.fastq").write_text(fastq_text(inter, "p"))
    r1 = run(["--reads", str(tmp_path / "a.fq"), "--pair", str(tmp_path / "b.fq"), "--out", str(tmp_path / "p.bg")])
    assert r1.returncode == 0, r1.stderr
    r2 = run(["--reads", str(tmp_path / "i.fastq"), "--interleaved", "--out", str(tmp_path / "i.bg")])
    assert r2.returncode == 0, r2.stderr
    zs = []
    for d in ("p.bg", "i.bg"):
        info = json.loads((tmp_path / d / "metadata" / "bg_info.json").read_text())
        sha = list(info["samples"].values())[0]
        zs.append((RS.SpiralZip(tmp_path / d / "seqset"), RS.SpiralZip(tmp_path / d / "coverage" / f"{sha}.readmap")))
    for k in (0, 1):
        for n in zs[0][k].namelist():
            if n not in ("file_info.json", "readmap.json"):
                assert zs[0][k].read(n) == zs[1][k].read(n), n
    # the same reads in another order: the seqset is the golden one; the readmap has mate loops (4 rows per kept pair)
    assert json.loads(zs[0][0].read("seqset.json")) == {"num_entries": 19935}
    rows = json.loads(zs[0][1].read("mate_loop_ptr/packed_varbit_vector.json"))["element_count"]
    assert rows == 16888


def test_biograph_dispatcher(tmp_path):
    """`biograph create ...` / `biograph merge ...` (modules/biograph/main.cpp) reach the two executables"""
    exe = os.path.join(ROOT, "biograph_b200", "biograph")
    r = subprocess.run([exe, "create", "--kmer-size", "99", "--reads", "x.fq", "--out", str(tmp_path / "o.bg")], capture_output=True, text=True)
    assert r.returncode == 1 and "--kmer-size must specify an integer <= 32" in r.stderr
    r = subprocess.run([exe, "merge", "--out", str(tmp_path / "m.bg")], capture_output=True, text=True)
    assert r.returncode == 1 and "the option '--in' is required but missing" in r.stderr
    r = subprocess.run([exe, "variants"], capture_output=True, text=True)
    assert r.returncode == 1 and "not part of the B200 seqset path" in r.stderr
    assert subprocess.run([exe, "help"], capture_output=True, text=True).returncode == 0


"""bgx-create's host FASTQ record parser (paired files, filters, the read dump: `next_record` in
biograph_b200/cli/bgx_create.cpp) against the REFERENCE'S OWN fastq_reader (modules/bio_format/fastq.cpp through its
file_reader, compiled into oracle/_ref) on well-formed, odd and malformed files: the same reads, or the same error
text with the same line number.  CPU only (--dump-reads needs no GPU)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from oracle import ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "biograph_b200", "bgx-create")
pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref.so not built (no reference checkout)")

GOOD = "@r0\nACGTACGTAC\n+\nIIIIIIIIII\n@r1 comment\nGGGTTTAACC\n+r1\nIIIIIIIIII\n"
CASES = {
    "good": GOOD,
    "crlf": GOOD.replace("\n", "\r\n"),
    "blank_lines_between": "@r0\nACGT\n+\nIIII\n\n\n@r1\nGGGA\n+\nIIII\n",
    "leading_blank": "\n@r0\nACGT\n+\nIIII\n",
    "trailing_blanks": "@r0\nACGT\n+\nIIII\n\n\n",
    "no_final_newline": "@r0\nACGT\n+\nIIII",
    "missing_at": "r0\nACGT\n+\nIIII\n",
    "short_id": "@\nACGT\n+\nIIII\n",
    "empty_seq": "@r0\n\n+\n\n",
    "bad_char": "@r0\nACGX\n+\nIIII\n",
    "lowercase": "@r0\nacgt\n+\nIIII\n",
    "n_bases": "@r0\nACNNGT\n+\nIIIIII\n",
    "missing_plus": "@r0\nACGT\nIIII\nIIII\n",
    "empty_plus": "@r0\nACGT\n\nIIII\n",
    "qual_short": "@r0\nACGT\n+\nIII\n",
    "qual_long": "@r0\nACGT\n+\nIIIII\n",
    "qual_odd_chars": "@r0\nACGT\n+\nII I\n",          # qualities are not kept on import: only the length counts
    "eof_after_id": "@r0\n",
    "eof_after_seq": "@r0\nACGT\n",
    "eof_after_plus": "@r0\nACGT\n+\n",
    "second_record_bad": GOOD + "@r2\nACGT\n+\nIII\n",
    "partial_id_line": GOOD + "@r2",
    "at_in_quality": "@r0\nACGT\n+\n@@@@\n@r1\nGGCC\n+\n@III\n",
    "empty_file": "",
    "only_blank_lines": "\n\n\n",
}


def bgx(path, extra=()):
    r = subprocess.run([EXE, "--dump-reads", "--out", "/nonexistent/x.bg", "--reads", str(path), *extra], capture_output=True, text=True, timeout=120)
    if r.returncode != 0:
        return None, r.stderr.strip().splitlines()[-1]
    return r.stdout.splitlines()[:-1], ""


def same_as_reference(path, gz=False):
    plain = path
    if gz:   # the reference reads .gz through its zip_reader into the same fastq_reader: compare on the inflated text
        plain = str(path) + ".plain"
        open(plain, "wb").write(gzip.open(path, "rb").read())
    want_reads, want_err = R.read_fastq(plain)
    got_reads, got_err = bgx(path)
    if want_err:
        assert got_reads is None and got_err == want_err, (got_err, want_err)
    else:
        assert got_err == "" and got_reads == want_reads
    return want_err


@pytest.mark.parametrize("name", sorted(CASES))
def test_fastq_case(tmp_path, name):
    p = tmp_path / "x.fastq"
    p.write_bytes(CASES[name].encode())
    err = same_as_reference(p)
    assert bool(err) == (name in {"no_final_newline", "missing_at", "short_id", "empty_seq", "bad_char", "lowercase", "missing_plus",
                                  "empty_plus", "qual_short", "qual_long", "eof_after_id", "eof_after_seq", "eof_after_plus",
                                  "second_record_bad", "partial_id_line"})


def test_fastq_random_damage(tmp_path):
    """a well-formed file of 40 records with one random edit each time (a byte changed, dropped or inserted; a line
    dropped or doubled): whatever the reference's reader makes of it, the host parser makes the same"""
    rng = np.random.default_rng(5)
    recs = []
    for i in range(40):
        n = int(rng.integers(1, 60))
        seq = "".join("ACGTN"[j] for j in rng.integers(0, 5, n))
        recs.append(f"@read{i}\n{seq}\n+\n{'F' * n}\n")
    text = "".join(recs)
    errors = 0
    for t in range(150):
        b = bytearray(text.encode())
        kind = int(rng.integers(0, 5))
        pos = int(rng.integers(0, len(b)))
        if kind == 0:
            b[pos] = int(rng.choice(list(b"ACGTN@+\n\r xI")))
        elif kind == 1:
            del b[pos]
        elif kind == 2:
            b.insert(pos, int(rng.choice(list(b"ACGTN@+\n\r x"))))
        else:
            lines = bytes(b).split(b"\n")
            li = int(rng.integers(0, len(lines) - 1))
            lines = lines[:li] + ([] if kind == 3 else [lines[li], lines[li]]) + lines[li + 1:]
            b = bytearray(b"\n".join(lines))
        p = tmp_path / f"d{t}.fastq"
        p.write_bytes(bytes(b))
        errors += bool(same_as_reference(p))
    assert 30 < errors < 150


def test_fastq_gzip_and_pairs(tmp_path):
    with gzip.open(tmp_path / "x.fq.gz", "wb") as f:
        f.write(CASES["blank_lines_between"].encode())
    same_as_reference(tmp_path / "x.fq.gz", gz=True)
    # paired files go through the same record parser, file by file: an error in the second file is reported as is
    (tmp_path / "a.fq").write_text(GOOD)
    (tmp_path / "b.fq").write_text(CASES["second_record_bad"][:-4] + "II\n")
    reads, err = bgx(tmp_path / "a.fq", ["--pair", str(tmp_path / "b.fq")])
    assert reads is None and err == R.read_fastq(str(tmp_path / "b.fq"))[1] != ""


# ---- the text path: whole chunks to the device parser, the host parser taking over where that one refuses ----------
@pytest.fixture(scope="module")
def mock_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("mockdev"))
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "cpp", "mock_tables_bgx.cpp"), "-o", os.path.join(d, "libbgx.so")])
    return d


def bgx_text_path(mock_dir, path, tmp_path, extra=()):
    """bgx-create proper (no --dump-reads): plain FASTQ goes to bgx_add_reads_fastq in whole-record chunks.  The mock
    of the C ABI implements the device parser's contract and logs every read it is handed; the run ends after the
    import stage (the mock serves no k-mers).  Returns (reads that reached the device in order, error or "")."""
    log = str(tmp_path / (os.path.basename(str(path)) + ".log"))
    env = dict(os.environ, BGX_MOCK_LOG=log, BGX_MOCK_TABLES=str(tmp_path / "nothing_served"),
               LD_LIBRARY_PATH=mock_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    out = str(tmp_path / (os.path.basename(str(path)) + ".bg"))
    r = subprocess.run([EXE, "--reads", str(path), "--out", out, "--force", *extra], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    last = r.stderr.strip().splitlines()[-1]
    reads = [l[2:] for l in open(log).read().splitlines()] if os.path.exists(log) else []
    kinds = {l[0] for l in open(log).read().splitlines()} if os.path.exists(log) else set()
    if last.startswith("mock: cannot open km_kmers.bin") or "No reads were imported" in last:
        return reads, "", kinds
    return reads, last, kinds


@pytest.mark.parametrize("name", sorted(CASES))
def test_fastq_text_path_case(mock_dir, tmp_path, name):
    p = tmp_path / "x.fastq"
    p.write_bytes(CASES[name].encode())
    want_reads, want_err = R.read_fastq(str(p))
    got_reads, got_err, kinds = bgx_text_path(mock_dir, p, tmp_path)
    assert got_err == want_err
    if not want_err:
        assert got_reads == want_reads
    if name == "good":
        assert kinds == {"F"}            # the well-formed file went through the device parser alone
    if name in ("crlf", "blank_lines_between"):
        assert "A" in kinds               # ... these needed the host parser, and give the reference's reads


def test_fastq_text_path_random_damage(mock_dir, tmp_path):
    rng = np.random.default_rng(6)
    recs = []
    for i in range(60):
        n = int(rng.integers(1, 60))
        seq = "".join("ACGTN"[j] for j in rng.integers(0, 5, n))
        recs.append(f"@read{i}\n{seq}\n+\n{'F' * n}\n")
    text = "".join(recs)
    for t in range(60):
        b = bytearray(text.encode())
        for _ in range(int(rng.integers(1, 3))):
            kind, pos = int(rng.integers(0, 4)), int(rng.integers(0, len(b)))
            if kind == 0:
                b[pos] = int(rng.choice(list(b"ACGTN@+\n\r xI")))
            elif kind == 1:
                del b[pos]
            elif kind == 2:
                b.insert(pos, int(rng.choice(list(b"ACGTN@+\n\r x"))))
            else:
                b[pos:pos] = b"\n\n"
        p = tmp_path / f"t{t}.fastq"
        p.write_bytes(bytes(b))
        want_reads, want_err = R.read_fastq(str(p))
        got_reads, got_err, _ = bgx_text_path(mock_dir, p, tmp_path)
        assert got_err == want_err, (t, got_err, want_err)
        if not want_err:
            assert got_reads == want_reads, t


def test_fastq_text_path_interleaved(mock_dir, tmp_path):
    # pairs through the text path, with blank lines that send the tail to the host parser: mates stay together
    reads = ["ACGTACGTAC" * 3 + "T" * i for i in range(1, 13)]
    text = "".join(f"@p{i // 2}/{1 + i % 2}\n{r}\n+\n{'I' * len(r)}\n" + ("\n" if i == 7 else "") for i, r in enumerate(reads))
    p = tmp_path / "i.fastq"
    p.write_text(text)
    got, err, kinds = bgx_text_path(mock_dir, p, tmp_path, ["--interleaved"])
    assert err == "" and got == reads and "A" in kinds


def test_fastq_text_path_interleaved_odd_and_even(mock_dir, tmp_path):
    # read_importer.cpp:688-693: the odd last read of an interleaved file is counted and dropped (a warning, no error);
    # the device takes whole pairs only, the left-over record reaches the host parser, which drops it
    reads = ["ACGTACGTAC" * 3 + "G" * i for i in range(1, 10)]
    fq = lambda rs: "".join(f"@p{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(rs))  # noqa: E731
    (tmp_path / "odd.fastq").write_text(fq(reads))
    got, err, kinds = bgx_text_path(mock_dir, tmp_path / "odd.fastq", tmp_path, ["--interleaved"])
    assert err == "" and got == reads[:8] and kinds == {"F"}
    (tmp_path / "even.fastq").write_text(fq(reads[:8]))
    got, err, kinds = bgx_text_path(mock_dir, tmp_path / "even.fastq", tmp_path, ["--interleaved"])
    assert err == "" and got == reads[:8] and kinds == {"F"}


def test_long_read_message(mock_dir, tmp_path):
    # read_importer_state::process (biograph_create.cpp:133-139): a read of more than 255 bases is refused with the
    # reference's words, on the text path (the device parser refuses the chunk, the host parser takes over) and on the
    # paired path
    long_read = "ACGT" * 70
    fq = f"@ok\nACGTACGT\n+\nIIIIIIII\n@long\n{long_read}\n+\n{'I' * len(long_read)}\n"
    (tmp_path / "l.fastq").write_text(fq)
    msg = "Encountered read of length 280, which is larger than the maximum read length 255"
    _, err, _ = bgx_text_path(mock_dir, tmp_path / "l.fastq", tmp_path)
    assert err == msg
    _, err, _ = bgx_text_path(mock_dir, tmp_path / "l.fastq", tmp_path, ["--pair", str(tmp_path / "l.fastq")])
    assert err == msg

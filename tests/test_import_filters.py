"""bgx-create's import stage on the host (biograph_b200/cli/bgx_create.cpp: import_inputs, RecordFilter) through
the --dump-reads test hook (no GPU): FASTQ single / --pair / --interleaved / gzip record parsing, --cut-reads
(read_batch::cut_reads, modules/build_seqset/read_importer.cpp:157-169; validate_cut_param,
modules/biograph/biograph_create.cpp:376-412) and --sample-reads (read_importer_state::process, :125-132)."""
import gzip
import os
import subprocess

import pytest

from tests.test_bam_import import bam_record, write_bam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "biograph_b200", "bgx-create")


def dump(args):
    r = subprocess.run([EXE, "--dump-reads", "--out", "/nonexistent/x.bg"] + [str(a) for a in args], capture_output=True, text=True, timeout=120)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.strip())
    lines = r.stdout.splitlines()
    tail = lines[-1].split()
    return lines[:-1], int(tail[2]), int(tail[4])


def fastq(reads, tag="r"):
    return "".join(f"@{tag}{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads))


READS = ["ACGTACGTAC" * 5 + "T" * i for i in range(1, 21)]


def test_fastq_forms(tmp_path):
    (tmp_path / "a.fq").write_text(fastq(READS[0::2]))
    (tmp_path / "b.fq").write_text(fastq(READS[1::2]))
    (tmp_path / "i.fastq").write_text(fastq(READS))
    with gzip.open(tmp_path / "s.fq.gz", "wt") as f:
        f.write(fastq(READS).replace("\n", "\r\n"))          # gzip, CRLF line ends
    assert dump(["--reads", tmp_path / "i.fastq"]) == (READS, 20, 0)
    assert dump(["--reads", tmp_path / "s.fq.gz"]) == (READS, 20, 0)
    pairs = [a + "\t" + b for a, b in zip(READS[0::2], READS[1::2])]
    assert dump(["--reads", tmp_path / "a.fq", "--pair", tmp_path / "b.fq"]) == (pairs, 20, 1)
    assert dump(["--reads", tmp_path / "i.fastq", "--interleaved"]) == (pairs, 20, 1)
    (tmp_path / "odd.fastq").write_text(fastq(READS[:5]))
    # read_importer.cpp:680-725: an odd last read of an interleaved file is counted and dropped; what one pair file
    # holds beyond the other is imported unpaired
    assert dump(["--reads", tmp_path / "odd.fastq", "--interleaved"]) == ([READS[0] + "\t" + READS[1], READS[2] + "\t" + READS[3]], 5, 1)
    a_reads = READS[0::2]
    assert dump(["--reads", tmp_path / "a.fq", "--pair", tmp_path / "odd.fastq"]) == (
        [x + "\t" + y for x, y in zip(a_reads, READS[:5])] + a_reads[5:], 15, 1)
    assert dump(["--reads", tmp_path / "odd.fastq", "--pair", tmp_path / "a.fq"]) == (
        [x + "\t" + y for x, y in zip(READS[:5], a_reads)] + a_reads[5:], 15, 1)
    (tmp_path / "cut.fq").write_text("@r0\nACGT\n+\n")
    with pytest.raises(RuntimeError, match="line 3: End of file while reading quality line"):   # fastq.cpp:98-101
        dump(["--reads", tmp_path / "cut.fq"])
    with pytest.raises(RuntimeError, match="Cannot determine the input file type"):
        dump(["--reads", tmp_path / "reads.txt"])


def test_cut_reads(tmp_path):
    (tmp_path / "i.fastq").write_text(fastq(READS))
    lines, n, _ = dump(["--reads", tmp_path / "i.fastq", "--cut-reads", "10-55"])
    assert n == 20 and lines == [r[10:55] for r in READS]     # substr(start, min(end, len) - start)
    lines, _, _ = dump(["--reads", tmp_path / "i.fastq", "--interleaved", "--cut-reads=0-3"])
    assert lines == [a[:3] + "\t" + b[:3] for a, b in zip(READS[0::2], READS[1::2])]
    write_bam(tmp_path / "t.bam", [bam_record(f"r{i}", r, 0) for i, r in enumerate(READS)])
    lines, _, _ = dump(["--reads", tmp_path / "t.bam", "--cut-reads", "50-60"])
    assert lines == [r[50:60] for r in READS]
    for value, msg in [("10", "cut-reads must specify a range separated by a dash"),
                       ("x-10", "cut-reads must specify a numerical range; couldn't parse x as a number"),
                       ("10-y", "cut-reads must specify a numerical range; couldn't parse y as a number"),
                       ("10-10", "cut-reads must specify a nonzero range; 10 must be less than 10")]:
        with pytest.raises(RuntimeError, match=msg):
            dump(["--reads", tmp_path / "i.fastq", "--cut-reads", value])
    with pytest.raises(RuntimeError, match="this_end > start"):   # a read shorter than the start: CHECK_GT in the reference
        dump(["--reads", tmp_path / "i.fastq", "--cut-reads", "60-80"])


def test_sample_reads(tmp_path):
    (tmp_path / "i.fastq").write_text(fastq(READS))
    # the accumulator of read_importer_state::process: += p per record, a record is kept when it passes 1
    def expect(p, n):
        acc, keep = 0.5, []                                      # m_sample_accum starts at 0.5 (:191)
        for i in range(n):
            acc += float(__import__("numpy").float32(p))       # the flag is parsed as a float (validate_float_param)
            if acc > 1:
                acc -= 1
                keep.append(i)
        return keep
    lines, n, _ = dump(["--reads", tmp_path / "i.fastq", "--sample-reads", "0.25"])
    assert n == 20 and lines == [READS[i] for i in expect(0.25, 20)] and len(lines) == 5
    lines, n, paired = dump(["--reads", tmp_path / "i.fastq", "--interleaved", "--sample-reads", "0.5"])
    assert (n, paired) == (20, 1) and lines == [READS[2 * i] + "\t" + READS[2 * i + 1] for i in expect(0.5, 10)]   # per record = per pair
    with pytest.raises(RuntimeError, match="sample-reads must specify a floating point number <= 1.000000"):
        dump(["--reads", tmp_path / "i.fastq", "--sample-reads", "1.5"])

"""bgx-create's BAM importer (biograph_b200/cli/bgx_create.cpp: BamReader / import_alignments), the reference's
read_importer_base::queue_bam + bam_process_line + bam1_to_unaligned_read
(modules/build_seqset/read_importer.cpp:182-266,483-575; htslib there, BGZF through zlib here).  Runs through
the --dump-reads test hook, which needs no GPU.

  * the reference's own test BAM (golden/ftest/seqset/hiv_test.bam, fixture tests/golden/hiv_test_bam.npz):
    the importer must hand back exactly the reads of the FASTQ it was aligned from, mates joined by name;
  * hand-written BAMs for the flag rules: secondary / supplementary skipped, reverse strand restored, paired
    records joined by read name across BGZF blocks, orphans last, IUPAC codes refused."""
import collections
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "biograph_b200", "bgx-create")
COMP = str.maketrans("ACGTN", "TGCAN")


def dump(path):
    r = subprocess.run([EXE, "--dump-reads", "--reads", str(path), "--out", "/nonexistent/x.bg"], capture_output=True, text=True, timeout=120)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.strip())
    lines = r.stdout.splitlines()
    tail = lines[-1].split()
    return lines[:-1], int(tail[2]), int(tail[4]), r.stderr


# ---- a minimal BAM writer (SAM spec 4.2; BGZF = gzip members with a 'BC' extra field) --------------------------
def bgzf_block(data):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def bam_record(qname, seq, flag, cigar_ops=1, tags=b""):
    code = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    packed = bytearray((len(seq) + 1) // 2)
    for i, ch in enumerate(seq):
        packed[i >> 1] |= code[ch] << (4 if i % 2 == 0 else 0)
    name = qname.encode() + b"\0"
    body = struct.pack("<iiBBHHHiiii", 0, 100, len(name), 30, 4680, cigar_ops, flag, len(seq), -1, -1, 0) + name + \
        struct.pack("<I", (len(seq) << 4) | 0) * cigar_ops + bytes(packed) + b"\x1e" * len(seq) + tags
    return struct.pack("<i", len(body)) + body


def write_bam(path, records, block=700):
    text = b"@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:100000\n"
    raw = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", 5) + b"chr1\0" + struct.pack("<i", 100000)
    raw += b"".join(records)
    with open(path, "wb") as f:   # small blocks: records straddle BGZF block boundaries
        for i in range(0, len(raw), block):
            f.write(bgzf_block(raw[i:i + block]))
        f.write(bgzf_block(b""))  # the EOF marker block


def rc(s):
    return s.translate(COMP)[::-1]


def test_built():
    assert os.path.exists(EXE), "run __graft_entry__.build()"


def test_reference_test_bam_gives_back_its_fastq(tmp_path):
    z = np.load(os.path.join(ROOT, "tests", "golden", "hiv_test_bam.npz"))
    p = tmp_path / "hiv_test.bam"
    p.write_bytes(z["bam"].tobytes())
    lines, n, paired, _ = dump(p)
    names, reads = [str(x) for x in z["names"]], [str(x) for x in z["reads"]]
    assert n == 999 and paired == 1
    by_name = collections.defaultdict(list)
    for nm, r in zip(names, reads):
        by_name[nm].append(r)
    want_pairs = collections.Counter(frozenset(v) if len(set(v)) == 2 else (v[0], v[1]) for v in by_name.values() if len(v) == 2)
    want_single = collections.Counter(v[0] for v in by_name.values() if len(v) == 1)
    got_pairs = collections.Counter()
    got_single = collections.Counter()
    for l in lines:
        f = l.split("\t")
        if len(f) == 2:
            got_pairs[frozenset(f) if f[0] != f[1] else (f[0], f[1])] += 1
        else:
            got_single[f[0]] += 1
    assert sum(got_pairs.values()) == 499 and sum(got_single.values()) == 1
    assert got_pairs == want_pairs and got_single == want_single


def test_flag_rules(tmp_path):
    rng = np.random.default_rng(3)
    seq = lambda n: "".join(rng.choice(list("ACGT"), n))
    a1, a2, b1, b2, s1, s2, orphan, dup = seq(100), seq(101), seq(37), seq(150), seq(99), seq(1), seq(64), seq(80)
    recs = [
        bam_record("pairA", a1, 0x1 | 0x40),
        bam_record("pairB", rc(b1), 0x1 | 0x10 | 0x40),           # reverse strand: stored reverse-complemented
        bam_record("sec", dup, 0x100),                            # secondary: skipped
        bam_record("pairA", rc(a2), 0x1 | 0x10 | 0x80, cigar_ops=3, tags=b"NMC\x00"),
        bam_record("supp", dup, 0x800 | 0x1),                     # supplementary: skipped
        bam_record("lonely", orphan, 0x1 | 0x80),                 # its mate never shows up
        bam_record("pairB", b2, 0x1 | 0x80),
        bam_record("u1", s1 + "N", 0x1 | 0x8),                    # paired flag, mate unmapped and absent: an orphan too
    ]
    p = tmp_path / "t.bam"
    write_bam(p, recs)
    lines, n, paired, _ = dump(p)
    assert (n, paired) == (6, 1)
    # add_paired_read(qname, this record, the cached mate): the later record first
    assert lines == [a2 + "\t" + a1, b2 + "\t" + b1, orphan, s1 + "N"]   # leftovers in name order: lonely, u1

    # unpaired file: reads in file order; a one-base read and odd lengths decode cleanly
    write_bam(p, [bam_record("r1", s1, 0), bam_record("r2", rc(s2), 0x10), bam_record("r3", a2, 0x4)], block=64)
    lines, n, paired, _ = dump(p)
    assert (lines, n, paired) == ([s1, s2, a2], 3, 0)

    # no records at all
    write_bam(p, [])
    lines, n, paired, err = dump(p)
    assert (lines, n, paired) == ([], 0, 0) and "no records present" in err


def test_refusals(tmp_path):
    p = tmp_path / "t.bam"
    write_bam(p, [bam_record("r1", "ACGTMACGT", 0)])
    with pytest.raises(RuntimeError, match="Failed conversion of dna_base, c = 'M'"):
        dump(p)
    p.write_bytes(gzip.compress(b"@HD\tVN:1.6\nthis is SAM text, not BAM\n"))
    with pytest.raises(RuntimeError, match="is not a valid BAM file"):
        dump(p)
    good = tmp_path / "g.bam"
    write_bam(good, [bam_record("r1", "ACGT" * 20, 0)], block=1 << 16)
    raw = good.read_bytes()
    p.write_bytes(raw[:len(raw) // 2])                            # truncated in the middle of a block
    with pytest.raises(RuntimeError, match="sam_read1 returned|not a valid BAM"):
        dump(p)


def test_bam_input_reaches_the_gpu_stages_like_fastq_input(tmp_path):
    """What goes to the device for a BAM is read for read, in order, what goes there for the same reads as FASTQ
    (single file, and --pair): so the BioGraph bgx-create builds from a BAM is the one it builds from the FASTQ forms,
    which tests/test_cli.py checks on the B200 against the reference's golden .bg (this check needs no GPU)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_10000snp.npz"))
    reads = [bytes(r).decode() for r in z["reads"]]

    def dump_any(args):
        r = subprocess.run([EXE, "--dump-reads", "--out", "/nonexistent/x.bg"] + [str(a) for a in args], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        return r.stdout.splitlines()
    fq = lambda rs, tag: "".join(f"@{tag}{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(rs))
    # unpaired: every third record on the reverse strand (stored reverse-complemented), secondary alignments in between
    recs = []
    for i, r in enumerate(reads):
        recs.append(bam_record(f"r{i}", rc(r), 0x10) if i % 3 == 0 else bam_record(f"r{i}", r, 0))
        if i % 50 == 0:
            recs.append(bam_record(f"r{i}", r[:20], 0x100))
    write_bam(tmp_path / "g.bam", recs, block=60000)
    (tmp_path / "g.fq").write_text(fq(reads, "r"))
    assert dump_any(["--reads", tmp_path / "g.bam"]) == dump_any(["--reads", tmp_path / "g.fq"])
    # paired: second mate first and on the reverse strand, supplementary pieces in between
    a, b = reads[0::2], reads[1::2]
    recs = []
    for i, (x, y) in enumerate(zip(a, b)):
        recs += [bam_record(f"q{i}", rc(y), 0x1 | 0x10 | 0x80), bam_record(f"q{i}", x, 0x1 | 0x40)]
        if i % 97 == 0:
            recs.append(bam_record(f"q{i}", x[:25], 0x1 | 0x800))
    write_bam(tmp_path / "p.bam", recs, block=60000)
    (tmp_path / "a.fq").write_text(fq(a, "p"))
    (tmp_path / "b.fq").write_text(fq(b, "p"))
    assert dump_any(["--reads", tmp_path / "p.bam"]) == dump_any(["--reads", tmp_path / "a.fq", "--pair", tmp_path / "b.fq"])

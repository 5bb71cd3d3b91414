"""bgx-create's CRAM reader (biograph_b200/cli/cram_reader.hpp; the reference reads CRAM through htslib's
sam_read1 with CRAM_OPT_REFERENCE = <ref dir>/source.fasta, modules/build_seqset/read_importer.cpp:498-509),
through the --dump-reads test hook (no GPU).

Fixture (tests/golden/e_coli_test_cram.npz, made by tests/golden/make_cram_fixture.py): the first data container
of the reference's own test file datasets/bams/e_coli/e_coli_test.cram as a stand-alone CRAM 3.0 file (rANS
order-0 / order-1 and gzip blocks, substitution / insertion / deletion / soft-clip features, delta-coded
positions, mates linked inside the slice), the reference bases it lies on, and the same 10 000 alignments as an
independent BAM parse gives them.  The reader must reproduce those reads and pair them the way
bam_process_line does."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "biograph_b200", "bgx-create")
COMP = str.maketrans("ACGTN", "TGCAN")


def dump(args):
    r = subprocess.run([EXE, "--dump-reads", "--out", "/nonexistent/x.bg"] + [str(a) for a in args], capture_output=True, text=True, timeout=120)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.strip())
    lines = r.stdout.splitlines()
    tail = lines[-1].split()
    return lines[:-1], int(tail[2]), int(tail[4])


@pytest.fixture(scope="module")
def fx(tmp_path_factory):
    z = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_test_cram.npz"))
    d = tmp_path_factory.mktemp("cram")
    (d / "t.cram").write_bytes(z["cram"].tobytes())
    os.makedirs(d / "ref")
    seq = str(z["ref"])
    (d / "ref" / "source.fasta").write_text(">F dna:plasmid\n" + "\n".join(seq[i:i + 60] for i in range(0, len(seq), 60)) + "\n")
    return d, z


def expected_import(names, flags, seqs):
    """read_importer_base::bam_process_line + bam_output_unpaired on the fixture's alignments"""
    lines, cache, n = [], {}, 0
    for nm, fl, sq in zip(names, flags, seqs):
        if fl & 0x900:
            continue
        if fl & 0x10:
            sq = sq.translate(COMP)[::-1]
        n += 1
        if fl & 0x1:
            if nm in cache:
                lines.append(sq + "\t" + cache.pop(nm))
            else:
                cache[nm] = sq
        else:
            lines.append(sq)
    lines += [cache[k] for k in sorted(cache)]
    return lines, n


def test_first_container_of_the_reference_test_cram(fx):
    d, z = fx
    names, flags, seqs = [str(x) for x in z["names"]], [int(x) for x in z["flags"]], [str(x) for x in z["seqs"]]
    want, n = expected_import(names, flags, seqs)
    lines, got_n, paired = dump(["--reads", d / "t.cram", "--ref", d / "ref"])
    assert (got_n, paired) == (n, 1) == (10000, 1)
    assert lines == want
    assert sum("\t" in l for l in lines) == 4951 and sum("\t" not in l for l in lines) == 98   # pairs, and mates beyond the slice
    # a FASTA path works as well as a reference directory
    assert dump(["--reads", d / "t.cram", "--ref", d / "ref" / "source.fasta"])[0] == want
    # the same alignments as a BAM file import identically (so the BioGraph built from either is the same)
    from tests.test_bam_import import bam_record, write_bam
    write_bam(d / "t.bam", [bam_record(nm, sq, fl) for nm, sq, fl in zip(names, seqs, flags)], block=60000)
    assert dump(["--reads", d / "t.bam"])[0] == want


def test_reference_checks(fx, tmp_path):
    d, z = fx
    with pytest.raises(RuntimeError, match="a reference FASTA is needed"):
        dump(["--reads", d / "t.cram"])
    seq = str(z["ref"])
    os.makedirs(tmp_path / "bad")
    (tmp_path / "bad" / "source.fasta").write_text(">F\n" + seq[:5000] + ("A" if seq[5000] != "A" else "C") + seq[5001:] + "\n")
    with pytest.raises(RuntimeError, match="the reference does not match .* wrong --ref"):
        dump(["--reads", d / "t.cram", "--ref", tmp_path / "bad"])
    os.makedirs(tmp_path / "other")
    (tmp_path / "other" / "source.fasta").write_text(">Chromosome\nACGT\n")
    with pytest.raises(RuntimeError, match="reference sequence F is not in"):
        dump(["--reads", d / "t.cram", "--ref", tmp_path / "other"])


def test_refusals(fx, tmp_path):
    d, z = fx
    raw = z["cram"].tobytes()
    (tmp_path / "x.cram").write_bytes(b"CRAX" + raw[4:])
    with pytest.raises(RuntimeError, match="is not a valid CRAM file"):
        dump(["--reads", tmp_path / "x.cram", "--ref", d / "ref"])
    (tmp_path / "v2.cram").write_bytes(raw[:4] + b"\x02\x01" + raw[6:])
    with pytest.raises(RuntimeError, match="version 2.1 is not supported"):
        dump(["--reads", tmp_path / "v2.cram", "--ref", d / "ref"])
    (tmp_path / "v31.cram").write_bytes(raw[:4] + b"\x03\x01" + raw[6:])
    with pytest.raises(RuntimeError, match="version 3.1 codecs are not supported"):
        dump(["--reads", tmp_path / "v31.cram", "--ref", d / "ref"])
    (tmp_path / "cut.cram").write_bytes(raw[:len(raw) // 2])
    with pytest.raises(RuntimeError, match="CRAM: truncated"):
        dump(["--reads", tmp_path / "cut.cram", "--ref", d / "ref"])
    # --cut-reads applies to CRAM input like to any other
    lines, _, _ = dump(["--reads", d / "t.cram", "--ref", d / "ref", "--cut-reads", "0-10"])
    assert all(len(x) == 10 for l in lines for x in l.split("\t"))


# ---- a tiny CRAM 3.0 writer (raw blocks only) for the parts the reference's test file does not exercise -------
import struct  # noqa: E402
import hashlib  # noqa: E402


def itf8(v):
    v &= 0xFFFFFFFF
    if v < 0x80: return bytes([v])
    if v < 0x4000: return bytes([0x80 | (v >> 8), v & 0xFF])
    if v < 0x200000: return bytes([0xC0 | (v >> 16), (v >> 8) & 0xFF, v & 0xFF])
    if v < 0x10000000: return bytes([0xE0 | (v >> 24), (v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF])
    return bytes([0xF0 | (v >> 28), (v >> 20) & 0xFF, (v >> 12) & 0xFF, (v >> 4) & 0xFF, v & 0x0F])


def ltf8(v):
    return itf8(v) if v < 0x10000000 else bytes([0xFF]) + struct.pack(">Q", v)


def block(ctype, cid, data):
    return bytes([0, ctype]) + itf8(cid) + itf8(len(data)) + itf8(len(data)) + data + b"\0\0\0\0"


def container(ref_id, start, span, n_rec, blocks, n_blocks):
    body = b"".join(blocks)
    head = itf8(ref_id) + itf8(start) + itf8(span) + itf8(n_rec) + ltf8(0) + ltf8(0) + itf8(n_blocks) + itf8(0) + b"\0\0\0\0"
    return struct.pack("<i", len(body)) + head + body


def enc_external(cid):
    p = itf8(cid)
    return itf8(1) + itf8(len(p)) + p


def enc_byte_array_stop(cid, stop=0):
    p = bytes([stop]) + itf8(cid)
    return itf8(5) + itf8(len(p)) + p


def enc_beta(offset, nbits):
    p = itf8(offset) + itf8(nbits)
    return itf8(6) + itf8(len(p)) + p


def enc_huffman_const(value):
    p = itf8(1) + itf8(value) + itf8(1) + itf8(0)
    return itf8(3) + itf8(len(p)) + p


INT_SERIES = ["BF", "CF", "RI", "RL", "AP", "RG", "MF", "NS", "NP", "TS", "NF", "TL", "FN", "FP", "DL", "RS", "PD", "HC", "MQ"]
BYTE_SERIES = ["FC", "BA", "QS", "BS"]
ARRAY_SERIES = ["RN", "IN", "SC", "BB", "QQ"]


class CramWriter:
    """records are dicts; every data series goes to its own raw external block (or the core block for BETA)"""

    def __init__(self, sq, read_names=True, ap_delta=False, rl_beta=False):
        self.sq, self.read_names, self.ap_delta, self.rl_beta = sq, read_names, ap_delta, rl_beta
        self.ids = {k: i + 1 for i, k in enumerate(INT_SERIES + BYTE_SERIES + ARRAY_SERIES)}

    def file(self, slices):
        text = ("@HD\tVN:1.5\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in self.sq)).encode()
        hdr = block(0, 0, struct.pack("<i", len(text)) + text)
        out = b"CRAM\x03\x00" + b"test".ljust(20, b"\0") + container(0, 0, 0, 0, [hdr], 1)
        for sl in slices:
            out += self.data_container(**sl)
        return out + container(-1, 4542278, 0, 0, [block(1, 0, bytes([1, 0, 1, 0, 1, 0]))], 1)

    def data_container(self, ref_id, start, span, records, embedded=None, md5=b"\0" * 16):
        streams = {k: bytearray() for k in self.ids}
        core_bits = []
        for r in records:
            def put(k, v):
                if k == "RL" and self.rl_beta:
                    core_bits.extend(((v + 3) >> i) & 1 for i in range(8, -1, -1))   # BETA(offset 3, 9 bits)
                elif k in INT_SERIES: streams[k] += itf8(v)
                elif k in BYTE_SERIES: streams[k].append(v if isinstance(v, int) else ord(v))
                else: streams[k] += v.encode() + b"\0"
            put("BF", r["bf"]); put("CF", r.get("cf", 0))
            if ref_id == -2: put("RI", r["ri"])
            put("RL", r["rl"]); put("AP", r["ap"]); put("RG", 0)
            if self.read_names: put("RN", r["name"])
            if r.get("cf", 0) & 2:
                put("MF", 0)
                if not self.read_names: put("RN", r["name"])
                put("NS", 0); put("NP", 0); put("TS", 0)
            elif r.get("cf", 0) & 4:
                put("NF", r["nf"])
            put("TL", 0)
            if not r["bf"] & 4:
                put("FN", len(r.get("features", [])))
                last = 0
                for f in r.get("features", []):
                    code, pos = f[0], f[1]
                    put("FC", code); put("FP", pos - last); last = pos
                    if code == "B": put("BA", f[2]); put("QS", 30)
                    elif code == "X": put("BS", f[2])
                    elif code in "IS": put("IN" if code == "I" else "SC", f[2])
                    elif code == "i": put("BA", f[2])
                    elif code == "b": put("BB", f[2])
                    elif code == "q": put("QQ", f[2])
                    elif code == "Q": put("QS", 30)
                    elif code in "DNPH": put({"D": "DL", "N": "RS", "P": "PD", "H": "HC"}[code], f[2])
                put("MQ", 60)
            else:
                for ch in r["bases"]: put("BA", ch)
        pres = itf8(4) + b"RN" + bytes([int(self.read_names)]) + b"AP" + bytes([int(self.ap_delta)]) + b"RR" + bytes([1]) + b"TD" + itf8(1) + b"\0"
        series = b""
        for k, cid in self.ids.items():
            if k == "RL" and self.rl_beta: e = enc_beta(3, 9)
            elif k == "RG": e = enc_huffman_const(0)
            elif k in ARRAY_SERIES: e = enc_byte_array_stop(cid)
            else: e = enc_external(cid)
            series += k.encode() + e
        series = itf8(len(self.ids)) + series
        comp = block(1, 0, itf8(len(pres)) + pres + itf8(len(series)) + series + itf8(1) + itf8(0))
        blocks = []
        while len(core_bits) % 8: core_bits.append(0)
        core = bytes(int("".join(map(str, core_bits[i:i + 8])), 2) for i in range(0, len(core_bits), 8))
        blocks.append(block(5, 0, core))
        ext_ids = []
        for k, cid in self.ids.items():
            if k == "RG" or (k == "RL" and self.rl_beta): continue
            blocks.append(block(4, cid, bytes(streams[k])))
            ext_ids.append(cid)
        emb_id = -1
        if embedded is not None:
            emb_id = 99
            blocks.append(block(4, emb_id, embedded.encode()))
            ext_ids.append(emb_id)
        sh = itf8(ref_id) + itf8(start) + itf8(span) + itf8(len(records)) + ltf8(0) + itf8(len(blocks)) + itf8(len(ext_ids)) + \
            b"".join(itf8(i) for i in ext_ids) + itf8(emb_id) + md5
        return container(ref_id, start, span, len(records), [comp, block(2, 0, sh)] + blocks, 2 + len(blocks))


def rc(s):
    return s.translate(COMP)[::-1]


def test_features_flags_and_layouts_from_a_written_cram(tmp_path):
    rng = np.random.default_rng(11)
    ref1 = "".join(rng.choice(list("ACGT"), 300))
    ref2 = "".join(rng.choice(list("ACGT"), 200))
    os.makedirs(tmp_path / "ref")
    (tmp_path / "ref" / "source.fasta").write_text(f">chr1 first\n{ref1}\n>chr2\n{ref2.lower()}\n")
    md5 = hashlib.md5(ref1[10 - 1:10 - 1 + 200].encode()).digest()

    def sub_code(ref_base, read_base):   # default substitution matrix (no SM entry): codes in ACGTN order without the ref base
        return [b for b in "ACGTN" if b != ref_base].index(read_base)
    # slice 1: one reference, names kept, absolute positions
    r_plain = ref1[19:59]                                                     # AP 20, 40 bases, no features
    sub_at = 5
    new_base = "A" if ref1[29 + sub_at - 1] != "A" else "C"
    r_sub = ref1[29:29 + sub_at - 1] + new_base + ref1[29 + sub_at:29 + 30]     # substitution at read position 5
    r_ins = ref1[49:59] + "TTT" + ref1[59:69]                                 # insertion of TTT after 10 bases
    r_del = ref1[69:79] + ref1[84:94]                                         # deletion of 5 reference bases
    r_clip = "GGGG" + ref1[99:115] + "G"                                      # soft clip, then match, then a single-base insertion
    r_skip = ref1[119:129] + ref1[149:159]                                    # reference skip (N) of 20
    r_bases = ref1[159:164] + "ACGTA" + ref1[169:174]                         # 'b' stretch of 5 bases replacing reference bases
    r_b1 = ref1[179:183] + "N" + ref1[184:190]                                # 'B' base+quality
    recs1 = [
        dict(name="plain", bf=0, rl=40, ap=20),
        dict(name="sub", bf=0x10, rl=30, ap=30, features=[("X", sub_at, sub_code(ref1[29 + sub_at - 1], new_base))]),
        dict(name="ins", bf=0, rl=23, ap=50, features=[("I", 11, "TTT")]),
        dict(name="del", bf=0, rl=20, ap=70, features=[("D", 11, 5)]),
        dict(name="clip", bf=0, rl=21, ap=100, features=[("S", 1, "GGGG"), ("H", 5, 3), ("i", 21, "G")]),
        dict(name="skip", bf=0, rl=20, ap=120, features=[("N", 11, 20), ("P", 11, 2), ("Q", 12), ("q", 13, "III")]),
        dict(name="bases", bf=0, rl=15, ap=160, features=[("b", 6, "ACGTA")]),
        dict(name="b1", bf=0, rl=11, ap=180, features=[("B", 5, "N")]),
        dict(name="sec", bf=0x100, rl=10, ap=20),                             # secondary: skipped by the importer
        dict(name="unm", bf=0x4, rl=7, ap=0, bases="ACGTNAC"),                # unmapped: bases verbatim
    ]
    # slice 2: several references in one slice (RI per record), names NOT kept, delta positions, RL in the core
    # block (BETA): mates linked by NF share a generated name; a detached record carries its own
    m1, m2 = ref2[9:39], ref2[99:129]
    recs2 = [
        dict(ri=1, name="", bf=0x1 | 0x40, cf=0x4, nf=1, rl=30, ap=10),       # mate is two records on
        dict(ri=0, name="", bf=0, rl=12, ap=5 - 10),                          # single read on chr1 (delta from 10 to 5)
        dict(ri=1, name="", bf=0x1 | 0x80 | 0x10, cf=0, rl=30, ap=100 - 5),   # the mate, reverse strand
        dict(ri=1, name="det", bf=0x1 | 0x40, cf=0x2, rl=8, ap=150 - 100),    # detached: mate elsewhere -> an orphan
    ]
    # slice 3: embedded reference (no FASTA needed for it)
    emb = "".join(rng.choice(list("ACGT"), 60))
    recs3 = [dict(name="emb", bf=0, rl=20, ap=1010)]
    w1 = CramWriter([("chr1", 300), ("chr2", 200)])
    w2 = CramWriter([("chr1", 300), ("chr2", 200)], read_names=False, ap_delta=True, rl_beta=True)
    raw = w1.file([dict(ref_id=0, start=10, span=200, records=recs1, md5=md5)])
    eof = raw[-len(container(-1, 4542278, 0, 0, [block(1, 0, bytes([1, 0, 1, 0, 1, 0]))], 1)):]
    raw = raw[:-len(eof)] + w2.data_container(ref_id=-2, start=0, span=0, records=recs2) + \
        w1.data_container(ref_id=0, start=1000, span=60, records=recs3, embedded=emb) + eof
    (tmp_path / "w.cram").write_bytes(raw)
    lines, n, paired = dump(["--reads", tmp_path / "w.cram", "--ref", tmp_path / "ref"])
    want = [r_plain, rc(r_sub), r_ins, r_del, r_clip, r_skip, r_bases, r_b1, "ACGTNAC",
            ref1[4:16], rc(m2) + "\t" + m1, emb[10:30], ref2[149:157]]
    assert (n, paired) == (14, 1)   # 13 lines: one of them is a pair
    assert lines == want


def test_stdin(fx):
    """`--reads -` (SEQSETMain reads BAM from STDIN by default, CRAM with --format cram, biograph_create.cpp:583-606)"""
    d, z = fx
    want = dump(["--reads", d / "t.cram", "--ref", d / "ref"])[0]
    for fmt, path in (("cram", d / "t.cram"),):
        with open(path, "rb") as f:
            r = subprocess.run([EXE, "--dump-reads", "--out", "/nonexistent/x.bg", "--reads", "-", "--format", fmt, "--ref", str(d / "ref")],
                               stdin=f, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        assert r.stdout.splitlines()[:-1] == want

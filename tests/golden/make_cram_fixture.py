"""Generates tests/golden/e_coli_test_cram.npz from the reference checkout:
  cram   : the file definition, the header container, the FIRST data container (10 000 records, rANS order-0 /
           order-1 and gzip blocks) and the EOF container of datasets/bams/e_coli/e_coli_test.cram -- itself a
           valid CRAM 3.0 file;
  ref    : the part of the reference those records lie on (sequence "F" of
           datasets/reference/e_coli_k12_ASM584v1/source.fasta, first 22 000 bases; the slice's MD5 covers
           F:7+21871);
  names / flags / seqs: the first 10 000 records of the BAM twin of that file
           (datasets/bams/e_coli/e_coli_test.bam), parsed here in Python straight from the BAM format -- the
           answer an independent reader gives for the same alignments.
Run once in the dev container; nothing at test time reads /root/reference."""
import gzip
import struct
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
d = open(f"{REF}/datasets/bams/e_coli/e_coli_test.cram", "rb").read()


def itf8(b, p):
    v = b[p]
    if v < 0x80: return v, p + 1
    if v < 0xc0: return ((v & 0x3f) << 8) | b[p + 1], p + 2
    if v < 0xe0: return ((v & 0x1f) << 16) | (b[p + 1] << 8) | b[p + 2], p + 3
    if v < 0xf0: return ((v & 0x0f) << 24) | (b[p + 1] << 16) | (b[p + 2] << 8) | b[p + 3], p + 4
    return ((v & 0x0f) << 28) | (b[p + 1] << 20) | (b[p + 2] << 12) | (b[p + 3] << 4) | (b[p + 4] & 0x0f), p + 5


def ltf8_skip(b, p):
    v, n = b[p], 0
    while n < 8 and (v & (0x80 >> n)): n += 1
    return p + 1 + n


def container_end(pos):
    length = struct.unpack("<i", d[pos:pos + 4])[0]
    p = pos + 4
    for _ in range(4): _, p = itf8(d, p)
    p = ltf8_skip(d, p); p = ltf8_skip(d, p)
    _, p = itf8(d, p)
    nl, p = itf8(d, p)
    for _ in range(nl): _, p = itf8(d, p)
    return p + 4 + length


c0 = container_end(26)        # header container
c1 = container_end(c0)        # first data container
eof = d[-38:]                 # the CRAM 3.0 EOF container
assert struct.unpack("<i", eof[:4])[0] == 15
cram = d[:c1] + eof

fasta = open(f"{REF}/datasets/reference/e_coli_k12_ASM584v1/source.fasta").read().split(">")
f_seq = "".join([x for x in fasta if x.startswith("F ")][0].split("\n")[1:])[:22000]

raw = gzip.open(f"{REF}/datasets/bams/e_coli/e_coli_test.bam").read()
assert raw[:4] == b"BAM\1"
p = 4
l_text = struct.unpack("<i", raw[p:p + 4])[0]; p += 4 + l_text
n_ref = struct.unpack("<i", raw[p:p + 4])[0]; p += 4
for _ in range(n_ref):
    l_name = struct.unpack("<i", raw[p:p + 4])[0]; p += 4 + l_name + 4
names, flags, seqs = [], [], []
while len(names) < 10000:
    bs = struct.unpack("<i", raw[p:p + 4])[0]; b = raw[p + 4:p + 4 + bs]; p += 4 + bs
    l_read_name, n_cigar, flag, l_seq = b[8], struct.unpack("<H", b[12:14])[0], struct.unpack("<H", b[14:16])[0], struct.unpack("<i", b[16:20])[0]
    names.append(b[32:32 + l_read_name - 1].decode())
    so = 32 + l_read_name + 4 * n_cigar
    seqs.append("".join("=ACMGRSVTWYHKDBN"[(b[so + (i >> 1)] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq)))
    flags.append(flag)
np.savez_compressed("tests/golden/e_coli_test_cram.npz", cram=np.frombuffer(cram, dtype=np.uint8), ref=np.array(f_seq),
                    names=np.array(names), flags=np.array(flags, dtype=np.uint16), seqs=np.array(seqs))
print(len(cram), len(f_seq), len(names), sum(1 for f in flags if f & 0x900))

"""Generates tests/golden/hiv_test_bam.npz from the reference checkout: the bytes of
golden/ftest/seqset/hiv_test.bam (999 paired-end records, written by the reference's own test tooling)
and the read sequences of the FASTQ it was aligned from (golden/ftest/seqset/hiv_test.fastq).  A BAM
importer must hand back exactly those reads (reverse-strand records restored to read orientation), mates
joined by name.  Run once in the dev container; nothing at test time reads /root/reference."""
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
d = f"{REF}/golden/ftest/seqset"
bam = np.frombuffer(open(f"{d}/hiv_test.bam", "rb").read(), dtype=np.uint8)
lines = open(f"{d}/hiv_test.fastq").read().split("\n")
names = [l[1:] for l in lines[0::4] if l]
reads = [l for l in lines[1::4] if l]
assert len(names) == len(reads) == 999
np.savez_compressed("tests/golden/hiv_test_bam.npz", bam=bam, names=np.array(names), reads=np.array(reads))
print(len(bam), len(reads), names[:3])

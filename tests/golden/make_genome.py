"""Packs the reference's golden/e_coli.fasta (NC_008253.1, pure ACGT) 2-bit into
tests/golden/e_coli_genome.npz so bench.py / tests can simulate "E. coli 100x" reads on the GPU
box, where /root/reference does not exist.  Run once in the dev container."""
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
seq = "".join(l.strip() for l in open(f"{REF}/golden/e_coli.fasta") if not l.startswith(">")).upper()
assert set(seq) <= set("ACGT"), set(seq)
n = len(seq)
code = np.zeros(256, dtype=np.uint8)
for i, ch in enumerate(b"ACGT"):
    code[ch] = i
c = code[np.frombuffer(seq.encode(), dtype=np.uint8)]
c = np.concatenate([c, np.zeros((-n) % 4, dtype=np.uint8)]).reshape(-1, 4)
packed = ((c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]).astype(np.uint8)
np.savez_compressed("tests/golden/e_coli_genome.npz", packed=packed, length=np.int64(n))
print(n, packed.nbytes)

"""Generates tests/golden/e_coli_10000snp.npz from the reference checkout.

Run once in the dev container (needs /root/reference); the .npz is committed so that nothing
at test time reads /root/reference.  Contents:
  reads        uint8[10000, 35]  ASCII bases of golden/e_coli_10000snp.fq (no N, fixed length)
  fixed, entry_sizes, shared, prev_{A,C,G,T}_{bits,subaccum,accum}
               raw payload members of golden/e_coli_10000snp.bg/seqset (stored zip members,
               read by offset because the reference leaves CRC fields unset)
"""
import struct
import sys
import zipfile

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
fq = open(f"{REF}/golden/e_coli_10000snp.fq").read().split("\n")
reads = [fq[i + 1] for i in range(0, len(fq) - 1, 4)]
assert len(reads) == 10000 and all(len(r) == 35 for r in reads)
arr = np.frombuffer("".join(reads).encode(), dtype=np.uint8).reshape(10000, 35)

path = f"{REF}/golden/e_coli_10000snp.bg/seqset"
raw = open(path, "rb").read()
z = zipfile.ZipFile(path)
out = {"reads": arr}
for info in z.infolist():
    o = info.header_offset
    sig, ver, flag, comp, mt, md, crc, cs, us, nl, el = struct.unpack("<IHHHHHIIIHH", raw[o:o + 30])
    assert sig == 0x04034B50 and comp == 0
    data = raw[o + 30 + nl + el:o + 30 + nl + el + info.file_size]
    name = info.filename
    if name.endswith(".json"):
        out["json:" + name] = np.frombuffer(data, dtype=np.uint8)
    elif name in ("entry_sizes", "shared"):
        out[name] = np.frombuffer(data, dtype=np.uint8)
    else:
        out[name.replace("/", "_")] = np.frombuffer(data, dtype=np.uint64)
np.savez_compressed("tests/golden/e_coli_10000snp.npz", **out)
print({k: v.shape for k, v in out.items()})

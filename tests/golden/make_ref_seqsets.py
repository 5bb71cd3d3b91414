"""Extracts the payload members of the seqsets the REFERENCE itself built and checked in
(datasets/lambdaToyData/benchmark/*_lambda.bg, datasets/hiv/biograph/*.bg; v1.1.0 layout,
150 bp and 250 bp reads, real data) into tests/golden/ref_seqsets.npz.

Run in the build container (reads /root/reference); the tests only read the .npz.  The inputs
of these builds are not in the reference tree, but a seqset is a fixed point of its own
construction: feeding its entries back as (uncorrected) reads must reproduce every member byte
for byte -- sizes, shared, the four prev bit vectors with their bitcount index, and fixed.
"""
import glob
import os
import struct
import zipfile

import numpy as np

REF = "/root/reference"
paths = sorted(glob.glob(REF + "/datasets/lambdaToyData/benchmark/*_lambda.bg/seqset") +
               glob.glob(REF + "/datasets/hiv/biograph/*.bg/seqset"))
out = {}
names = []
for p in paths:
    name = p.split("/")[-2].replace(".bg", "")
    z = zipfile.ZipFile(p)
    d = open(p, "rb").read()

    def member(fn):  # members are stored raw; CRC fields are not valid, so read by offset
        i = z.getinfo(fn)
        off = i.header_offset
        n, e = struct.unpack("<HH", d[off + 26:off + 30])
        return d[off + 30 + n + e: off + 30 + n + e + i.file_size]

    names.append(name)
    for fn in ["seqset.json", "part_info.json", "fixed", "entry_sizes/packed_varbit_vector.json", "entry_sizes/elements",
               "shared/packed_varbit_vector.json", "shared/elements"] + \
              [f"prev_{b}/{m}" for b in "ACGT" for m in ("bitcount.json", "bits", "subaccum", "accum")]:
        out[f"{name}|{fn}"] = np.frombuffer(member(fn), dtype=np.uint8)
out["names"] = np.array(names)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_seqsets.npz")
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst), names)

"""Generates tests/golden/e_coli_10000snp_readmap.npz from the reference checkout: the payload
members of golden/e_coli_10000snp.bg/coverage/<sha1>.readmap (stored zip members, read by offset
because the reference leaves CRC fields unset).  Run once in the dev container; committed so that
nothing at test time reads /root/reference.  The golden file is the v3.1.1 layout: read_lengths as
raw uint8, mate_loop_ptr as a 32-bit packed_vector, is_forward as a 1-bit packed_vector."""
import glob
import struct
import sys
import zipfile

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
path = glob.glob(f"{REF}/golden/e_coli_10000snp.bg/coverage/*.readmap")[0]
raw = open(path, "rb").read()
out = {}
for info in zipfile.ZipFile(path).infolist():
    o = info.header_offset
    sig, ver, flag, comp, mt, md, crc, cs, us, nl, el = struct.unpack("<IHHHHHIIIHH", raw[o:o + 30])
    assert sig == 0x04034B50 and comp == 0
    data = raw[o + 30 + nl + el:o + 30 + nl + el + info.file_size]
    name = info.filename
    if name == "file_info.json":
        continue
    out[name.replace("/", "|")] = np.frombuffer(data, dtype=np.uint8)
np.savez_compressed("tests/golden/e_coli_10000snp_readmap.npz", **out)
print({k: v.shape for k, v in out.items()})

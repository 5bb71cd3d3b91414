"""Extracts the ZIP framing (every local header, the central directory, the end record) and the JSON
members of the reference's golden seqset spiral file into tests/golden/zip_framing.json.
Run in the build container (needs /root/reference); the fixture is what travels."""
import json
import struct
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/golden/e_coli_10000snp.bg/seqset"
d = open(src, "rb").read()
eocd = d.rfind(b"PK\x05\x06")
n_entries, cd_size, cd_off = struct.unpack("<HII", d[eocd + 10:eocd + 20])
members, off = [], 0
for _ in range(n_entries):
    sig, ver, flag, meth, dd, crc, cs, us, n, e = struct.unpack("<IHHHIIIIHH", d[off:off + 30])
    assert sig == 0x04034b50 and cs == us
    name = d[off + 30:off + 30 + n].decode()
    data = d[off + 30 + n + e:off + 30 + n + e + cs]
    m = {"name": name, "offset": off, "size": cs, "header": d[off:off + 30 + n + e].hex()}
    if name.endswith(".json"):
        m["text"] = data.decode()
    members.append(m)
    off += 30 + n + e + cs
assert off == cd_off
json.dump({"source": "golden/e_coli_10000snp.bg/seqset", "file_size": len(d), "members": members,
           "central_directory": d[cd_off:cd_off + cd_size].hex(), "end": d[cd_off + cd_size:].hex()},
          open(__file__.rsplit("/", 1)[0] + "/zip_framing.json", "w"), indent=0)
print(len(members), "members")

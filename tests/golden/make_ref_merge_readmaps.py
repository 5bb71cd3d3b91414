"""Generates tests/golden/ref_merge_readmaps.npz from the reference checkout: the readmaps of the
three lambda samples before and after `biograph merge` (datasets/lambdaToyData/benchmark/
{proband,father,mother}_lambda.bg/coverage/*.readmap -> family_lambda.bg/coverage/*.readmap, which
make_readmap::fast_migrate wrote, modules/bio_mapred/make_readmap.cpp:459-520).

fast_migrate re-targets read_ids/source_to_mid through the part's mergemap and copies everything
else; the generator asserts the "copies everything else" half here (every other payload member is
byte-identical between the old and the migrated file), so the fixture only keeps the old and the new
source_to_mid members.  Run once in the dev container; nothing at test time reads /root/reference."""
import json
import struct
import sys
import zipfile

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
BASE = f"{REF}/datasets/lambdaToyData/benchmark"


def members(path):
    raw = open(path, "rb").read()
    out = {}
    for info in zipfile.ZipFile(path).infolist():
        o = info.header_offset
        nl, el = struct.unpack("<HH", raw[o + 26:o + 30])
        out[info.filename] = raw[o + 30 + nl + el:o + 30 + nl + el + info.file_size]
    return out


family = json.load(open(f"{BASE}/family_lambda.bg/metadata/bg_info.json"))["samples"]
out = {}
for sample in ["proband", "father", "mother"]:
    own = json.load(open(f"{BASE}/{sample}_lambda.bg/metadata/bg_info.json"))["samples"][sample]
    old = members(f"{BASE}/{sample}_lambda.bg/coverage/{own}.readmap")
    new = members(f"{BASE}/family_lambda.bg/coverage/{family[sample]}.readmap")
    assert list(old) == list(new)
    for name in old:
        if name in ("file_info.json", "readmap.json") or name.startswith("read_ids/source_to_mid/"):
            continue
        assert old[name] == new[name], name  # copied verbatim by fast_migrate
    for which, m in (("old", old), ("new", new)):
        for f in ("bitcount.json", "bits", "subaccum", "accum"):
            out[f"{sample}|{which}|{f}"] = np.frombuffer(m[f"read_ids/source_to_mid/{f}"], dtype=np.uint8)
np.savez_compressed("tests/golden/ref_merge_readmaps.npz", **out)
print({k: v.shape for k, v in out.items()})

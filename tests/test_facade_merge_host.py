"""The host side of bgx_bs::seqset_merger (include/bgx_build_seqset.hpp) -- the .mergemap spiral file
(seqset_mergemap_builder, modules/bio_base/seqset_mergemap.cpp:5-20), the migrated readmap file
(make_readmap::fast_migrate, modules/bio_mapred/make_readmap.cpp:459-520: everything copied, source_to_mid and the
seqset uuid replaced), flat entries -- against a MOCK of the C ABI that returns canned tables
(tests/cpp/mock_bgx.cpp), so it runs without a GPU.  The device side of the same calls is checked on the B200 in
tests/test_zz_merge_gpu.py / tests/test_zz_cli_merge.py."""
import json
import os
import subprocess
import zipfile

import numpy as np
import pytest

from tests import refseqset as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEMBERS = ["seqset.json", "part_info.json", "fixed", "entry_sizes/packed_varbit_vector.json", "entry_sizes/elements",
           "shared/packed_varbit_vector.json", "shared/elements"] + [f"prev_{b}/{m}" for b in "ACGT" for m in ("bitcount.json", "bits", "subaccum", "accum")]


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("mock")
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", inc, "-shared", "-fPIC", os.path.join(ROOT, "tests", "cpp", "mock_bgx.cpp"),
                           "-o", str(d / "libbgx.so")])
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", inc, os.path.join(ROOT, "tests", "cpp", "facade_merge_host_test.cpp"),
                           "-o", str(d / "fmh"), "-L", str(d), "-lbgx", f"-Wl,-rpath,{d}"])
    return str(d / "fmh")


def bitcount(bits01):
    n = len(bits01)
    words = np.packbits(np.concatenate([bits01, np.zeros((-n) % 64, np.uint8)]), bitorder="little").view("<u8")
    pc = np.array([bin(int(w)).count("1") for w in words], dtype=np.uint64)
    sub_n, acc_n = (n + 511) // 512, (n + 1 + 511) // 512
    sub = np.zeros(sub_n, dtype=np.uint64)
    for g in range(sub_n):
        s = 0
        for j in range(8):
            s <<= 8
            if g * 8 + j < len(words):
                s |= int(pc[g * 8 + j])
        sub[g] = s
    cum = np.concatenate([[0], np.cumsum(pc)]).astype(np.uint64)
    acc = np.array([cum[min(g * 8, len(words))] for g in range(acc_n)], dtype=np.uint64)
    return words, sub, acc


def test_mergemap_and_migrated_readmap_files(exe, tmp_path):
    paths = []
    for k, name in enumerate(("father_lambda", "mother_lambda")):
        p = tmp_path / f"s{k}"
        with zipfile.ZipFile(p, "w", zipfile.ZIP_STORED) as z:
            z.writestr("file_info.json", json.dumps({"uuid": f"uuid-{k}"}, separators=(",", ":")))
            for fn in MEMBERS:
                z.writestr(fn, RS.member(name, fn))
        paths.append(str(p))
    n_old = RS.tables("father_lambda")["n"]
    rng = np.random.default_rng(5)
    old01 = (rng.random(n_old) < 0.3).astype(np.uint8)
    ow, osub, oacc = bitcount(old01)
    rm = tmp_path / "old.readmap"
    other = {"read_lengths/elements": rng.integers(0, 255, 333, dtype=np.uint8).tobytes(), "is_forward/packed_data": b"\x01\x02\x03\x04" * 10,
             "mate_loop_ptr/packed_varbit_vector.json": '{"bits_per_value":17,"element_count":9,"max_value":99999}'}
    with zipfile.ZipFile(rm, "w", zipfile.ZIP_STORED) as z:
        z.writestr("file_info.json", '{"uuid":"old"}')
        z.writestr("part_info.json", '{"part_type":"readmap"}')
        z.writestr("readmap.json", '{"seqset_uuid":"uuid-0"}')
        z.writestr("read_ids/source_to_mid/bitcount.json", json.dumps({"nbits": n_old}, separators=(",", ":")))
        z.writestr("read_ids/source_to_mid/bits", ow.tobytes())
        z.writestr("read_ids/source_to_mid/subaccum", osub.tobytes())
        z.writestr("read_ids/source_to_mid/accum", oacc.tobytes())
        z.writestr("read_ids/dest_to_mid/bits", b"\xff" * 64)
        for k, v in other.items():
            z.writestr(k, v)
    out = tmp_path / "out"
    os.makedirs(out)
    r = subprocess.run([exe] + paths + [str(rm), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    j = json.loads(r.stdout)
    assert j == {"need_build": True, "total": 1000, "n_bits": 1000, "n_set": 334, "flat": ["ACGT" * 4, "ACGT" * 5, "ACGT", "ACGT" * 2], "no_such": True}
    # the .mergemap files: mergemap.json + the merged_entries bitcount, in the reference's member order
    for p in (0, 1):
        z = RS.SpiralZip(out / f"{p}.mergemap")
        assert z.namelist() == ["file_info.json", "part_info.json", "mergemap.json", "merged_entries/part_info.json", "merged_entries/bitcount.json",
                                "merged_entries/bits", "merged_entries/subaccum", "merged_entries/accum"]
        assert json.loads(z.read("part_info.json"))["part_type"] == "mergemap"
        assert json.loads(z.read("mergemap.json")) == {"merged_seqset_uuid": "merged-uuid", "orig_seqset_uuid": f"uuid-{p}"}
        assert json.loads(z.read("merged_entries/bitcount.json")) == {"nbits": 1000}
        w, sub, acc = bitcount((np.arange(1000) % (p + 2) == 0).astype(np.uint8))
        assert z.read("merged_entries/bits") == w.tobytes() and z.read("merged_entries/subaccum") == sub.tobytes()
        assert z.read("merged_entries/accum") == acc.tobytes()
        assert all(z.crc_ok(n) for n in z.namelist() if n.endswith(".json"))
    # the migrated readmap: same members in the same order; source_to_mid and the uuid replaced, the rest verbatim
    old, new = RS.SpiralZip(rm), RS.SpiralZip(out / "migrated.readmap")
    assert old.namelist() == new.namelist()
    assert json.loads(new.read("readmap.json")) == {"seqset_uuid": "merged-uuid"}
    want01 = np.zeros(1000, dtype=np.uint8)
    want01[0:1000:2] = old01[:500]
    w, sub, acc = bitcount(want01)
    assert json.loads(new.read("read_ids/source_to_mid/bitcount.json")) == {"nbits": 1000}
    assert new.read("read_ids/source_to_mid/bits") == w.tobytes()
    assert new.read("read_ids/source_to_mid/subaccum") == sub.tobytes() and new.read("read_ids/source_to_mid/accum") == acc.tobytes()
    for name in old.namelist():
        if name not in ("file_info.json", "readmap.json") and not name.startswith("read_ids/source_to_mid/"):
            assert old.read(name) == new.read(name), name
    # a readmap of another seqset is refused
    with zipfile.ZipFile(tmp_path / "wrong.readmap", "w", zipfile.ZIP_STORED) as z:
        z.writestr("read_ids/source_to_mid/bitcount.json", '{"nbits":5}')
        z.writestr("read_ids/source_to_mid/bits", b"\0" * 8)
    r = subprocess.run([exe] + paths + [str(tmp_path / "wrong.readmap"), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "does not belong to" in r.stderr

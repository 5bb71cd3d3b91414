"""The facade's spiral-file READER (spiral_file_reader, seqset_file in include/bgx_build_seqset.hpp; what
bgx-merge opens its inputs with): member lookup by offset in classic and ZIP64 archives, and the decoding
of a seqset's entry sizes (packed_varbit_vector from seqset 1.1.0 on, raw uint8 before:
modules/bio_base/seqset.cpp:58-62) and prev bits.  CPU only."""
import json
import os
import subprocess
import zipfile

import numpy as np
import pytest

from tests import refseqset as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "biograph_b200")


def fnv(b):
    h = 1469598103934665603
    for x in bytes(b):
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def compile_cpp(tmp, name):
    exe = str(tmp / name)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, "-L", LIB, "-lbgx", f"-Wl,-rpath,{LIB}"])
    return exe


@pytest.fixture(scope="module")
def tools(tmp_path_factory):
    if not os.path.exists(os.path.join(LIB, "libbgx.so")):
        pytest.fail("libbgx.so is missing: run __graft_entry__.build()")
    d = tmp_path_factory.mktemp("spiral")
    return compile_cpp(d, "spiral_reader_test"), compile_cpp(d, "zipwriter_test")


SEQSET_MEMBERS = ["seqset.json", "part_info.json", "fixed", "entry_sizes/packed_varbit_vector.json", "entry_sizes/elements",
                  "shared/packed_varbit_vector.json", "shared/elements"] + \
                 [f"prev_{b}/{m}" for b in "ACGT" for m in ("bitcount.json", "bits", "subaccum", "accum")]


@pytest.mark.parametrize("name", ["father_lambda", "ERR732130"])
def test_seqset_file_decodes_a_reference_built_seqset(tools, tmp_path, name):
    """the members of a seqset the reference built (v1.1.0: varbit entry sizes of 8 bits), re-zipped"""
    path = tmp_path / "seqset"
    with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
        z.writestr("file_info.json", json.dumps({"uuid": "1234-abcd", "command_line": ["biograph", "create"]}, separators=(",", ":")))
        for fn in SEQSET_MEMBERS:
            z.writestr(fn, RS.member(name, fn))
    got = json.loads(subprocess.check_output([tools[0], "seqset", str(path)]))
    t = RS.tables(name)
    words = (t["n"] + 63) // 64
    assert got["n"] == t["n"] and got["uuid"] == "1234-abcd" and got["max_read_len"] == int(t["sizes"].max())
    assert got["sizes"] == fnv(t["sizes"].astype("<u2").tobytes())
    for b in range(4):
        assert got[f"prev{b}"] == fnv(t["prev"][b][:words].astype("<u8").tobytes())


def test_seqset_file_decodes_narrow_varbit_and_raw_sizes(tools, tmp_path):
    """entry sizes at 6 bits per value (35-base reads in the 1.1.0 layout) and as raw uint8 (the 1.0.0 layout
    of the reference's 2018 golden .bg)"""
    rng = np.random.default_rng(1)
    n = 1000
    sizes = rng.integers(1, 36, n).astype(np.uint16)
    sizes[7] = 35
    prev = [rng.integers(0, 2**63, (n + 63) // 64, dtype=np.uint64) for _ in range(4)]
    bits = 6
    el = np.zeros((n * bits + 63) // 64, dtype=np.uint64)
    for i, v in enumerate(sizes):
        pos = i * bits
        el[pos >> 6] |= np.uint64(int(v) << (pos & 63) & 0xFFFFFFFFFFFFFFFF)
        if (pos & 63) + bits > 64:
            el[(pos >> 6) + 1] |= np.uint64(int(v) >> (64 - (pos & 63)))
    for layout in ("varbit", "raw"):
        path = tmp_path / f"seqset_{layout}"
        with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
            z.writestr("file_info.json", '{"uuid":"u"}')
            z.writestr("seqset.json", json.dumps({"num_entries": n}, separators=(",", ":")))
            if layout == "varbit":
                z.writestr("entry_sizes/packed_varbit_vector.json", '{"bits_per_value":6,"element_count":1000,"max_value":35}')
                z.writestr("entry_sizes/elements", el.astype("<u8").tobytes())
            else:
                z.writestr("entry_sizes", sizes.astype(np.uint8).tobytes())
            for b, ch in enumerate("ACGT"):
                z.writestr(f"prev_{ch}/bits", prev[b].astype("<u8").tobytes())
        got = json.loads(subprocess.check_output([tools[0], "seqset", str(path)]))
        assert got["n"] == n and got["max_read_len"] == 35
        assert got["sizes"] == fnv(sizes.astype("<u2").tobytes()), layout


def test_reader_reads_what_the_writer_wrote_incl_zip64(tools, tmp_path):
    """round trip through the facade's own writer (the reference's minizip framing): ZIP64 extra fields in
    every local header, a member above 4 GiB (a hole), members whose headers start beyond 4 GiB"""
    big = (5 << 30) + 123
    man, out = tmp_path / "manifest.txt", tmp_path / "out.zip"
    man.write_text("\n".join(["J file_info.json " + b'{"uuid":"x"}'.hex(), "A fixed 40",
                              f"R big/elements {big} {b'tail-of-big'.hex()}",
                              "J after/part_info.json " + b'{"part_type":"x"}'.hex(), "A after/data 24"]) + "\n")
    subprocess.check_output([tools[1], str(out), str(man)])
    got = json.loads(subprocess.check_output([tools[0], "members", str(out)]))
    z = zipfile.ZipFile(out)
    assert [m["name"] for m in got] == [i.filename for i in z.infolist()]
    by = {m["name"]: m for m in got}
    assert by["big/elements"]["size"] == big and by["after/data"]["offset"] > (5 << 30)
    assert by["file_info.json"]["fnv"] == fnv(b'{"uuid":"x"}') and by["after/part_info.json"]["fnv"] == fnv(b'{"part_type":"x"}')
    assert by["fixed"]["fnv"] == fnv(b"\0" * 40)
    with open(out, "rb") as f:   # the data offsets are where the bytes are
        f.seek(by["big/elements"]["offset"] + big - 11)
        assert f.read(11) == b"tail-of-big"


def test_reader_errors(tools, tmp_path):
    p = tmp_path / "junk"
    p.write_bytes(b"not a zip at all, just some text that is long enough to look at")
    r = subprocess.run([tools[0], "members", str(p)], capture_output=True, text=True)
    assert r.returncode == 1 and "not a spiral file" in r.stderr
    with zipfile.ZipFile(tmp_path / "z", "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("a.json", "{}" * 100)
    r = subprocess.run([tools[0], "members", str(tmp_path / "z")], capture_output=True, text=True)
    assert r.returncode == 1 and "compressed member" in r.stderr

"""The spiral-file (ZIP64) writer of the C++ facade against the reference's own framing.

* every local header, the central directory and the end record of the reference's golden seqset file
  (tests/golden/zip_framing.json, extracted by tests/golden/make_zip_framing.py) are reproduced byte
  for byte from the member names, sizes and JSON texts alone;
* a member larger than 4 GiB and members whose headers start beyond 4 GiB (written as holes) come
  back through a ZIP64-aware reader with the right offsets and sizes -- the layout minizip's unzip
  side, which the reference reads with (modules/io/spiral_file_mmap.cpp:82-125), expects."""
import json
import os
import struct
import subprocess
import zipfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def writer(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("zipw") / "zipwriter_test")
    lib = os.path.join(ROOT, "biograph_b200")
    if not os.path.exists(os.path.join(lib, "libbgx.so")):
        pytest.fail("libbgx.so is missing: run __graft_entry__.build()")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "zipwriter_test.cpp"), "-o", exe, "-L", lib, "-lbgx",
                           f"-Wl,-rpath,{lib}"])
    return exe


def run_writer(exe, tmp_path, lines):
    man, out = tmp_path / "manifest.txt", tmp_path / "out.zip"
    man.write_text("\n".join(lines) + "\n")
    size = int(subprocess.check_output([exe, str(out), str(man)]).split()[-1])
    assert size == os.path.getsize(out)
    return str(out)


def test_framing_equals_the_reference_golden_file(writer, tmp_path):
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "zip_framing.json")))
    lines = []
    for m in fx["members"]:
        if "text" in m:
            lines.append(f"J {m['name']} {m['text'].encode().hex()}")
        else:
            lines.append(f"A {m['name']} {m['size']}")
    d = open(run_writer(writer, tmp_path, lines), "rb").read()
    assert len(d) == fx["file_size"]
    for m in fx["members"]:
        h = bytes.fromhex(m["header"])
        assert d[m["offset"]:m["offset"] + len(h)] == h, m["name"]
    cd = bytes.fromhex(fx["central_directory"])
    end = bytes.fromhex(fx["end"])
    assert d[-len(end) - len(cd):-len(end)] == cd
    assert d[-len(end):] == end


def test_members_beyond_4gib(writer, tmp_path):
    big = (5 << 30) + 123
    out = run_writer(writer, tmp_path, ["J file_info.json " + b'{"uuid":"x"}'.hex(), "A fixed 40",
                                        f"R big/elements {big} {b'tail-of-big'.hex()}",
                                        "J after/part_info.json " + b'{"part_type":"x"}'.hex(), "A after/data 24"])
    assert os.stat(out).st_blocks * 512 < (64 << 20), "the > 4 GiB member must stay a hole in this test"
    z = zipfile.ZipFile(out)
    info = {i.filename: i for i in z.infolist()}
    assert [i.filename for i in z.infolist()] == ["file_info.json", "fixed", "big/elements", "after/part_info.json", "after/data"]
    assert info["big/elements"].file_size == big == info["big/elements"].compress_size
    assert info["after/data"].header_offset > (5 << 30) and info["after/data"].file_size == 24
    assert info["fixed"].extract_version == 20 and info["big/elements"].extract_version == 45
    assert z.read("file_info.json") == b'{"uuid":"x"}' and z.read("after/part_info.json") == b'{"part_type":"x"}'
    with open(out, "rb") as f:
        i = info["big/elements"]
        f.seek(i.header_offset)
        h = f.read(30)
        assert h[:6] == b"PK\x03\x04\x2d\x00" and h[18:26] == b"\xff" * 8      # sizes live in the extra field
        n, e = struct.unpack("<HH", h[26:30])
        f.seek(i.header_offset + 30 + n)
        assert f.read(e) == struct.pack("<HHQQ", 1, 16, big, big)
        f.seek(i.header_offset + 30 + n + e + big - 11)
        assert f.read(11) == b"tail-of-big"
        f.seek(-98, 2)
        end = f.read(98)
        assert end[:4] == b"PK\x06\x06" and end[56:60] == b"PK\x06\x07" and end[76:80] == b"PK\x05\x06"

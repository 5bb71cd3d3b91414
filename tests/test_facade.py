"""The C++ host façade (include/bgx_build_seqset.hpp): reference-named stage classes over the C ABI.

CPU: the header and its test driver compile and link against libbgx.so.
GPU: tests/cpp/facade_test runs the reference's builder_test known answers through
seqset_for_reads, then the whole `biograph create` stage sequence on the reference's golden reads,
and writes a seqset spiral file whose members are compared with the golden .bg and the oracle."""
import json
import os
import subprocess
import zipfile

import numpy as np
import pytest

from oracle import oracle as O
from tests import refseqset as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "facade_test")


def build_facade_test():
    src = os.path.join(ROOT, "tests", "cpp", "facade_test.cpp")
    hdr = os.path.join(ROOT, "include", "bgx_build_seqset.hpp")
    lib = os.path.join(ROOT, "biograph_b200", "libbgx.so")
    assert os.path.exists(lib), "build libbgx.so first (python -c 'import __graft_entry__ as g; g.build()')"
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(lib)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-o", BIN,
                               "-L" + os.path.join(ROOT, "biograph_b200"), "-lbgx",
                               "-Wl,-rpath," + os.path.join(ROOT, "biograph_b200")])
    return BIN


def test_facade_compiles_and_links():
    assert os.path.exists(build_facade_test())


@pytest.mark.gpu
def test_facade_create_flow_on_golden(tmp_path, golden, golden_reads):
    exe = build_facade_test()
    reads_txt = tmp_path / "reads.txt"
    reads_txt.write_text("\n".join(golden_reads) + "\n")
    out = subprocess.run([exe, str(reads_txt), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    for l in lines[:6]:
        assert l["ok"] and l["entries"] == l["expected"]      # bs/builder_test.cpp:52-122
    c = lines[-1]
    rm = [l for l in lines if l.get("case") == "readmap"][0]
    assert (rm["rows"], rm["entries"]) == (16888, 19935)
    # golden/e_coli_10000snp.bg/qc/create_log.txt (normative counts)
    assert (c["reads"], c["kmers"], c["corrected_reads"], c["corrected_bases"], c["entries"], c["written_entries"]) == \
        (10000, 7108, 8444, 288464, 19935, 19935)
    z = RS.SpiralZip(tmp_path / "seqset")
    # the reference's CRC convention: JSON members carry theirs, array members carry 0
    assert all(z.crc_ok(n) for n in z.namelist() if n.endswith(".json"))
    assert all(z.info[n].CRC == 0 for n in z.namelist() if not n.endswith(".json"))
    names = z.namelist()
    assert names[:4] == ["file_info.json", "part_info.json", "seqset.json", "fixed"]
    assert json.loads(z.read("seqset.json")) == {"num_entries": 19935}
    assert z.read("part_info.json") == b'{"part_type":"seqset","version":{"build":"","major":1,"minor":1,"patch":0,"pre":""}}'
    assert z.read("fixed") == np.asarray(golden["fixed"], dtype="<u8").tobytes()
    for b in "ACGT":
        assert z.read(f"prev_{b}/bits") == np.asarray(golden[f"prev_{b}_bits"]).tobytes()
        assert z.read(f"prev_{b}/subaccum") == np.asarray(golden[f"prev_{b}_subaccum"]).tobytes()
        assert z.read(f"prev_{b}/accum") == np.asarray(golden[f"prev_{b}_accum"]).tobytes()
        assert z.read(f"prev_{b}/bitcount.json") == b'{"nbits":19935}'
    # the 2018 golden stores entry_sizes/shared as raw uint8; this commit's layout is a varbit part
    for part, gold in (("entry_sizes", "entry_sizes"), ("shared", "shared")):
        meta = json.loads(z.read(f"{part}/packed_varbit_vector.json"))
        vals = RS.varbit_decode(z.read(f"{part}/elements"), meta["bits_per_value"], 19935)
        assert np.array_equal(vals, np.asarray(golden[gold]).astype(np.uint16))
        el, bits = O.varbit_pack(np.asarray(golden[gold]).astype(np.uint16), meta["max_value"])
        assert bits == meta["bits_per_value"] and z.read(f"{part}/elements") == el.astype("<u8").tobytes()

    # ---- the readmap spiral file (make_readmap::do_make, unpaired) against the golden readmap -----------------
    gz = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_10000snp_readmap.npz"))
    zr = RS.SpiralZip(tmp_path / "readmap")
    assert all(zr.crc_ok(n) for n in zr.namelist() if n.endswith(".json"))
    assert zr.namelist() == [
        "file_info.json", "part_info.json", "readmap.json", "read_ids/part_info.json",
        "read_ids/source_to_mid/part_info.json", "read_ids/source_to_mid/bitcount.json", "read_ids/source_to_mid/bits",
        "read_ids/source_to_mid/subaccum", "read_ids/source_to_mid/accum",
        "read_ids/dest_to_mid/part_info.json", "read_ids/dest_to_mid/bitcount.json", "read_ids/dest_to_mid/bits",
        "read_ids/dest_to_mid/subaccum", "read_ids/dest_to_mid/accum",
        "read_lengths/part_info.json", "read_lengths/packed_varbit_vector.json", "read_lengths/elements",
        "mate_loop_ptr/part_info.json", "mate_loop_ptr/packed_varbit_vector.json", "mate_loop_ptr/elements",
        "is_forward/part_info.json", "is_forward/packed_data", "is_forward/packed_vector.json"]  # order of a reference-built v1.2.0 file
    assert zr.read("part_info.json") == b'{"part_type":"readmap","version":{"build":"","major":1,"minor":2,"patch":0,"pre":""}}'
    assert zr.read("readmap.json") == b'{"seqset_uuid":"test-uuid"}'
    for fn in ("read_ids/part_info.json", "read_ids/source_to_mid/part_info.json", "read_ids/source_to_mid/bitcount.json",
               "read_ids/dest_to_mid/part_info.json", "read_ids/dest_to_mid/bitcount.json", "is_forward/part_info.json",
               "is_forward/packed_vector.json"):
        assert zr.read(fn) == gz[fn.replace("/", "|")].tobytes(), fn
    for d in ("source_to_mid", "dest_to_mid"):
        for part in ("bits", "subaccum", "accum"):
            assert zr.read(f"read_ids/{d}/{part}") == gz[f"read_ids|{d}|{part}"].tobytes(), (d, part)
    assert zr.read("is_forward/packed_data") == gz["is_forward|packed_data"].tobytes()
    # the 2018 golden stores read_lengths raw and mate_loop_ptr as 32-bit values; this layout is varbit
    ml = json.loads(zr.read("read_lengths/packed_varbit_vector.json"))
    assert ml == {"bits_per_value": 6, "element_count": 16888, "max_value": 35}
    assert np.array_equal(RS.varbit_decode(zr.read("read_lengths/elements"), 6, 16888), gz["read_lengths"].astype(np.uint16))
    mp = json.loads(zr.read("mate_loop_ptr/packed_varbit_vector.json"))
    assert mp == {"bits_per_value": 15, "element_count": 16888, "max_value": 16888}
    want_ptr = gz["mate_loop_ptr|packed_data"].view("<u4")
    el, bits = O.varbit_pack(want_ptr.astype(np.uint16), 16888)
    assert bits == 15 and zr.read("mate_loop_ptr/elements") == el.astype("<u8").tobytes()


def build_multi_session_test(tmp_dir):
    exe = os.path.join(str(tmp_dir), "multi_session_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "multi_session_test.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "biograph_b200"), "-lbgx", "-Wl,-rpath," + os.path.join(ROOT, "biograph_b200")])
    return exe


def test_multi_session_compiles_and_links(tmp_path):
    assert os.path.exists(build_multi_session_test(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 4])
def test_multi_session_single_process(tmp_path, golden_reads, gpus):
    """bgx_bs::multi_session: ONE process, one host thread per GPU (the shape SEQSETMain needs), the sharded
    build over NCCL + peer access inside the process; tables equal to the single-GPU build"""
    import biograph_b200 as B
    if B.load_library().bgx_device_count() < gpus:
        pytest.skip(f"needs >= {gpus} GPUs (run under gpurun --gpus {gpus})")
    exe = build_multi_session_test(tmp_path)
    reads_txt = tmp_path / "reads.txt"
    reads_txt.write_text("\n".join(golden_reads) + "\n")
    out = subprocess.run([exe, str(reads_txt), str(gpus)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"] and res["entries"] == res["expected"] == 19935

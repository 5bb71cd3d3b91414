"""Drop-in check of the `.bg` seqset output: the facade's spiral-file writer (bgx_bs::builder::make_seqset, the code
bgx-create uses) writes a seqset file, and the REFERENCE'S OWN reader opens it (oracle/_ref: spiral_file_open_mmap
over the vendored minizip, part types / versions, seqset::seqset(open state), bitcount, packed_varbit_vector) and
returns the same tables; seqset_flat then walks every entry's sequence through the reference's pop_front logic.
The tables come from the CPU oracle through a mock of the C ABI (tests/cpp/mock_tables_bgx.cpp), so no GPU is needed:
what is under test is the file, not the tables.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R
from tests.test_ref_vs_oracle import reads_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref.so not built (no reference checkout)")


@pytest.fixture(scope="module")
def writer(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("facade_writer"))
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", inc, "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "cpp", "mock_tables_bgx.cpp"), "-o", os.path.join(d, "libbgx.so")])
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", inc, os.path.join(ROOT, "tests", "cpp", "facade_write_test.cpp"),
                           "-L", d, "-lbgx", "-Wl,-rpath," + d, "-lpthread", "-o", os.path.join(d, "facade_write_test")])
    return d


def write_tables(d, ss):
    n = ss["n"]
    mx = int(ss["sizes"].max())
    se, sb = O.varbit_pack(ss["sizes"], mx)
    he, hb = O.varbit_pack(ss["shared"], mx - 1)
    acc_words = None
    for b in range(4):
        sub, acc, _ = O.bitcount_finalize(ss["prev"][b], n)
        ss["prev"][b].tofile(os.path.join(d, f"prev_bits_{b}.bin"))
        sub.tofile(os.path.join(d, f"prev_sub_{b}.bin"))
        acc.tofile(os.path.join(d, f"prev_acc_{b}.bin"))
        acc_words = len(acc)
    ss["fixed"].astype(np.uint64).tofile(os.path.join(d, "fixed.bin"))
    se.tofile(os.path.join(d, "sizes_elements.bin"))
    he.tofile(os.path.join(d, "shared_elements.bin"))
    bits = lambda v: max(1, int(v).bit_length())  # noqa: E731
    with open(os.path.join(d, "meta.txt"), "w") as f:
        f.write(f"{n} {mx} {ss['prev'].shape[1]} {(n + 511) // 512} {acc_words} {bits(mx)} {mx} {bits(mx - 1)} {mx - 1}\n")


@pytest.mark.parametrize("case", ["golden", "random", "long_reads"])
def test_reference_reader_opens_the_facade_written_seqset(writer, tmp_path, golden_reads, case):
    reads = {"golden": lambda: golden_reads, "random": lambda: reads_of(8000, 5000, 100, 0.01, seed=51),
             "long_reads": lambda: reads_of(3000, 1500, 250, 0.005, seed=52)}[case]()
    solid = O.solid_set(O.count_kmers(reads, 30), 5)
    cr = O.correct_reads(reads, solid, 30)
    ss = O.seqset_closed_form((cr["seq"], cr["offs"]))
    tables = str(tmp_path / "tables")
    os.mkdir(tables)
    write_tables(tables, ss)
    path = str(tmp_path / "seqset")
    uuid = "0f0e0d0c-1111-2222-3333-444455556666"
    env = dict(os.environ, BGX_MOCK_TABLES=tables)
    out = subprocess.run([os.path.join(writer, "facade_write_test"), path, uuid], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    with R.Run(2) as r:
        got = r.open_seqset_file(path)           # the reference's reader accepts the file ...
        assert r.seqset_uuid() == uuid
        assert got["n"] == ss["n"]
        for t in ("sizes", "shared", "prev", "fixed"):   # ... and sees the tables that went in
            assert np.array_equal(got[t], ss[t]), t
        flat = r.flat()                           # every entry's sequence through the reference's own traversal
    ents = sorted(set(O.entries_closed_form_py(O.corrected_list(cr)))) if case != "golden" else None
    assert flat == sorted(flat) and len(flat) == ss["n"]
    assert [len(e) for e in flat] == ss["sizes"].tolist()
    if ents is not None:
        assert [e.decode() for e in flat] == ents


def test_reference_reader_refuses_a_damaged_file(writer, tmp_path, golden_reads):
    # the same reader is not lenient: a truncated file is an error, not garbage tables
    solid = O.solid_set(O.count_kmers(golden_reads, 30), 5)
    cr = O.correct_reads(golden_reads, solid, 30)
    ss = O.seqset_closed_form((cr["seq"], cr["offs"]))
    tables = str(tmp_path / "tables")
    os.mkdir(tables)
    write_tables(tables, ss)
    path = str(tmp_path / "seqset")
    subprocess.check_call([os.path.join(writer, "facade_write_test"), path, "u"], env=dict(os.environ, BGX_MOCK_TABLES=tables),
                          stdout=subprocess.DEVNULL)
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[:len(raw) // 2])
    with R.Run(1) as r:
        with pytest.raises(RuntimeError):
            r.open_seqset_file(path)


# ---- readmap files -----------------------------------------------------------------------------------------------
def _readmap_case(seed, is_paired):
    import bisect

    from oracle import readmap as RM
    from tests.test_oracle_readmap import _random_paired_case
    reads, kept, ents = _random_paired_case(seed, 400, 500, 45, 0.3, 0.1)

    def lookup(s):
        i = bisect.bisect_left(ents, s)
        assert ents[i].startswith(s)
        return i

    fwd = [lookup(r) if k else 0 for r, k in zip(reads, kept)]
    rc = [lookup(O.revcomp(r)) if k else 0 for r, k in zip(reads, kept)]
    lens = [len(r) for r in reads]
    if is_paired:
        t = RM.readmap_tables_paired(*RM.pair_records(fwd, rc, lens, kept), len(ents))
    else:
        k = np.asarray(kept, dtype=bool)
        t = RM.readmap_tables(np.asarray(fwd)[k], np.asarray(rc)[k], np.asarray(lens)[k], len(ents))
    return t, len(ents), max(len(e) for e in ents)


@pytest.mark.parametrize("is_paired", [False, True])
def test_reference_reader_opens_the_facade_written_readmap(writer, tmp_path, is_paired):
    from oracle import readmap as RM
    t, n_entries, max_len = _readmap_case(61 + is_paired, is_paired)
    d = str(tmp_path / "tables")
    os.mkdir(d)
    n = t["n_rows"]
    open(os.path.join(d, "meta.txt"), "w").write(f"{n_entries} {max_len} 0 0 0 1 1 1 1\n")  # bgx_seqset_layout: entry count
    open(os.path.join(d, "rm_meta.txt"), "w").write(f"{n}\n")
    t["read_lengths"].astype(np.uint16).tofile(os.path.join(d, "rm_lens.bin"))
    t["mate_loop_ptr"].astype(np.uint64).tofile(os.path.join(d, "rm_ptr.bin"))
    RM.pack_bits(t["is_forward"]).tofile(os.path.join(d, "rm_fwd.bin"))
    for name, bits in (("src", t["source_to_mid"]), ("dst", t["dest_to_mid"])):
        words = RM.pack_bits(bits)
        sub, acc, _ = O.bitcount_finalize(words, len(bits))
        for i, a in enumerate((words, sub, acc)):
            a.tofile(os.path.join(d, f"rm_{name}_{i}.bin"))
    path = str(tmp_path / "x.readmap")
    uuid = "aaaaaaaa-bbbb-cccc-dddd-eeeeeeeeeeee"
    out = subprocess.run([os.path.join(writer, "facade_write_test"), "readmap", path, uuid, "1" if is_paired else "0", str(max_len)],
                         env=dict(os.environ, BGX_MOCK_TABLES=d), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = R.read_readmap_file(path)   # readmap::open_anonymous_readmap + its accessors, row by row
    assert got["seqset_uuid"] == uuid
    assert np.array_equal(got["entry_id"], t["entry_id"])
    assert np.array_equal(got["read_lengths"], t["read_lengths"].astype(np.int32))
    assert np.array_equal(got["is_forward"], t["is_forward"])
    assert np.array_equal(got["mate_loop_ptr"], t["mate_loop_ptr"])


# ---- the other direction: files the REFERENCE wrote, read by the facade -------------------------------------------
def _fnv(b):
    h = 1469598103934665603
    for x in bytes(b):
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


@pytest.fixture(scope="module")
def reader_exe(tmp_path_factory):
    lib = os.path.join(ROOT, "biograph_b200")
    if not os.path.exists(os.path.join(lib, "libbgx.so")):
        pytest.fail("libbgx.so is missing: run __graft_entry__.build()")
    exe = str(tmp_path_factory.mktemp("rd") / "spiral_reader_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "spiral_reader_test.cpp"), "-o", exe, "-L", lib, "-lbgx", f"-Wl,-rpath,{lib}"])
    return exe


def test_facade_reader_opens_a_reference_written_seqset(reader_exe, writer, tmp_path):
    """spiral_file_create_mmap (minizip, zip64 = 1, CRC fields unset) wrote it; spiral_file_reader / seqset_file -- what
    bgx-merge opens its inputs with -- read it; and the facade's own writer frames the same members the same way"""
    import json
    reads = reads_of(6000, 4000, 100, 0.01, seed=71)
    path = str(tmp_path / "ref_seqset")
    with R.Run(2) as r:
        solid = O.solid_set(O.count_kmers(reads, 30), 5)
        cr = O.correct_reads(reads, solid, 30)
        r.seed(O.corrected_list(cr))
        t = r.make_seqset(path)          # written by the reference, reopened by the reference
        uuid = r.seqset_uuid()
    got = json.loads(subprocess.check_output([reader_exe, "seqset", path]))
    words = (t["n"] + 63) // 64
    assert got["n"] == t["n"] and got["uuid"] == uuid and got["max_read_len"] == int(t["sizes"].max())
    assert got["sizes"] == _fnv(t["sizes"].astype("<u2").tobytes())
    for b in range(4):
        assert got[f"prev{b}"] == _fnv(t["prev"][b][:words].astype("<u8").tobytes())
    # member for member: names in the reference's order, payload sizes, and the local-header framing of the
    # facade's writer over the same tables (JSON texts differ in key order / build stamp, so compare payload members)
    ref_members = json.loads(subprocess.check_output([reader_exe, "members", path]))
    tables = str(tmp_path / "tables")
    os.mkdir(tables)
    ss = dict(t)
    write_tables(tables, ss)
    mine = str(tmp_path / "my_seqset")
    subprocess.check_call([os.path.join(writer, "facade_write_test"), mine, uuid], env=dict(os.environ, BGX_MOCK_TABLES=tables),
                          stdout=subprocess.DEVNULL)
    my_members = json.loads(subprocess.check_output([reader_exe, "members", mine]))
    assert [m["name"] for m in my_members] == [m["name"] for m in ref_members]
    for a, b_ in zip(my_members, ref_members):
        if not a["name"].endswith(".json"):
            assert (a["size"], a["fnv"]) == (b_["size"], b_["fnv"]), a["name"]
        # same distance from a member's local header to its data: same header layout incl. the ZIP64 extra field
    gaps = lambda ms, raw: [m["offset"] - raw.rfind(b"PK\x03\x04", 0, m["offset"]) - len(m["name"]) for m in ms]  # noqa: E731
    assert gaps(my_members, open(mine, "rb").read()) == gaps(ref_members, open(path, "rb").read())


# ---- a whole .bg directory written by bgx-create, opened by the reference ------------------------------------------
def _serve_create_results(d, reads, paired):
    """everything bgx-create asks the device for, computed by the oracle and written where the mock C ABI reads it"""
    import bisect

    from oracle import readmap as RM
    oc = O.count_kmers(reads, 30)
    solid = O.solid_set(oc, 5)
    cr = O.correct_reads(reads, solid, 30)
    ss = O.seqset_closed_form((cr["seq"], cr["offs"]))
    write_tables(d, ss)
    oc["kmers"].astype(np.uint64).tofile(os.path.join(d, "km_kmers.bin"))
    oc["fwd"].astype(np.uint32).tofile(os.path.join(d, "km_fwd.bin"))
    oc["rev"].astype(np.uint32).tofile(os.path.join(d, "km_rev.bin"))
    oc["flags"].astype(np.uint8).tofile(os.path.join(d, "km_flags.bin"))
    lens = np.diff(cr["offs"]).astype(np.uint16)
    lens.tofile(os.path.join(d, "cr_lens.bin"))
    open(os.path.join(d, "cr_bases.bin"), "wb").write(cr["seq"])
    cr["corrections"].astype(np.uint8).tofile(os.path.join(d, "cr_corr.bin"))
    open(os.path.join(d, "stats.json"), "w").write('{"entries":%d,"entries_round1":%d,"walk_new_records":0}' % (ss["n"], ss["n"]))
    # the readmap the device would build: entry of every kept read and of its reverse complement
    seqs = [cr["seq"][cr["offs"][i]:cr["offs"][i + 1]].decode() for i in range(len(lens))]
    ents = sorted(O.entries_closed_form_py([s for s in seqs if s]))
    assert len(ents) == ss["n"]

    def lookup(s):
        i = bisect.bisect_left(ents, s)
        assert ents[i].startswith(s)
        return i

    kept = lens > 0
    fwd = [lookup(s) if k else 0 for s, k in zip(seqs, kept)]
    rc = [lookup(O.revcomp(s)) if k else 0 for s, k in zip(seqs, kept)]
    if paired:
        t = RM.readmap_tables_paired(*RM.pair_records(fwd, rc, lens.astype(np.int64), kept), len(ents))
    else:
        t = RM.readmap_tables(np.asarray(fwd)[kept], np.asarray(rc)[kept], lens[kept].astype(np.int64), len(ents))
    open(os.path.join(d, "rm_meta.txt"), "w").write(f"{t['n_rows']}\n")
    t["read_lengths"].astype(np.uint16).tofile(os.path.join(d, "rm_lens.bin"))
    t["mate_loop_ptr"].astype(np.uint64).tofile(os.path.join(d, "rm_ptr.bin"))
    RM.pack_bits(t["is_forward"]).tofile(os.path.join(d, "rm_fwd.bin"))
    for name, bits in (("src", t["source_to_mid"]), ("dst", t["dest_to_mid"])):
        words = RM.pack_bits(bits)
        sub, acc, _ = O.bitcount_finalize(words, len(bits))
        for i, a in enumerate((words, sub, acc)):
            a.tofile(os.path.join(d, f"rm_{name}_{i}.bin"))
    return ss, cr, t, lens


@pytest.mark.parametrize("paired", [False, True])
def test_reference_opens_the_bg_directory_bgx_create_writes(writer, tmp_path, paired):
    """bgx-create end to end on the host (flags, import, stages, seqset + readmap files, sha1-named readmap,
    metadata/bg_info.json, qc/create_stats.json) with the device results served by the mock; then the reference's
    biograph_dir + seqset + readmap open the directory as any of its tools would"""
    import json
    exe = os.path.join(ROOT, "biograph_b200", "bgx-create")
    if not os.path.exists(exe):
        pytest.fail("bgx-create is missing: run __graft_entry__.build()")
    reads = reads_of(7000, 4000, 100, 0.01, seed=81)
    d = str(tmp_path / "served")
    os.mkdir(d)
    ss, cr, t, lens = _serve_create_results(d, reads, paired)
    fq = tmp_path / "reads.fastq"
    fq.write_text("".join(f"@r{i}/{1 + i % 2}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)))
    out = str(tmp_path / "sample1.bg")
    env = dict(os.environ, BGX_MOCK_TABLES=d, LD_LIBRARY_PATH=writer + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    cmd = [exe, "--reads", str(fq), "--out", out, "--id", "NA0001"] + (["--interleaved"] if paired else [])
    run = subprocess.run(cmd, env=env, capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    assert "mock-tables" in open(os.path.join(out, "qc", "create_log.txt")).read()   # the mock was the library in use
    bg = R.open_biograph(out)
    n_kept = int((lens > 0).sum())
    assert bg["accession_id"] == "NA0001" and bg["sample_accession"] == "NA0001" and bg["samples"] == 1
    assert bg["biograph_id"] == bg["seqset_uuid"] == bg["readmap_seqset_uuid"] and len(bg["biograph_id"]) == 36
    assert bg["seqset_entries"] == ss["n"] and bg["max_read_len"] == int(ss["sizes"].max())
    assert bg["readmap_rows"] == t["n_rows"] == 2 * n_kept
    assert bg["num_bases"] == int(lens.sum())
    if paired:
        both = int(np.sum((lens[0::2] > 0) & (lens[1::2] > 0)))
        assert bg["paired_reads"] == 2 * both and bg["unpaired_reads"] == n_kept - 2 * both
    else:
        assert bg["paired_reads"] == 0 and bg["unpaired_reads"] == n_kept
    assert bg["readmap_max_read_len"] == int(lens.max()) and bg["min_read_len"] == int(lens[lens > 0].min())
    # the readmap is named by its sha1 and listed under the accession id (biograph_create.cpp:827-828,798)
    import hashlib
    sha = os.path.basename(bg["readmap_path"])[:-len(".readmap")]
    assert hashlib.sha1(open(bg["readmap_path"], "rb").read()).hexdigest() == sha
    st = json.load(open(os.path.join(out, "qc", "create_stats.json")))
    assert st["imported_reads"] == len(reads) and st["corrected_reads"] == n_kept and st["corrected_bases"] == int(lens.sum())
    assert st["entries"] == ss["n"] and st["uuid"] == bg["biograph_id"]


# ---- a merged .bg directory written by bgx-merge, opened by the reference -----------------------------------------
def _bgx_create_under_mock(writer, tmp_path, name, reads, paired):
    d = str(tmp_path / (name + "_served"))
    os.mkdir(d)
    ss, cr, t, lens = _serve_create_results(d, reads, paired)
    fq = tmp_path / (name + ".fastq")
    fq.write_text("".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)))
    out = str(tmp_path / (name + ".bg"))
    env = dict(os.environ, BGX_MOCK_TABLES=d, LD_LIBRARY_PATH=writer + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    cmd = [os.path.join(ROOT, "biograph_b200", "bgx-create"), "--reads", str(fq), "--out", out, "--id", name] + (["--interleaved"] if paired else [])
    run = subprocess.run(cmd, env=env, capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    seqs = [cr["seq"][cr["offs"][i]:cr["offs"][i + 1]] for i in range(len(lens)) if lens[i]]
    return out, ss, sorted(e.encode() for e in O.entries_closed_form_py([s.decode() for s in seqs])), t


def _reference_written_bg(tmp_path, name, reads, paired):
    """an input BioGraph as the REFERENCE writes it (biograph_dir, spiral_file_create_mmap, make_readmap, save_metadata)
    from reads taken as corrected; returns what _bgx_create_under_mock returns"""
    import bisect

    from oracle import readmap as RM
    out = str(tmp_path / (name + ".bg"))
    ro = list(range(0, len(reads) + 1, 2)) if paired else None
    ss = R.write_biograph(out, name, reads, ro, paired)
    ents = sorted(e.encode() for e in O.entries_closed_form_py(reads))
    assert len(ents) == ss["n"]
    look = [e.decode() for e in ents]

    def lookup(s_):
        i = bisect.bisect_left(look, s_)
        assert look[i].startswith(s_)
        return i

    fwd = np.array([lookup(r) for r in reads])
    rc = np.array([lookup(O.revcomp(r)) for r in reads])
    lens = np.array([len(r) for r in reads])
    kept = np.ones(len(reads), dtype=bool)
    t = RM.readmap_tables_paired(*RM.pair_records(fwd, rc, lens, kept), len(ents)) if paired else RM.readmap_tables(fwd, rc, lens, len(ents))
    return out, ss, ents, t


@pytest.mark.parametrize("inputs", ["bgx-create", "reference"])
def test_reference_opens_the_bg_directory_bgx_merge_writes(writer, tmp_path, inputs):
    """two BioGraphs (one paired, one not) written by bgx-create -- or by the REFERENCE itself: BioGraphs users already
    have --, merged by bgx-merge -- host side end to end: the
    facade reads the input .bg directories (the mock checks what it was handed), writes the merged seqset, migrates
    both readmaps, writes the two-sample metadata -- with the device results served from the merge oracle; then the
    reference's biograph_dir + seqset + readmap open the merged directory sample by sample, and every migrated row
    still points at an entry that starts with its read"""
    from oracle import merge as M
    from tests.test_ref_merge import words
    exe = os.path.join(ROOT, "biograph_b200", "bgx-merge")
    if not os.path.exists(exe):
        pytest.fail("bgx-merge is missing: run __graft_entry__.build()")
    genome_reads = reads_of(6000, 5000, 100, 0.01, seed=95)
    if inputs == "bgx-create":
        a_bg, a_ss, a_ents, a_rm = _bgx_create_under_mock(writer, tmp_path, "SAMPLE_A", genome_reads[:2600], True)
        b_bg, b_ss, b_ents, b_rm = _bgx_create_under_mock(writer, tmp_path, "SAMPLE_B", genome_reads[2000:], False)
    else:
        clean = reads_of(6000, 3000, 100, 0.0, seed=96)
        a_bg, a_ss, a_ents, a_rm = _reference_written_bg(tmp_path, "SAMPLE_A", clean[:1600], True)
        other = reads_of(5000, 1500, 90, 0.0, seed=97)          # another genome, plus a share of the first one's reads
        b_bg, b_ss, b_ents, b_rm = _reference_written_bg(tmp_path, "SAMPLE_B", other + clean[1200:1240], False)
    merged, bits = M.make_mergemap([a_ents, b_ents])
    assert len(merged) > max(len(a_ents), len(b_ents))          # overlapping, neither contains the other
    tb = M.merge_tables(merged)
    d = str(tmp_path / "merge_served")
    os.mkdir(d)
    write_tables(d, {"n": tb["n"], "sizes": tb["sizes"], "shared": tb["shared"], "fixed": tb["fixed"],
                     "prev": np.stack([words(tb["prev"][b]) for b in range(4)])})
    for p, (ss, mm) in enumerate(((a_ss, bits[0]), (b_ss, bits[1]))):
        ss["sizes"].astype(np.uint16).tofile(os.path.join(d, f"in_{p}_sizes.bin"))
        words(mm).tofile(os.path.join(d, f"mm_{p}.bin"))
    out = str(tmp_path / "family.bg")
    env = dict(os.environ, BGX_MOCK_TABLES=d, LD_LIBRARY_PATH=writer + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    run = subprocess.run([exe, "--in", a_bg, "--in", b_bg, "--out", out], env=env, capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    with R.Run(2) as r:
        got = r.open_seqset_file(os.path.join(out, "seqset"))
        assert got["n"] == tb["n"] and np.array_equal(got["sizes"], tb["sizes"]) and np.array_equal(got["shared"], tb["shared"])
        flat = r.flat()
    assert flat == merged
    for name, ents, rm in (("SAMPLE_A", a_ents, a_rm), ("SAMPLE_B", b_ents, b_rm)):
        bg = R.open_biograph(out, name)                    # readmap(seqset, path) CHECKs the uuid link
        assert bg["samples"] == 2 and bg["sample_accession"] == name and bg["accession_id"] == "SAMPLE_A+SAMPLE_B"
        assert bg["biograph_id"] == bg["seqset_uuid"] == bg["readmap_seqset_uuid"]
        assert bg["seqset_entries"] == len(merged) and bg["readmap_rows"] == rm["n_rows"]
        rows = R.read_readmap_file(bg["readmap_path"])
        assert np.array_equal(rows["read_lengths"], rm["read_lengths"].astype(np.int32))
        assert np.array_equal(rows["mate_loop_ptr"], rm["mate_loop_ptr"]) and np.array_equal(rows["is_forward"], rm["is_forward"])
        for i in range(len(rows["entry_id"])):
            ln = int(rows["read_lengths"][i])
            assert flat[int(rows["entry_id"][i])][:ln] == ents[int(rm["entry_id"][i])][:ln]

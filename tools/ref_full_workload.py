"""One-off (CPU only, minutes): the restated oracle against oracle/_ref -- the reference's OWN classes compiled from
its sources -- over a WHOLE bench workload, every stage output.  bench.py checks the CUDA path against the oracle
over the whole workload on the GPU box; this closes the chain at the same size: CUDA path == oracle == reference.

  python tools/ref_full_workload.py ecoli100x > profiles/r2u_oracle_vs_reference_ecoli100x.json
  python tools/ref_full_workload.py ecoli100x --ranks 2 > profiles/r2u_oracle_vs_reference_ecoli100x_x2.json
      (the reads of ranks 0 .. N-1 of a sharded N-GPU bench run, concatenated: what its parity block describes)
  --no-oracle: the reference alone (digests only)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from biograph_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.view(np.uint8).reshape(-1).data if a.size else b"").hexdigest()[:16]


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    name = argv[0] if argv else "ecoli100x"
    ranks = int(sys.argv[sys.argv.index("--ranks") + 1]) if "--ranks" in sys.argv else 1
    with_oracle = "--no-oracle" not in sys.argv
    if "--ranks" in sys.argv:
        argv = [a for a in argv if a != str(ranks)] or argv
    threads = bench.host_threads()
    reads = np.concatenate([bench.make_workload(name, r, None) for r in range(ranks)])
    cov = bench.WORKLOADS[name]["coverage"]
    buf, offs = synth.as_buffer(reads)
    rb = (buf.tobytes(), offs)
    del buf
    t0 = time.perf_counter()
    with R.Run(threads) as r:
        counts, solid = r.count_kmers(rb, 30, 5, genome_bases=max(1, reads.size // cov))
        t1 = time.perf_counter()
        rcr = r.correct(rb, 8, 2, 0.7)
        t2 = time.perf_counter()
        rss = r.make_seqset()
        t3 = time.perf_counter()
    mism = []
    m = (counts["fwd"].astype(np.int64) + counts["rev"]) >= 5
    if with_oracle:
        oc = O.count_kmers(rb, 30, threads=threads, prefilter_min=5)
        osol = O.solid_set(oc, 5)
        del oc
        ocr = O.correct_reads(rb, osol, 30, 8, 2, 0.7, threads=threads)
        oss = O.seqset_staged((ocr["seq"], ocr["offs"]), ocr["next_fwd"], ocr["next_rev"], threads=threads)
        if not np.array_equal(solid["kmers"], osol["kmers"]):
            mism.append("solid/kmers")
        for f in ("kmers", "fwd", "rev", "flags"):
            if not np.array_equal(counts[f][m], osol[f]):
                mism.append("counts/" + f)
        if rcr["seq"] != ocr["seq"]:
            mism.append("corrected/bases")
        for f in ("offs", "kept"):
            if not np.array_equal(rcr[f], ocr[f]):
                mism.append("corrected/" + f)
        if rss["n"] != oss["n"]:
            mism.append("seqset/num_entries")
        for t in ("sizes", "shared", "prev", "fixed"):
            if not np.array_equal(rss[t], oss[t]):
                mism.append("seqset/" + t)
        del osol, ocr, oss
    t4 = time.perf_counter()
    # digests of the REFERENCE's output, under the names and in the forms bench.py's `parity.sha256_16` uses for the
    # CUDA path's output on the GPU box: equal digests = equal bytes, without the two ever meeting on one machine
    digests = {"solid_kmers": sha(counts["kmers"][m]), "solid_counts": sha(np.stack([counts["fwd"][m], counts["rev"][m]])),
               "corrected_bases": sha(np.frombuffer(rcr["seq"], dtype=np.uint8))}
    rss["subaccum"], rss["accum"] = [], []
    for b in range(4):
        sub, acc, _ = O.bitcount_finalize(rss["prev"][b], rss["n"])
        rss["subaccum"].append(sub)
        rss["accum"].append(acc)
    for member, a in bench.seqset_members(rss):
        digests["seqset/" + member] = sha(a)
    print(json.dumps({
        "what": "oracle port vs oracle/_ref (the reference's own classes) over a whole bench workload, CPU only",
        "workload": name, "ranks": ranks, "reads": int(reads.shape[0]), "read_len": int(reads.shape[1]), "bases": int(reads.size),
        "host_threads": threads, "solid_kmers": int(len(solid["kmers"])), "corrected_reads": int(rcr["kept"].sum()),
        "entries": int(rss["n"]), "equal": (not mism) if with_oracle else None, "mismatches": mism, "sha256_16_reference": digests,
        "reference_seconds": {"count": round(t1 - t0, 1), "correct": round(t2 - t1, 1), "seqset": round(t3 - t2, 1),
                              "total": round(t3 - t0, 1), "bases_per_s": reads.size / (t3 - t0)},
        "oracle_port_seconds": round(t4 - t3, 1)}))


if __name__ == "__main__":
    main()

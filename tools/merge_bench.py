"""Times bgx_merge_seqsets on the GPU: two (or more) seqsets built from synthetic reads of overlapping
genome halves are merged; prints one JSON line with the stage times of the merge (CUDA events on the
context's stream, bgx_stats_json), the record counts and a check of the result against a GPU build over
the union of the reads (parallel_splits = 1: every table must be equal).

  python tools/merge_bench.py [--genome 2000000] [--cov 20] [--parts 2] [--reps 3]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biograph_b200 as B  # noqa: E402
from biograph_b200 import synth  # noqa: E402


def build(reads):
    with B.Bgx() as g:
        g.add_reads(synth.as_buffer(reads))
        g.seed_uncorrected()
        g.build_seqset()
        return g.export_seqset()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=2000000)
    ap.add_argument("--cov", type=float, default=20.0)
    ap.add_argument("--parts", type=int, default=2)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    genome = synth.random_genome(a.genome, seed=31)
    span = a.genome * 2 // (a.parts + 1)           # neighbouring parts overlap by half
    reads = []
    for p in range(a.parts):
        lo = p * span // 2
        sub = genome[lo:lo + span]
        n = int(a.cov * len(sub) / 150)
        reads.append(synth.simulate_reads(sub, n, read_len=150, error_rate=0.0, seed=40 + p, paired=False))
    t0 = time.time()
    parts = [build(r) for r in reads]
    whole = build(np.concatenate(reads))
    t_build = time.time() - t0
    best = None
    with B.Bgx() as g:
        for _ in range(a.reps):
            g.merge_seqsets(parts, parallel_splits=1)
            st = g.stats()
            if best is None or st["ms_merge_total"] < best["ms_merge_total"]:
                best = st
            m = g.export_seqset()
            g.reset_results()
        with_chunks = None
        g.merge_seqsets(parts)   # the reference's chunk rule: one more kernel (merge_prev_kernel)
        with_chunks = g.stats()
    equal = m["n"] == whole["n"] and all(np.array_equal(m[k], whole[k]) for k in ("sizes", "shared", "prev", "fixed"))
    n_in = int(sum(p["n"] for p in parts))
    bases_in = int(sum(int(p["sizes"].astype(np.int64).sum()) for p in parts))
    keys = ["ms_merge_total", "ms_merge_flatten", "ms_sort_radix", "ms_sort_ties", "ms_dedup", "ms_walk", "ms_tables"]
    out = {"bench": "bgx_merge_seqsets", "inputs": a.parts, "input_entries": n_in, "input_bases": bases_in,
           "merged_entries": int(m["n"]), "equals_build_over_all_reads": bool(equal),
           "stage_ms": {k[3:]: round(best.get(k, 0.0), 3) for k in keys},
           "stage_ms_reference_chunk_rule": {k[3:]: round(with_chunks.get(k, 0.0), 3) for k in keys},
           "input_entries_per_s": n_in / (best["ms_merge_total"] * 1e-3),
           # flatten: reads sizes + prev bits, five doubling rounds (2 gathers + 2 stores of 12 B), writes bases/4 + 16 B records
           "flatten_alg_bytes": int(n_in * (2 + 0.5 + 5 * 2 * 12 + 16) + bases_in / 4),
           "build_s_for_inputs": round(t_build, 2), "peak_device_bytes": best.get("peak_device_bytes")}
    out["flatten_gbs"] = out["flatten_alg_bytes"] / (best["ms_merge_flatten"] * 1e-3) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()

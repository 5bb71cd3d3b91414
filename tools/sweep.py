"""Scaling sweep (BASELINE.json configs[4]): k-mer counting + suffix sort + the whole path at a growing
number of reads, constant 30x coverage (the genome grows with the read count), one GPU.

  python tools/sweep.py [reads_in_millions ...]        default: 3.3 6.5 13 26

Reads are simulated ON THE GPU with torch (uniform starts, uniform strand, 0.5 % substitutions) so
that the sweep spends its time in the path, not in numpy; they reach the library as ASCII through
bgx_add_reads_ascii (packed on the device).  One JSON line per size: CUDA-event stage times,
bases/s for counting, for sort + dedup and for the whole path, algorithmic GB/s of the radix
passes.  Not a bench line: bench.py is the contract; this is the sweep table under profiles/."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biograph_b200 as B  # noqa: E402


def simulate_reads_gpu(n_reads, read_len=150, coverage=30, error=0.005, seed=0, chunk=1 << 20):
    """uint8 ASCII [n_reads, read_len] on the host (pinned), from a random genome with 5 % repeats"""
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    glen = max(10000, n_reads * read_len // coverage)
    genome = torch.randint(0, 4, (glen,), dtype=torch.uint8, device=dev, generator=gen)
    # 5 % of the length overwritten by copies of earlier segments (1-10 kb), as the chr20 workload
    n_rep = max(1, int(0.05 * glen / 5000))
    src = torch.randint(0, glen - 10000, (n_rep,), device=dev, generator=gen).tolist()
    dst = torch.randint(0, glen - 10000, (n_rep,), device=dev, generator=gen).tolist()
    ln = torch.randint(1000, 10000, (n_rep,), device=dev, generator=gen).tolist()
    for s, d, l in zip(src, dst, ln):
        genome[d:d + l] = genome[s:s + l].clone()
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    out = torch.empty((n_reads, read_len), dtype=torch.uint8).pin_memory()
    ar = torch.arange(read_len, device=dev)
    for lo in range(0, n_reads, chunk):
        m = min(chunk, n_reads - lo)
        starts = torch.randint(0, glen - read_len, (m,), device=dev, generator=gen)
        b = genome[starts[:, None] + ar[None, :]]
        rc = torch.rand((m,), device=dev, generator=gen) < 0.5
        b = torch.where(rc[:, None], 3 - b.flip(1), b)
        err = torch.rand((m, read_len), device=dev, generator=gen) < error
        sub = torch.randint(1, 4, (m, read_len), dtype=torch.uint8, device=dev, generator=gen)
        b = torch.where(err, (b + sub) & 3, b)
        out[lo:lo + m].copy_(lut[b.long()], non_blocking=True)
    torch.cuda.synchronize()
    del genome, b, err, sub, starts
    torch.cuda.empty_cache()
    return out, glen


def main():
    sizes = [float(a) for a in sys.argv[1:]] or [3.3, 6.5, 13, 26]
    for millions in sizes:
        n = int(millions * 1e6)
        t0 = time.perf_counter()
        reads, glen = simulate_reads_gpu(n, seed=int(millions * 10))
        t_gen = time.perf_counter() - t0
        L = reads.shape[1]
        offs = np.arange(n + 1, dtype=np.uint64) * L
        g = B.Bgx()
        t0 = time.perf_counter()
        g.add_reads((reads.numpy().reshape(-1), offs))
        t_up = time.perf_counter() - t0
        best = None
        for rep in range(2):  # first run warms the allocator
            g.reset_results()
            g.timer_start()
            g.run()
            ms = g.timer_stop()
            st = g.stats()
            best = (ms, st)
        ms, st = best
        bases = n * L
        sort_ms = st["ms_sort_radix"] + st["ms_sort_ties"] + st["ms_dedup"]
        line = {
            "reads": n, "genome": glen, "bases": bases, "ms_total": round(ms, 2),
            "gbases_per_s": round(bases / ms / 1e6, 2),
            "count_ms": round(st["ms_count_total"], 2), "count_gbases_per_s": round(bases / st["ms_count_total"] / 1e6, 2),
            "count_batches": int(st.get("count_batches", 1)),
            "correct_ms": round(st["ms_correct_total"], 2),
            "seqset_ms": round(st["ms_seqset_total"], 2),
            "sort_dedup_ms": round(sort_ms, 2),
            "sort_records": int(st["seeds"]),
            "sort_radix_ms": round(st["ms_sort_radix"], 2), "sort_radix_passes": int(st["sort_radix_passes"]),
            "sort_radix_alg_gbs": round(st["alg_bytes_sort_radix"] / st["ms_sort_radix"] / 1e6, 1),
            "entries": int(st["entries"]), "kmer_distinct": int(st["kmer_distinct"]), "kmer_solid": int(st["kmer_solid"]),
            "table_slots": int(st["count_table_slots"]),
            "host_s": {"simulate": round(t_gen, 1), "upload_ascii": round(t_up, 2)},
        }
        print(json.dumps(line), flush=True)
        g.close()
        del reads


if __name__ == "__main__":
    main()

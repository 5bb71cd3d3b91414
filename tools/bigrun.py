#!/usr/bin/env python
"""BASELINE configs 4 and 5: GRCh38-scale (and swept) synthetic builds on N GPUs, reads generated ON
THE DEVICE so that no host ever holds the ASCII.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 \
      tools/bigrun.py --genome-bases 3100000000 --coverage 30 [--reads R] [--runs 2] [--verify-reads 8000000]

* genome: `--genome-bases` i.i.d. uniform bases with 5 % of its length overwritten by copies of earlier
  1-10 kb segments (the read model of SURVEY 8d, seeds 38/39), generated identically on every rank's GPU;
* reads: rank r draws its share of ceil(coverage * G / 150) reads (150 bp, uniform start and strand, 0.5 %
  substitutions, seed 40 + rank), 2-bit packs them on the GPU and hands DEVICE pointers to
  bgx_add_reads_packed;
* one sharded build over all ranks (bgx_dist_init + bgx_run), timed with CUDA events, max over ranks;
* --verify-reads V: the same pipeline on a down-sampled input (genome and reads scaled to V reads in
  total) checked against a single-GPU build over the concatenated reads (bench.py's N>1 parity check).
Prints one JSON line per configuration (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import biograph_b200 as B  # noqa: E402

READ_LEN = 150


def make_genome(n_bases, seed, repeat_seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    genome = torch.randint(0, 4, (n_bases,), dtype=torch.uint8, device=device, generator=g)
    # 5 % of the length overwritten by copies of earlier segments of 1-10 kb
    rng = np.random.default_rng(repeat_seed)
    target, done = n_bases // 20, 0
    while done < target and n_bases > 20000:
        ln = int(rng.integers(1000, 10001))
        dst = int(rng.integers(ln, n_bases - ln))
        src = int(rng.integers(0, dst - ln + 1))
        genome[dst:dst + ln] = genome[src:src + ln].clone()
        done += ln
    return genome


def gen_reads_packed(genome, n_reads, seed, device, chunk=1 << 20, error=0.005):
    """[n_reads, 40] uint8 packed reads (dna_sequence byte order: 4 bases per byte, first base in the high
    bits; 150 bases + 10 zero pad = five 8-byte words) generated chunk by chunk on the device"""
    G = genome.numel()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_reads, 40), dtype=torch.uint8, device=device)
    ar = torch.arange(READ_LEN, device=device, dtype=torch.int64)
    for s0 in range(0, n_reads, chunk):
        n = min(chunk, n_reads - s0)
        start = torch.randint(0, G - READ_LEN + 1, (n,), device=device, generator=g, dtype=torch.int64)
        rc = torch.randint(0, 2, (n,), device=device, generator=g, dtype=torch.uint8).bool()
        bases = genome[start[:, None] + ar[None, :]]
        bases = torch.where(rc[:, None], (3 - bases).flip(1), bases)
        err = torch.rand((n, READ_LEN), device=device, generator=g) < error
        sub = torch.randint(1, 4, (n, READ_LEN), device=device, generator=g, dtype=torch.uint8)
        bases = torch.where(err, (bases + sub) & 3, bases)
        pad = torch.zeros((n, 160), dtype=torch.uint8, device=device)
        pad[:, :READ_LEN] = bases
        q = pad.view(n, 40, 4)
        out[s0:s0 + n] = (q[:, :, 0] << 6) | (q[:, :, 1] << 4) | (q[:, :, 2] << 2) | q[:, :, 3]
        del start, rc, bases, err, sub, pad, q
    return out


def build(world, rank, local, packed, runs, free_input=None):
    n = packed.shape[0]
    lens = torch.full((n,), READ_LEN, dtype=torch.int16, device=packed.device)
    g = B.Bgx(device=local)
    if world > 1:
        ids = [B.Bgx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        g.dist_init(world, rank, ids[0])
    torch.cuda.synchronize()
    g.add_reads_packed_ptr(packed.data_ptr(), None, None, lens.data_ptr(), n)
    torch.cuda.synchronize()
    del lens, packed
    if free_input is not None:   # the library holds its own copy now
        free_input()
    best, stats, all_ms = None, None, []
    for i in range(runs):
        g.reset_results()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        g.timer_start()
        g.run()
        ms = g.timer_stop()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        all_ms.append(round(t.item(), 2))
        if best is None or t.item() < best:
            best, stats = t.item(), g.stats()
    stats["all_runs_ms"] = all_ms   # the first runs map the peers' exchange buffers (CUDA IPC): cold
    return g, best, stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-bases", type=int, default=3_100_000_000)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--reads", type=int, default=0, help="total reads (default: coverage * genome / 150)")
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--verify-reads", type=int, default=0)
    ap.add_argument("--tag", default="grch38_30x")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def one(genome_bases, total_reads, tag, verify):
        t0 = time.time()
        genome = make_genome(genome_bases, 38, 39, dev)
        per = -(-total_reads // world)
        packed = gen_reads_packed(genome, per, 40 + rank, dev)
        del genome
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        t_gen = time.time() - t0
        host_copy = packed.cpu().reshape(-1) if verify else None
        holder = [packed]
        del packed

        def free_input():
            holder.clear()
            torch.cuda.empty_cache()
        g, ms, st = build(world, rank, local, holder[0], args.runs, free_input)
        peak = torch.tensor([float(st.get("peak_device_bytes", 0))], dtype=torch.float64, device="cuda")
        ent = torch.tensor([float(st.get("entries", 0))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(peak, op=dist.ReduceOp.MAX)
            dist.all_reduce(ent, op=dist.ReduceOp.SUM)
        parity = None
        if verify:
            pinned = host_copy
            lens = torch.full((per,), READ_LEN, dtype=torch.int16).view(torch.uint8)
            if world > 1:
                parity = bench.verify_against_single_gpu(g, B, dist, rank, world, local, pinned, None, lens, per)
            else:
                parity = {"checked": False, "why": "one GPU"}
        bases = per * world * READ_LEN
        peak_hbm, _ = bench.measured_peak_gbs()
        line = {"config": tag, "n_gpus": world, "reads": per * world, "read_len": READ_LEN, "genome_bases": genome_bases,
                "bases": bases, "ms": ms, "value": bases / (ms / 1e3), "unit": "bases/s", "entries": int(ent.item()),
                "all_runs_ms": st.pop("all_runs_ms"), "gen_s": round(t_gen, 1), "peak_device_gb": round(peak.item() / 2**30, 2), "parity": parity,
                "stage_ms": {k[3:]: round(v, 2) for k, v in st.items() if k.startswith("ms_")},
                "counters": {k: v for k, v in st.items() if not k.startswith(("ms_", "hostms_", "alg_bytes_")) and not isinstance(v, list)}}
        # sort/dedup roofline of rank 0: radix passes + tie groups + dedup, algorithmic bytes over time
        sd_ms = sum(st.get("ms_" + k, 0.0) for k in ("sort_radix", "sort_ties", "dedup"))
        sd_bytes = st.get("alg_bytes_sort_radix", 0.0) + st.get("alg_bytes_dedup", 0.0)
        if sd_ms:
            line["sort_dedup"] = {"ms": round(sd_ms, 2), "alg_bytes": sd_bytes, "achieved_gbs": sd_bytes / sd_ms / 1e6,
                                  "frac_of_measured_hbm_peak": sd_bytes / sd_ms / 1e6 / peak_hbm}
        g.close()
        torch.cuda.empty_cache()
        if rank == 0:
            print(json.dumps(line), flush=True)

    total = args.reads or int(-(-args.coverage * args.genome_bases // READ_LEN))
    if args.verify_reads:
        frac = args.verify_reads / total
        one(max(200000, int(args.genome_bases * frac)), args.verify_reads, args.tag + "_downsampled_verify", True)
    one(args.genome_bases, total, args.tag, False)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

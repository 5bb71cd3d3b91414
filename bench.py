#!/usr/bin/env python
"""bench.py -- seqset construction throughput (input bases/sec to finished seqset) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload chr20_30x|ecoli100x|small] [--no-verify]
  python bench.py --impl reference ...      # the restated CPU path (oracle) on the host cores

The default workload is BASELINE.json configs[2], the largest single-GPU configuration (synthetic human
chr20 at 30x); configs[1] (E. coli 100x) is --workload ecoli100x.  After the timed region the output
of the whole workload is checked (--no-verify skips it): at N=1 against the CPU oracle run over the
WHOLE workload (its time is the cpu_baseline), at N>1 against a 1-GPU build over the concatenated
reads of all ranks; the JSON line carries the verdict and a sha256 per seqset member ("parity").

A "step" is one pass of the hot path (k-mer count -> correct -> seqset tables) over the whole
synthetic read set of the workload.
  value : bases/s with the 2-bit packed reads already resident in HBM when the timed region
          starts (tables left in HBM).
  e2e   : bases/s through the C ABI with HOST buffers: H2D of the packed reads from pinned
          memory, the three stages, and D2H of every payload member of the seqset file (entry sizes and
          shared lengths as their packed_varbit_vector elements), all inside the timed region.
Timing: CUDA events recorded on the library's own stream (bgx_timer_start/stop), max over ranks;
L2 is flushed between steps and the working set (GBs of table + records) is far larger than L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: synthetic E. coli 100x 150bp paired reads, 0.5% error
    "ecoli100x": dict(genome="ecoli", coverage=100, read_len=150, error=0.005, seed=20260101, paired=True),
    # configs[2]: synthetic human chr20 30x
    "chr20_30x": dict(genome=("random", 64444167, 20, 21), coverage=30, read_len=150, error=0.005, seed=22,
                      paired=False),
    "small": dict(genome=("random", 200000, 5, 6), coverage=100, read_len=150, error=0.005, seed=7, paired=True),
}


def make_workload(name, rank=0, reads_override=None, genome_prefix_reads=None):
    """genome_prefix_reads: build the reads, at the workload's coverage and read model, over only a
    prefix of the genome sized to give that many reads (the bounded CPU-baseline sample)."""
    from biograph_b200 import synth
    w = WORKLOADS[name]
    if w["genome"] == "ecoli" and rank > 0:
        # weak scaling: rank r sequences its OWN genome of the same size (random, 5 % repeats), so the
        # sharded build over all ranks sees N genomes: k-mers, seeds and entries all grow with N
        genome = synth.random_genome(4938920, seed=5000 + rank, repeat_frac=0.05, repeat_seed=6000 + rank)
    elif w["genome"] == "ecoli":
        z = np.load(os.path.join(ROOT, "tests", "golden", "e_coli_genome.npz"))
        genome = synth.unpack_genome(z["packed"], int(z["length"]))
    elif rank > 0:
        _, glen, s1, s2 = w["genome"]
        genome = synth.random_genome(glen, seed=s1 + 100 * rank, repeat_frac=0.05, repeat_seed=s2 + 100 * rank)
    else:
        _, glen, s1, s2 = w["genome"]
        genome = synth.random_genome(glen, seed=s1, repeat_frac=0.05, repeat_seed=s2)
    if genome_prefix_reads:
        glen = max(2000, genome_prefix_reads * w["read_len"] // w["coverage"])
        genome = genome[:min(len(genome), glen)]
    n_reads = -(-w["coverage"] * len(genome) // w["read_len"])
    if reads_override:
        n_reads = reads_override
    # weak scaling: every rank brings a read set of the same shape from its own genome
    reads = synth.simulate_reads(genome, n_reads, read_len=w["read_len"], error_rate=w["error"],
                                 seed=w["seed"] + 1000 * rank, paired=w["paired"])
    return reads


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    """all the host cores this process may use -- regardless of OMP_NUM_THREADS (torchrun sets it to 1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_port_run(reads2d, threads, k=30, keep=False):
    """The restated CPU path (oracle port) on a read set, the reference's way: two-stage k-mer counter
    (probabilistic pass, then exact CAS table) -> filter -> recursive correction -> staged seqset (seed +
    stride-7/255 + stride-1/6 rounds).  Returns (seconds, entries[, results])."""
    from oracle import oracle as O
    from biograph_b200 import synth
    buf, offs = synth.as_buffer(reads2d)
    rb = (buf.tobytes(), offs)
    del buf
    t0 = time.perf_counter()
    counts = O.count_kmers(rb, k, threads=threads, prefilter_min=5)
    solid = O.solid_set(counts, 5)
    del counts
    cr = O.correct_reads(rb, solid, k, 8, 2, 0.7, threads=threads)
    ss = O.seqset_staged((cr["seq"], cr["offs"]), cr["next_fwd"], cr["next_rev"], threads=threads)
    dt = time.perf_counter() - t0
    if keep:
        return dt, ss["n"], {"solid": solid, "corrected": cr, "seqset": ss}
    return dt, ss["n"]


def cpu_ref_run(reads2d, threads, coverage, k=30, keep=False, workload_bases=None):
    """The REFERENCE'S OWN classes (oracle/_ref/libref.so, compiled from /root/reference: kmer_counter -> kmer_set ->
    build_seqset::correct_reads -> expander x 4 -> builder -> seqset; oracle/ref_shim.cpp) on a read set, with the
    create flow's defaults (k 30, min count 5, 8 corrections, good run 2, trim 0.7).  Returns (seconds, entries,
    stage seconds[, results])."""
    from oracle import ref as R
    from biograph_b200 import synth
    buf, offs = synth.as_buffer(reads2d)
    rb = (buf.tobytes(), offs)
    del buf
    # The create flow gives the k-mer counter the process's whole memory budget (biograph_create.cpp:546: --max-mem,
    # 48 GiB by default), which sizes its tables.  A SAMPLE of the workload gets the same share of that budget as it is
    # of the workload: the whole budget on a 1/20 sample means tables 20 times sparser than in the run being modelled
    # (measured on 600 k reads of chr20: count stage 20.7 s with 48 GiB, 14.6 s with the sample's share).
    budget = 0
    if workload_bases and workload_bases > reads2d.size:
        budget = max(256 << 20, int((48 << 30) * (reads2d.size / float(workload_bases))))
    t0 = time.perf_counter()
    with R.Run(threads) as r:
        counts, solid = r.count_kmers(rb, k, 5, counter_max_memory_bytes=budget,
                                      genome_bases=max(1, int(reads2d.size // max(1, coverage))))
        t1 = time.perf_counter()
        cr = r.correct(rb, 8, 2, 0.7)
        t2 = time.perf_counter()
        ss = r.make_seqset()
        t3 = time.perf_counter()
    st = {"count_s": round(t1 - t0, 2), "correct_s": round(t2 - t1, 2), "seqset_s": round(t3 - t2, 2)}
    if keep:
        return t3 - t0, ss["n"], st, {"counts": counts, "solid": solid, "corrected": cr, "seqset": ss}
    return t3 - t0, ss["n"], st


def arm_config(workload, n_reads, read_len, world):
    """`config` of the JSON line: the job both arms are measured on (the reference arm prints the same dict, as the
    contract asks: "on your arm's config"; how ITS run was parallelised is in `host_parallelism`)"""
    return {"workload": workload, "reads_per_gpu": int(n_reads), "read_len": int(read_len),
            "bases_per_gpu": int(n_reads) * int(read_len), "kmer_size": 30,
            "parallelism": (f"one sharded build over {world} GPUs: reads split by rank ({world} genomes of this size), k-mers "
                            "routed by hash partition, suffix records by prefix range (NCCL all-to-all)")
            if world > 1 else "1 gpu",
            "l2": "flushed between steps (256 MiB memset); working set >> L2"}


def cpu_ref_run_isolated(workload, sample_reads, threads, coverage):
    """cpu_ref_run(keep=True) in a child process (`bench.py --_ref-sample`): the reference's code CHECK-fails by
    design where it finds something it does not expect, and that must not take the bench line with it.  The child
    regenerates the (deterministic) sample, runs the reference's classes and leaves what the parity check needs in an
    .npz.  Returns (seconds, entries, stage seconds, results) like cpu_ref_run."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory(prefix="bgx_bench_ref_") as d:
        out = os.path.join(d, "ref.npz")
        cmd = [sys.executable, os.path.abspath(__file__), "--_ref-sample", out, "--workload", workload,
               "--cpu-sample-reads", str(sample_reads), "--_threads", str(threads)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
        if r.returncode != 0 or not os.path.exists(out):
            raise RuntimeError(f"reference sample run failed (rc {r.returncode}): {r.stderr[-300:]}")
        z = np.load(out)
        meta = json.loads(str(z["meta"]))
        res = {"counts": {f: z["c_" + f] for f in ("kmers", "fwd", "rev", "flags")},
               "solid": {"kmers": z["s_kmers"], "flags": z["s_flags"]},
               "corrected": {"seq": z["cr_seq"].tobytes(), "offs": z["cr_offs"], "kept": z["cr_kept"]},
               "seqset": {"n": int(meta["entries"]), "sizes": z["ss_sizes"], "shared": z["ss_shared"], "prev": z["ss_prev"],
                          "fixed": z["ss_fixed"]}}
        return meta["seconds"], int(meta["entries"]), meta["stage_s"], res


def ref_sample_child(args):
    """the child of cpu_ref_run_isolated"""
    w = WORKLOADS[args.workload]
    sub = make_workload(args.workload, 0, None, genome_prefix_reads=args.cpu_sample_reads)
    total = -(-w["coverage"] * (4938920 if w["genome"] == "ecoli" else w["genome"][1]) // w["read_len"])
    dt, n_ent, stage, res = cpu_ref_run(sub, args._threads or host_threads(), w["coverage"], keep=True,
                                        workload_bases=int(total) * int(sub.shape[1]))
    c = res["counts"]
    m = (c["fwd"].astype(np.int64) + c["rev"]) >= 5   # the parity check reads the solid part only
    np.savez(args._ref_sample, meta=json.dumps({"seconds": dt, "entries": int(n_ent), "stage_s": stage}),
             c_kmers=c["kmers"][m], c_fwd=c["fwd"][m], c_rev=c["rev"][m], c_flags=c["flags"][m],
             s_kmers=res["solid"]["kmers"], s_flags=res["solid"]["flags"],
             cr_seq=np.frombuffer(res["corrected"]["seq"], dtype=np.uint8), cr_offs=res["corrected"]["offs"],
             cr_kept=res["corrected"]["kept"], ss_sizes=res["seqset"]["sizes"], ss_shared=res["seqset"]["shared"],
             ss_prev=res["seqset"]["prev"], ss_fixed=res["seqset"]["fixed"])


def reference_available():
    try:
        from oracle import ref as R
        return R.available()
    except Exception:
        return False


def verify_sample_against_reference(B, device, sub, ref):
    """A GPU build of the CPU sample against what the reference's own code computed for it: solid k-mers with their
    counts, surviving corrected reads, every seqset table.  Never raises: a failure is reported in the line."""
    try:
        from biograph_b200 import synth
        buf, offs = synth.as_buffer(sub)
        mism = []
        with B.Bgx(device=device) as g2:
            g2.add_reads((buf, offs))
            g2.run()
            gs = g2.export_kmers(5)
            cr = g2.export_corrected()
            ss = g2.export_seqset()
        c = ref["counts"]
        m = (c["fwd"].astype(np.int64) + c["rev"]) >= 5
        if not np.array_equal(ref["solid"]["kmers"], gs["kmers"]):
            mism.append("solid_kmers/kmers")
        for f in ("kmers", "fwd", "rev", "flags"):
            if not np.array_equal(c[f][m], gs[f]):
                mism.append("counts/" + f)
        rcr = ref["corrected"]
        if rcr["seq"] != cr["seq"]:
            mism.append("corrected/bases")
        for f in ("offs", "kept"):
            if not np.array_equal(rcr[f], cr[f]):
                mism.append("corrected/" + f)
        rss = ref["seqset"]
        if rss["n"] != ss["n"]:
            mism.append("seqset/num_entries")
        for t in ("sizes", "shared", "prev", "fixed"):
            if not np.array_equal(rss[t], ss[t]):
                mism.append("seqset/" + t)
        return {"checked": True, "against": "oracle/_ref (the reference's own classes), the cpu_baseline sample",
                "reads": int(sub.shape[0]), "entries": int(ss["n"]), "members_equal": not mism, "mismatches": mism}
    except Exception as e:  # noqa: BLE001
        return {"checked": False, "error": f"{type(e).__name__}: {e}"[:300]}


def reference_full_workload_digests(workload, parity, ranks=1):
    """The reference's own code over the WHOLE workload takes minutes on a CPU, so it was run once in the development
    container (tools/ref_full_workload.py) and its per-member digests committed under profiles/: equal digests = equal
    bytes.  ranks > 1: the concatenated reads of a sharded run's ranks (its parity block holds the digests of the
    assembled tables under the bare member names).  Returns the comparison for the bench line (None: no committed
    digests for this workload)."""
    try:
        rf = os.path.join(ROOT, "profiles", f"r2u_oracle_vs_reference_{workload}" + (f"_x{ranks}" if ranks > 1 else "") + ".json")
        if not os.path.exists(rf) or os.path.getsize(rf) == 0:
            return None
        rj = json.load(open(rf))
        want = rj["sha256_16_reference"]
        have = parity["sha256_16"]
        if ranks > 1:   # the sharded line digests the seqset members only, without the "seqset/" prefix
            want = {k_[len("seqset/"):]: v for k_, v in want.items() if k_.startswith("seqset/")}
        diff = sorted(k_ for k_ in want if have.get(k_) != want[k_])
        return {"against": "oracle/_ref (the reference's own classes) over the whole workload; digests committed in "
                           + os.path.relpath(rf, ROOT), "reads": rj["reads"], "entries": rj["entries"],
                "digests_compared": len(want), "digests_equal": not diff and rj["entries"] == parity["entries"],
                "differing": diff}
    except Exception as e:  # noqa: BLE001
        return {"digests_equal": None, "error": f"{type(e).__name__}: {e}"[:300]}


def sha(a):
    import hashlib
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.view(np.uint8).reshape(-1).data if a.size else b"").hexdigest()[:16]


def seqset_members(ss, lo=None, n=None, lay=None):
    """(name, array) of every payload member of a seqset (or of the slice a rank of a sharded build
    holds: entries [lo, lo + n), bit vectors and their bitcount index from the same 512-entry groups)"""
    if lo is None:
        out = [("sizes", ss["sizes"]), ("shared", ss["shared"]), ("fixed", ss["fixed"])]
        for b in range(4):
            out += [(f"prev_{'ACGT'[b]}/bits", ss["prev"][b]), (f"prev_{'ACGT'[b]}/subaccum", ss["subaccum"][b]),
                    (f"prev_{'ACGT'[b]}/accum", ss["accum"][b])]
        return out
    w0, g0 = lo // 64, lo // 512
    out = [("sizes", ss["sizes"][lo:lo + n]), ("shared", ss["shared"][lo:lo + n]), ("fixed", ss["fixed"])]
    for b in range(4):
        out += [(f"prev_{'ACGT'[b]}/bits", ss["prev"][b][w0:w0 + lay["prev_words"]]),
                (f"prev_{'ACGT'[b]}/subaccum", ss["subaccum"][b][g0:g0 + lay["sub_words"]]),
                (f"prev_{'ACGT'[b]}/accum", ss["accum"][b][g0:g0 + lay["acc_words"]])]
    return out


def verify_against_oracle(g, reads, threads):
    """N=1: the oracle over the WHOLE workload, once; compares the solid k-mers (counts, flags), the
    corrected reads (+ seed counts) and every seqset member.  Returns (parity dict, cpu seconds)."""
    from oracle import oracle as O
    dt, n_ent, ref = cpu_port_run(reads, threads, keep=True)
    mism = []
    gs = g.export_kmers(5)
    for f in ("kmers", "fwd", "rev", "flags"):
        if not np.array_equal(ref["solid"][f], gs[f]):
            mism.append("solid_kmers/" + f)
    digests = {"solid_kmers": sha(gs["kmers"]), "solid_counts": sha(np.stack([gs["fwd"], gs["rev"]]))}
    del gs
    cr, ocr = g.export_corrected(), ref["corrected"]
    if ocr["seq"] != cr["seq"]:
        mism.append("corrected/bases")
    for f in ("offs", "kept", "next_fwd", "next_rev", "corrections"):
        if not np.array_equal(ocr[f], cr[f]):
            mism.append("corrected/" + f)
    digests["corrected_bases"] = sha(np.frombuffer(cr["seq"], dtype=np.uint8))
    del cr, ocr
    ss, oss = g.export_seqset(), ref["seqset"]
    if oss["n"] != ss["n"]:
        mism.append("seqset/num_entries")
    oss["subaccum"], oss["accum"] = [], []
    for b in range(4):
        sub, acc, _ = O.bitcount_finalize(oss["prev"][b], oss["n"])
        oss["subaccum"].append(sub)
        oss["accum"].append(acc)
    for (name, a), (_, b_) in zip(seqset_members(ss), seqset_members(oss)):
        digests["seqset/" + name] = sha(a)
        if not np.array_equal(a, b_):
            mism.append("seqset/" + name)
    par = {"checked": True, "against": "oracle port, full workload", "members_equal": not mism, "mismatches": mism,
           "entries": int(ss["n"]), "sha256_16": digests}
    return par, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  Where
    oracle/_ref/libref.so exists (the reference's classes compiled from its sources, oracle/Makefile) that is what
    runs ("kind": "reference"); the reference's Bazel build as a whole is not possible offline, so
    without the library the restated port runs instead ("kind": "port").  All host threads; each step a bounded
    sample of the workload sized so that the whole --steps/--warmup run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = host_threads()
    w = WORKLOADS[args.workload]
    total = args.reads or -(-w["coverage"] * (4938920 if w["genome"] == "ecoli" else w["genome"][1]) // w["read_len"])
    use_ref = reference_available()
    sample_reads = args.cpu_sample_reads
    if use_ref:  # ~3 Mbases/s on 16 threads: keep (warmup + steps) x sample near four minutes
        sample_reads = min(sample_reads, max(50000, int(args.cpu_sample_reads * 8 / max(1, args.warmup + args.steps))))
    # bounded sample: the workload's read model at the workload's coverage over a genome prefix
    sub = make_workload(args.workload, 0, None, genome_prefix_reads=min(total, sample_reads))
    sample = sub.shape[0]
    reads = sub
    times = []
    stage = None
    fell_back = None
    for i in range(args.warmup + args.steps):
        if use_ref:
            try:
                dt, n_ent, stage = cpu_ref_run(sub, threads, w["coverage"], workload_bases=int(total) * int(sub.shape[1]))
            except Exception as e:  # noqa: BLE001 -- the arm must print a line: finish on the restated port and say so
                use_ref, stage, times = False, None, []
                fell_back = f"{type(e).__name__}: {e}"[:200]
        if not use_ref:
            dt, n_ent = cpu_port_run(sub, threads)
        if i >= args.warmup:
            times.append(dt)
    bases = sub.size
    ms = 1e3 * float(np.mean(times))
    val = bases / (ms / 1e3)
    what = ("oracle/_ref: the reference's own classes compiled from its sources (kmer_counter, kmer_set, correct_reads, "
            "expander, builder, seqset; driver sequence restated in oracle/ref_shim.cpp)" if use_ref else
            "oracle port (oracle/_ref not built)" if fell_back is None else
            "oracle port (oracle/_ref failed: " + fell_back + ")")
    line = {"impl": "reference", "metric": "input bases/sec to finished seqset", "value": val, "unit": "bases/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": arm_config(args.workload, total, reads.shape[1], args.gpus),
            "host_parallelism": f"{threads} host threads (CPU path; no GPU)",
            "cpu_baseline": {"value": val, "unit": "bases/s", "cores": threads, "kind": "reference" if use_ref else "port",
                             "sample": f"{sample} reads at the workload's coverage over a genome prefix "
                                       f"(workload has {total}), whole path (2-stage count + correct + expand/sort/dedup "
                                       f"+ build), {what}", "entries": int(n_ent), "stage_s": stage},
            "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def verify_against_single_gpu(g, B, dist, rank, world, local, pinned, pinned_mask, pinned_lens, n_reads):
    """N>1: the tables the sharded build left on the ranks against ONE single-GPU build (rank 0's GPU,
    a fresh context without NCCL) over all ranks' reads concatenated in rank order.  Every rank hashes
    the members of its slice; rank 0 hashes the same slices of the single-GPU tables."""
    import torch
    part = g.export_seqset()
    lay = g.seqset_layout()
    mine = {"first": int(part["first"]), "n": int(part["n"]), "lay": lay,
            "sha": {name: sha(a) for name, a in seqset_members(part)}}
    del part
    # the sharded context goes away on EVERY rank before rank 0 starts its own build: closing it unmaps the
    # peers' exchange buffers and returns this rank's device memory (a block must not be freed while a peer
    # still has it mapped)
    g.close()
    torch.cuda.empty_cache()
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    # all ranks' packed reads to rank 0 (equal sizes by construction of the weak-scaling workload)
    have_mask = torch.tensor([0 if pinned_mask is None else 1], device="cuda")
    dist.all_reduce(have_mask, op=dist.ReduceOp.MAX)
    if have_mask.item():
        return {"checked": False, "why": "reads with N: the gather below assumes no N mask"}
    bufs = []
    for t in (pinned, pinned_lens):
        d = t.cuda(non_blocking=True)
        lst = [torch.empty_like(d) for _ in range(world)] if rank == 0 else None
        dist.gather(d, lst, dst=0)
        if rank == 0:
            bufs.append(torch.cat(lst).cpu().numpy())
        del d, lst
    torch.cuda.empty_cache()
    res = None
    if rank == 0:
        try:
            g1 = B.Bgx(device=local)
            lens_all = bufs[1].view(np.uint16)
            g1.add_reads_packed(bufs[0], None, None, lens_all)
            g1.run()
            ss = g1.export_seqset()
            mism = []
            total = sum(p_["n"] for p_ in parts)
            if ss["n"] != total:
                mism.append(f"num_entries {total} != {ss['n']}")
            else:
                for r_, p_ in enumerate(parts):
                    for name, a in seqset_members(ss, p_["first"], p_["n"], p_["lay"]):
                        if sha(a) != p_["sha"][name]:
                            mism.append(f"rank{r_}/{name}")
            res = {"checked": True, "against": f"1-GPU build over the {world * n_reads} concatenated reads (same library, no NCCL)",
                   "members_equal": not mism, "mismatches": mism[:16], "entries": int(ss["n"]),
                   "sha256_16": {name: sha(a) for name, a in seqset_members(ss)}}
            g1.close()
        except Exception as e:  # e.g. the concatenated input does not fit one GPU
            res = {"checked": False, "why": f"1-GPU build of the concatenated reads failed: {e}"}
    out = [res]
    dist.broadcast_object_list(out, src=0)
    return out[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bgx", choices=["bgx", "reference"])
    ap.add_argument("--workload", default="chr20_30x", choices=sorted(WORKLOADS))
    ap.add_argument("--no-verify", action="store_true", help="skip the parity check of the whole workload")
    ap.add_argument("--reads", type=int, default=None, help="override the read count (debug)")
    ap.add_argument("--cpu-sample-reads", type=int, default=600000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--_ref-sample", dest="_ref_sample", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--_threads", dest="_threads", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args._ref_sample:
        return ref_sample_child(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import biograph_b200 as B
    from biograph_b200 import bgx as bgxmod

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- workload: synthetic reads, 2-bit packed into pinned host memory (T0 of the clock) ----
    reads = make_workload(args.workload, rank, args.reads)
    n_reads, read_len = reads.shape
    bases = int(reads.size)
    packed, nmask, woffs, lens = bgxmod.pack_reads_2bit(reads)
    pinned = torch.empty(packed.nbytes, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = packed
    pinned_mask = None
    if nmask is not None:
        pinned_mask = torch.empty(nmask.nbytes, dtype=torch.uint8).pin_memory()
        pinned_mask.numpy()[:] = nmask.view(np.uint8)
    h2d = packed.nbytes + (nmask.nbytes if nmask is not None else 0) + lens.nbytes
    del packed

    g = B.Bgx(device=local)
    if world > 1:
        # ONE sharded build over all ranks' reads: NCCL inside libbgx, id carried by torch.distributed
        ids = [B.Bgx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        g.dist_init(world, rank, ids[0])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    pinned_lens = torch.empty(lens.nbytes, dtype=torch.uint8).pin_memory()
    pinned_lens.numpy()[:] = lens.view(np.uint8)

    def upload(overlap=False):
        g.add_reads_packed_ptr(pinned.data_ptr(), None if pinned_mask is None else pinned_mask.data_ptr(),
                               None, pinned_lens.data_ptr(), n_reads, overlap=overlap)

    # ---- device-resident arm ----------------------------------------------------------------------
    upload()
    stats_acc = {}
    for i in range(args.warmup):
        g.reset_results(); flush_l2(); g.run()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = g.launch_count()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for i in range(args.steps):
        g.reset_results()
        flush_l2()
        g.timer_start()
        g.run()
        dev_ms += g.timer_stop()
        st = g.stats()
        for k_, v in st.items():
            if isinstance(v, (int, float)) and not isinstance(v, bool):
                stats_acc[k_] = stats_acc.get(k_, 0.0) + v
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = g.launch_count() - launches0
    clocks = sampler.stop()
    ms_per_step = max_over_ranks(dev_ms / args.steps)
    total_bases = sum_over_ranks(bases)
    value = total_bases / (ms_per_step / 1e3)
    st_mean = {k_: v / args.steps for k_, v in stats_acc.items()}
    ss_n = int(sum_over_ranks(st_mean.get("entries", 0)))

    # ---- end-to-end arm: host buffers in, host tables out ------------------------------------------------
    e2e_ms = 0.0
    d2h = 0
    for i in range(max(1, min(args.warmup, 2)) + args.steps):
        g.clear_reads(); g.reset_results(); flush_l2()
        g.timer_start()
        upload(overlap=True)   # bgx_add_reads_packed_async: the PCIe copy runs under pass 1 of counting
        g.run()
        # every payload member of the seqset spiral file back on the host, in the file's own form: entry_sizes and
        # shared as packed_varbit_vector elements (seqset.cpp:27-33), prev bits + bitcount indexes, fixed
        out = g.export_seqset(per_entry=False)
        vb = [g.export_varbit(0), g.export_varbit(1)]
        ms = g.timer_stop()
        if i >= max(1, min(args.warmup, 2)):
            e2e_ms += ms
        d2h = sum(int(np.asarray(v).nbytes) for v in (out["prev"], out["fixed"], vb[0]["elements"], vb[1]["elements"])) + \
            sum(int(a.nbytes) for a in out["subaccum"]) + sum(int(a.nbytes) for a in out["accum"])
        del vb
    barrier()
    e2e_ms_step = max_over_ranks(e2e_ms / args.steps)
    e2e_val = total_bases / (e2e_ms_step / 1e3)

    # ---- roofline of the dominant kernel (live CUDA-event time inside the timed steps) -------------------------
    # Every stage below is ONE kernel (or one kernel per radix pass) timed by CUDA events on the library's
    # stream; algorithmic bytes as DESIGN.md section 3 defines them (reported by the library per run).
    peak, peak_src = measured_peak_gbs()
    kern_ms = {k_[3:]: v for k_, v in st_mean.items() if k_.startswith("ms_")}
    kernel_of = {"count_partition": "kmer_partition_kernel", "count_split": "kmer_subhist_kernel + kmer_split_kernel",
                 "count_kernel": "kmer_count_bins_kernel", "correct_probe": "probe_kernel",
                 "sort_radix": "radix_hist_kernel + onesweep_kernel x passes",
                 "dedup": "dedup_flag_kernel + scan + compact_pairs_kernel (2 rounds)"}
    K_inst = st_mean.get("kmer_instances", 0.0) if world == 1 else bases * (read_len - 29) / read_len
    alg = {
        "count_partition": st_mean.get("alg_bytes_count_partition", 0.0),
        "count_split": st_mean.get("alg_bytes_count_split", 0.0),
        "count_kernel": st_mean.get("alg_bytes_count_kernel", 0.0),
        # probe pass: the packed reads once + one 32-byte sector of the solid set per k-mer (SURVEY 8d "correct")
        "correct_probe": bases / 4.0 + 32.0 * K_inst,
        "sort_radix": st_mean.get("alg_bytes_sort_radix", 0.0),
        "dedup": st_mean.get("alg_bytes_dedup", 0.0),
    }
    stages = {}
    for name, b_ in alg.items():
        ms = kern_ms.get(name, 0.0)
        if ms > 0:
            stages[name] = {"ms": ms, "alg_bytes": b_, "achieved_gbs": b_ / ms / 1e6, "frac": b_ / ms / 1e6 / peak}
    # the sort/dedup stage as SURVEY 8d evaluates it: radix passes + tie groups + dedup, and the one-touch floor
    sd_ms = sum(kern_ms.get(k_, 0.0) for k_ in ("sort_radix", "sort_ties", "dedup"))
    if sd_ms > 0:
        sd_bytes = alg["sort_radix"] + alg["dedup"]
        stages["sort_dedup_stage"] = {"ms": sd_ms, "alg_bytes": sd_bytes, "achieved_gbs": sd_bytes / sd_ms / 1e6,
                                      "frac": sd_bytes / sd_ms / 1e6 / peak,
                                      "one_touch_floor_frac": st_mean.get("useful_bytes_sort", 0.0) / sd_ms / 1e6 / peak}
    dom = max(alg, key=lambda n_: kern_ms.get(n_, 0.0))
    roof = {"bound": "hbm", "kernel": kernel_of[dom], "stage": dom, "achieved": stages.get(dom, {}).get("achieved_gbs"), "peak": peak,
            "unit": "GB/s", "frac": stages.get(dom, {}).get("frac"), "traffic": None, "peak_source": peak_src,
            "stages": stages, "share_of_step": kern_ms.get(dom, 0.0) / (dev_ms / args.steps)}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr) and world == 1:  # ncu DRAM bytes per launch, captured per workload on one GPU
        try:
            t_ = json.load(open(tr)).get(args.workload, {})
            roof["traffic"] = t_.get(dom)
            roof["traffic_over_alg"] = {k_: t_[k_] / stages[k_]["alg_bytes"] for k_ in t_ if k_ in stages and stages[k_]["alg_bytes"]}
        except Exception:
            pass

    # ---- parity of the whole workload + CPU baseline (outside the timed region) ---------------------------------
    # The last end-to-end step left its results resident.  N=1: the oracle runs over the WHOLE workload
    # once: it is the checker and, timed, the cpu_baseline (same config).  N>1: a 1-GPU build over the
    # concatenated reads of all ranks must give the same tables as the sharded build just timed.
    cpu = None
    parity = {"checked": False, "why": "--no-verify"}
    if world == 1 and not args.no_verify:
        threads = host_threads()
        parity, dt = verify_against_oracle(g, reads, threads)
        cpu = {"value": n_reads * read_len / dt, "unit": "bases/s", "cores": threads, "kind": "port", "same_config": True,
               "sample": f"the whole workload ({n_reads} reads) once, whole path (2-stage count + correct + staged "
                         f"seqset), {dt:.1f} s; oracle port (the restated CPU path); this run is "
                         "also the parity check"}
        if args.reads is None:
            rfw = reference_full_workload_digests(args.workload, parity)
            if rfw is not None:
                parity["reference_full_workload"] = rfw
        if reference_available() and not args.no_cpu_baseline:
            # the reference's own code on a bounded sample: the CPU baseline proper, and a second parity anchor
            try:
                cov = WORKLOADS[args.workload]["coverage"]
                sub = make_workload(args.workload, 0, None, genome_prefix_reads=min(n_reads, args.cpu_sample_reads))
                rdt, r_ent, rstage, rres = cpu_ref_run_isolated(args.workload, sub.shape[0], threads, cov)
                cpu = {"value": sub.size / rdt, "unit": "bases/s", "cores": threads, "kind": "reference",
                       "same_config": False, "entries": int(r_ent), "stage_s": rstage,
                       "sample": f"{sub.shape[0]} reads at the workload's coverage over a genome prefix (workload has "
                                 f"{n_reads}), whole path once, {rdt:.1f} s; oracle/_ref = the reference's own classes "
                                 "compiled from its sources (oracle/ref_shim.cpp says what is and is not the reference's)",
                       "oracle_port_whole_workload": {"value": n_reads * read_len / dt, "unit": "bases/s",
                                                      "seconds": round(dt, 1)}}
                parity["reference_sample"] = verify_sample_against_reference(B, local, sub, rres)
                del rres
            except Exception as e:  # noqa: BLE001
                parity["reference_sample"] = {"checked": False, "error": f"{type(e).__name__}: {e}"[:300]}
    elif world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        sub = make_workload(args.workload, 0, None, genome_prefix_reads=min(n_reads, args.cpu_sample_reads))
        sample = sub.shape[0]
        dt, _ = cpu_port_run(sub, threads)
        cpu = {"value": sample * read_len / dt, "unit": "bases/s", "cores": threads, "kind": "port", "same_config": False,
               "sample": f"{sample} reads at the workload's coverage over a genome prefix (workload has {n_reads}), "
                         f"whole path (2-stage count + correct + staged seqset) once, {dt:.1f} s; oracle port "
                         "(oracle/_ref not built)"}
    elif world > 1 and not args.no_verify:
        parity = verify_against_single_gpu(g, B, dist, rank, world, local, pinned, pinned_mask, pinned_lens, n_reads)
        if rank == 0 and args.reads is None and isinstance(parity, dict) and "sha256_16" in parity:
            rfw = reference_full_workload_digests(args.workload, parity, world)
            if rfw is not None:
                parity["reference_full_workload"] = rfw

    if rank == 0:
        line = {
            "metric": "input bases/sec to finished seqset", "value": value, "unit": "bases/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": arm_config(args.workload, n_reads, read_len, world),
            "entries": ss_n,
            "e2e": {"value": e2e_val, "unit": "bases/s", "ms_per_step": e2e_ms_step, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity": parity,
            "clocks": clocks,
            "wall_s_timed_region": wall_s,
            "stage_ms": {k_: round(v, 3) for k_, v in sorted(kern_ms.items())},
            # host wall clock over the same spans: a gap to the device time is host-side stall (syncs, NCCL setup)
            "host_stage_ms": {k_[7:]: round(v, 3) for k_, v in sorted(st_mean.items()) if k_.startswith("hostms_")},
            "counters": {k_: v for k_, v in sorted(st_mean.items())
                         if not k_.startswith(("ms_", "hostms_", "alg_bytes_"))},
        }
        print(json.dumps(line))
    g.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""One rank of a multi-GPU parity run (launched by torchrun from tests/test_dist.py).

Every rank takes a contiguous block of the reads, the build runs collectively through the C ABI
(NCCL inside libbgx), rank 0 gathers the per-rank outputs, assembles them and compares with the CPU
oracle run on ALL reads: k-mer counts, corrected reads and every seqset table must be identical
to the single-process result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import biograph_b200 as B  # noqa: E402
from biograph_b200 import synth  # noqa: E402


def make_reads(case):
    if case == "small":
        genome = synth.random_genome(30000, seed=5, repeat_frac=0.1, repeat_seed=6)
        return synth.simulate_reads(genome, 12000, read_len=100, error_rate=0.005, seed=3, paired=True, frag_mean=250,
                                    frag_sd=20)
    if case == "tiny":  # fewer entries than ranks * 512: some ranks end up empty after the rebalance
        genome = synth.random_genome(300, seed=9)
        return synth.simulate_reads(genome, 400, read_len=60, error_rate=0.0, seed=4, paired=False)
    if case == "n_and_ragged":
        genome = synth.random_genome(20000, seed=15)
        return synth.simulate_reads(genome, 9000, read_len=120, error_rate=0.01, seed=8, paired=True, frag_mean=300,
                                    frag_sd=30, n_rate=0.002)
    if case == "medium":
        genome = synth.random_genome(400000, seed=21, repeat_frac=0.05, repeat_seed=22)
        return synth.simulate_reads(genome, 200000, read_len=150, error_rate=0.005, seed=23, paired=True)
    raise ValueError(case)


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "small"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")  # only carries the NCCL id and the gathered results
    torch.cuda.set_device(local)
    reads = make_reads(case)
    n = reads.shape[0]
    lo, hi = n * rank // world, n * (rank + 1) // world
    ids = [B.Bgx.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    g = B.Bgx(device=local)
    g.dist_init(world, rank, ids[0])
    buf, offs = synth.as_buffer(reads[lo:hi])
    g.add_reads((buf, offs))
    g.count_kmers()
    km = g.export_kmers(1)
    g.correct()
    cr = g.export_corrected()
    g.build_seqset()
    ss = g.export_seqset()
    st = g.stats()
    g.close()
    gathered = [None] * world
    dist.gather_object({"km": km, "cr": cr, "ss": ss, "st": st}, gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        from oracle import oracle as O
        abuf, aoffs = synth.as_buffer(reads)
        rb = (abuf.tobytes(), aoffs)
        # k-mers: the union of what the ranks own, sorted
        ok_ = O.count_kmers(rb, 30)
        order = np.argsort(np.concatenate([p["km"]["kmers"] for p in gathered]), kind="stable")
        for f in ("kmers", "fwd", "rev", "flags"):
            got = np.concatenate([p["km"][f] for p in gathered])[order]
            assert np.array_equal(got, ok_[f]), f"k-mer {f} differs"
        # corrected reads: rank order == read order
        ocr = O.correct_reads(rb, O.solid_set(ok_, 5), 30)
        seq = b"".join(p["cr"]["seq"] for p in gathered)
        lens = np.concatenate([p["cr"]["lens"] for p in gathered])
        assert seq == ocr["seq"], "corrected bases differ"
        assert np.array_equal(lens, np.diff(ocr["offs"])), "corrected lengths differ"
        for f in ("next_fwd", "next_rev", "corrections"):
            assert np.array_equal(np.concatenate([p["cr"][f] for p in gathered]), ocr[f]), f
        # seqset
        whole = B.assemble_seqset([p["ss"] for p in gathered])
        oss = O.seqset_staged((ocr["seq"], ocr["offs"]), ocr["next_fwd"], ocr["next_rev"])
        assert whole["n"] == oss["n"], (whole["n"], oss["n"])
        for f in ("sizes", "shared", "prev", "fixed"):
            assert np.array_equal(whole[f], oss[f]), f"seqset {f} differs"
        for b in range(4):
            sub, acc, _ = O.bitcount_finalize(oss["prev"][b], oss["n"])
            assert np.array_equal(whole["subaccum"][b], sub), "subaccum differs"
            assert np.array_equal(whole["accum"][b], acc), "accum differs"
        parts = [p["ss"]["n"] for p in gathered]
        print(f"dist parity ok: case={case} world={world} reads={n} entries={whole['n']} per-rank={parts} "
              f"routed={[int(p['st'].get('route_records_out', 0)) for p in gathered]}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

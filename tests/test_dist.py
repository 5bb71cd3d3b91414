"""Multi-GPU path.

* CPU (`-m "not gpu"`): the host-side logic of the sharded build -- the rank layout rule
  (512-entry-aligned ranges) and the assembly of per-rank tables -- exercised over a world_size-2
  `gloo` group, with the oracle's tables standing in for what each rank's GPU would hand back.
* GPU (`-m gpu`, needs >= 2 devices): tests/dist_worker.py under torchrun, the real NCCL build,
  compared bit for bit with the oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_ranges(n_total, world):
    """the layout rule of stage_build_seqset_dist step 5"""
    chunk = max(512, (-(-n_total // world) + 511) // 512 * 512)
    return [(min(r * chunk, n_total), min((r + 1) * chunk, n_total)) for r in range(world)]


def _slice_part(ss, sub, acc, lo, hi, is_last_owner, n_total):
    n = hi - lo
    words = (n + 63) // 64
    subw = (n + 511) // 512
    accw = ((n_total + 1 + 511) // 512 - lo // 512) if is_last_owner else subw
    return {"n": n, "n_global": n_total, "first": lo, "max_entry_len": int(ss["sizes"].max()) if n_total else 0,
            "fixed": ss["fixed"], "sizes": ss["sizes"][lo:hi], "shared": ss["shared"][lo:hi],
            "prev": np.stack([ss["prev"][b][lo // 64: lo // 64 + words] for b in range(4)]),
            "subaccum": [sub[b][lo // 512: lo // 512 + subw] for b in range(4)],
            "accum": [acc[b][lo // 512: lo // 512 + accw] for b in range(4)]}


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import biograph_b200 as B
    from biograph_b200 import synth
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the id broadcast that precedes bgx_dist_init (the id itself is opaque bytes)
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    genome = synth.random_genome(3000, seed=1)
    reads = synth.simulate_reads(genome, 1500, read_len=80, error_rate=0.0, seed=2, paired=False)
    buf, offs = synth.as_buffer(reads)
    ss = O.seqset_closed_form((buf.tobytes(), offs))
    fin = [O.bitcount_finalize(ss["prev"][b], ss["n"]) for b in range(4)]
    sub, acc = [f[0] for f in fin], [f[1] for f in fin]
    ranges = _rank_ranges(ss["n"], world)
    last_owner = max(r for r in range(world) if ranges[r][1] > ranges[r][0])
    lo, hi = ranges[rank]
    part = _slice_part(ss, sub, acc, lo, hi, rank == last_owner, ss["n"])
    gathered = [None] * world
    dist.gather_object(part, gathered if rank == 0 else None, dst=0)
    if rank == 0:
        whole = B.assemble_seqset(gathered)
        ok = whole["n"] == ss["n"] and all(np.array_equal(whole[f], ss[f]) for f in ("sizes", "shared", "prev", "fixed"))
        ok = ok and all(np.array_equal(whole["subaccum"][b], sub[b]) and np.array_equal(whole["accum"][b], acc[b])
                        for b in range(4))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_layout_rule():
    for n_total, world in [(0, 2), (1, 2), (511, 2), (512, 2), (513, 4), (1024, 2), (100000, 8), (4096, 8)]:
        rr = _rank_ranges(n_total, world)
        assert rr[0][0] == 0 and rr[-1][1] == n_total
        for (a, b), (c, d) in zip(rr, rr[1:]):
            assert b == c
        for a, b in rr:
            assert a % 512 == 0 or a == n_total  # every non-empty range starts on a bitcount group boundary


@pytest.mark.parametrize("world", [2, 4])
def test_assemble_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", ["small", "tiny", "n_and_ragged"])
def test_multi_gpu_build_matches_oracle(case, world):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < world:
        pytest.skip(f"needs >= {world} GPUs (run under gpurun --gpus {world})")
    port = 29600 + (os.getpid() + 11 * world) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist parity ok" in r.stdout

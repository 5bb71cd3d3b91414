"""NCCL point-to-point bandwidth probe (torchrun, 2+ ranks): sizes the expectations for the
exchange steps of the sharded build."""
import os, time, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
peer = (rank + 1) % world
for mb in (1, 16, 128, 1024):
    a = torch.empty(mb << 20, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
    for nops in (1, 64):
        chunks_a, chunks_b = a.chunk(nops), b.chunk(nops)
        def step():
            ops = [dist.P2POp(dist.isend, x, peer) for x in chunks_a] + [dist.P2POp(dist.irecv, y, (rank - 1) % world) for y in chunks_b]
            for w in dist.batch_isend_irecv(ops): w.wait()
        for _ in range(3): step()
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        for _ in range(5): step()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        if rank == 0: print(f"{mb:5d} MiB in {nops:3d} ops: {dt*1e3:8.3f} ms  {mb/1024/dt:7.1f} GiB/s per direction", flush=True)
dist.destroy_process_group()

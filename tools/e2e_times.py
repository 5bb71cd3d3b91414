"""Wall-clock breakdown of the end-to-end arm (upload / run / export), per iteration.
usage: python tools/e2e_times.py [workload] [runs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import biograph_b200 as B  # noqa: E402
from biograph_b200 import bgx as bgxmod  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "ecoli100x"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
overlap = len(sys.argv) > 3 and sys.argv[3] == "overlap"
reads = bench.make_workload(wl)
packed, nmask, woffs, lens = bgxmod.pack_reads_2bit(reads)
pinned = torch.empty(packed.nbytes, dtype=torch.uint8).pin_memory()
pinned.numpy()[:] = packed
pl = torch.empty(lens.nbytes, dtype=torch.uint8).pin_memory()
pl.numpy()[:] = lens.view(np.uint8)
g = B.Bgx()
for i in range(runs):
    g.clear_reads(); g.reset_results()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g.timer_start()
    g.add_reads_packed_ptr(pinned.data_ptr(), None, None, pl.data_ptr(), len(lens), overlap=overlap)
    t1 = time.perf_counter()
    g.run()
    t2 = time.perf_counter()
    out = g.export_seqset()
    t3 = time.perf_counter()
    ms = g.timer_stop()
    st = g.stats()
    print(f"iter {i}: event {ms:.2f} ms | wall upload {1e3*(t1-t0):.2f} run {1e3*(t2-t1):.2f} export {1e3*(t3-t2):.2f} | "
          f"count {st['hostms_count_total']:.2f} correct {st['hostms_correct_total']:.2f} seqset {st['hostms_seqset_total']:.2f}", flush=True)
    del out
g.close()

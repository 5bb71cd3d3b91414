"""One upload + N runs of the hot path on a workload, for ncu captures (never a bench number).
usage: python tools/profile_step.py [workload] [runs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import biograph_b200 as B  # noqa: E402
from biograph_b200 import bgx as bgxmod  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "ecoli100x"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reads = bench.make_workload(wl)
packed, nmask, woffs, lens = bgxmod.pack_reads_2bit(reads)
g = B.Bgx()
g.add_reads_packed(packed, nmask, woffs, lens)
for i in range(runs):
    g.reset_results()
    n0 = g.launch_count()
    g.run()
    print("run", i, "launches", g.launch_count() - n0, flush=True)
print({k: round(v, 3) for k, v in g.stats().items() if k.startswith("ms_")})
g.close()

"""Prints per-stage CUDA-event times of N runs on a workload (quick experiment helper)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import biograph_b200 as B  # noqa: E402
from biograph_b200 import bgx as bgxmod  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "ecoli100x"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reads = bench.make_workload(wl)
packed, nmask, woffs, lens = bgxmod.pack_reads_2bit(reads)
g = B.Bgx()
g.add_reads_packed(packed, nmask, woffs, lens)
for i in range(runs):
    g.reset_results()
    g.timer_start()
    g.run()
    ms = g.timer_stop()
    st = g.stats()
    print(f"run {i}: {ms:.2f} ms  " + " ".join(f"{k[3:]}={v:.2f}" for k, v in st.items() if k.startswith("ms_")), flush=True)
print({k: v for k, v in st.items() if not k.startswith("ms_") and not k.startswith("hostms_")})
print("host:", {k[7:]: round(v, 2) for k, v in st.items() if k.startswith("hostms_")})
g.close()

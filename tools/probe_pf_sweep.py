"""A/B of the solid-set probe loads: L2 prefetch-size hint (BGX_PROBE_PF) x device L2 fetch granularity
(BGX_L2_FETCH).  usage: python tools/probe_pf_sweep.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import biograph_b200 as B  # noqa: E402
from biograph_b200 import bgx as bgxmod  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "chr20_30x"
reads = bench.make_workload(wl)
packed, nmask, woffs, lens = bgxmod.pack_reads_2bit(reads)
for l2 in ("32", "128"):
    os.environ["BGX_L2_FETCH"] = l2
    g = B.Bgx()
    g.add_reads_packed(packed, nmask, woffs, lens)
    for pf in ("0", "64", "128", "256", "1"):
        os.environ["BGX_PROBE_PF"] = pf
        best = None
        for i in range(2):
            g.reset_results()
            g.timer_start()
            g.run()
            ms = g.timer_stop()
            st = g.stats()
            if best is None or st["ms_correct_probe"] < best["ms_correct_probe"]:
                best = dict(st, total=ms)
        print(f"l2_fetch={l2} pf={pf}: total {best['total']:.2f} probe {best['ms_correct_probe']:.2f} correct_kernel "
              f"{best['ms_correct_kernel']:.2f} walk {best['ms_walk']:.2f} tables {best['ms_tables']:.2f} dedup {best['ms_dedup']:.2f} "
              f"ties {best['ms_sort_ties']:.2f}", flush=True)
    g.close()

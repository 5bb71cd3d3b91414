set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "pytest exit ${PIPESTATUS[0]}" ) > gpurun_out/q_pytest.log 2>&1
timeout 200 python - > gpurun_out/q_lookup_time.log 2>&1 <<'P'
import sys; sys.path.insert(0, '.')
import bench, biograph_b200 as B
from biograph_b200 import bgx as bgxmod
reads = bench.make_workload("ecoli100x")
packed, nmask, woffs, lens = bgxmod.pack_reads_2bit(reads)
g = B.Bgx(); g.add_reads_packed(packed, nmask, woffs, lens); g.run()
for i in range(3):
    f, r = g.lookup_reads()
    print("lookup_reads ms", g.stats().get("ms_lookup_reads"), "reads", len(f), "kept", int((f != 2**64-1).sum()), flush=True)
g.close()
P
tail -3 gpurun_out/q_pytest.log; tail -3 gpurun_out/q_lookup_time.log

"""DRAM traffic per stage (dram__bytes_read.sum + dram__bytes_write.sum, summed over the stage's launches of ONE
build) from `ncu --set full` reports -> profiles/traffic.json, keyed by workload; bench.py reads it for
roofline.traffic.   usage: python tools/make_traffic.py ecoli100x=gpurun_out/a.ncu-rep chr20_30x=gpurun_out/b.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

STAGE = [("kmer_partition_kernel", "count_partition"), ("kmer_subhist_kernel", "count_split"), ("kmer_split_kernel", "count_split"),
         ("kmer_count_bins_kernel", "count_kernel"), ("probe_kernel", "correct_probe"), ("correct_kernel", "correct_kernel"),
         ("radix_hist_kernel", "sort_radix"), ("onesweep_kernel", "sort_radix"), ("dedup_flag_kernel", "dedup"),
         ("compact_pairs_kernel", "dedup"), ("tie_small_kernel", "sort_ties"), ("walk_phase1_kernel", "walk"),
         ("tables_kernel", "tables")]


def units(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


out = {}
p = os.environ.get("TRAFFIC_JSON") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
if os.path.exists(p):
    out = {k: v for k, v in json.load(open(p)).items() if isinstance(v, dict)}
for arg in sys.argv[1:]:
    wl, rep = arg.split("=")
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, un, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    acc, launches = {}, {}
    for r in data:
        name = r[col["Kernel Name"]]
        for pat, stage in STAGE:
            if pat in name:
                b = units(r[col["dram__bytes_read.sum"]], un[col["dram__bytes_read.sum"]]) + \
                    units(r[col["dram__bytes_write.sum"]], un[col["dram__bytes_write.sum"]])
                acc[stage] = acc.get(stage, 0.0) + b
                launches[stage] = launches.get(stage, 0) + 1
                break
    out[wl] = {k: int(v) for k, v in acc.items()}
    out[wl]["_launches"] = launches
    out[wl]["_source"] = os.path.basename(rep)
out["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum of every launch of a stage in ONE build, from ncu --set full "
                   "--clock-control none (cold cache, serialised); written by tools/make_traffic.py, read by bench.py")
json.dump(out, open(p, "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])

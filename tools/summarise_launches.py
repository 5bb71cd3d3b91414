"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarise_launches.py gpurun_out/launches.csv [runs] > profiles/<name>.md
`runs` = number of hot-path runs in the capture (per-run columns are divided by it)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"source: {path}; {sum(v[0] for v in agg.values())} launches, {runs} run(s) of the hot path; "
      f"{tot / runs:.3f} ms of kernel time per run (ncu: cold-cache, serialised -- compare SHARES)\n")
print("| kernel | launches/run | ms/run | share |")
print("|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0] / runs:g} | {v[1] / runs:.3f} | {100 * v[1] / tot:.1f}% |")

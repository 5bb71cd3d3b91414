"""Key counters of every kernel in an `ncu --set full` report, as a markdown table.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/<name>.md   (needs ncu on PATH; no GPU)"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp insts"),
]

path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print(f"source: {path} (ncu --set full --clock-control none; per-launch, cold cache, serialised)\n")
print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
print("|---|" + "---:|" * len(WANT))
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("unnamed>::", "").replace("void ", "")
    cells = []
    for key, _ in WANT:
        if key not in col:
            cells.append("-")
            continue
        v, u = r[col[key]], units[col[key]]
        try:
            v = f"{float(v.replace(',', '')):.4g}"
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip() if u not in ("", "%") else v)
    print(f"| `{name}` | " + " | ".join(cells) + " |")

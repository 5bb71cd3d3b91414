"""Warp-stall samples of one kernel of an ncu report, summed per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [min_pct]"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}",
                      "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, hdr, agg, src = None, None, collections.Counter(), {}
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
    elif r[0] == "Line No":
        hdr, ci = r, r.index("# Samples")
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        s = int(r[ci] or 0)
        if s:
            agg[(cur, int(r[0]))] += s
            src[(cur, int(r[0]))] = r[1]
tot = sum(agg.values()) or 1
print(f"{tot} samples")
for (f, l), s in sorted(agg.items()):
    if 100 * s / tot >= min_pct:
        print(f"{f}:{l} {100 * s / tot:5.1f}%  {src[(f, l)].strip()[:110]}")

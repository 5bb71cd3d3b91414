"""Top SASS lines by warp-stall samples for one kernel of an ncu report.
usage: python tools/ncu_hot_sass.py report.ncu-rep kernel_regex [n]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break
    body.append(r)
c_s, c_src, c_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[c_s]) for r in body) or 1
print(f"{len(body)} SASS lines, {tot} samples")
for i, r in sorted(enumerate(body), key=lambda t: -int(t[1][c_s]))[:n]:
    print(f"{i:5d} {100*int(r[c_s])/tot:5.1f}%  ex={int(r[c_ex]):>10d}  {r[c_src].strip()[:110]}")

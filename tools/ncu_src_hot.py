"""Hot source lines of one kernel from an ncu report (warp-stall samples per CUDA line).
python tools/ncu_src_hot.py rep [kernel-regex] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else ".*"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]
si = hdr.index("# Samples")
ei = hdr.index("Instructions Executed")
lines = []
for r in rows[hi + 1:]:
    if len(r) > si and r[2] == "-" and r[si].isdigit():       # CUDA source line rows
        lines.append((int(r[si]), int(r[ei]) if r[ei].isdigit() else 0, r[0], r[1]))
tot = sum(l[0] for l in lines)
print("total samples", tot)
for s, e, no, src in sorted(lines, key=lambda x: -x[0])[:top]:
    print(f"{s:7d} {s / max(tot, 1):6.3f} inst={e:9d} L{no:>5s} {src.strip()[:110]}")

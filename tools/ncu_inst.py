"""Aggregate executed warp-instructions per CUDA source line (cuda,sass source page)."""
import csv, sys, collections, subprocess
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[2]
ie = hdr.index("Instructions Executed")
src = open(rows[0][1]).read().splitlines()
agg = collections.Counter(); line = None
for r in rows[3:]:
    if r and r[0].strip().isdigit():
        line = int(r[0]); continue
    if line is None or len(r) <= ie: continue
    try: agg[line] += int(r[ie])
    except ValueError: pass
tot = sum(agg.values())
print("total warp-instructions", tot)
for ln, s in agg.most_common(top):
    print(f"{ln:5d} {100.0*s/tot:5.1f}% | {src[ln-1].strip()[:110] if ln-1 < len(src) else ''}")

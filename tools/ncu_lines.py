"""Aggregate ncu warp-stall samples per CUDA source line (cuda,sass source page)."""
import csv, sys, collections, subprocess
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
src = None
try:
    path = rows[0][1]
    src = open(path).read().splitlines()
except Exception:
    pass
agg = collections.Counter(); nins = collections.Counter()
line = None
for r in rows[3:]:
    if r and r[0].strip().isdigit():
        line = int(r[0]); continue
    if line is None or len(r) < 5: continue
    try: s = int(r[4])
    except ValueError: continue
    agg[line] += s; nins[line] += 1
tot = sum(agg.values())
print("total samples", tot)
for ln, s in agg.most_common(top):
    text = src[ln - 1].strip()[:110] if src and ln - 1 < len(src) else ""
    print(f"{ln:5d} {100.0*s/tot:5.1f}% {nins[ln]:4d} ins | {text}")

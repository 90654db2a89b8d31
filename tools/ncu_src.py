"""Per-source-line totals of an ncu source page (cuda,sass): executed warp instructions, shared-memory
wavefronts and stall samples, keyed by (file, line).  Optional line ranges aggregate phases.
  python tools/ncu_src.py report.ncu-rep kernel_regex [top] [file_substr:lo-hi ...]"""
import csv, sys, collections, subprocess, io

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ranges = []
for a in sys.argv[4:]:
    f, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((f, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, line = None, None, None
ins = collections.Counter(); wav = collections.Counter(); smp = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]; hdr = None; line = None; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); iw = hdr.index("L1 Wavefronts Shared"); isx = hdr.index("# Samples")
        continue
    if hdr is None:
        continue
    if r[0].strip().isdigit():
        line = int(r[0]); continue
    if line is None or len(r) <= max(ie, iw):
        continue
    key = (cur_file, line)
    for col, acc in ((ie, ins), (iw, wav), (isx, smp)):
        try:
            acc[key] += int(r[col])
        except ValueError:
            pass
ti, tw, ts = sum(ins.values()), sum(wav.values()), sum(smp.values())
print(f"total warp-instructions {ti}  shared wavefronts {tw}  samples {ts}")
srcs = {}
def text(f, ln):
    if f not in srcs:
        try: srcs[f] = open(f).read().splitlines()
        except Exception: srcs[f] = []
    s = srcs[f]
    return s[ln - 1].strip()[:90] if 0 < ln <= len(s) else ""
for (f, ln), n in ins.most_common(top):
    print(f"{f.split('/')[-1]:>16s}:{ln:<4d} ins {100.0*n/ti:5.1f}%  wav {100.0*wav[(f,ln)]/max(tw,1):5.1f}%  smp {100.0*smp[(f,ln)]/max(ts,1):5.1f}% | {text(f, ln)}")
for f, lo, hi in ranges:
    a = sum(n for (ff, ln), n in ins.items() if f in ff and lo <= ln <= hi)
    b = sum(n for (ff, ln), n in wav.items() if f in ff and lo <= ln <= hi)
    c = sum(n for (ff, ln), n in smp.items() if f in ff and lo <= ln <= hi)
    print(f"range {f}:{lo}-{hi}: ins {100.0*a/ti:5.1f}%  wav {100.0*b/max(tw,1):5.1f}%  smp {100.0*c/max(ts,1):5.1f}%")

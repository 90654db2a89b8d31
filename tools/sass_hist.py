"""Per-kernel SASS opcode histogram of libvgpa_b200.so (cuobjdump -sass): the opcodes that show which
hardware paths a kernel uses -- DMMA (FP64 tensor core), UBLKCP / UBLKPF (1-D bulk TMA copy / L2 prefetch),
LDGSTS (cp.async), SYNCS (mbarrier), BAR, SHFL, MUFU, DFMA/DMUL/DADD, LDS/STS, LDG/STG -- plus the code size.
    python tools/sass_hist.py [lib] > profiles/sass_r02.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "vgpa_b200/libvgpa_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["DMMA", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "BAR", "SHFL", "MUFU", "DFMA", "DMUL", "DADD", "DSETP",
         "LDS", "STS", "LDG", "STG", "LDL", "STL", "UTMALDG", "UTCMMA", "HMMA", "WARPSYNC", "NOP"]
kernels = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("LDS", "STS", "LDG", "STG") and op.startswith(w)):
                kernels[cur][w] += 1
                break
print(f"# SASS opcode histogram of {lib} (sm_100a), one line per kernel: instructions, code bytes (16 B each), watched opcodes")
demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
for (name, c), pretty in zip(kernels.items(), demangle or kernels):
    short = pretty.replace("vgpa::(anonymous namespace)::", "").replace("vgpa::<unnamed>::", "").replace("(int)", "")
    short = re.sub(r"\(.*", "", short).replace("void ", "")
    ops = "  ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{short:<46s} {c['_total']:6d} ins {16 * c['_total']:7d} B   {ops}")
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print("TOTAL".ljust(46), f"{tot['_total']:6d} ins {16 * tot['_total']:7d} B   " + "  ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))
